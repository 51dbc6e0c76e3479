// tools/gpu/ubench_gather.cu -- what bounds K1's leaf gathers on this B200?
// Random 4-byte gathers, 6 independent ones per thread and iteration (like K1), from tables of several sizes
// through (a) ld.global.nc, (b) the texture path, (c) shared memory (64 KB table), (d) ld.global.nc.v4.
// Prints gathers/s and the implied 32-byte-sector bandwidth; the "L2 gather peak" DESIGN.md quotes.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ unsigned mix(unsigned x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x;
}

template <int MODE>
__global__ void __launch_bounds__(1024, 2) k(const float* __restrict__ tab, cudaTextureObject_t tex, unsigned mask, int iters,
                                             float* out) {
  extern __shared__ float s_tab[];
  if (MODE == 2) {
    for (int i = threadIdx.x; i < 16384; i += blockDim.x) s_tab[i] = tab[i];
    __syncthreads();
  }
  unsigned h = mix(blockIdx.x * blockDim.x + threadIdx.x + 1);
  float acc = 0.f;
  for (int it = 0; it < iters; it++) {
    unsigned idx[6];
#pragma unroll
    for (int g = 0; g < 6; g++) { h = h * 1664525u + 1013904223u; idx[g] = mix(h) & mask; }
#pragma unroll
    for (int g = 0; g < 6; g++) {
      if (MODE == 0) acc += __ldg(tab + idx[g]);
      else if (MODE == 1) acc += tex1Dfetch<float>(tex, (int)idx[g]);
      else if (MODE == 2) acc += s_tab[idx[g] & 16383];
      else { float4 v = __ldg(reinterpret_cast<const float4*>(tab) + (idx[g] >> 2)); acc += v.x + v.w; }
    }
  }
  if (acc == 1234.5f) out[0] = acc;
}

int main() {
  int dev = 0; cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, dev));
  const int grid = p.multiProcessorCount * 2, iters = 256;
  float* out; CK(cudaMalloc(&out, 4));
  const size_t sizes[] = {64u << 10, 256u << 10, 768u << 10, 1536u << 10, 16u << 20};
  const char* names[] = {"ld.global.nc.f32", "tex1Dfetch f32", "lds f32 (64KB)", "ld.global.nc.v4"};
  for (size_t sz : sizes) {
    const size_t n = sz / 4;  // floats; power of two or 3*2^k -> use mask on next lower pow2
    unsigned mask = 1; while ((size_t)mask * 2 <= n) mask *= 2; mask -= 1;
    float* tab; CK(cudaMalloc(&tab, sz)); CK(cudaMemset(tab, 0, sz));
    cudaResourceDesc rd = {}; rd.resType = cudaResourceTypeLinear; rd.res.linear.devPtr = tab;
    rd.res.linear.desc = cudaCreateChannelDesc<float>(); rd.res.linear.sizeInBytes = sz;
    cudaTextureDesc td = {}; td.readMode = cudaReadModeElementType;
    cudaTextureObject_t tex; CK(cudaCreateTextureObject(&tex, &rd, &td, NULL));
    for (int mode = 0; mode < 4; mode++) {
      if (mode == 2 && sz != (64u << 10)) continue;
      cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
      float best = 1e30f;
      for (int rep = 0; rep < 4; rep++) {
        CK(cudaEventRecord(a));
        if (mode == 0) k<0><<<grid, 1024>>>(tab, tex, mask, iters, out);
        if (mode == 1) k<1><<<grid, 1024>>>(tab, tex, mask, iters, out);
        if (mode == 2) { CK(cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536)); k<2><<<grid, 1024, 65536>>>(tab, tex, mask, iters, out); }
        if (mode == 3) k<3><<<grid, 1024>>>(tab, tex, mask, iters, out);
        CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b)); if (ms < best) best = ms;
      }
      const double gathers = (double)grid * 1024 * iters * 6;
      printf("table %6zu KB  %-18s %8.3f ms  %7.1f Ggather/s  (%.2f TB/s of 32B sectors, %.1f cyc/warp-gather/SM @1.9GHz)\n", sz >> 10,
             names[mode], best, gathers / best / 1e6, gathers * 32 / best / 1e9, 1.9e9 * best * 1e-3 * p.multiProcessorCount / (gathers / 32));
    }
    CK(cudaDestroyTextureObject(tex)); CK(cudaFree(tab));
  }
  return 0;
}
