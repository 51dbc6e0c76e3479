// tools/gpu/ubench_l2.cu -- the L2 roofline denominators of this B200 ("L2 peak not in
// MEASURED_PEAKS.json -> measure once and record it next to the results").  Prints one JSON object:
//   l2_stream_gbs      16-byte ld.global.cg (L1 bypassed) streaming reads of a 48 MB buffer that stays L2-resident
//   l2_gather_sector_gbs   random 4-byte ld.global.nc gathers from an L2-resident 1.5 MB table (K1's leaf table size before
//                      the tables moved to shared memory), counted as 32-byte sectors -- what lts__t_bytes counts
//   hbm_stream_gbs     the same streaming kernel over 4 GB (no reuse), for reference beside MEASURED_PEAKS.json
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/gpu/ubench_l2 tools/gpu/ubench_l2.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__global__ void __launch_bounds__(512) k_stream(const uint4* __restrict__ p, size_t n, int reps, unsigned* out) {
  unsigned acc = 0;
  for (int r = 0; r < reps; r++)
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
      const uint4 v = __ldcg(p + i);
      acc ^= v.x ^ v.y ^ v.z ^ v.w;
    }
  if (acc == 0x12345678u) out[0] = acc;
}

__device__ __forceinline__ unsigned mix(unsigned x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x;
}
__global__ void __launch_bounds__(1024, 2) k_gather(const float* __restrict__ tab, unsigned mask, int iters, float* out) {
  unsigned h = mix(blockIdx.x * blockDim.x + threadIdx.x + 1);
  float acc = 0.f;
  for (int it = 0; it < iters; it++) {
    unsigned idx[8];
#pragma unroll
    for (int g = 0; g < 8; g++) { h = h * 1664525u + 1013904223u; idx[g] = mix(h) & mask; }
#pragma unroll
    for (int g = 0; g < 8; g++) acc += __ldg(tab + idx[g]);
  }
  if (acc == 1234.5f) out[0] = acc;
}

static float time_ms(void (*launch)(void*), void* arg, int reps) {
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  float best = 1e30f;
  for (int r = 0; r < reps; r++) {
    CK(cudaEventRecord(a)); launch(arg); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b)); if (ms < best) best = ms;
  }
  return best;
}

struct StreamArg { const uint4* p; size_t n; int reps, grid; unsigned* out; };
static void launch_stream(void* a) { StreamArg* s = (StreamArg*)a; k_stream<<<s->grid, 512>>>(s->p, s->n, s->reps, s->out); }
struct GatherArg { const float* t; unsigned mask; int iters, grid; float* out; };
static void launch_gather(void* a) { GatherArg* s = (GatherArg*)a; k_gather<<<s->grid, 1024>>>(s->t, s->mask, s->iters, s->out); }

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  void* out; CK(cudaMalloc(&out, 64));
  // streaming, L2-resident: 48 MB read 40 times per launch (first pass from HBM: 2.5 % of the bytes)
  const size_t l2_bytes = 48u << 20;
  uint4* buf; CK(cudaMalloc(&buf, l2_bytes)); CK(cudaMemset(buf, 1, l2_bytes));
  StreamArg sa = {buf, l2_bytes / 16, 40, p.multiProcessorCount * 4, (unsigned*)out};
  launch_stream(&sa); CK(cudaDeviceSynchronize());
  const float ms_l2 = time_ms(launch_stream, &sa, 5);
  const double l2_stream = (double)l2_bytes * 40 / ms_l2 / 1e6;
  // streaming from HBM: 4 GB once
  const size_t big = (size_t)4 << 30;
  uint4* bbuf; CK(cudaMalloc(&bbuf, big)); CK(cudaMemset(bbuf, 1, big));
  StreamArg sb = {bbuf, big / 16, 1, p.multiProcessorCount * 8, (unsigned*)out};
  const float ms_hbm = time_ms(launch_stream, &sb, 5);
  const double hbm_stream = (double)big / ms_hbm / 1e6;
  // random 4-byte gathers from a 1.5 MB (2^18 + 2^17 floats -> mask to 1 MB) and a 32 MB table, both L2-resident
  double gather[2]; size_t gsz[2] = {(size_t)1 << 20, (size_t)32 << 20};
  for (int t = 0; t < 2; t++) {
    float* tab; CK(cudaMalloc(&tab, gsz[t])); CK(cudaMemset(tab, 0, gsz[t]));
    GatherArg ga = {tab, (unsigned)(gsz[t] / 4 - 1), 256, p.multiProcessorCount * 2, (float*)out};
    launch_gather(&ga); CK(cudaDeviceSynchronize());
    const float ms = time_ms(launch_gather, &ga, 5);
    gather[t] = (double)ga.grid * 1024 * 256 * 8 * 32 / ms / 1e6;  // 32-byte sectors
    CK(cudaFree(tab));
  }
  int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("{\"gpu\": \"%s\", \"sm_count\": %d, \"sm_clock_khz\": %d, \"l2_stream_gbs\": %.1f, \"hbm_stream_gbs\": %.1f, "
         "\"l2_gather_sector_gbs_1mb_table\": %.1f, \"l2_gather_sector_gbs_32mb_table\": %.1f, "
         "\"how\": \"tools/gpu/ubench_l2.cu: 16-byte ld.global.cg over 48 MB x 40 passes (L2-resident) / 4 GB once (HBM); random 4-byte "
         "ld.global.nc gathers counted as 32-byte sectors; best of 5, CUDA events\"}\n",
         p.name, p.multiProcessorCount, clk, l2_stream, hbm_stream, gather[0], gather[1]);
  return 0;
}
