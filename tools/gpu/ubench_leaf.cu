// tools/gpu/ubench_leaf.cu -- where should K1's leaf probabilities live?  Random 4-byte reads, 6 independent
// ones per thread and iteration (like K1), 1024-thread CTAs:
//   0  shared memory, 128 KB table                     (1 CTA/SM)
//   1  61 % of the lanes shared memory (128 KB), 39 % ld.global.nc from a 128 KB table (1 CTA/SM, 136 KB smem)
//   2  ld.global.nc only, 128 KB table, 136 KB of shared memory allocated (L1 ~ 120 KB)
//   3  ld.global.nc only, 256 KB table, 8 KB shared memory, 2 CTAs/SM (what k1_planes_phased does)
//   4  2-CTA cluster: 256 KB table split over the two CTAs' shared memory, remote half through DSMEM
//   5  like 3 with a 1.5 MB table (L2-resident: the unphased kernel)
// Prints SM cycles per warp-wide gather instruction (lower is better).
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ unsigned mix(unsigned x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x;
}

template <int MODE>
__global__ void __launch_bounds__(1024) k(const float* __restrict__ tab, int iters, float* out) {
  extern __shared__ float s_tab[];
  constexpr int NS = 32768;  // floats of shared table (128 KB)
  if (MODE == 0 || MODE == 1 || MODE == 4) {
    for (int i = threadIdx.x; i < NS; i += blockDim.x) s_tab[i] = tab[i];
    __syncthreads();
  }
  const float* remote = s_tab;
  if (MODE == 4) {
    cg::cluster_group cl = cg::this_cluster();
    remote = cl.map_shared_rank(s_tab, cl.block_rank() ^ 1);
    cl.sync();
  }
  unsigned h = mix(blockIdx.x * blockDim.x + threadIdx.x + 1);
  float acc = 0.f;
  for (int it = 0; it < iters; it++) {
    unsigned idx[6];
#pragma unroll
    for (int g = 0; g < 6; g++) { h = h * 1664525u + 1013904223u; idx[g] = mix(h); }
#pragma unroll
    for (int g = 0; g < 6; g++) {
      const unsigned r = idx[g];
      if (MODE == 0) acc += s_tab[r & (NS - 1)];
      else if (MODE == 1) acc += ((r >> 20) % 100 < 61) ? s_tab[r & (NS - 1)] : __ldg(tab + (r & (NS - 1)));
      else if (MODE == 2) acc += __ldg(tab + (r & (NS - 1)));
      else if (MODE == 3) acc += __ldg(tab + (r & 65535));
      else if (MODE == 4) acc += (r & 65536) ? remote[r & (NS - 1)] : s_tab[r & (NS - 1)];
      else acc += __ldg(tab + (r % 393216));
    }
  }
  if (acc == 1234.5f) out[0] = acc;
  if (MODE == 4) cg::this_cluster().sync();
}

template <int MODE>
static void run(const char* name, const float* tab, float* out, int sms, size_t smem, int ctas_per_sm, int cluster) {
  const int iters = 512, grid = sms * ctas_per_sm;
  CK(cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(1024); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  float best = 1e30f;
  for (int rep = 0; rep < 4; rep++) {
    CK(cudaEventRecord(a));
    CK(cudaLaunchKernelEx(&cfg, k<MODE>, tab, iters, out));
    CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b)); if (ms < best) best = ms;
  }
  int khz; CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
  const double warp_gathers_per_sm = (double)iters * 6 * 32 * ctas_per_sm;
  printf("%-70s %8.3f ms  %6.2f cycles per warp gather (at %d MHz)\n", name, best, best * 1e-3 * khz * 1e3 / warp_gathers_per_sm,
         khz / 1000);
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  float *tab, *out; CK(cudaMalloc(&tab, 4 << 20)); CK(cudaMemset(tab, 0, 4 << 20)); CK(cudaMalloc(&out, 4));
  const int sms = p.multiProcessorCount & ~1;
  run<0>("0 lds 128 KB", tab, out, sms, 128 << 10, 1, 1);
  run<1>("1 61% lds 128 KB + 39% ldg 128 KB (smem 136 KB)", tab, out, sms, 136 << 10, 1, 1);
  run<2>("2 ldg 128 KB table, 136 KB smem allocated", tab, out, sms, 136 << 10, 1, 1);
  run<3>("3 ldg 256 KB table, 8 KB smem, 2 CTAs/SM", tab, out, sms, 8 << 10, 2, 1);
  run<4>("4 cluster 2: 50% local lds + 50% DSMEM (128 KB each)", tab, out, sms, 128 << 10, 1, 2);
  run<5>("5 ldg 1.5 MB table, 8 KB smem, 2 CTAs/SM", tab, out, sms, 8 << 10, 2, 1);
  return 0;
}
