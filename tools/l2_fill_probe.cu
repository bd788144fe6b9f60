// L2 -> shared-memory fill bandwidth of a B200, measured the way conv_tc_kernel's producer warp loads its operands:
// one elected thread per CTA issues bulk async copies (TMA engine, mbarrier completion) into a ring of shared-memory
// stages, one persistent CTA per SM, source resident in L2.  Prints one JSON line per case:
//   private   every CTA streams its own 512 KB window (37 .. 76 MB in total: L2-resident after the warm-up pass)
//   shared    all CTAs stream the SAME 512 KB window (the weight operand of a conv: every CTA reads the same bytes)
//   half      CTA i streams window i/2 (two CTAs per window: what a 2-CTA cluster would share by multicast)
//   mcast2    clusters of two CTAs: each CTA loads HALF of every chunk of window i/2 and multicasts it to both, so each
//             CTA still receives 32 KB per stage but L2 is read once per pair
// for grids of 148 / 74 / 16 / 1 CTAs.  Not part of the product; build + run:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o build/l2_fill_probe tools/l2_fill_probe.cu && build/l2_fill_probe
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

constexpr int kStages = 6;
constexpr int kChunk = 32 * 1024;          // bytes per bulk copy (conv_tc stages are 16 KB boxes x 2..4 per barrier)
constexpr int kWindow = 512 * 1024;        // bytes a CTA cycles through

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int n) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ unsigned int g_timeouts;
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {   // bounded: a probe must never hang the box
  if (*reinterpret_cast<volatile unsigned int*>(&g_timeouts) != 0u) return;   // something is wrong: get out fast
  for (int spin = 0; spin < (1 << 22); ++spin) {
    uint32_t ok;
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
    if (ok) return;
  }
  atomicAdd(&g_timeouts, 1u);
}
__device__ __forceinline__ void bulk_load_mc(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar)), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__global__ void __launch_bounds__(128, 1) fill_kernel(const uint8_t* src, int mode, int iters, long long* clk_out) {
  extern __shared__ __align__(1024) uint8_t ring[];
  __shared__ __align__(8) uint64_t full[kStages];
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (mode == 3) cluster_sync_all();   // the peer's barriers exist before its first multicast signal can arrive
  if (threadIdx.x == 0) {
    uint32_t rank = 0;
    if (mode == 3) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const int win = mode == 0 ? blockIdx.x : mode == 1 ? 0 : blockIdx.x / 2;
    const uint8_t* base = src + static_cast<size_t>(win) * kWindow;
    const int per_win = kWindow / kChunk;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const int s = it % kStages;
      if (it >= kStages) mbar_wait(&full[s], ((it / kStages) - 1) & 1);   // the stage's previous copy has landed
      mbar_expect_tx(&full[s], kChunk);
      const uint8_t* g = base + static_cast<size_t>(it % per_win) * kChunk;
      if (mode == 3)
        bulk_load_mc(ring + s * kChunk + rank * (kChunk / 2), g + rank * (kChunk / 2), kChunk / 2, &full[s], 3);
      else
        bulk_load(ring + s * kChunk, g, kChunk, &full[s]);
    }
    for (int it = iters; it < iters + kStages; ++it) {   // drain
      const int s = it % kStages;
      mbar_wait(&full[s], ((it / kStages) - 1) & 1);
    }
    clk_out[blockIdx.x] = clock64() - t0;
  }
  if (mode == 3) cluster_sync_all();   // no CTA retires while its peer may still be writing into it
}

int main() {
  int dev = 0;
  CK(cudaSetDevice(dev));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, dev));
  const int sms = prop.multiProcessorCount;
  const size_t bytes = static_cast<size_t>(sms) * kWindow;
  uint8_t* src;
  CK(cudaMalloc(&src, bytes));
  CK(cudaMemset(src, 1, bytes));
  long long* clk;
  CK(cudaMalloc(&clk, sizeof(long long) * sms));
  const int smem = kStages * kChunk;
  CK(cudaFuncSetAttribute(fill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  const char* names[4] = {"private", "shared", "half", "mcast2"};
  auto launch = [&](int grid, int mode, int iters) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = mode == 3 ? 2 : 1;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    CK(cudaLaunchKernelEx(&cfg, fill_kernel, static_cast<const uint8_t*>(src), mode, iters, clk));
  };
  const int grids[4] = {sms, sms / 2, 16, 1};
  const int iters = 4096;   // 128 MB per CTA
  long long* h = static_cast<long long*>(malloc(sizeof(long long) * sms));
  for (int mode = 0; mode < 4; ++mode) {
    for (int gi = 0; gi < 4; ++gi) {
      const int grid = (mode == 3 && grids[gi] < 2) ? 2 : grids[gi];
      launch(grid, mode, 256);   // warm-up: brings the windows into L2
      CK(cudaDeviceSynchronize());
      CK(cudaEventRecord(e0));
      launch(grid, mode, iters);
      CK(cudaEventRecord(e1));
      CK(cudaDeviceSynchronize());
      float ms = 0.f;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      CK(cudaMemcpy(h, clk, sizeof(long long) * grid, cudaMemcpyDeviceToHost));
      long long mx = 0, mn = 1LL << 62;
      double sum = 0;
      for (int i = 0; i < grid; ++i) { mx = h[i] > mx ? h[i] : mx; mn = h[i] < mn ? h[i] : mn; sum += static_cast<double>(h[i]); }
      const double per_cta = static_cast<double>(iters) * kChunk;
      printf("{\"case\": \"%s\", \"ctas\": %d, \"ms\": %.4f, \"TBps\": %.3f, \"bytes_per_clk_per_sm_mean\": %.2f, "
             "\"bytes_per_clk_per_sm_min\": %.2f, \"bytes_per_clk_chip\": %.1f, \"sm_mhz_effective\": %.0f}\n",
             names[mode], grid, ms, per_cta * grid / (ms * 1e-3) / 1e12, per_cta / (sum / grid), per_cta / static_cast<double>(mx),
             per_cta * grid / (sum / grid), (sum / grid) / (ms * 1e-3) / 1e6);
      fflush(stdout);
    }
  }
  unsigned int to = 0;
  CK(cudaMemcpyFromSymbol(&to, g_timeouts, sizeof(to)));
  printf("{\"wait_timeouts\": %u}\n", to);
  return 0;
}
