// DeepSORT appearance cost: L2-normalise gallery / detection rows into tensor-core operands, and the
// segmented row-min that turns the [S, N] similarity GEMM into the [T, N] cost matrix.
// The GEMM itself runs on the tcgen05 conv kernel (1x1 conv, split precision).
//
// Reference ops replaced: nn_matching._cosine_distance (deep_sort/nn_matching.py:31-54),
// _nn_cosine_distance (:78-96), NearestNeighborDistanceMetric.distance (:156-177).
#include "common.h"
#include "kernels.h"

namespace b2 {
namespace {

// one warp per row: x / ||x||  -> (hi, lo) planes, row stride ld (zero padded to ld)
__global__ void cosine_normalize_kernel(const float* __restrict__ src, int rows, int D, __half* __restrict__ hi,
                                        __half* __restrict__ lo, int ld) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* x = src + static_cast<size_t>(row) * D;
  float ss = 0.f;
  for (int c = lane; c < D; c += 32) ss = fmaf(x[c], x[c], ss);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float nrm = sqrtf(ss);
  for (int c = lane; c < ld; c += 32) {
    const float v = c < D ? __fdiv_rn(x[c], nrm) : 0.f;
    const __half h = __float2half_rn(v);
    hi[static_cast<size_t>(row) * ld + c] = h;
    if (lo) lo[static_cast<size_t>(row) * ld + c] = __float2half_rn((v - __half2float(h)) * kLoScale);
  }
}

// cost[t][n] = min over gallery rows s in [off[t], off[t+1]) of (1 - dots[s][n])
__global__ void cosine_segmin_kernel(const float* __restrict__ dots, int ld, const int* __restrict__ off, int T, int N,
                                     float* __restrict__ cost) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= T * N) return;
  const int t = idx / N, n = idx - t * N;
  float best = 3.4e38f;
  for (int s = off[t]; s < off[t + 1]; ++s) best = fminf(best, __fsub_rn(1.f, dots[static_cast<size_t>(s) * ld + n]));
  cost[idx] = best;
}

}  // namespace

int cosine_normalize_rows(const float* src, int rows, int D, __half* hi, __half* lo, int ld, cudaStream_t s) {
  if (rows <= 0) return 0;
  const int threads = 256;
  const int blocks = (rows * 32 + threads - 1) / threads;
  cosine_normalize_kernel<<<blocks, threads, 0, s>>>(src, rows, D, hi, lo, ld);
  B2_CUDA(cudaGetLastError());
  return 0;
}

int cosine_segmin(const float* dots, int ld, const int* seg_offsets, int T, int N, float* cost, cudaStream_t s) {
  if (T * N <= 0) return 0;
  cosine_segmin_kernel<<<(T * N + 255) / 256, 256, 0, s>>>(dots, ld, seg_offsets, T, N, cost);
  B2_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace b2
