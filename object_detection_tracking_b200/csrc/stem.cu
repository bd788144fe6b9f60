// Frame ingest kernels: fused preprocess + 7x7/2 stem convolution, 3x3/2 max-pool, and the
// fp32 <-> (hi, lo) fp16 plane conversions.
//
// Reference ops replaced: build_preprocess (models.py:337-357), resnet_fpn_backbone stem
// (nn.py:871-900: pad [3, 2+pad32], conv0 7x7 s2 VALID + BN + ReLU, pad [1,0], MaxPooling 3x3 s2 VALID).
#include "common.h"

namespace b2 {
namespace {

__global__ void f32_to_planes_kernel(const float* __restrict__ src, __half* __restrict__ hi, __half* __restrict__ lo,
                                     size_t n) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float v = src[i];
    const __half h = __float2half_rn(v);
    hi[i] = h;
    if (lo) lo[i] = __float2half_rn((v - __half2float(h)) * kLoScale);
  }
}

__global__ void planes_to_f32_kernel(const __half* __restrict__ hi, const __half* __restrict__ lo,
                                     float* __restrict__ dst, size_t n) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    float v = __half2float(hi[i]);
    if (lo) v = fmaf(__half2float(lo[i]), kLoInv, v);
    dst[i] = v;
  }
}

// Stem: each thread = one output pixel x 16 output channels; weights [147][64] fp32 in shared memory.
// Input is the raw frame (uint8 or float32, HWC, BGR); normalisation follows models.py:345-355 op by op:
// x * (1/255), - mean, / std  (mean/std reversed to BGR order).
constexpr int kStemTaps = 7 * 7 * 3;

template <typename TIn>
__global__ void __launch_bounds__(256) stem_kernel(const TIn* __restrict__ img, int B, int H, int W,
                                                   const float* __restrict__ wgt /*[147][64]*/,
                                                   const float* __restrict__ bias /*[64]*/, __half* __restrict__ out_hi,
                                                   __half* __restrict__ out_lo, int Ho, int Wo) {
  __shared__ __align__(16) float sw[kStemTaps * 64];
  __shared__ float sb[64];
  for (int i = threadIdx.x; i < kStemTaps * 64; i += blockDim.x) sw[i] = wgt[i];
  if (threadIdx.x < 64) sb[threadIdx.x] = bias[threadIdx.x];
  __syncthreads();
  const float mean[3] = {0.406f, 0.456f, 0.485f};   // BGR order (models.py:350-352)
  const float stdv[3] = {0.225f, 0.224f, 0.229f};
  const int cg = threadIdx.x & 3;                   // 16-channel group
  const size_t npix = static_cast<size_t>(B) * Ho * Wo;
  for (size_t pix = blockIdx.x * 64ull + (threadIdx.x >> 2); pix < npix; pix += static_cast<size_t>(gridDim.x) * 64) {
    const int b = static_cast<int>(pix / (static_cast<size_t>(Ho) * Wo));
    const int rem = static_cast<int>(pix % (static_cast<size_t>(Ho) * Wo));
    const int p = rem / Wo, q = rem % Wo;
    float acc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = 0.f;
    for (int r = 0; r < 7; ++r) {
      const int ih = p * 2 - 3 + r;
      if (ih < 0 || ih >= H) continue;
      for (int s = 0; s < 7; ++s) {
        const int iw = q * 2 - 3 + s;
        if (iw < 0 || iw >= W) continue;
        const TIn* px = img + ((static_cast<size_t>(b) * H + ih) * W + iw) * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          float x = static_cast<float>(px[c]);
          x = __fmul_rn(x, 1.0f / 255);
          x = __fsub_rn(x, mean[c]);
          x = __fdiv_rn(x, stdv[c]);
          const float4* w4 = reinterpret_cast<const float4*>(&sw[((r * 7 + s) * 3 + c) * 64 + cg * 16]);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 wv = w4[j];
            acc[4 * j + 0] = fmaf(x, wv.x, acc[4 * j + 0]);
            acc[4 * j + 1] = fmaf(x, wv.y, acc[4 * j + 1]);
            acc[4 * j + 2] = fmaf(x, wv.z, acc[4 * j + 2]);
            acc[4 * j + 3] = fmaf(x, wv.w, acc[4 * j + 3]);
          }
        }
      }
    }
    __align__(16) __half hbuf[16];
    __align__(16) __half lbuf[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float v = fmaxf(acc[j] + sb[cg * 16 + j], 0.f);
      hbuf[j] = __float2half_rn(v);
      lbuf[j] = __float2half_rn((v - __half2float(hbuf[j])) * kLoScale);
    }
    uint4* oh = reinterpret_cast<uint4*>(out_hi + pix * 64 + cg * 16);
    oh[0] = reinterpret_cast<uint4*>(hbuf)[0];
    oh[1] = reinterpret_cast<uint4*>(hbuf)[1];
    if (out_lo) {
      uint4* ol = reinterpret_cast<uint4*>(out_lo + pix * 64 + cg * 16);
      ol[0] = reinterpret_cast<uint4*>(lbuf)[0];
      ol[1] = reinterpret_cast<uint4*>(lbuf)[1];
    }
  }
}

// 3x3 stride-2 max pool with one zero row/column of padding on top/left (inputs are post-ReLU, >= 0).
// Each thread handles 8 channels of one output pixel; the (hi, lo) pair of the max element is kept.
__global__ void maxpool_kernel(const __half* __restrict__ in_hi, const __half* __restrict__ in_lo, int B, int H, int W,
                               int C, __half* __restrict__ out_hi, __half* __restrict__ out_lo, int Ho, int Wo) {
  const int cvec = C / 8;
  const size_t total = static_cast<size_t>(B) * Ho * Wo * cvec;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int cv = static_cast<int>(idx % cvec);
    const size_t pix = idx / cvec;
    const int b = static_cast<int>(pix / (static_cast<size_t>(Ho) * Wo));
    const int rem = static_cast<int>(pix % (static_cast<size_t>(Ho) * Wo));
    const int p = rem / Wo, q = rem % Wo;
    float best[8];
    __half bh[8], bl[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      best[j] = 0.f;
      bh[j] = __float2half_rn(0.f);
      bl[j] = __float2half_rn(0.f);
    }
    for (int r = 0; r < 3; ++r) {
      const int ih = p * 2 - 1 + r;
      if (ih < 0 || ih >= H) continue;
      for (int s = 0; s < 3; ++s) {
        const int iw = q * 2 - 1 + s;
        if (iw < 0 || iw >= W) continue;
        const size_t off = ((static_cast<size_t>(b) * H + ih) * W + iw) * C + cv * 8;
        const uint4 vh = *reinterpret_cast<const uint4*>(in_hi + off);
        uint4 vl = make_uint4(0, 0, 0, 0);
        if (in_lo) vl = *reinterpret_cast<const uint4*>(in_lo + off);
        const __half* h = reinterpret_cast<const __half*>(&vh);
        const __half* l = reinterpret_cast<const __half*>(&vl);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float v = fmaf(__half2float(l[j]), kLoInv, __half2float(h[j]));
          if (v > best[j]) {
            best[j] = v;
            bh[j] = h[j];
            bl[j] = l[j];
          }
        }
      }
    }
    const size_t o = pix * C + cv * 8;
    *reinterpret_cast<uint4*>(out_hi + o) = *reinterpret_cast<uint4*>(bh);
    if (out_lo) *reinterpret_cast<uint4*>(out_lo + o) = *reinterpret_cast<uint4*>(bl);
  }
}

inline unsigned grid_for(size_t total, int threads, unsigned cap = 148 * 32) {
  size_t b = (total + threads - 1) / threads;
  if (b > cap) b = cap;
  if (b == 0) b = 1;
  return static_cast<unsigned>(b);
}

}  // namespace

int f32_to_planes(const float* src, __half* hi, __half* lo, size_t n, cudaStream_t s) {
  f32_to_planes_kernel<<<grid_for(n, 256), 256, 0, s>>>(src, hi, lo, n);
  B2_CUDA(cudaGetLastError());
  return 0;
}
int planes_to_f32(const __half* hi, const __half* lo, float* dst, size_t n, cudaStream_t s) {
  planes_to_f32_kernel<<<grid_for(n, 256), 256, 0, s>>>(hi, lo, dst, n);
  B2_CUDA(cudaGetLastError());
  return 0;
}

int stem_launch(const void* img, int is_u8, int B, int H, int W, const float* wgt, const float* bias, __half* out_hi,
                __half* out_lo, int Ho, int Wo, cudaStream_t s) {
  const size_t npix = static_cast<size_t>(B) * Ho * Wo;
  const unsigned grid = grid_for(npix, 64, 148 * 16);
  if (is_u8)
    stem_kernel<uint8_t><<<grid, 256, 0, s>>>(static_cast<const uint8_t*>(img), B, H, W, wgt, bias, out_hi, out_lo, Ho, Wo);
  else
    stem_kernel<float><<<grid, 256, 0, s>>>(static_cast<const float*>(img), B, H, W, wgt, bias, out_hi, out_lo, Ho, Wo);
  B2_CUDA(cudaGetLastError());
  return 0;
}

int maxpool_launch(const __half* in_hi, const __half* in_lo, int B, int H, int W, int C, __half* out_hi,
                   __half* out_lo, int Ho, int Wo, cudaStream_t s) {
  const size_t total = static_cast<size_t>(B) * Ho * Wo * (C / 8);
  maxpool_kernel<<<grid_for(total, 256), 256, 0, s>>>(in_hi, in_lo, B, H, W, C, out_hi, out_lo, Ho, Wo);
  B2_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace b2
