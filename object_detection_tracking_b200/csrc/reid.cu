// Non-GEMM kernels of the OSNet ReID embedding (HBM-bound byte/element work; the 1x1 convolutions and
// the fc run on the tcgen05 implicit-GEMM kernel).
//
// Reference ops replaced (torchreid/models/osnet.py): LightConv3x3's depthwise 3x3 + BN + ReLU (:128-156),
// ChannelGate (:162-220: global avg-pool -> fc1 -> ReLU -> fc2 -> sigmoid -> scale), the 4-stream gated sum
// of OSBlock.forward (:262-268), AvgPool2d(2) of the transitions (:375-381), AdaptiveAvgPool2d(1) (:428).
#include "common.h"
#include "kernels.h"

namespace b2 {
namespace {

__device__ __forceinline__ void unpack8(const uint4& h, const uint4& l, bool has_lo, float (&v)[8]) {
  const __half2* hh = reinterpret_cast<const __half2*>(&h);
  const __half2* ll = reinterpret_cast<const __half2*>(&l);
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    float2 f = __half22float2(hh[t]);
    if (has_lo) {
      const float2 g = __half22float2(ll[t]);
      f.x = fmaf(g.x, kLoInv, f.x);
      f.y = fmaf(g.y, kLoInv, f.y);
    }
    v[2 * t] = f.x;
    v[2 * t + 1] = f.y;
  }
}

__device__ __forceinline__ void pack8(const float (&v)[8], uint4& h, uint4& l) {
  __half2* hh = reinterpret_cast<__half2*>(&h);
  __half2* ll = reinterpret_cast<__half2*>(&l);
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    hh[t] = __floats2half2_rn(v[2 * t], v[2 * t + 1]);
    const float2 f = __half22float2(hh[t]);
    ll[t] = __floats2half2_rn((v[2 * t] - f.x) * kLoScale, (v[2 * t + 1] - f.y) * kLoScale);
  }
}

// depthwise 3x3, stride 1, pad 1, + folded BN + ReLU.  w: [9][C] fp32 (BN scale folded), bias [C].
__global__ void dwconv3x3_kernel(const __half* __restrict__ in_hi, const __half* __restrict__ in_lo, int B, int H, int W,
                                 int C, const float* __restrict__ w, const float* __restrict__ bias,
                                 __half* __restrict__ out_hi, __half* __restrict__ out_lo) {
  const int cvec = C / 8;
  const size_t total = static_cast<size_t>(B) * H * W * cvec;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int cv = static_cast<int>(idx % cvec);
    const size_t pix = idx / cvec;
    const int b = static_cast<int>(pix / (static_cast<size_t>(H) * W));
    const int rem = static_cast<int>(pix % (static_cast<size_t>(H) * W));
    const int y = rem / W, x = rem % W;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int iy = y + r - 1;
      if (iy < 0 || iy >= H) continue;
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const int ix = x + s - 1;
        if (ix < 0 || ix >= W) continue;
        const size_t off = ((static_cast<size_t>(b) * H + iy) * W + ix) * C + cv * 8;
        const uint4 h = __ldg(reinterpret_cast<const uint4*>(in_hi + off));
        uint4 l = make_uint4(0, 0, 0, 0);
        if (in_lo) l = __ldg(reinterpret_cast<const uint4*>(in_lo + off));
        float v[8];
        unpack8(h, l, in_lo != nullptr, v);
        const float4* wp = reinterpret_cast<const float4*>(w + static_cast<size_t>(r * 3 + s) * C + cv * 8);
        const float4 w0 = __ldg(wp), w1 = __ldg(wp + 1);
        acc[0] = fmaf(v[0], w0.x, acc[0]); acc[1] = fmaf(v[1], w0.y, acc[1]);
        acc[2] = fmaf(v[2], w0.z, acc[2]); acc[3] = fmaf(v[3], w0.w, acc[3]);
        acc[4] = fmaf(v[4], w1.x, acc[4]); acc[5] = fmaf(v[5], w1.y, acc[5]);
        acc[6] = fmaf(v[6], w1.z, acc[6]); acc[7] = fmaf(v[7], w1.w, acc[7]);
      }
    }
    const float4* bp = reinterpret_cast<const float4*>(bias + cv * 8);
    const float4 b0 = __ldg(bp), b1 = __ldg(bp + 1);
    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = fmaxf(acc[j] + bb[j], 0.f);
    uint4 oh, ol;
    pack8(acc, oh, ol);
    const size_t o = pix * C + cv * 8;
    *reinterpret_cast<uint4*>(out_hi + o) = oh;
    if (out_lo) *reinterpret_cast<uint4*>(out_lo + o) = ol;
  }
}

// global average pool: one block per (image, 64-channel group); out fp32 [B][C]
__global__ void __launch_bounds__(256) gap_kernel(const __half* __restrict__ in_hi, const __half* __restrict__ in_lo,
                                                  int HW, int C, float* __restrict__ out, int out_ld) {
  const int b = blockIdx.y, cg = blockIdx.x;          // 64 channels per block: 8 vectors x 32 pixel lanes
  const int cv = threadIdx.x & 7, pl = threadIdx.x >> 3;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  for (int p = pl; p < HW; p += 32) {
    const size_t off = (static_cast<size_t>(b) * HW + p) * C + cg * 64 + cv * 8;
    const uint4 h = __ldg(reinterpret_cast<const uint4*>(in_hi + off));
    uint4 l = make_uint4(0, 0, 0, 0);
    if (in_lo) l = __ldg(reinterpret_cast<const uint4*>(in_lo + off));
    float v[8];
    unpack8(h, l, in_lo != nullptr, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] += v[j];
  }
  __shared__ float red[32][65];
#pragma unroll
  for (int j = 0; j < 8; ++j) red[pl][cv * 8 + j] = acc[j];
  __syncthreads();
  if (threadIdx.x < 64) {
    float s = 0.f;
    for (int i = 0; i < 32; ++i) s += red[i][threadIdx.x];
    out[static_cast<size_t>(b) * out_ld + cg * 64 + threadIdx.x] = s / static_cast<float>(HW);
  }
}

// ChannelGate MLP on pooled vectors: one block per pooled row.  g [rows][C] -> gates [rows][C]
// fc1: [Cr][C] + b1, ReLU, fc2: [C][Cr] + b2, sigmoid.  Only the first Creal channels are live.
__global__ void __launch_bounds__(128) gate_mlp_kernel(const float* __restrict__ g, int C, int Creal, int Cr,
                                                       const float* __restrict__ w1, const float* __restrict__ b1,
                                                       const float* __restrict__ w2, const float* __restrict__ b2,
                                                       float* __restrict__ gates) {
  __shared__ float sx[512];
  __shared__ float sh[64];
  const int row = blockIdx.x;
  for (int i = threadIdx.x; i < Creal; i += blockDim.x) sx[i] = g[static_cast<size_t>(row) * C + i];
  __syncthreads();
  if (threadIdx.x < Cr) {
    float a = b1[threadIdx.x];
    for (int i = 0; i < Creal; ++i) a = fmaf(w1[threadIdx.x * Creal + i], sx[i], a);
    sh[threadIdx.x] = fmaxf(a, 0.f);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float out = 0.f;
    if (c < Creal) {
      float a = b2[c];
      for (int i = 0; i < Cr; ++i) a = fmaf(w2[c * Cr + i], sh[i], a);
      out = 1.f / (1.f + expf(-a));
    }
    gates[static_cast<size_t>(row) * C + c] = out;
  }
}

// x2 = gate(a)*a + gate(b)*b + gate(c)*c + gate(d)*d ; gates [B][4][C] fp32
__global__ void gated_sum4_kernel(const __half* __restrict__ a_hi, const __half* __restrict__ a_lo,
                                  const __half* __restrict__ b_hi, const __half* __restrict__ b_lo,
                                  const __half* __restrict__ c_hi, const __half* __restrict__ c_lo,
                                  const __half* __restrict__ d_hi, const __half* __restrict__ d_lo,
                                  const float* __restrict__ gates, int B, int HW, int C, __half* __restrict__ out_hi,
                                  __half* __restrict__ out_lo) {
  const int cvec = C / 8;
  const size_t total = static_cast<size_t>(B) * HW * cvec;
  const __half* his[4] = {a_hi, b_hi, c_hi, d_hi};
  const __half* los[4] = {a_lo, b_lo, c_lo, d_lo};
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int cv = static_cast<int>(idx % cvec);
    const size_t pix = idx / cvec;
    const int b = static_cast<int>(pix / HW);
    const size_t off = pix * C + cv * 8;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const uint4 h = __ldg(reinterpret_cast<const uint4*>(his[s] + off));
      uint4 l = make_uint4(0, 0, 0, 0);
      if (los[s]) l = __ldg(reinterpret_cast<const uint4*>(los[s] + off));
      float v[8];
      unpack8(h, l, los[s] != nullptr, v);
      const float* gp = gates + (static_cast<size_t>(b) * 4 + s) * C + cv * 8;
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = __fadd_rn(acc[j], __fmul_rn(v[j], __ldg(gp + j)));
    }
    uint4 oh, ol;
    pack8(acc, oh, ol);
    *reinterpret_cast<uint4*>(out_hi + off) = oh;
    if (out_lo) *reinterpret_cast<uint4*>(out_lo + off) = ol;
  }
}

// 2x2 average pool, stride 2
__global__ void avgpool2_kernel(const __half* __restrict__ in_hi, const __half* __restrict__ in_lo, int B, int H, int W,
                                int C, __half* __restrict__ out_hi, __half* __restrict__ out_lo) {
  const int Ho = H / 2, Wo = W / 2, cvec = C / 8;
  const size_t total = static_cast<size_t>(B) * Ho * Wo * cvec;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int cv = static_cast<int>(idx % cvec);
    const size_t pix = idx / cvec;
    const int b = static_cast<int>(pix / (static_cast<size_t>(Ho) * Wo));
    const int rem = static_cast<int>(pix % (static_cast<size_t>(Ho) * Wo));
    const int y = rem / Wo, x = rem % Wo;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const size_t off = ((static_cast<size_t>(b) * H + 2 * y + dy) * W + 2 * x + dx) * C + cv * 8;
        const uint4 h = __ldg(reinterpret_cast<const uint4*>(in_hi + off));
        uint4 l = make_uint4(0, 0, 0, 0);
        if (in_lo) l = __ldg(reinterpret_cast<const uint4*>(in_lo + off));
        float v[8];
        unpack8(h, l, in_lo != nullptr, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += v[j];
      }
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] *= 0.25f;
    uint4 oh, ol;
    pack8(acc, oh, ol);
    const size_t o = pix * C + cv * 8;
    *reinterpret_cast<uint4*>(out_hi + o) = oh;
    if (out_lo) *reinterpret_cast<uint4*>(out_lo + o) = ol;
  }
}

inline unsigned grid_for(size_t total, int threads, unsigned cap = 148 * 32) {
  size_t b = (total + threads - 1) / threads;
  if (b > cap) b = cap;
  if (b == 0) b = 1;
  return static_cast<unsigned>(b);
}

}  // namespace

int dwconv3x3_launch(const __half* in_hi, const __half* in_lo, int B, int H, int W, int C, const float* w,
                     const float* bias, __half* out_hi, __half* out_lo, cudaStream_t s) {
  const size_t total = static_cast<size_t>(B) * H * W * (C / 8);
  dwconv3x3_kernel<<<grid_for(total, 256), 256, 0, s>>>(in_hi, in_lo, B, H, W, C, w, bias, out_hi, out_lo);
  B2_CUDA(cudaGetLastError());
  return 0;
}

int gap_launch(const __half* in_hi, const __half* in_lo, int B, int HW, int C, float* out, int out_ld, cudaStream_t s) {
  B2_CHECK(C % 64 == 0, "gap: C must be a multiple of 64");
  gap_kernel<<<dim3(C / 64, B), 256, 0, s>>>(in_hi, in_lo, HW, C, out, out_ld);
  B2_CUDA(cudaGetLastError());
  return 0;
}

int gate_mlp_launch(const float* g, int rows, int C, int Creal, int Cr, const float* w1, const float* b1,
                    const float* w2, const float* b2, float* gates, cudaStream_t s) {
  B2_CHECK(Creal <= 512 && Cr <= 64, "gate_mlp: channel counts out of range");
  gate_mlp_kernel<<<rows, 128, 0, s>>>(g, C, Creal, Cr, w1, b1, w2, b2, gates);
  B2_CUDA(cudaGetLastError());
  return 0;
}

int gated_sum4_launch(const __half* const hi[4], const __half* const lo[4], const float* gates, int B, int HW, int C,
                      __half* out_hi, __half* out_lo, cudaStream_t s) {
  const size_t total = static_cast<size_t>(B) * HW * (C / 8);
  gated_sum4_kernel<<<grid_for(total, 256), 256, 0, s>>>(hi[0], lo[0], hi[1], lo[1], hi[2], lo[2], hi[3], lo[3], gates,
                                                           B, HW, C, out_hi, out_lo);
  B2_CUDA(cudaGetLastError());
  return 0;
}

int avgpool2_launch(const __half* in_hi, const __half* in_lo, int B, int H, int W, int C, __half* out_hi,
                    __half* out_lo, cudaStream_t s) {
  const size_t total = static_cast<size_t>(B) * (H / 2) * (W / 2) * (C / 8);
  avgpool2_kernel<<<grid_for(total, 256), 256, 0, s>>>(in_hi, in_lo, B, H, W, C, out_hi, out_lo);
  B2_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace b2
