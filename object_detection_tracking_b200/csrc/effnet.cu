// Non-GEMM kernels of the EfficientNet backbone (EfficientDet config) and of the wrapper's input pre-processing.
// The 1x1 expand / project convolutions and the (im2col-packed) stem run on the tcgen05 conv kernel.
//
// Reference ops replaced: EfficientDet.build_preprocess (efficientdet_wrapper.py:40-61) + InputProcessor
// (efficientdet/dataloader.py:56-123: convert_image_dtype, mean/std, tf.image.resize_images BILINEAR,
// pad_to_bounding_box); MBConvBlock's DepthwiseConv2D SAME + BN + swish (efficientnet_model.py:255-268,363-364) and
// squeeze-excite (_call_se :300-326); Model's stem Conv2D 3x3/2 SAME (:519-531).
#include "common.h"
#include "kernels.h"

namespace b2 {
namespace {

__device__ __forceinline__ void ld8(const __half* hi, const __half* lo, size_t off, float (&v)[8]) {
  const uint4 h = __ldg(reinterpret_cast<const uint4*>(hi + off));
  const __half2* hh = reinterpret_cast<const __half2*>(&h);
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const float2 f = __half22float2(hh[t]);
    v[2 * t] = f.x;
    v[2 * t + 1] = f.y;
  }
  if (lo) {
    const uint4 l = __ldg(reinterpret_cast<const uint4*>(lo + off));
    const __half2* ll = reinterpret_cast<const __half2*>(&l);
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const float2 f = __half22float2(ll[t]);
      v[2 * t] = fmaf(f.x, kLoInv, v[2 * t]);
      v[2 * t + 1] = fmaf(f.y, kLoInv, v[2 * t + 1]);
    }
  }
}

__device__ __forceinline__ void st8(__half* hi, __half* lo, size_t off, const float (&v)[8]) {
  uint4 oh, ol;
  __half2* hh = reinterpret_cast<__half2*>(&oh);
  __half2* ll = reinterpret_cast<__half2*>(&ol);
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    hh[t] = __floats2half2_rn(v[2 * t], v[2 * t + 1]);
    const float2 f = __half22float2(hh[t]);
    ll[t] = __floats2half2_rn((v[2 * t] - f.x) * kLoScale, (v[2 * t + 1] - f.y) * kLoScale);
  }
  *reinterpret_cast<uint4*>(hi + off) = oh;
  if (lo) *reinterpret_cast<uint4*>(lo + off) = ol;
}

// uint8 BGR frame -> RGB, * 1/255, (x - mean) / std, legacy bilinear resize (align_corners=False, no half-pixel
// centres) to (sh, sw), zero-padded bottom/right to [H, W, 3] fp32.  The reference normalises first and resizes the
// float image; lerp of normalised corner values is done in the same order here.
__global__ void effnet_preprocess_kernel(const uint8_t* __restrict__ img, int h, int w, int sh, int sw, float ys, float xs,
                                         float* __restrict__ out, int H, int W) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= H * W) return;
  const int y = idx / W, x = idx - y * W;
  float r[3] = {0.f, 0.f, 0.f};
  if (y < sh && x < sw) {
    const float in_y = __fmul_rn(static_cast<float>(y), ys), in_x = __fmul_rn(static_cast<float>(x), xs);
    const float fy = floorf(in_y), fx = floorf(in_x);
    const int y0 = static_cast<int>(fy), x0 = static_cast<int>(fx);
    const int y1 = min(static_cast<int>(ceilf(in_y)), h - 1), x1 = min(static_cast<int>(ceilf(in_x)), w - 1);
    const float ly = __fsub_rn(in_y, fy), lx = __fsub_rn(in_x, fx);
    const float mean[3] = {0.485f, 0.456f, 0.406f}, stdv[3] = {0.229f, 0.224f, 0.225f};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int sc = 2 - c;   // BGR -> RGB
      auto px = [&](int yy, int xx) {
        const float v = __fmul_rn(static_cast<float>(img[(static_cast<size_t>(yy) * w + xx) * 3 + sc]), 1.0f / 255.0f);
        return __fdiv_rn(__fsub_rn(v, mean[c]), stdv[c]);
      };
      const float tl = px(y0, x0), tr = px(y0, x1), bl = px(y1, x0), br = px(y1, x1);
      const float top = __fadd_rn(tl, __fmul_rn(__fsub_rn(tr, tl), lx));
      const float bot = __fadd_rn(bl, __fmul_rn(__fsub_rn(br, bl), lx));
      r[c] = __fadd_rn(top, __fmul_rn(__fsub_rn(bot, top), ly));
    }
  }
  float* o = out + static_cast<size_t>(idx) * 3;
  o[0] = r[0]; o[1] = r[1]; o[2] = r[2];
}

// stem operand: 3x3/2 SAME patches of the fp32 image as one 64-channel pixel (k = (r*3 + s)*3 + c, 27 live)
__global__ void stem_im2col_kernel(const float* __restrict__ img, int H, int W, int Ho, int Wo, int pad_t, int pad_l,
                                   __half* __restrict__ out_hi, __half* __restrict__ out_lo) {
  const size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;   // (pixel, octet)
  if (idx >= static_cast<size_t>(Ho) * Wo * 8) return;
  const int oct = static_cast<int>(idx & 7);
  const size_t pix = idx >> 3;
  const int y = static_cast<int>(pix / Wo), x = static_cast<int>(pix - static_cast<size_t>(y) * Wo);
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int k = oct * 8 + j;
    float t = 0.f;
    if (k < 27) {
      const int c = k % 3, rs = k / 3, r = rs / 3, s = rs - r * 3;
      const int iy = 2 * y - pad_t + r, ix = 2 * x - pad_l + s;
      if (iy >= 0 && iy < H && ix >= 0 && ix < W) t = __ldg(img + (static_cast<size_t>(iy) * W + ix) * 3 + c);
    }
    v[j] = t;
  }
  st8(out_hi, out_lo, pix * 64 + oct * 8, v);
}

// depthwise KxK, stride 1/2, TF SAME padding; BatchNorm folded into w [K*K][C] and bias [C]; swish
template <int K>
__global__ void dwconv_bn_swish_kernel(const __grid_constant__ DwConvParams p) {
  const int cvec = p.C / 8;
  const size_t total = static_cast<size_t>(p.Ho) * p.Wo * cvec;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int cv = static_cast<int>(idx % cvec);
    const size_t pix = idx / cvec;
    const int y = static_cast<int>(pix / p.Wo), x = static_cast<int>(pix - static_cast<size_t>(y) * p.Wo);
    float acc[8];
    {
      const float4* bp = reinterpret_cast<const float4*>(p.bias + cv * 8);
      const float4 b0 = __ldg(bp), b1 = __ldg(bp + 1);
      acc[0] = b0.x; acc[1] = b0.y; acc[2] = b0.z; acc[3] = b0.w;
      acc[4] = b1.x; acc[5] = b1.y; acc[6] = b1.z; acc[7] = b1.w;
    }
    const int y0 = y * p.stride - p.pad_t, x0 = x * p.stride - p.pad_l;
#pragma unroll
    for (int r = 0; r < K; ++r) {
      const int iy = y0 + r;
      if (iy < 0 || iy >= p.H) continue;
#pragma unroll
      for (int s = 0; s < K; ++s) {
        const int ix = x0 + s;
        if (ix < 0 || ix >= p.W) continue;
        float v[8];
        ld8(p.in_hi, p.in_lo, (static_cast<size_t>(iy) * p.W + ix) * p.C + cv * 8, v);
        const float4* wp = reinterpret_cast<const float4*>(p.w + static_cast<size_t>(r * K + s) * p.C + cv * 8);
        const float4 w0 = __ldg(wp), w1 = __ldg(wp + 1);
        acc[0] = fmaf(v[0], w0.x, acc[0]); acc[1] = fmaf(v[1], w0.y, acc[1]);
        acc[2] = fmaf(v[2], w0.z, acc[2]); acc[3] = fmaf(v[3], w0.w, acc[3]);
        acc[4] = fmaf(v[4], w1.x, acc[4]); acc[5] = fmaf(v[5], w1.y, acc[5]);
        acc[6] = fmaf(v[6], w1.z, acc[6]); acc[7] = fmaf(v[7], w1.w, acc[7]);
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = __fmul_rn(acc[j], 1.f / (1.f + expf(-acc[j])));
    st8(p.out_hi, p.out_lo, pix * p.C + cv * 8, acc);
  }
}

// squeeze: per-channel partial sums over a pixel chunk.  grid (C/64, nchunks); partial [nchunks][C]
__global__ void __launch_bounds__(256) se_partial_kernel(const __half* __restrict__ in_hi, const __half* __restrict__ in_lo,
                                                         int HW, int C, int chunk, float* __restrict__ partial) {
  const int cg = blockIdx.x, ch = blockIdx.y;
  const int cv = threadIdx.x & 7, pl = threadIdx.x >> 3;
  const int p0 = ch * chunk, p1 = min(p0 + chunk, HW);
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  for (int p = p0 + pl; p < p1; p += 32) {
    float v[8];
    ld8(in_hi, in_lo, static_cast<size_t>(p) * C + cg * 64 + cv * 8, v);
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
    partial[static_cast<size_t>(ch) * C + cg * 64 + threadIdx.x] = s;
  }
}

// excite: mean -> 1x1 reduce + bias + swish -> 1x1 expand + bias -> sigmoid.  One block.
//   w1 [C][nr] (HWIO of [1,1,C,nr]), w2 [nr][C]; gate [Cpad] (0 beyond Creal)
__global__ void __launch_bounds__(256) se_excite_kernel(const float* __restrict__ partial, int nchunks, int Cpad, int C,
                                                        int nr, float inv_hw, const float* __restrict__ w1,
                                                        const float* __restrict__ b1, const float* __restrict__ w2,
                                                        const float* __restrict__ b2, float* __restrict__ gate) {
  extern __shared__ float sm[];
  float* g = sm;            // [C]
  float* r = sm + C;        // [nr]
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (int k = 0; k < nchunks; ++k) s += partial[static_cast<size_t>(k) * Cpad + c];
    g[c] = s * inv_hw;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int j = wid; j < nr; j += nw) {
    float a = 0.f;
    for (int c = lane; c < C; c += 32) a = fmaf(g[c], __ldg(w1 + static_cast<size_t>(c) * nr + j), a);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) {
      a += b1[j];
      r[j] = a * (1.f / (1.f + expf(-a)));
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < Cpad; c += blockDim.x) {
    float a = 0.f;
    if (c < C) {
      a = b2[c];
      for (int j = 0; j < nr; ++j) a = fmaf(r[j], __ldg(w2 + static_cast<size_t>(j) * C + c), a);
      a = 1.f / (1.f + expf(-a));
    }
    gate[c] = a;
  }
}

// fold the per-frame SE gate into the projection weights: w'[o][c] = w[o][c] * gate[c] -> (hi, lo) operand planes
__global__ void se_scale_weights_kernel(const float* __restrict__ w, const float* __restrict__ gate, int rows, int K,
                                        __half* __restrict__ hi, __half* __restrict__ lo) {
  const size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (idx >= static_cast<size_t>(rows) * K) return;
  const int c = static_cast<int>(idx % K);
  const float v = __fmul_rn(w[idx], gate[c]);
  const __half h = __float2half_rn(v);
  hi[idx] = h;
  if (lo) lo[idx] = __float2half_rn((v - __half2float(h)) * kLoScale);
}

inline unsigned grid_for(size_t total, int threads, unsigned cap) {
  size_t b = (total + threads - 1) / threads;
  if (b > cap) b = cap;
  if (b == 0) b = 1;
  return static_cast<unsigned>(b);
}

}  // namespace

int effnet_preprocess_launch(const uint8_t* img, int h, int w, int sh, int sw, float* out, int H, int W, cudaStream_t s) {
  B2_CHECK(sh >= 1 && sw >= 1 && h >= 1 && w >= 1, "effnet preprocess: empty frame");
  const float ys = static_cast<float>(h) / static_cast<float>(sh), xs = static_cast<float>(w) / static_cast<float>(sw);
  effnet_preprocess_kernel<<<(H * W + 255) / 256, 256, 0, s>>>(img, h, w, sh, sw, ys, xs, out, H, W);
  B2_CUDA(cudaGetLastError());
  return 0;
}

int stem_im2col_launch(const float* img, int H, int W, int Ho, int Wo, int pad_t, int pad_l, __half* out_hi, __half* out_lo,
                       cudaStream_t s) {
  const size_t total = static_cast<size_t>(Ho) * Wo * 8;
  stem_im2col_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, s>>>(img, H, W, Ho, Wo, pad_t, pad_l, out_hi, out_lo);
  B2_CUDA(cudaGetLastError());
  return 0;
}

int dwconv_bn_swish_launch(const DwConvParams& p, cudaStream_t s) {
  B2_CHECK(p.C % 8 == 0 && (p.K == 3 || p.K == 5) && (p.stride == 1 || p.stride == 2), "dwconv: unsupported geometry");
  const size_t total = static_cast<size_t>(p.Ho) * p.Wo * (p.C / 8);
  const unsigned grid = grid_for(total, 256, 148 * 32);
  if (p.K == 3) dwconv_bn_swish_kernel<3><<<grid, 256, 0, s>>>(p);
  else dwconv_bn_swish_kernel<5><<<grid, 256, 0, s>>>(p);
  B2_CUDA(cudaGetLastError());
  return 0;
}

int se_chunks(int HW) {
  int n = (HW + 2047) / 2048;
  return n < 1 ? 1 : (n > 512 ? 512 : n);
}

int se_gate_launch(const __half* in_hi, const __half* in_lo, int HW, int Cpad, int C, int nr, float* partial, const float* w1,
                   const float* b1, const float* w2, const float* b2, float* gate, cudaStream_t s) {
  B2_CHECK(Cpad % 64 == 0 && (C + nr) * 4 <= 48 * 1024, "se gate: channel count out of range");
  const int nch = se_chunks(HW), chunk = (HW + nch - 1) / nch;
  se_partial_kernel<<<dim3(Cpad / 64, nch), 256, 0, s>>>(in_hi, in_lo, HW, Cpad, chunk, partial);
  se_excite_kernel<<<1, 256, static_cast<size_t>(C + nr) * 4, s>>>(partial, nch, Cpad, C, nr, 1.0f / static_cast<float>(HW), w1, b1,
                                                                     w2, b2, gate);
  B2_CUDA(cudaGetLastError());
  return 0;
}

int se_scale_weights_launch(const float* w, const float* gate, int rows, int K, __half* hi, __half* lo, cudaStream_t s) {
  const size_t total = static_cast<size_t>(rows) * K;
  se_scale_weights_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, s>>>(w, gate, rows, K, hi, lo);
  B2_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace b2
