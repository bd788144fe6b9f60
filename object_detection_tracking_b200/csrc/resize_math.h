// Per-pixel arithmetic of the frame-resize ingest kernel: cv2.resize(im, (neww, newh), interpolation=cv2.INTER_LINEAR)
// on a float32 frame, as the reference drivers call it (nn.py:1540-1545 resizeImage; obj_detect_tracking.py:597-605,
// enqueuer_thread.py:259-263: uint8 frame -> astype(float32) -> resize).  OpenCV is a third-party dependency absent from
// /root/reference; its published bilinear algorithm (imgproc resize.cpp, CV_32F path) is restated: half-pixel centres,
// source coordinate in float from a double scale, horizontal taps clamped with the weight reset to 0 at the borders,
// vertical taps clamped with the weight kept, horizontal pass first, plain float multiply-adds.
// Shared by the CUDA kernel (stem.cu) and a host harness the CPU tests compile (tests/test_resize_cpu.py) so that the
// kernel's arithmetic is checked against the cv2-pinned oracle without a GPU.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define B2_HD __host__ __device__ __forceinline__
#else
#define B2_HD inline
#endif

#if defined(__CUDA_ARCH__)
#define B2_FMUL(a, b) __fmul_rn((a), (b))
#define B2_FADD(a, b) __fadd_rn((a), (b))
#define B2_DMUL(a, b) __dmul_rn((a), (b))     // nvcc would contract the double multiply-subtract into one DFMA
#define B2_DSUB(a, b) __dsub_rn((a), (b))
#else
#define B2_FMUL(a, b) ((a) * (b))
#define B2_FADD(a, b) ((a) + (b))
#define B2_DMUL(a, b) ((a) * (b))
#define B2_DSUB(a, b) ((a) - (b))
#endif

namespace b2 {

struct ResizeTap {
  int i0, i1;      // source indices
  float w0, w1;    // weights
};

// horizontal: border taps collapse to one sample with weight 1
B2_HD ResizeTap resize_tap_x(int d, int src, int dst) {
  const double scale = 1.0 / (static_cast<double>(dst) / static_cast<double>(src));
  float f = static_cast<float>(B2_DSUB(B2_DMUL(d + 0.5, scale), 0.5));
  int s = static_cast<int>(floorf(f));
  f -= static_cast<float>(s);
  if (s < 0) {
    s = 0;
    f = 0.f;
  }
  if (s >= src - 1) {
    s = src - 1;
    f = 0.f;
  }
  ResizeTap t;
  t.i0 = s;
  t.i1 = s + 1 < src ? s + 1 : src - 1;
  t.w0 = B2_FADD(1.f, -f);
  t.w1 = f;
  return t;
}

// vertical: rows are clamped, the weight is not reset
B2_HD ResizeTap resize_tap_y(int d, int src, int dst) {
  const double scale = 1.0 / (static_cast<double>(dst) / static_cast<double>(src));
  float f = static_cast<float>(B2_DSUB(B2_DMUL(d + 0.5, scale), 0.5));
  const int s = static_cast<int>(floorf(f));
  f -= static_cast<float>(s);
  ResizeTap t;
  t.i0 = s < 0 ? 0 : (s > src - 1 ? src - 1 : s);
  t.i1 = s + 1 < 0 ? 0 : (s + 1 > src - 1 ? src - 1 : s + 1);
  t.w0 = B2_FADD(1.f, -f);
  t.w1 = f;
  return t;
}

B2_HD float resize_sample(float p00, float p01, float p10, float p11, const ResizeTap& tx, const ResizeTap& ty) {
  const float r0 = B2_FADD(B2_FMUL(p00, tx.w0), B2_FMUL(p01, tx.w1));
  const float r1 = B2_FADD(B2_FMUL(p10, tx.w0), B2_FMUL(p11, tx.w1));
  return B2_FADD(B2_FMUL(r0, ty.w0), B2_FMUL(r1, ty.w1));
}

}  // namespace b2
