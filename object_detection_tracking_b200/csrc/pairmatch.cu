// Multi-camera ReID track-pair costs (BASELINE config 5): for a pair of cameras, the appearance cost of every
// (track of camera 1, track of camera 2) pair = min over the tracks' crop embeddings of the squared euclidean distance,
// gated by the spatial (top-down trajectory) distance.
//
// Reference (Python loops, one sklearn euclidean_distances call per track pair): multi_video_reid.py:308-324
// compute_feature_dist, :260-305 compute_spatial_dist; the assignment that follows (:512) is b2_lapjv (csrc/tmot.cpp).
// Here: ONE tensor-core GEMM over the concatenated galleries ([Sa,D] x [Sb,D]^T on conv_tc_kernel) and one segmented
// min kernel; the trajectory distance is plain host code (float64, as the reference).
#include <math.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <vector>

#include "../../include/b200det.h"
#include "common.h"
#include "kernels.h"

#include <mutex>

namespace b2 {
namespace {

// Grow-only workspace per device (the pair costs of config 5 are computed pair after pair with similar shapes): operand
// planes, products, norms, segment tables, gate and the GEMM plans cached per padded shape.  Round 1 allocated 13 buffers,
// queried the device properties and encoded the tensor maps on every call (13.7 ms per camera pair on an 8-GPU box, of
// which the GEMM + segmented min are well under a millisecond).
struct PairWs {
  int num_sms = 0;
  cudaStream_t stream = nullptr;
  size_t cap_a = 0, cap_b = 0, cap_sp = 0, cap_np = 0, cap_dp = 0, cap_seg = 0, cap_nm = 0;
  float *d_a = nullptr, *d_b = nullptr, *d_dots = nullptr, *d_out = nullptr, *d_bias = nullptr, *d_na2 = nullptr, *d_nb2 = nullptr;
  int *d_sa = nullptr, *d_sb = nullptr;
  unsigned char* d_gate = nullptr;
  __half *a_hi = nullptr, *a_lo = nullptr, *b_hi = nullptr, *b_lo = nullptr;
  std::map<std::vector<int>, ConvPlan*> plans;
  void release() {
    for (auto& kv : plans) conv_tc_plan_destroy(kv.second);
    plans.clear();
    void* ptrs[] = {d_a, d_b, d_dots, d_out, d_bias, d_na2, d_nb2, d_sa, d_sb, d_gate, a_hi, a_lo, b_hi, b_lo};
    for (void* q : ptrs) if (q) cudaFree(q);
    d_a = d_b = d_dots = d_out = d_bias = d_na2 = d_nb2 = nullptr;
    d_sa = d_sb = nullptr;
    d_gate = nullptr;
    a_hi = a_lo = b_hi = b_lo = nullptr;
  }
};
std::map<int, PairWs> g_pair_ws;
std::mutex g_pair_mutex;

int pair_ws_reserve(PairWs& w, int device, size_t need_a, size_t need_b, int Sp, int Np, int Dp, int nseg, size_t nm) {
  if (!w.stream) {
    B2_CUDA(cudaDeviceGetAttribute(&w.num_sms, cudaDevAttrMultiProcessorCount, device));
    B2_CUDA(cudaStreamCreateWithFlags(&w.stream, cudaStreamNonBlocking));
  }
  if (need_a <= w.cap_a && need_b <= w.cap_b && static_cast<size_t>(Sp) <= w.cap_sp && static_cast<size_t>(Np) <= w.cap_np &&
      static_cast<size_t>(Dp) <= w.cap_dp && static_cast<size_t>(nseg) <= w.cap_seg && nm <= w.cap_nm)
    return 0;
  B2_CUDA(cudaStreamSynchronize(w.stream));
  w.release();                                            // the plans hold the old addresses
  auto grow = [](size_t cap, size_t need) { return need > cap ? need + need / 2 : cap; };
  w.cap_a = grow(w.cap_a, std::max<size_t>(need_a, 1));
  w.cap_b = grow(w.cap_b, std::max<size_t>(need_b, 1));
  w.cap_sp = grow(w.cap_sp, static_cast<size_t>(Sp));
  w.cap_np = grow(w.cap_np, static_cast<size_t>(Np));
  w.cap_dp = std::max(w.cap_dp, static_cast<size_t>(Dp));
  w.cap_seg = grow(w.cap_seg, static_cast<size_t>(nseg));
  w.cap_nm = grow(w.cap_nm, std::max<size_t>(nm, 1));
  B2_CUDA(cudaMalloc(&w.d_a, w.cap_a * 4));
  B2_CUDA(cudaMalloc(&w.d_b, w.cap_b * 4));
  B2_CUDA(cudaMalloc(&w.d_dots, w.cap_sp * w.cap_np * 4));
  B2_CUDA(cudaMalloc(&w.d_out, w.cap_nm * 4));
  B2_CUDA(cudaMalloc(&w.d_bias, w.cap_np * 4));
  B2_CUDA(cudaMalloc(&w.d_na2, w.cap_sp * 4));
  B2_CUDA(cudaMalloc(&w.d_nb2, w.cap_np * 4));
  B2_CUDA(cudaMalloc(&w.d_sa, w.cap_seg * 4));
  B2_CUDA(cudaMalloc(&w.d_sb, w.cap_seg * 4));
  B2_CUDA(cudaMalloc(&w.d_gate, w.cap_nm));
  B2_CUDA(cudaMalloc(&w.a_hi, w.cap_sp * w.cap_dp * 2));
  B2_CUDA(cudaMalloc(&w.a_lo, w.cap_sp * w.cap_dp * 2));
  B2_CUDA(cudaMalloc(&w.b_hi, w.cap_np * w.cap_dp * 2));
  B2_CUDA(cudaMalloc(&w.b_lo, w.cap_np * w.cap_dp * 2));
  B2_CUDA(cudaMemsetAsync(w.d_bias, 0, w.cap_np * 4, w.stream));
  B2_CUDA(cudaMemsetAsync(w.d_na2, 0, w.cap_sp * 4, w.stream));
  B2_CUDA(cudaMemsetAsync(w.d_nb2, 0, w.cap_np * 4, w.stream));
  B2_CUDA(cudaMemsetAsync(w.a_hi, 0, w.cap_sp * w.cap_dp * 2, w.stream));
  B2_CUDA(cudaMemsetAsync(w.a_lo, 0, w.cap_sp * w.cap_dp * 2, w.stream));
  B2_CUDA(cudaMemsetAsync(w.b_hi, 0, w.cap_np * w.cap_dp * 2, w.stream));
  B2_CUDA(cudaMemsetAsync(w.b_lo, 0, w.cap_np * w.cap_dp * 2, w.stream));
  B2_CUDA(cudaStreamSynchronize(w.stream));
  return 0;
}

}  // namespace
}  // namespace b2

using namespace b2;

extern "C" {

// Shared body: `a` / `b` are host pointers (a_dev == 0: uploaded first) or device pointers -- b may be PEER memory of
// another GPU opened through b2_gallery_open: rows_to_planes then reads it over NVLink while converting to the fp16 operand
// planes, i.e. the "exchange" is the operand load of the GEMM's first stage, no staged copy of the peer gallery exists.
// Operand rows beyond Sa / Sb (padding up to the plan's shape) hold finite leftovers of earlier calls: they only produce
// products nobody reads.
static int pair_cost_impl(int device, const float* a, bool a_dev, const int32_t* seg_a, int N, const float* b, bool b_dev,
                          const int32_t* seg_b, int M, int D, const uint8_t* gate, float fill, int precision, float* out) {
  B2_CHECK(N >= 0 && M >= 0 && D > 0, "b2_track_pair_cost: bad argument");
  if (N == 0 || M == 0) return 0;
  B2_CHECK(a && seg_a && b && seg_b && out, "b2_track_pair_cost: null argument");
  const int Sa = seg_a[N], Sb = seg_b[M];
  B2_CHECK(seg_a[0] == 0 && seg_b[0] == 0, "b2_track_pair_cost: segment offsets must start at 0");
  for (int i = 0; i < N; ++i) B2_CHECK(seg_a[i + 1] >= seg_a[i], "b2_track_pair_cost: seg_a not monotone");
  for (int j = 0; j < M; ++j) B2_CHECK(seg_b[j + 1] >= seg_b[j], "b2_track_pair_cost: seg_b not monotone");
  if (Sa == 0 || Sb == 0) {
    for (size_t k = 0; k < static_cast<size_t>(N) * M; ++k) out[k] = fill;
    return 0;
  }
  B2_CUDA(cudaSetDevice(device));
  const bool split = precision == 1;
  const int Dp = (D + 63) / 64 * 64, Np = (Sb + 15) / 16 * 16, Sp = (Sa + 127) / 128 * 128;
  std::lock_guard<std::mutex> lock(g_pair_mutex);
  PairWs& w = g_pair_ws[device];
  if (pair_ws_reserve(w, device, a_dev ? 0 : static_cast<size_t>(Sa) * D, b_dev ? 0 : static_cast<size_t>(Sb) * D, Sp, Np, Dp,
                      std::max(N, M) + 1, static_cast<size_t>(N) * M))
    return -1;
  cudaStream_t st = w.stream;
  const float* d_a = a;
  const float* d_b = b;
  if (!a_dev) {
    B2_CUDA(cudaMemcpyAsync(w.d_a, a, sizeof(float) * Sa * D, cudaMemcpyHostToDevice, st));
    d_a = w.d_a;
  }
  if (!b_dev) {
    B2_CUDA(cudaMemcpyAsync(w.d_b, b, sizeof(float) * Sb * D, cudaMemcpyHostToDevice, st));
    d_b = w.d_b;
  }
  B2_CUDA(cudaMemcpyAsync(w.d_sa, seg_a, sizeof(int) * (N + 1), cudaMemcpyHostToDevice, st));
  B2_CUDA(cudaMemcpyAsync(w.d_sb, seg_b, sizeof(int) * (M + 1), cudaMemcpyHostToDevice, st));
  if (gate) B2_CUDA(cudaMemcpyAsync(w.d_gate, gate, static_cast<size_t>(N) * M, cudaMemcpyHostToDevice, st));
  if (rows_to_planes(d_a, Sa, D, w.a_hi, w.a_lo, Dp, w.d_na2, st) || rows_to_planes(d_b, Sb, D, w.b_hi, w.b_lo, Dp, w.d_nb2, st))
    return -1;
  const std::vector<int> key = {Sp, Np, Dp, split ? 1 : 0};
  auto it = w.plans.find(key);
  if (it == w.plans.end()) {
    ConvDesc d;
    d.B = 1; d.in_H = 1; d.in_W = Sp; d.Cin = Dp; d.in_pitch_H = 1; d.in_pitch_W = Sp; d.in_ld = Dp;
    d.Cout = Np; d.out_H = 1; d.out_W = Sp; d.ldc = Np;
    ConvWeights cw;
    cw.w_hi = w.b_hi; cw.w_lo = split ? w.b_lo : nullptr; cw.bias = w.d_bias; cw.Cout_pad = Np; cw.K = Dp;
    ConvIO io;
    io.in_hi = w.a_hi; io.in_lo = split ? w.a_lo : nullptr; io.out_f32 = w.d_dots;
    ConvPlan* plan = conv_tc_plan_create(d, cw, io, split, w.num_sms);
    B2_CHECK(plan != nullptr, std::string("b2_track_pair_cost: ") + last_error());
    it = w.plans.emplace(key, plan).first;
  }
  if (conv_tc_launch(it->second, st)) return -1;
  if (pair_segmin(w.d_dots, Np, w.d_na2, w.d_nb2, w.d_sa, N, w.d_sb, M, gate ? w.d_gate : nullptr, fill, w.d_out, st)) return -1;
  B2_CUDA(cudaMemcpyAsync(out, w.d_out, sizeof(float) * N * M, cudaMemcpyDeviceToHost, st));
  B2_CUDA(cudaStreamSynchronize(st));
  return 0;
}

int b2_track_pair_cost(int device, const float* a, const int32_t* seg_a, int N, const float* b, const int32_t* seg_b,
                       int M, int D, const uint8_t* gate, float fill, int precision, float* out) {
  return pair_cost_impl(device, a, false, seg_a, N, b, false, seg_b, M, D, gate, fill, precision, out);
}

int b2_track_pair_cost_dev(int device, const float* a_dev, const int32_t* seg_a, int N, const float* b_dev,
                           const int32_t* seg_b, int M, int D, const uint8_t* gate, float fill, int precision, float* out) {
  return pair_cost_impl(device, a_dev, true, seg_a, N, b_dev, true, seg_b, M, D, gate, fill, precision, out);
}

// ---- galleries in device memory that other processes of the node can map (one process per GPU) ----
int b2_gallery_create(int device, const float* feats_host, int rows, int D, float** dev_out, uint8_t handle_out[64]) {
  B2_CHECK(dev_out && handle_out && rows >= 0 && D > 0 && (rows == 0 || feats_host), "b2_gallery_create: bad argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  *dev_out = nullptr;
  B2_CUDA(cudaSetDevice(device));
  float* p = nullptr;
  B2_CUDA(cudaMalloc(&p, std::max<size_t>(static_cast<size_t>(rows) * D, 1) * sizeof(float)));
  cudaError_t e = rows ? cudaMemcpy(p, feats_host, static_cast<size_t>(rows) * D * sizeof(float), cudaMemcpyHostToDevice)
                       : cudaSuccess;
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    B2_CUDA(e);
  }
  memcpy(handle_out, &h, 64);
  *dev_out = p;
  return 0;
}

int b2_gallery_open(int device, const uint8_t handle[64], float** peer_out) {
  B2_CHECK(handle && peer_out, "b2_gallery_open: null argument");
  *peer_out = nullptr;
  B2_CUDA(cudaSetDevice(device));
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, 64);
  void* p = nullptr;
  B2_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));   // enables P2P (NVLink) access to the owner GPU
  *peer_out = static_cast<float*>(p);
  return 0;
}

int b2_gallery_close(int device, float* peer) {
  if (!peer) return 0;
  B2_CUDA(cudaSetDevice(device));
  B2_CUDA(cudaIpcCloseMemHandle(peer));
  return 0;
}

int b2_gallery_free(int device, float* dev) {
  if (!dev) return 0;
  B2_CUDA(cudaSetDevice(device));
  B2_CUDA(cudaFree(dev));
  return 0;
}

// multi_video_reid.py:260-305 compute_spatial_dist: trajectories as (frame, x, y) rows grouped per track; for every
// track pair the mean point distance over the frames both tracks have (camera-2 frames shifted by frame_offset), kept
// if <= tol, else 9999.  A frame that occurs twice in a track keeps its LAST point (dict semantics, :269-275).
int b2_track_spatial_dist(const int32_t* frames1, const double* pts1, const int32_t* seg1, int N, const int32_t* frames2,
                          const double* pts2, const int32_t* seg2, int M, int frame_offset, double tol, double* out) {
  B2_CHECK(N >= 0 && M >= 0, "b2_track_spatial_dist: bad argument");
  if (N == 0 || M == 0) return 0;
  B2_CHECK(frames1 && pts1 && seg1 && frames2 && pts2 && seg2 && out, "b2_track_spatial_dist: null argument");
  std::vector<std::map<int, const double*>> t2(M);
  for (int j = 0; j < M; ++j)
    for (int k = seg2[j]; k < seg2[j + 1]; ++k) t2[j][frames2[k] + frame_offset] = pts2 + 2 * k;
  for (int i = 0; i < N; ++i) {
    std::map<int, const double*> t1;
    for (int k = seg1[i]; k < seg1[i + 1]; ++k) t1[frames1[k]] = pts1 + 2 * k;
    for (int j = 0; j < M; ++j) {
      double sum = 0;
      int cnt = 0;
      for (const auto& kv : t1) {
        auto it = t2[j].find(kv.first);
        if (it == t2[j].end()) continue;
        const double dx = kv.second[0] - it->second[0], dy = kv.second[1] - it->second[1];
        sum += sqrt(dx * dx + dy * dy);
        ++cnt;
      }
      double v = 9999.0;
      if (cnt > 0) {
        const double mean = sum / cnt;
        if (mean <= tol) v = mean;
      }
      out[static_cast<size_t>(i) * M + j] = v;
    }
  }
  return 0;
}

}  // extern "C"
