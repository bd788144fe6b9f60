// Association primitives shared by the native trackers (tracker.cpp: DeepSORT, tmot.cpp: JDE/TMOT): the float64
// Kalman filter of deep_sort/kalman_filter.py:23-232 (tmot/kalman_filter.py is the same arithmetic plus the
// vectorised multi_predict, whose covariance product associates as (F P) F^T instead of F (P F^T)) and the
// rectangular linear-sum-assignment solver restated from SciPy (linear_assignment.py:5,60).
#pragma once
#include <math.h>
#include <string.h>

#include <vector>

namespace b2 {

constexpr double kChi2Inv95_4 = 9.4877;     // kalman_filter.py:11-20 (4 degrees of freedom)
constexpr double kStdPos = 1.0 / 20, kStdVel = 1.0 / 160;   // kalman_filter.py:51-52

// scipy.optimize.linear_sum_assignment on a row-major [nr,nc] matrix; (rows, cols) sorted by row; -1 = infeasible.
int lsap(int nr, int nc, const double* cost_in, std::vector<int>& rows, std::vector<int>& cols);

inline double sq(double x) { return x * x; }

inline void kf_initiate(const double z[4], double mean[8], double cov[64]) {   // :55-87
  for (int i = 0; i < 4; ++i) {
    mean[i] = z[i];
    mean[4 + i] = 0;
  }
  const double h = z[3];
  const double std_[8] = {2 * kStdPos * h, 2 * kStdPos * h, 1e-2, 2 * kStdPos * h,
                          10 * kStdVel * h, 10 * kStdVel * h, 1e-5, 10 * kStdVel * h};
  memset(cov, 0, sizeof(double) * 64);
  for (int i = 0; i < 8; ++i) cov[i * 9] = sq(std_[i]);
}

inline void kf_predict(double mean[8], double cov[64]) {   // :89-124: mean = F mean ; cov = F (cov F^T) + Q
  const double h = mean[3];
  const double q[8] = {sq(kStdPos * h), sq(kStdPos * h), sq(1e-2), sq(kStdPos * h),
                       sq(kStdVel * h), sq(kStdVel * h), sq(1e-5), sq(kStdVel * h)};
  for (int i = 0; i < 4; ++i) mean[i] = mean[i] + mean[i + 4];
  double x[64];
  for (int i = 0; i < 8; ++i)
    for (int j = 0; j < 8; ++j) x[i * 8 + j] = j < 4 ? cov[i * 8 + j] + cov[i * 8 + j + 4] : cov[i * 8 + j];
  for (int i = 0; i < 8; ++i)
    for (int j = 0; j < 8; ++j) cov[i * 8 + j] = i < 4 ? x[i * 8 + j] + x[(i + 4) * 8 + j] : x[i * 8 + j];
  for (int i = 0; i < 8; ++i) cov[i * 9] += q[i];
}

// tmot/kalman_filter.py:154-194 multi_predict: same model, covariance = (F cov) F^T + Q
inline void kf_predict_fp_ft(double mean[8], double cov[64]) {
  const double h = mean[3];
  const double q[8] = {sq(kStdPos * h), sq(kStdPos * h), sq(1e-2), sq(kStdPos * h),
                       sq(kStdVel * h), sq(kStdVel * h), sq(1e-5), sq(kStdVel * h)};
  for (int i = 0; i < 4; ++i) mean[i] = mean[i] + mean[i + 4];
  double x[64];
  for (int i = 0; i < 8; ++i)
    for (int j = 0; j < 8; ++j) x[i * 8 + j] = i < 4 ? cov[i * 8 + j] + cov[(i + 4) * 8 + j] : cov[i * 8 + j];
  for (int i = 0; i < 8; ++i)
    for (int j = 0; j < 8; ++j) cov[i * 8 + j] = j < 4 ? x[i * 8 + j] + x[i * 8 + j + 4] : x[i * 8 + j];
  for (int i = 0; i < 8; ++i) cov[i * 9] += q[i];
}

inline void kf_project(const double mean[8], const double cov[64], double pm[4], double pc[16]) {   // :126-154
  const double h = mean[3];
  const double r[4] = {sq(kStdPos * h), sq(kStdPos * h), sq(1e-1), sq(kStdPos * h)};
  for (int i = 0; i < 4; ++i) {
    pm[i] = mean[i];
    for (int j = 0; j < 4; ++j) pc[i * 4 + j] = cov[i * 8 + j];
    pc[i * 5] += r[i];
  }
}

inline bool chol4(const double a[16], double L[16]) {   // lower Cholesky factor of a 4x4 SPD matrix
  memset(L, 0, sizeof(double) * 16);
  for (int j = 0; j < 4; ++j) {
    double d = a[j * 4 + j];
    for (int k = 0; k < j; ++k) d -= L[j * 4 + k] * L[j * 4 + k];
    if (!(d > 0)) return false;
    L[j * 4 + j] = sqrt(d);
    for (int i = j + 1; i < 4; ++i) {
      double s = a[i * 4 + j];
      for (int k = 0; k < j; ++k) s -= L[i * 4 + k] * L[j * 4 + k];
      L[i * 4 + j] = s / L[j * 4 + j];
    }
  }
  return true;
}

inline bool kf_update(double mean[8], double cov[64], const double z[4]) {   // :156-190
  double pm[4], pc[16], L[16];
  kf_project(mean, cov, pm, pc);
  if (!chol4(pc, L)) return false;
  // kalman_gain = cho_solve(pc, (cov H^T)^T)^T : for every state row i solve pc g = cov[i, :4]
  double gain[8 * 4];
  for (int i = 0; i < 8; ++i) {
    double y[4], g[4];
    for (int r = 0; r < 4; ++r) {
      double s = cov[i * 8 + r];
      for (int k = 0; k < r; ++k) s -= L[r * 4 + k] * y[k];
      y[r] = s / L[r * 5];
    }
    for (int r = 3; r >= 0; --r) {
      double s = y[r];
      for (int k = r + 1; k < 4; ++k) s -= L[k * 4 + r] * g[k];
      g[r] = s / L[r * 5];
    }
    for (int r = 0; r < 4; ++r) gain[i * 4 + r] = g[r];
  }
  double innov[4];
  for (int r = 0; r < 4; ++r) innov[r] = z[r] - pm[r];
  for (int i = 0; i < 8; ++i) {
    double s = 0;
    for (int r = 0; r < 4; ++r) s += innov[r] * gain[i * 4 + r];
    mean[i] += s;
  }
  // cov -= gain (pc gain^T)   (numpy multi_dot evaluates K (S K^T) for these shapes)
  double skt[4 * 8];
  for (int r = 0; r < 4; ++r)
    for (int j = 0; j < 8; ++j) {
      double s = 0;
      for (int k = 0; k < 4; ++k) s += pc[r * 4 + k] * gain[j * 4 + k];
      skt[r * 8 + j] = s;
    }
  for (int i = 0; i < 8; ++i)
    for (int j = 0; j < 8; ++j) {
      double s = 0;
      for (int r = 0; r < 4; ++r) s += gain[i * 4 + r] * skt[r * 8 + j];
      cov[i * 8 + j] -= s;
    }
  return true;
}

// squared Mahalanobis distance of measurement z to the projected state (:192-232)
inline bool kf_gating(const double mean[8], const double cov[64], const double* zs, int n, double* out) {
  double pm[4], pc[16], L[16];
  kf_project(mean, cov, pm, pc);
  if (!chol4(pc, L)) return false;
  for (int d = 0; d < n; ++d) {
    double y[4], acc = 0;
    for (int r = 0; r < 4; ++r) {
      double s = zs[d * 4 + r] - pm[r];
      for (int k = 0; k < r; ++k) s -= L[r * 4 + k] * y[k];
      y[r] = s / L[r * 5];
    }
    for (int r = 0; r < 4; ++r) acc += y[r] * y[r];
    out[d] = acc;
  }
  return true;
}


}  // namespace b2
