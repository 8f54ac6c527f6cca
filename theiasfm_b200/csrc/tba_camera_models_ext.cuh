// tba_camera_models_ext.cuh -- FISHEYE, FOV and DIVISION_UNDISTORTION camera models (the reference
// instantiates ReprojectionError<Model> for them at create_reprojection_error_cost_function.h:72-90).
// These models are not on the benchmarked path, so instead of hand-derived Jacobians the projection is written ONCE as a
// template over the scalar type and differentiated with a small forward-mode dual number (the same mathematical object
// as the reference's ceres::Jet): T = double gives Camera::ProjectPoint / the candidate cost, T = Dual<13> gives
// d pixel / d (camera-frame point, 10 intrinsics), which the common chain of tba_camera_models.cuh turns into the
// compact per-observation linearisation.  The PINHOLE / PINHOLE_RADIAL_TANGENTIAL kernels are separate template
// instantiations (EXT = false) and do not contain any of this code.
//   FISHEYE                fisheye_camera_model.h:160-187 (CameraToPixelCoordinates), :224-270 (DistortPoint, 3-D input)
//   FOV                    fov_camera_model.h:157-181, :212-258 (three branches on omega and r_u^2)
//   DIVISION_UNDISTORTION  division_undistortion_camera_model.h:171-202, :256-286 (distortion applied in pixel units)
// and their inverses for the viewing rays of the track estimator (UndistortPoint: fisheye :272-335 iterative,
// fov :260-300 closed form, division :288-311 closed form).
#pragma once
#include <cfloat>
#include <cmath>

namespace tba {

constexpr int kModelFisheye = 2;
constexpr int kModelFov = 3;
constexpr int kModelDivisionUndistortion = 4;

__host__ __device__ constexpr int model_num_parameters(int model) {
  return model == 0 ? 7 : model == 1 ? 10 : model == 2 ? 9 : (model == 3 || model == 4) ? 5 : 0;
}

// ---------------------------------------------------------------- forward-mode dual numbers
template <int N>
struct Dual {
  double v;
  double d[N];
};

__host__ __device__ inline double value_of(double a) { return a; }
template <int N>
__host__ __device__ inline double value_of(const Dual<N>& a) { return a.v; }

#define TBA_DUAL_FOR for (int i_ = 0; i_ < N; ++i_)
template <int N> __host__ __device__ inline Dual<N> operator+(const Dual<N>& a, const Dual<N>& b) { Dual<N> r; r.v = a.v + b.v; TBA_DUAL_FOR r.d[i_] = a.d[i_] + b.d[i_]; return r; }
template <int N> __host__ __device__ inline Dual<N> operator-(const Dual<N>& a, const Dual<N>& b) { Dual<N> r; r.v = a.v - b.v; TBA_DUAL_FOR r.d[i_] = a.d[i_] - b.d[i_]; return r; }
template <int N> __host__ __device__ inline Dual<N> operator-(const Dual<N>& a) { Dual<N> r; r.v = -a.v; TBA_DUAL_FOR r.d[i_] = -a.d[i_]; return r; }
template <int N> __host__ __device__ inline Dual<N> operator*(const Dual<N>& a, const Dual<N>& b) { Dual<N> r; r.v = a.v * b.v; TBA_DUAL_FOR r.d[i_] = a.v * b.d[i_] + a.d[i_] * b.v; return r; }
template <int N> __host__ __device__ inline Dual<N> operator/(const Dual<N>& a, const Dual<N>& b) {
  Dual<N> r; const double inv = 1.0 / b.v; r.v = a.v * inv;
  TBA_DUAL_FOR r.d[i_] = (a.d[i_] - r.v * b.d[i_]) * inv;
  return r;
}
template <int N> __host__ __device__ inline Dual<N> operator+(const Dual<N>& a, double b) { Dual<N> r = a; r.v += b; return r; }
template <int N> __host__ __device__ inline Dual<N> operator+(double b, const Dual<N>& a) { Dual<N> r = a; r.v += b; return r; }
template <int N> __host__ __device__ inline Dual<N> operator-(const Dual<N>& a, double b) { Dual<N> r = a; r.v -= b; return r; }
template <int N> __host__ __device__ inline Dual<N> operator-(double b, const Dual<N>& a) { Dual<N> r; r.v = b - a.v; TBA_DUAL_FOR r.d[i_] = -a.d[i_]; return r; }
template <int N> __host__ __device__ inline Dual<N> operator*(const Dual<N>& a, double b) { Dual<N> r; r.v = a.v * b; TBA_DUAL_FOR r.d[i_] = a.d[i_] * b; return r; }
template <int N> __host__ __device__ inline Dual<N> operator*(double b, const Dual<N>& a) { return a * b; }
template <int N> __host__ __device__ inline Dual<N> operator/(const Dual<N>& a, double b) { return a * (1.0 / b); }
template <int N> __host__ __device__ inline Dual<N> dsqrt(const Dual<N>& a) { Dual<N> r; r.v = sqrt(a.v); const double t = 0.5 / r.v; TBA_DUAL_FOR r.d[i_] = a.d[i_] * t; return r; }
template <int N> __host__ __device__ inline Dual<N> dabs(const Dual<N>& a) { return a.v < 0.0 ? -a : a; }
template <int N> __host__ __device__ inline Dual<N> dtan(const Dual<N>& a) { Dual<N> r; r.v = tan(a.v); const double t = 1.0 + r.v * r.v; TBA_DUAL_FOR r.d[i_] = a.d[i_] * t; return r; }
template <int N> __host__ __device__ inline Dual<N> datan(const Dual<N>& a) { Dual<N> r; r.v = atan(a.v); const double t = 1.0 / (1.0 + a.v * a.v); TBA_DUAL_FOR r.d[i_] = a.d[i_] * t; return r; }
template <int N> __host__ __device__ inline Dual<N> datan2(const Dual<N>& y, const Dual<N>& x) {
  Dual<N> r; r.v = atan2(y.v, x.v);
  const double t = 1.0 / (x.v * x.v + y.v * y.v);
  TBA_DUAL_FOR r.d[i_] = t * (x.v * y.d[i_] - y.v * x.d[i_]);
  return r;
}
#undef TBA_DUAL_FOR
__host__ __device__ inline double dsqrt(double a) { return sqrt(a); }
__host__ __device__ inline double dabs(double a) { return fabs(a); }
__host__ __device__ inline double dtan(double a) { return tan(a); }
__host__ __device__ inline double datan(double a) { return atan(a); }
__host__ __device__ inline double datan2(double y, double x) { return atan2(y, x); }

// ---------------------------------------------------------------- CameraToPixelCoordinates<T>
// q = point in the camera frame, k = the model's intrinsics (layouts: fisheye f, ar, skew, cx, cy, k1..k4;
// fov / division f, ar, cx, cy, omega | k).  Branches are taken on values, as the reference's Jet code does.
template <class T>
__host__ __device__ inline void camera_to_pixel_ext(int model, const T* k, const T q[3], T pix[2]) {
  if (model == kModelFisheye) {
    T ud, vd;
    const T r_sq = q[0] * q[0] + q[1] * q[1];
    if (value_of(r_sq) < 1e-8) { ud = q[0]; vd = q[1]; }
    else {
      const T r_num = dsqrt(r_sq);
      const T theta = datan2(r_num, dabs(q[2]));
      const T t2 = theta * theta;
      const T theta_d = theta * (1.0 + k[5] * t2 + k[6] * t2 * t2 + k[7] * t2 * t2 * t2 + k[8] * t2 * t2 * t2 * t2);
      ud = theta_d * q[0] / r_num; vd = theta_d * q[1] / r_num;
      if (value_of(q[2]) < 0.0) { ud = -ud; vd = -vd; }
    }
    pix[0] = k[0] * ud + k[2] * vd + k[3];
    pix[1] = k[0] * k[1] * vd + k[4];
  } else if (model == kModelFov) {
    const T u = q[0] / q[2], v = q[1] / q[2];
    const T omega = k[4];
    const T r_u_sq = u * u + v * v;
    T r_d;
    if (value_of(omega) < 1e-3) r_d = (omega * omega * r_u_sq) / 3.0 - omega * omega / 12.0 + 1.0;
    else if (value_of(r_u_sq) < 1e-3) { const T th = dtan(omega / 2.0); r_d = (-2.0 * th * (4.0 * r_u_sq * th * th - 3.0)) / (3.0 * omega); }
    else { const T r_u = dsqrt(r_u_sq); r_d = datan(2.0 * r_u * dtan(omega / 2.0)) / (r_u * omega); }
    pix[0] = k[0] * (r_d * u) + k[2];
    pix[1] = k[0] * k[1] * (r_d * v) + k[3];
  } else {
    const T u = q[0] / q[2], v = q[1] / q[2];
    const T up0 = k[0] * u, up1 = k[0] * k[1] * v;
    const T r_u_sq = up0 * up0 + up1 * up1;
    const T denom = 2.0 * k[4] * r_u_sq;
    const T inner = 1.0 - 4.0 * k[4] * r_u_sq;
    T d0 = up0, d1 = up1;
    if (!(fabs(value_of(denom)) < DBL_EPSILON || value_of(inner) < 0.0)) {
      const T scale = (1.0 - dsqrt(inner)) / denom;
      d0 = up0 * scale; d1 = up1 * scale;
    }
    pix[0] = d0 + k[2];
    pix[1] = d1 + k[3];
  }
}

// PixelToCameraCoordinates (ray with z = 1) of the three models.
__host__ __device__ inline void pixel_to_camera_ext(int model, const double* __restrict__ k, double x, double y, double& xu, double& yu) {
  if (model == kModelFisheye) {
    const double yd = (y - k[4]) / (k[0] * k[1]);
    const double xd = (x - k[3] - yd * k[2]) / k[0];
    xu = xd; yu = yd;
    for (int i = 0; i < 100; ++i) {
      const double px = xu, py = yu;
      const double r = sqrt(xu * xu + yu * yu);
      if (r < 1e-8) { xu = xd; yu = yd; return; }
      const double theta = atan2(r, 1.0), t2 = theta * theta;
      const double theta_d = theta * (1.0 + k[5] * t2 + k[6] * t2 * t2 + k[7] * t2 * t2 * t2 + k[8] * t2 * t2 * t2 * t2);
      xu = r * xd / theta_d; yu = r * yd / theta_d;
      if (fabs(xu - px) < 1e-10 && fabs(yu - py) < 1e-10) break;
    }
  } else if (model == kModelFov) {
    const double d0 = (x - k[2]) / k[0], d1 = (y - k[3]) / (k[0] * k[1]), omega = k[4], r_d_sq = d0 * d0 + d1 * d1;
    double r_u;
    if (omega < 1e-3) r_u = (omega * omega * r_d_sq) / 3.0 - omega * omega / 12.0 + 1.0;
    else if (r_d_sq < 1e-3) r_u = (omega * (omega * omega * r_d_sq + 3.0)) / (6.0 * tan(omega / 2.0));
    else { const double r_d = sqrt(r_d_sq); r_u = tan(r_d * omega) / (2.0 * r_d * tan(omega / 2.0)); }
    xu = r_u * d0; yu = r_u * d1;
  } else {
    const double d0 = x - k[2], d1 = y - k[3];
    const double und = 1.0 / (1.0 + k[4] * (d0 * d0 + d1 * d1));
    xu = d0 * und / k[0]; yu = d1 * und / (k[0] * k[1]);
  }
}

}  // namespace tba
