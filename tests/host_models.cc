// host_models.cc -- compiles the PRODUCT's device-side camera-model code (theiasfm_b200/csrc/tba_camera_models.cuh)
// for the host, so that the analytic Jacobian / residual / loss code that runs on the GPU is also checked by the
// CPU-only test suite against the committed golden vectors (tests/test_device_math_on_host.py).  The header is plain
// C++ once the CUDA qualifiers are defined away; nothing here is used by the product.
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#ifndef _GNU_SOURCE
#define _GNU_SOURCE
#endif
#include <cmath>
using std::atan;
using std::atan2;
using std::fabs;
using std::fmax;
using std::sqrt;
using std::tan;

#include "../theiasfm_b200/csrc/tba_camera_models.cuh"

extern "C" {

// One observation through cam_prep + linearize_obs_any<all 10 intrinsics columns, EXT> (the five camera models; models 0/1
// take exactly the code of the EXT = false kernels).
// Outputs: r[2], rho0, J[2][20] in the golden column order [C(3) w(3) | intr(10) | X(3) h].
int host_linearize(int model, const double* ext, const double* intr, const double* pt, const double* xy, int loss_type,
                   double loss_width, double* r, double* rho0, double* J) {
  double rec[tba::kCamRec];
  tba::cam_prep(ext + 3, rec);
  double Ja[6], Jw[6], Jh[2], Ji[20];
  const bool ok = tba::linearize_obs_any<0x3FFu, true>(model, ext, rec, intr, pt[0], pt[1], pt[2], pt[3], xy[0], xy[1], loss_type, loss_width,
                                             r, *rho0, Ja, Jw, Jh, Ji);
  if (!ok) return 0;
  for (int row = 0; row < 2; ++row) {
    double* Jr = J + row * 20;
    for (int j = 0; j < 3; ++j) Jr[j] = -pt[3] * Ja[row * 3 + j];  // J_C = -h J_a
    for (int j = 0; j < 3; ++j) Jr[3 + j] = Jw[row * 3 + j];
    for (int j = 0; j < 10; ++j) Jr[6 + j] = Ji[row * 10 + j];
    for (int j = 0; j < 3; ++j) Jr[16 + j] = Ja[row * 3 + j];
    Jr[19] = Jh[row];
  }
  return 1;
}

int host_reproject(int model, const double* ext, const double* intr, const double* pt, const double* xy, double* r) {
  double rec[tba::kCamRec];
  tba::cam_prep(ext + 3, rec);
  return tba::reproject_any<true>(model, ext, rec, intr, pt[0], pt[1], pt[2], pt[3], xy[0], xy[1], r[0], r[1]) ? 1 : 0;
}

// PixelToCameraCoordinates of FISHEYE / FOV / DIVISION_UNDISTORTION (the viewing rays of the track estimator).
void host_pixel_to_camera_ext(int model, const double* intr, const double* pix, double* out) {
  tba::pixel_to_camera_ext(model, intr, pix[0], pix[1], out[0], out[1]);
  out[2] = 1.0;
}

void host_loss(int type, double a, double s, double* rho) { tba::loss_evaluate(type, a, s, rho); }

}  // extern "C"
