// host_inner.cc -- the PRODUCT's inner-iteration machinery on the host: the lockstep per-block LM driver (tba_block_lm.h, the
// very code the engine runs) with the observation passes evaluated by the per-slot device bodies (tba_inner.cuh, compiled for
// the host) over the packed layout of tba_debug_pack, then the point stage (tba_point_lm.cuh).  tests/test_inner_iterations.py
// compares the result with the oracle's recursive sub-solves.
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#ifndef _GNU_SOURCE
#define _GNU_SOURCE
#endif
#include <cmath>
#include <vector>
using std::atan; using std::atan2; using std::fabs; using std::fmax; using std::fmin; using std::isfinite; using std::sqrt; using std::tan;

#include "../include/theia_ba_b200.h"
#include "../theiasfm_b200/csrc/tba_point_lm.cuh"
#include "../theiasfm_b200/csrc/tba_inner.cuh"
#include "../theiasfm_b200/csrc/tba_block_lm.h"

using namespace tba;

namespace {
struct Packed {
  int n_cam, n_group, n_pk; long long n_slots;
  double *ext, *intr, *pt;   // in/out (pt: packed order)
  const unsigned char* ext_const; const int* cam_group; const int* group_model; const unsigned int* group_mask; const double* mask;
  const double* xy; const int* slot_cam; const int* slot_pt; const long long* pt_slot; const int* pt_len; const unsigned char* pt_const;
  int loss_type; double loss_width;
};

template <int KIND>
void run_stage(Packed& P) {
  constexpr int ND = block_dim(KIND), NA = block_acc(KIND), NHD = ND * (ND + 1) / 2;
  const int nb = KIND == kBlockCamera ? P.n_cam : P.n_group;
  const BlockLmOptions lo;
  std::vector<BlockLm> B((size_t)nb);
  std::vector<int> dims((size_t)nb, ND);
  double* vals0 = KIND == kBlockCamera ? P.ext : P.intr;
  for (int b = 0; b < nb; ++b) {
    bool fr[kBlkMaxN];
    const int N = KIND == kBlockCamera ? 6 : TBA_MODEL_NUM_PARAMETERS(P.group_model[b]);
    dims[b] = N;
    for (int j = 0; j < N; ++j) fr[j] = P.mask[(KIND == kBlockCamera ? (size_t)b * 6 : (size_t)P.n_cam * 6 + (size_t)b * 10) + j] != 0.0;
    block_lm_init(B[b], N, fr, vals0 + (size_t)b * ND, lo);
  }
  std::vector<double> rec((size_t)P.n_cam * kCamRec);
  auto exec = [&](int pass, const std::vector<uint8_t>& active, const std::vector<double>& pv, std::vector<double>& sum) -> int {
    std::fill(sum.begin(), sum.end(), 0.0);
    const double* ext = KIND == kBlockCamera ? pv.data() : P.ext;
    const double* intr = KIND == kBlockCamera ? P.intr : pv.data();
    for (int c = 0; c < P.n_cam; ++c) cam_prep(ext + (size_t)c * 6 + 3, rec.data() + (size_t)c * kCamRec);
    for (long long s = 0; s < P.n_slots; ++s) {   // body of k_block_pass, one slot at a time
      const int cam = P.slot_cam[s];
      if (cam < 0) continue;
      const int grp = P.cam_group[cam];
      const int blk = KIND == kBlockCamera ? cam : grp;
      if (!active[blk]) continue;
      double* acc = &sum[(size_t)blk * NA];
      const long long wq = s >> 5; const int l = (int)(s & 31);
      const double x = P.xy[(size_t)(wq * 2) * 32 + l], y = P.xy[(size_t)(wq * 2 + 1) * 32 + l];
      const double* X = P.pt + (size_t)P.slot_pt[s] * 4;
      const int model = P.group_model[grp];
      if (pass == 1) {
        double r0, r1, rho[3];
        if (!reproject_any<true>(model, ext + (size_t)cam * 6, rec.data() + (size_t)cam * kCamRec, intr + (size_t)grp * 10, X[0], X[1], X[2], X[3], x, y, r0, r1)) { acc[NHD + ND + 1] += 1.0; continue; }
        loss_evaluate(P.loss_type, P.loss_width, r0 * r0 + r1 * r1, rho);
        acc[NHD + ND] += 0.5 * rho[0];
        continue;
      }
      double r[2], rho0, Jb[2][10];
      const uint32_t bits = KIND == kBlockCamera ? (uint32_t)P.ext_const[cam] : P.group_mask[grp];
      if (!block_obs_linearize<KIND, true>(model, ext + (size_t)cam * 6, rec.data() + (size_t)cam * kCamRec, intr + (size_t)grp * 10, X, x, y, P.loss_type,
                                           P.loss_width, bits, r, rho0, Jb)) { acc[NHD + ND + 1] += 1.0; continue; }
      int n = 0;
      for (int a = 0; a < ND; ++a) for (int b = a; b < ND; ++b) acc[n++] += Jb[0][a] * Jb[0][b] + Jb[1][a] * Jb[1][b];
      for (int a = 0; a < ND; ++a) acc[NHD + a] += Jb[0][a] * r[0] + Jb[1][a] * r[1];
      acc[NHD + ND] += 0.5 * rho0;
    }
    return 0;
  };
  block_lm_run_lockstep(B, dims, ND, NA, lo, exec);
  for (int b = 0; b < nb; ++b) for (int j = 0; j < dims[b]; ++j) vals0[(size_t)b * ND + j] = B[b].x[j];
}
}  // namespace

extern "C" void host_inner_iterations(int n_cam, int n_group, int n_pk, long long n_slots, double* ext, double* intr, double* pt_packed,
                                      const unsigned char* ext_const, const int* cam_group, const int* group_model, const unsigned int* group_mask,
                                      const double* mask, const double* xy, const int* slot_cam, const int* slot_pt, const long long* pt_slot,
                                      const int* pt_len, const unsigned char* pt_const, int loss_type, double loss_width) {
  Packed P{n_cam, n_group, n_pk, n_slots, ext, intr, pt_packed, ext_const, cam_group, group_model, group_mask, mask, xy, slot_cam, slot_pt, pt_slot,
           pt_len, pt_const, loss_type, loss_width};
  run_stage<kBlockCamera>(P);
  run_stage<kBlockGroup>(P);
  std::vector<double> rec((size_t)n_cam * kCamRec);
  for (int c = 0; c < n_cam; ++c) cam_prep(ext + (size_t)c * 6 + 3, rec.data() + (size_t)c * kCamRec);
  FilterView V;
  V.ext = ext; V.cam_rec = rec.data(); V.intr = intr; V.pt = pt_packed; V.xy = xy; V.slot_cam = slot_cam; V.cam_group = cam_group; V.group_model = group_model;
  const BlockLmOptions lo;
  PointLmOptions po;
  po.loss_type = loss_type; po.loss_width = loss_width; po.max_num_iterations = lo.max_num_iterations;
  po.function_tolerance = lo.function_tolerance; po.gradient_tolerance = lo.gradient_tolerance; po.parameter_tolerance = lo.parameter_tolerance;
  po.initial_radius = lo.initial_radius; po.max_radius = lo.max_radius; po.min_radius = lo.min_radius; po.min_relative_decrease = lo.min_relative_decrease;
  po.min_diag = lo.min_diag; po.max_diag = lo.max_diag; po.jacobi_scaling = 1; po.max_consecutive_invalid = lo.max_consecutive_invalid;
  for (int k = 0; k < n_pk; ++k) {
    if (pt_const[k]) continue;
    point_lm<true>(V, pt_slot[k], pt_len[k], pt_packed + (size_t)k * 4, po);
  }
}
