// tba_inner.cuh -- observation passes of Ceres' inner iterations (N4).  Theia enables them by default
// (bundle_adjustment.h:114; SetBundleAdjustmentOptions, reconstruction_estimator_utils.cc:118) and reverses the ordering
// (bundle_adjuster.cc:196-200) so that coordinate descent visits EXTRINSICS, then INTRINSICS GROUPS, then POINTS.
// Within one of the first two sets every block is an independent small dense problem over the block's own observations;
// the per-block trust-region logic runs on the host in lockstep (tba_block_lm.h), the observation passes here:
//   k_block_normal<KIND>:  acc[block] += { J_b^T J_b (upper triangle), J_b^T r, cost }   at the block's current value
//   k_block_cost<KIND>:    acc[block] += { cost }                                          at the block's trial value
// one thread per observation slot (all slots of the device-resident problem; slots of inactive blocks return at once),
// fp64 RED accumulation per block, 256 replica rows when a single intrinsics group owns every observation.
// The point set is the third stage: point_lm (tba_point_lm.cuh) with the same default options.
// The per-slot bodies are host/device and are run on the CPU by tests/host_inner.cc.
#pragma once
#include <cstdint>

#include "tba_camera_models.cuh"

namespace tba {

constexpr int kBlockCamera = 0;
constexpr int kBlockGroup = 1;
__host__ __device__ constexpr int block_dim(int kind) { return kind == kBlockCamera ? 6 : 10; }
__host__ __device__ constexpr int block_acc(int kind) { return block_dim(kind) * (block_dim(kind) + 1) / 2 + block_dim(kind) + 2; }  // H, g, cost, failures

// Residual (robustified) and the masked Jacobian rows of one observation with respect to ITS camera's extrinsics
// (KIND camera: [J_C = -h J_a | J_w], masked by the ext_const bits) or ITS group's intrinsics (KIND group: 10 columns,
// masked by group_const_mask).  rho0 = loss value (cost contribution 0.5 rho0).
template <int KIND, bool EXT>
__host__ __device__ inline bool block_obs_linearize(int model, const double* __restrict__ C, const double* __restrict__ rec, const double* __restrict__ k,
                                                    const double* __restrict__ X, double x, double y, int loss_type, double loss_width,
                                                    uint32_t const_bits, double r[2], double& rho0, double Jb[2][10]) {
  double Ja[6], Jw[6], Jh[2], Ji[20];
  if (KIND == kBlockCamera) {
    if (!linearize_obs_any<0u, EXT>(model, C, rec, k, X[0], X[1], X[2], X[3], x, y, loss_type, loss_width, r, rho0, Ja, Jw, Jh, nullptr)) return false;
    const bool pos_free = !(const_bits & 1u), rot_free = !(const_bits & 2u);
    for (int row = 0; row < 2; ++row)
      for (int j = 0; j < 3; ++j) {
        Jb[row][j] = pos_free ? -X[3] * Ja[row * 3 + j] : 0.0;
        Jb[row][3 + j] = rot_free ? Jw[row * 3 + j] : 0.0;
      }
  } else {
    if (!linearize_obs_any<0x3FFu, EXT>(model, C, rec, k, X[0], X[1], X[2], X[3], x, y, loss_type, loss_width, r, rho0, Ja, Jw, Jh, Ji)) return false;
    const int K = model_num_parameters(model);
    for (int row = 0; row < 2; ++row)
      for (int j = 0; j < 10; ++j) Jb[row][j] = (j < K && !((const_bits >> j) & 1u)) ? Ji[row * 10 + j] : 0.0;
  }
  return true;
}

#if defined(__CUDACC__) || defined(TBA_EMULATE)
struct BlockPassArgs {
  const double* ext;     // [n_cam][6]   values to evaluate with (block values already substituted by the host driver)
  const double* rec;     // [n_cam][kCamRec] for ext
  const double* intr;    // [n_group][10]
  const double* pt;      // [n_pt][4]
  const uint8_t* active; // [n_block]
  const uint8_t* ext_const;        // [n_cam]  (KIND camera)
  const uint32_t* group_const;     // [n_group] (KIND group)
  double* acc;           // [n_rep][n_block][NA]
  int n_rep;             // replica rows (power of two); > 1 only when one group owns every observation
};

__device__ __forceinline__ void blk_red(double* a, double v) { atomicAdd(a, v); }

template <int KIND, bool EXT, bool COST_ONLY>
__global__ void __launch_bounds__(256) k_block_pass(DevProblem P, BlockPassArgs A) {
  constexpr int N = block_dim(KIND), NH = N * (N + 1) / 2, NA = block_acc(KIND);
  const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int cam = P.slot_cam[s];  // grid = tiles: s < n_slots always
  if (cam < 0) return;
  const int grp = P.cam_group[cam];
  const int blk = KIND == kBlockCamera ? cam : grp;
  if (!A.active[blk]) return;
  const int n_block = KIND == kBlockCamera ? P.n_cam : P.n_group;
  double* acc = A.acc + ((size_t)(blockIdx.x & (A.n_rep - 1)) * n_block + blk) * NA;
  const long long wq = s >> 5;
  const int l = (int)(s & 31);
  const double x = P.xy[(size_t)(wq * 2) * 32 + l], y = P.xy[(size_t)(wq * 2 + 1) * 32 + l];
  const double* X = A.pt + (size_t)P.slot_pt[s] * 4;
  const int model = P.group_model[grp];
  if (COST_ONLY) {
    double r0, r1, rho[3];
    if (!reproject_any<EXT>(model, A.ext + (size_t)cam * 6, A.rec + (size_t)cam * kCamRec, A.intr + (size_t)grp * 10, X[0], X[1], X[2], X[3], x, y, r0, r1)) {
      blk_red(acc + NH + N + 1, 1.0);
      return;
    }
    loss_evaluate(P.loss_type, P.loss_width, r0 * r0 + r1 * r1, rho);
    blk_red(acc + NH + N, 0.5 * rho[0]);
    return;
  }
  double r[2], rho0, Jb[2][10];
  const uint32_t bits = KIND == kBlockCamera ? (uint32_t)A.ext_const[cam] : A.group_const[grp];
  if (!block_obs_linearize<KIND, EXT>(model, A.ext + (size_t)cam * 6, A.rec + (size_t)cam * kCamRec, A.intr + (size_t)grp * 10, X, x, y, P.loss_type,
                                      P.loss_width, bits, r, rho0, Jb)) {
    blk_red(acc + NH + N + 1, 1.0);
    return;
  }
  int n = 0;
#pragma unroll
  for (int a = 0; a < N; ++a) {
#pragma unroll
    for (int b = a; b < N; ++b) {
      const double v = Jb[0][a] * Jb[0][b] + Jb[1][a] * Jb[1][b];
      if (v != 0.0) blk_red(acc + n, v);
      ++n;
    }
  }
#pragma unroll
  for (int a = 0; a < N; ++a) {
    const double v = Jb[0][a] * r[0] + Jb[1][a] * r[1];
    if (v != 0.0) blk_red(acc + NH + a, v);
  }
  blk_red(acc + NH + N, 0.5 * rho0);
}

// |x - x_c|^2 over the non-constant parameter blocks (ambient coordinates), after the inner iterations moved the candidate:
// scal[4] (camera side) and scal[5] (points), the slots stage_evaluate_candidate uses for the step norm.
__global__ void k_xdiff(DevProblem P, const double* __restrict__ blk_free, double* __restrict__ scal, int count_cs) {
  __shared__ double s_red[32];
  double a_cs = 0.0, a_pt = 0.0;
  const size_t stride = (size_t)gridDim.x * blockDim.x, t0 = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  for (size_t i = t0; i < (size_t)P.ne; i += stride) if (blk_free[i / 6] != 0.0) { const double d = P.ext_c[i] - P.ext[i]; a_cs += d * d; }
  for (size_t i = t0; i < (size_t)P.n_group * 10; i += stride) if (blk_free[P.n_cam + i / 10] != 0.0) { const double d = P.intr_c[i] - P.intr[i]; a_cs += d * d; }
  for (size_t i = t0; i < (size_t)P.n_pt * 4; i += stride) if (!P.pt_const[i >> 2]) { const double d = P.pt_c[i] - P.pt[i]; a_pt += d * d; }
  const double s1 = block_sum(a_cs, s_red);
  const double s2 = block_sum(a_pt, s_red);
  if (threadIdx.x == 0) { if (count_cs) red_add(scal + 4, s1); red_add(scal + 5, s2); }
}
#endif  // __CUDACC__

}  // namespace tba
