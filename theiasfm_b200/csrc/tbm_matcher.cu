// tbm_matcher.cu -- secondary path: brute-force descriptor matching on the GPU, C-ABI of
// include/theia_matcher_b200.h.  Round-1 kernel: CUDA cores, exact float arithmetic in the reference's order (so the
// match sets are bit-identical to the CPU restatement); NOT yet the tcgen05 distance GEMM.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstring>
#include <vector>

#include "../../include/theia_matcher_b200.h"
#include "tbm_top2.h"
#ifndef TBA_EMULATE
#include "tbm_matcher_tc.cuh"   // tcgen05 / TMA path (sm_100a); the SIMT emulation build keeps the exact CUDA-core kernel only
#endif
#include <cstdlib>

static double g_last_timing[4] = {0, 0, 0, 0};  // ms: candidate GEMM kernel, exact re-evaluation kernel, H2D of the descriptors; [3] = queries scanned exhaustively

namespace {

constexpr int ROWS = 32;    // query descriptors per CTA (= lanes of a warp)
constexpr int SLICES = 8;   // warps per CTA; warp w scans candidate rows w, w+8, ... of every tile
constexpr int TJ = 32;      // candidate descriptors per shared-memory tile

using tbm::Top2;
using tbm::top2_push;

// For every row i of A: the nearest (squared L2, ties -> lower index) and second-nearest distance among the rows of B.
// Distances are accumulated left to right in float WITHOUT fused multiply-add: s = s + (a-b)*(a-b), exactly
// L2::operator() evaluated term by term (distance.h:52-56).
__global__ void __launch_bounds__(ROWS* SLICES) k_nn2(const float* __restrict__ A, int nA, const float* __restrict__ B, int nB, int dim,
                                                      int* __restrict__ best_j, float* __restrict__ best_d, float* __restrict__ second_d) {
#ifdef TBA_EMULATE  // CPU emulation build (tests/emu)
  float* smem = emu::dyn_smem<float>();
#else
  extern __shared__ float smem[];
#endif
  const int dimp = dim + 1;                 // padded row stride of the query tile: conflict-free column access
  float* sA = smem;                         // [ROWS][dimp]
  float* sB = smem + ROWS * dimp;           // [TJ][dim]
  __shared__ Top2 s_merge[SLICES][ROWS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i0 = blockIdx.x * ROWS;
  for (int idx = threadIdx.x; idx < ROWS * dim; idx += ROWS * SLICES) {
    const int r = idx / dim, k = idx - r * dim;
    sA[r * dimp + k] = (i0 + r < nA) ? A[(size_t)(i0 + r) * dim + k] : 0.0f;
  }
  Top2 t;
  tbm::top2_init(t);
  for (int j0 = 0; j0 < nB; j0 += TJ) {
    __syncthreads();  // previous tile consumed (and sA written, first time)
    for (int idx = threadIdx.x; idx < TJ * dim; idx += ROWS * SLICES) {
      const int r = idx / dim, k = idx - r * dim;
      sB[idx] = (j0 + r < nB) ? B[(size_t)(j0 + r) * dim + k] : 0.0f;
    }
    __syncthreads();
    for (int jj = warp; jj < TJ; jj += SLICES) {
      const int j = j0 + jj;
      if (j >= nB) break;
      const float* a = sA + lane * dimp;
      const float* b = sB + jj * dim;
      float s = 0.0f;
      for (int k = 0; k < dim; ++k) {
        const float d = __fsub_rn(a[k], b[k]);
        s = __fadd_rn(s, __fmul_rn(d, d));
      }
      top2_push(t, s, j);
    }
  }
  s_merge[warp][lane] = t;
  __syncthreads();
  if (warp == 0 && i0 + lane < nA) {
    // merge the 8 scanners (tbm_top2.h)
    Top2 m = s_merge[0][lane];
    for (int w = 1; w < SLICES; ++w) tbm::top2_merge(m, s_merge[w][lane]);
    best_j[i0 + lane] = m.bj;
    best_d[i0 + lane] = m.bd;
    second_d[i0 + lane] = m.has2 ? m.sd : 0.0f;
  }
}

struct DevF { float* p = nullptr; size_t n = 0; ~DevF() { if (p) cudaFree(p); } bool alloc(size_t c) { if (c <= n && p) return true; if (p) cudaFree(p); p = nullptr; n = 0; if (cudaMalloc(&p, (c ? c : 1) * sizeof(float)) != cudaSuccess) return false; n = c; return true; } };
struct DevI { int* p = nullptr; size_t n = 0; ~DevI() { if (p) cudaFree(p); } bool alloc(size_t c) { if (c <= n && p) return true; if (p) cudaFree(p); p = nullptr; n = 0; if (cudaMalloc(&p, (c ? c : 1) * sizeof(int)) != cudaSuccess) return false; n = c; return true; } };

// :58-59, :78-81: keep the best match when the ratio test is off, there is no second candidate, or it passes
inline bool passes(const tbm_options* o, float best, float second, int second_valid) {
  if (!o->use_lowes_ratio || !second_valid) return true;
  const float sqf = o->lowes_ratio * o->lowes_ratio;  // FeatureMatcherOptions::lowes_ratio is a float: the product is rounded to float, then widened (:58-59)
  const double sq = (double)sqf;
  return (double)best < sq * (double)second;
}

}  // namespace

extern "C" {

void tbm_options_init(tbm_options* o) { o->keep_only_symmetric_matches = 1; o->use_lowes_ratio = 1; o->lowes_ratio = 0.8f; o->min_num_feature_matches = 30; }

void tbm_debug_last_timing(double* out4) { for (int i = 0; i < 4; ++i) out4[i] = g_last_timing[i]; }

int tbm_debug_postprocess(const int32_t* f_best_j, const float* f_best_d, const float* f_second_d, int32_t n1, int f_second_valid,
                          const int32_t* r_best_j, const float* r_best_d, const float* r_second_d, int32_t n2, int r_second_valid,
                          const tbm_options* o, tbm_match* matches, int32_t* n_matches) {
  int n = 0;
  for (int i = 0; i < n1; ++i) {  // forward matches (:63-82)
    if (f_best_j[i] < 0) continue;
    if (passes(o, f_best_d[i], f_second_d[i], f_second_valid)) { matches[n].feature1_ind = i; matches[n].feature2_ind = f_best_j[i]; matches[n].distance = f_best_d[i]; ++n; }
  }
  *n_matches = n;
  if (n < o->min_num_feature_matches) return 0;  // :84-86
  if (o->keep_only_symmetric_matches) {          // :89-113 + IntersectMatches
    int kept = 0;
    for (int k = 0; k < n; ++k) {
      const int i = matches[k].feature1_ind, j = matches[k].feature2_ind;
      const bool rev = j >= 0 && j < n2 && r_best_j[j] == i && passes(o, r_best_d[j], r_second_d[j], r_second_valid);
      if (rev) matches[kept++] = matches[k];
    }
    n = kept;
    *n_matches = n;
  }
  return n >= o->min_num_feature_matches;  // :116
}

#ifndef TBA_EMULATE
// Tensor-core path (dim == 128): all pairs of a chunk in ONE launch of k_nn_candidates + ONE launch of k_exact_top2, one
// device-to-host copy per chunk, then MatchImagePair's ratio test / early exits / IntersectMatches per pair on the host.
static int match_all_tc(const float* descriptors, const int64_t* img_off, int32_t n_img, const int32_t* pairs, int64_t n_pairs,
                        const tbm_options* options, tbm_match* matches, int64_t cap, int64_t* match_off, uint8_t* pair_ok) {
  using namespace tbm_tc;
  const int64_t total = img_off[n_img];
  if (total >= (int64_t)1 << 31) return -1;
  cudaEvent_t ev[6];
  for (auto& e : ev) if (cudaEventCreate(&e) != cudaSuccess) return -3;
  struct EvGuard { cudaEvent_t* e; ~EvGuard() { for (int i = 0; i < 6; ++i) cudaEventDestroy(e[i]); } } ev_guard{ev};
  g_last_timing[0] = g_last_timing[1] = g_last_timing[2] = g_last_timing[3] = 0.0;
  cudaEventRecord(ev[0]);
  int n_sm = 148;
  { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev); }
  DevF d_desc, d_nrm, d_bd, d_sd;
  DevI d_cand, d_bj, d_qrow, d_brow0, d_brows;
  if (!d_desc.alloc((size_t)(total > 0 ? total : 1) * DIM) || !d_nrm.alloc((size_t)(total > 0 ? total : 1))) return -3;
  if (total > 0 && cudaMemcpy(d_desc.p, descriptors, (size_t)total * DIM * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) return -3;
  cudaEventRecord(ev[1]);
  int any_negative = 0;
  if (total > 0) {
    DevI d_neg;
    if (!d_neg.alloc(1) || cudaMemset(d_neg.p, 0, sizeof(int)) != cudaSuccess) return -3;
    k_row_norms<<<(unsigned)((total + 255) / 256), 256>>>(d_desc.p, total, d_nrm.p, d_neg.p);
    if (cudaPeekAtLastError() != cudaSuccess) return -3;
    if (cudaMemcpy(&any_negative, d_neg.p, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return -3;
  }
  CUtensorMap map;
  if (total > 0 && !make_desc_map(&map, d_desc.p, total)) return -3;
  unsigned long long* d_nex = nullptr;  // queries handed to the exhaustive exact scan (diagnostics: tbm_debug_last_timing)
  if (cudaMalloc(&d_nex, 8) != cudaSuccess || cudaMemset(d_nex, 0, 8) != cudaSuccess) return -3;
  struct NexGuard { unsigned long long* p; ~NexGuard() { cudaFree(p); } } nex_guard{d_nex};
  if (cudaFuncSetAttribute(k_nn_candidates<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes) != cudaSuccess ||
      cudaFuncSetAttribute(k_nn_candidates<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes) != cudaSuccess) return -3;
  struct DevItems { WorkItem* p = nullptr; size_t n = 0; ~DevItems() { if (p) cudaFree(p); } } d_items;
  std::vector<WorkItem> items;
  std::vector<int> h_qrow, h_brow0, h_brows, h_bj;
  std::vector<float> h_bd, h_sd;
  std::vector<tbm_match> tmp;
  const bool sym = options->keep_only_symmetric_matches != 0;
  int64_t written = 0;
  bool overflow = false;
  const int64_t kChunkQueries = (int64_t)4 << 20;  // queries per chunk (both directions): bounds the device / host staging
  for (int64_t p0 = 0; p0 < n_pairs;) {
    // ---- chunk [p0, p1): as many pairs as fit the query budget
    int64_t p1 = p0, nq_chunk = 0;
    while (p1 < n_pairs) {
      const int a = pairs[2 * p1], b = pairs[2 * p1 + 1];
      if (a < 0 || a >= n_img || b < 0 || b >= n_img) return -1;
      const int64_t n1 = img_off[a + 1] - img_off[a], n2 = img_off[b + 1] - img_off[b];
      const int64_t add = n1 + (sym ? n2 : 0);
      if (p1 > p0 && nq_chunk + add > kChunkQueries) break;
      nq_chunk += add; ++p1;
    }
    items.clear(); h_qrow.resize((size_t)nq_chunk); h_brow0.resize((size_t)nq_chunk); h_brows.resize((size_t)nq_chunk);
    std::vector<int64_t> q_off((size_t)(p1 - p0) * 2 + 1, 0);
    int64_t qo = 0;
    for (int64_t p = p0; p < p1; ++p) {
      const int a = pairs[2 * p], b = pairs[2 * p + 1];
      for (int dir = 0; dir < 2; ++dir) {
        q_off[(size_t)(p - p0) * 2 + dir] = qo;
        if (dir == 1 && !sym) continue;
        const int qa = dir == 0 ? a : b, cb = dir == 0 ? b : a;
        const int nq = (int)(img_off[qa + 1] - img_off[qa]), nc = (int)(img_off[cb + 1] - img_off[cb]);
        for (int i = 0; i < nq; ++i) { h_qrow[(size_t)qo + i] = (int)img_off[qa] + i; h_brow0[(size_t)qo + i] = (int)img_off[cb]; h_brows[(size_t)qo + i] = nc; }
        if (nc > 0)
          for (int m0 = 0; m0 < nq; m0 += BM) {
            WorkItem w;
            w.a_row0 = (int)img_off[qa] + m0; w.a_rows = nq - m0 < BM ? nq - m0 : BM; w.b_row0 = (int)img_off[cb]; w.b_rows = nc; w.out_row0 = qo + m0;
            items.push_back(w);
          }
        qo += nq;
      }
    }
    q_off.back() = qo;
    if (nq_chunk > 0) {
      if (!d_cand.alloc((size_t)nq_chunk * KC) || !d_bj.alloc((size_t)nq_chunk) || !d_bd.alloc((size_t)nq_chunk) || !d_sd.alloc((size_t)nq_chunk) ||
          !d_qrow.alloc((size_t)nq_chunk) || !d_brow0.alloc((size_t)nq_chunk) || !d_brows.alloc((size_t)nq_chunk)) return -3;
      if (cudaMemset(d_cand.p, 0xFF, (size_t)nq_chunk * KC * sizeof(int)) != cudaSuccess) return -3;  // -1: no candidate (empty other image)
      if (cudaMemcpy(d_qrow.p, h_qrow.data(), (size_t)nq_chunk * sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess ||
          cudaMemcpy(d_brow0.p, h_brow0.data(), (size_t)nq_chunk * sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess ||
          cudaMemcpy(d_brows.p, h_brows.data(), (size_t)nq_chunk * sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess) return -3;
      if (!items.empty()) {
        if (d_items.n < items.size()) { if (d_items.p) cudaFree(d_items.p); d_items.p = nullptr; d_items.n = 0;
          if (cudaMalloc(&d_items.p, items.size() * sizeof(WorkItem)) != cudaSuccess) return -3; d_items.n = items.size(); }
        if (cudaMemcpy(d_items.p, items.data(), items.size() * sizeof(WorkItem), cudaMemcpyHostToDevice) != cudaSuccess) return -3;
        const int grid = (int)(items.size() < (size_t)n_sm ? items.size() : (size_t)n_sm);
        cudaEventRecord(ev[2]);
        if (any_negative) k_nn_candidates<false><<<grid, THREADS, kSmemBytes>>>(map, d_items.p, (int)items.size(), d_nrm.p, d_cand.p);
        else k_nn_candidates<true><<<grid, THREADS, kSmemBytes>>>(map, d_items.p, (int)items.size(), d_nrm.p, d_cand.p);
        if (cudaPeekAtLastError() != cudaSuccess) return -3;
        cudaEventRecord(ev[3]);
      }
      cudaEventRecord(ev[4]);
      k_exact_top2<<<(unsigned)((nq_chunk + 31) / 32), 256>>>(d_desc.p, d_qrow.p, d_brow0.p, d_brows.p, d_cand.p, nq_chunk, d_bj.p, d_bd.p, d_sd.p, d_nex);
      if (cudaPeekAtLastError() != cudaSuccess) return -3;
      cudaEventRecord(ev[5]);
      h_bj.resize((size_t)nq_chunk); h_bd.resize((size_t)nq_chunk); h_sd.resize((size_t)nq_chunk);
      if (cudaMemcpy(h_bj.data(), d_bj.p, (size_t)nq_chunk * sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess ||
          cudaMemcpy(h_bd.data(), d_bd.p, (size_t)nq_chunk * sizeof(float), cudaMemcpyDeviceToHost) != cudaSuccess ||
          cudaMemcpy(h_sd.data(), d_sd.p, (size_t)nq_chunk * sizeof(float), cudaMemcpyDeviceToHost) != cudaSuccess) return -3;
      float ms = 0;
      if (!items.empty() && cudaEventElapsedTime(&ms, ev[2], ev[3]) == cudaSuccess) g_last_timing[0] += ms;
      if (cudaEventElapsedTime(&ms, ev[4], ev[5]) == cudaSuccess) g_last_timing[1] += ms;
    }
    // ---- MatchImagePair's decisions, per pair
    static const int kNone = -1; static const float kZero = 0.0f;
    for (int64_t p = p0; p < p1; ++p) {
      const int a = pairs[2 * p], b = pairs[2 * p + 1];
      const int n1 = (int)(img_off[a + 1] - img_off[a]), n2 = (int)(img_off[b + 1] - img_off[b]);
      const int64_t f0 = q_off[(size_t)(p - p0) * 2], r0 = q_off[(size_t)(p - p0) * 2 + 1];
      tmp.resize((size_t)(n1 > 0 ? n1 : 1));
      int32_t nm = 0;
      const int* fbj = n1 > 0 ? h_bj.data() + f0 : &kNone; const float* fbd = n1 > 0 ? h_bd.data() + f0 : &kZero; const float* fsd = n1 > 0 ? h_sd.data() + f0 : &kZero;
      const bool have_r = sym && n2 > 0;
      const int* rbj = have_r ? h_bj.data() + r0 : &kNone; const float* rbd = have_r ? h_bd.data() + r0 : &kZero; const float* rsd = have_r ? h_sd.data() + r0 : &kZero;
      pair_ok[p] = (uint8_t)tbm_debug_postprocess(fbj, fbd, fsd, n1, n2 >= 2, rbj, rbd, rsd, have_r ? n2 : 0, n1 >= 2, options, tmp.data(), &nm);
      match_off[p] = written;
      if (written + nm > cap || !matches) overflow = true;
      else memcpy(matches + written, tmp.data(), (size_t)nm * sizeof(tbm_match));
      written += nm;
    }
    p0 = p1;
  }
  if (cudaDeviceSynchronize() != cudaSuccess) return -3;
  { float ms = 0; if (cudaEventElapsedTime(&ms, ev[0], ev[1]) == cudaSuccess) g_last_timing[2] = ms; }
  { unsigned long long h = 0; if (cudaMemcpy(&h, d_nex, 8, cudaMemcpyDeviceToHost) == cudaSuccess) g_last_timing[3] = (double)h; }
  match_off[n_pairs] = written;
  return overflow ? -1 : 0;
}
#endif

int tbm_match_all(int device, const float* descriptors, const int64_t* img_off, int32_t n_img, int32_t dim, const int32_t* pairs,
                  int64_t n_pairs, const tbm_options* options, tbm_match* matches, int64_t cap, int64_t* match_off, uint8_t* pair_ok) {
  if (!descriptors || !img_off || !pairs || !options || !match_off || !pair_ok || n_img < 0 || dim <= 0 || dim > 512 || n_pairs < 0) return -1;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) { cudaGetLastError(); return -5; }
  if (cudaSetDevice(device) != cudaSuccess) return -3;
#ifndef TBA_EMULATE
  {
    // dim 128 (SIFT): the tensor-core path.  TBM_PATH=exact forces the CUDA-core kernel (the bit-exact checker of round 1).
    const char* e = getenv("TBM_PATH");
    for (int i = 0; i < n_img; ++i) if (img_off[i + 1] < img_off[i]) return -1;
    if (dim == tbm_tc::DIM && !(e != nullptr && e[0] == 'e')) return match_all_tc(descriptors, img_off, n_img, pairs, n_pairs, options, matches, cap, match_off, pair_ok);
  }
#endif
  const int64_t total = img_off[n_img];
  int64_t max_n = 0;
  for (int i = 0; i < n_img; ++i) { if (img_off[i + 1] < img_off[i]) return -1; max_n = img_off[i + 1] - img_off[i] > max_n ? img_off[i + 1] - img_off[i] : max_n; }
  DevF d_desc, d_bd[2], d_sd[2];
  DevI d_bj[2];
  if (!d_desc.alloc((size_t)total * dim)) return -3;
  for (int s = 0; s < 2; ++s) if (!d_bd[s].alloc((size_t)max_n) || !d_sd[s].alloc((size_t)max_n) || !d_bj[s].alloc((size_t)max_n)) return -3;
  if (cudaMemcpy(d_desc.p, descriptors, (size_t)total * dim * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) return -3;
  const size_t smem = ((size_t)ROWS * (dim + 1) + (size_t)TJ * dim) * sizeof(float);
  if (smem > 48 * 1024 && cudaFuncSetAttribute(k_nn2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -3;
  std::vector<int32_t> bj[2];
  std::vector<float> bd[2], sd[2];
  std::vector<tbm_match> tmp;
  int64_t written = 0;
  bool overflow = false;
  for (int64_t p = 0; p < n_pairs; ++p) {
    const int a = pairs[2 * p], b = pairs[2 * p + 1];
    if (a < 0 || a >= n_img || b < 0 || b >= n_img) return -1;
    const int n1 = (int)(img_off[a + 1] - img_off[a]), n2 = (int)(img_off[b + 1] - img_off[b]);
    const float* A = d_desc.p + (size_t)img_off[a] * dim;
    const float* B = d_desc.p + (size_t)img_off[b] * dim;
    for (int dir = 0; dir < 2; ++dir) {
      const int nq = dir == 0 ? n1 : n2, nc = dir == 0 ? n2 : n1;
      bj[dir].assign((size_t)nq, -1); bd[dir].assign((size_t)nq, 0.0f); sd[dir].assign((size_t)nq, 0.0f);
      if (nq == 0 || nc == 0) continue;
      if (dir == 1 && !options->keep_only_symmetric_matches) continue;
#ifdef TBA_EMULATE
      emu::launch((const void*)k_nn2, (unsigned)((nq + ROWS - 1) / ROWS), (unsigned)(ROWS * SLICES), smem,
                  [&] { k_nn2(dir == 0 ? A : B, nq, dir == 0 ? B : A, nc, dim, d_bj[dir].p, d_bd[dir].p, d_sd[dir].p); });
#else
      k_nn2<<<(nq + ROWS - 1) / ROWS, ROWS * SLICES, smem>>>(dir == 0 ? A : B, nq, dir == 0 ? B : A, nc, dim, d_bj[dir].p, d_bd[dir].p, d_sd[dir].p);
#endif
      if (cudaPeekAtLastError() != cudaSuccess) return -3;
      if (cudaMemcpy(bj[dir].data(), d_bj[dir].p, (size_t)nq * sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess ||
          cudaMemcpy(bd[dir].data(), d_bd[dir].p, (size_t)nq * sizeof(float), cudaMemcpyDeviceToHost) != cudaSuccess ||
          cudaMemcpy(sd[dir].data(), d_sd[dir].p, (size_t)nq * sizeof(float), cudaMemcpyDeviceToHost) != cudaSuccess) return -3;
    }
    tmp.resize((size_t)(n1 > 0 ? n1 : 1));
    int32_t nm = 0;
    pair_ok[p] = (uint8_t)tbm_debug_postprocess(bj[0].data(), bd[0].data(), sd[0].data(), n1, n2 >= 2, bj[1].data(), bd[1].data(), sd[1].data(), n2,
                                                n1 >= 2, options, tmp.data(), &nm);
    match_off[p] = written;
    if (written + nm > cap || !matches) overflow = true;
    else memcpy(matches + written, tmp.data(), (size_t)nm * sizeof(tbm_match));
    written += nm;
  }
  match_off[n_pairs] = written;
  return overflow ? -1 : 0;
}

}  // extern "C"
