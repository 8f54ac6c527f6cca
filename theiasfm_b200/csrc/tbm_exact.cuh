// tbm_exact.cuh -- pass 2 of the tensor-core matcher (tbm_matcher_tc.cuh): the exact float re-evaluation of the candidates the TF32
// pass selected, in the reference's summation order (distance.h:52-56), and the exhaustive scan of the queries whose candidate list
// overflowed.  Plain CUDA (no tcgen05 / TMA): also compiled by the SIMT emulation build, where tbm_debug_exact_top2 drives it with
// hand-made candidate lists (tests/test_matcher_host.py).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace tbm_tc {

constexpr int DIM = 128;       // descriptor length (SIFT); 4 swizzle atoms of 32 floats
constexpr int KC = 16;         // candidate slots per query handed to the exact pass (-1 = empty; typically 2..4 are filled)
constexpr int kOverflow = -2;  // cand[q*KC] marker: the exact pass scans every candidate of this query

// ------------------------------------------------------------------ pass 2: exact top-2 among the candidates
// One thread per (query, candidate): the exact squared distance -- float, term by term, no fused multiply-add: L2::operator()
// (distance.h:52-56) as the oracle and k_nn2 evaluate it; then thread 0 of the query picks the best two (ties: lower index).
// q_row[i] = global descriptor row of query i, b_row0[i] / b_rows[i] = first row / number of rows of the candidate image (indices are
// reported relative to b_row0).  A query whose candidate ring overflowed in pass 1 (cand[i*KC] == kOverflow) is scanned exhaustively.
// exact squared distance, float, term by term in index order without fused multiply-add (128-bit loads, same arithmetic)
__device__ __forceinline__ float exact_sqdist(const float* __restrict__ a, const float* __restrict__ b) {
  const float4* a4 = reinterpret_cast<const float4*>(a);
  const float4* b4 = reinterpret_cast<const float4*>(b);
  float s = 0.0f;
#pragma unroll 8
  for (int k = 0; k < DIM / 4; ++k) {
    const float4 x = __ldg(a4 + k), y = __ldg(b4 + k);
    float df = __fsub_rn(x.x, y.x); s = __fadd_rn(s, __fmul_rn(df, df));
    df = __fsub_rn(x.y, y.y); s = __fadd_rn(s, __fmul_rn(df, df));
    df = __fsub_rn(x.z, y.z); s = __fadd_rn(s, __fmul_rn(df, df));
    df = __fsub_rn(x.w, y.w); s = __fadd_rn(s, __fmul_rn(df, df));
  }
  return s;
}
// lexicographic (distance, index) order: what MatchImagePair's partial_sort over index-ordered candidates yields
__device__ __forceinline__ void top2_take(int& bj, float& bd, int& sj, float& sd, int jj, float dd) {
  if (jj < 0) return;
  if (bj < 0 || dd < bd || (dd == bd && jj < bj)) { sj = bj; sd = bd; bj = jj; bd = dd; }
  else if (sj < 0 || dd < sd || (dd == sd && jj < sj)) { sj = jj; sd = dd; }
}
// smem_sqdist: the same sum from shared-memory rows (the term order and the roundings of exact_sqdist)
__device__ __forceinline__ float smem_sqdist(const float* a, const float* b) {
  const float4* a4 = reinterpret_cast<const float4*>(a);
  const float4* b4 = reinterpret_cast<const float4*>(b);
  float s = 0.0f;
#pragma unroll 8
  for (int k = 0; k < DIM / 4; ++k) {
    const float4 x = a4[k], y = b4[k];
    float df = __fsub_rn(x.x, y.x); s = __fadd_rn(s, __fmul_rn(df, df));
    df = __fsub_rn(x.y, y.y); s = __fadd_rn(s, __fmul_rn(df, df));
    df = __fsub_rn(x.z, y.z); s = __fadd_rn(s, __fmul_rn(df, df));
    df = __fsub_rn(x.w, y.w); s = __fadd_rn(s, __fmul_rn(df, df));
  }
  return s;
}

// One CTA = 32 queries.  Every descriptor row the CTA touches goes through SHARED MEMORY, fetched with coalesced 128-bit loads (32
// consecutive threads = one 512-byte row) -- per-lane row loads (32 different rows per load instruction) made the first version of this
// kernel LSU-bound (ncu: 93 % LSU wavefronts) and the exhaustive scans 85 % of its time:
//   * the 32 query rows are staged once;
//   * candidate slots are processed two per query at a time (64 rows, row stride 132 floats: conflict-free 128-bit reads by 64 threads,
//     one row each, in the reference's term order); slot pairs that are empty for the whole CTA are skipped;
//   * a query whose list overflowed in pass 1 (cand[q*KC] or cand[q*KC + KC/2] == kOverflow) is scanned exhaustively by the WHOLE CTA
//     afterwards, 64 candidate rows per staged tile.
// Dynamic shared memory: (XT + 32) * XS floats = kExactSmemBytes.
constexpr int XT = 64;             // candidate rows per staged tile
constexpr int XS = DIM + 4;        // padded row stride (floats)
constexpr int kExactSmemBytes = (XT + 32) * XS * (int)sizeof(float);
__global__ void __launch_bounds__(256) k_exact_top2(const float* __restrict__ d, const int* __restrict__ q_row, const int* __restrict__ b_row0,
                                                    const int* __restrict__ b_rows, const int* __restrict__ cand, long long n_q,
                                                    int* __restrict__ best_j, float* __restrict__ best_d, float* __restrict__ second_d,
                                                    unsigned long long* __restrict__ n_exhaustive) {
#ifdef TBA_EMULATE
  float* smem = emu::dyn_smem<float>();
#else
  extern __shared__ __align__(16) float smem[];
#endif
  float* s_rows = smem;             // [XT][XS] candidate rows of the current round / tile
  float* s_q = smem + XT * XS;      // [32][XS] the CTA's query rows
  __shared__ float m_d[XT], m_d2[XT];
  __shared__ int m_j[XT], m_j2[XT];
  __shared__ int s_rowj[XT];        // global descriptor row staged in s_rows[r], or -1
  __shared__ int s_ovf[32];
  __shared__ int s_novf;
  __shared__ int s_any[2];        // "this round has rows", double-buffered: a round without rows has no barrier behind the read
  const int tid = threadIdx.x;
  const long long q0 = (long long)blockIdx.x * 32;
  const int nq_cta = (int)((n_q - q0) < 32 ? (n_q - q0) : 32);
  if (tid == 0) s_novf = 0;
  // ---- the query rows
  for (int e = tid; e < nq_cta * (DIM / 4); e += 256) {
    const int r = e / (DIM / 4), k4 = e % (DIM / 4);
    reinterpret_cast<float4*>(s_q + r * XS)[k4] = __ldg(reinterpret_cast<const float4*>(d + (size_t)q_row[q0 + r] * DIM) + k4);
  }
  __syncthreads();
  if (tid < nq_cta) {
    const long long qi = q0 + tid;
    if (cand[qi * KC] == kOverflow || cand[qi * KC + KC / 2] == kOverflow) {
      s_ovf[atomicAdd(&s_novf, 1)] = tid;
      if (n_exhaustive) atomicAdd(n_exhaustive, 1ull);
    }
  }
  __syncthreads();
  // ---- listed candidates: thread r < 64 owns (query r / 2, slots 2 g + (r & 1)) of every round g and keeps its own top-2
  int bj = -1, sj = -1; float bd = 0.0f, sd = 0.0f;
  int my_base = 0;
  bool my_listed = false;
  if (tid < XT && (tid >> 1) < nq_cta) {
    const long long qi = q0 + (tid >> 1);
    my_base = b_row0[qi];
    my_listed = !(cand[qi * KC] == kOverflow || cand[qi * KC + KC / 2] == kOverflow);
  }
  for (int g = 0; g < KC / 2; ++g) {
    if (tid == 0) s_any[g & 1] = 0;
    __syncthreads();  // the rows of the previous round are consumed; the flag of this round reset
    if (tid < XT) {
      int j = -1;
      if (my_listed) j = cand[(q0 + (tid >> 1)) * KC + 2 * g + (tid & 1)];
      s_rowj[tid] = j;
      if (j >= 0) s_any[g & 1] = 1;
    }
    __syncthreads();
    if (!s_any[g & 1]) continue;  // (uniform: read after the barrier; the next write to this flag is two barriers away)
    for (int e = tid; e < XT * (DIM / 4); e += 256) {
      const int r = e / (DIM / 4), k4 = e % (DIM / 4);
      const int j = s_rowj[r];
      if (j >= 0) reinterpret_cast<float4*>(s_rows + r * XS)[k4] = __ldg(reinterpret_cast<const float4*>(d + (size_t)j * DIM) + k4);
    }
    __syncthreads();
    if (tid < XT && s_rowj[tid] >= 0) top2_take(bj, bd, sj, sd, s_rowj[tid] - my_base, smem_sqdist(s_q + (tid >> 1) * XS, s_rows + tid * XS));
  }
  __syncthreads();
  if (tid < XT) { m_d[tid] = bd; m_j[tid] = bj; m_d2[tid] = sd; m_j2[tid] = sj; }
  __syncthreads();
  if (tid < nq_cta) {
    const long long qi = q0 + tid;
    if (!(cand[qi * KC] == kOverflow || cand[qi * KC + KC / 2] == kOverflow)) {
      int fj = -1, gj = -1; float fd = 0.0f, gd = 0.0f;
      for (int k = 2 * tid; k < 2 * tid + 2; ++k) { top2_take(fj, fd, gj, gd, m_j[k], m_d[k]); top2_take(fj, fd, gj, gd, m_j2[k], m_d2[k]); }
      best_j[qi] = fj; best_d[qi] = fd; second_d[qi] = gj >= 0 ? gd : 0.0f;
    }
  }
  // ---- exhaustive scans, one overflowed query of this CTA after the other, all 256 threads
  const int novf = s_novf;
  for (int o = 0; o < novf; ++o) {
    const int ql = s_ovf[o];
    const long long qx = q0 + ql;
    const int base = b_row0[qx], nb = b_rows[qx];
    int xbj = -1, xsj = -1; float xbd = 0.0f, xsd = 0.0f;
    for (int r0 = 0; r0 < nb; r0 += XT) {
      __syncthreads();  // the previous tile (or round, or merge) is consumed
      const int rows = nb - r0 < XT ? nb - r0 : XT;
      for (int e = tid; e < rows * (DIM / 4); e += 256) {
        const int r = e / (DIM / 4), k4 = e % (DIM / 4);
        reinterpret_cast<float4*>(s_rows + r * XS)[k4] = __ldg(reinterpret_cast<const float4*>(d + (size_t)(base + r0 + r) * DIM) + k4);
      }
      __syncthreads();
      if (tid < rows) top2_take(xbj, xbd, xsj, xsd, r0 + tid, smem_sqdist(s_q + ql * XS, s_rows + tid * XS));
    }
    __syncthreads();
    if (tid < XT) { m_d[tid] = xbd; m_j[tid] = xbj; m_d2[tid] = xsd; m_j2[tid] = xsj; }
    __syncthreads();
    if (tid == 0) {
      int fj = -1, gj = -1; float fd = 0.0f, gd = 0.0f;
      for (int k = 0; k < XT; ++k) { top2_take(fj, fd, gj, gd, m_j[k], m_d[k]); top2_take(fj, fd, gj, gd, m_j2[k], m_d2[k]); }
      best_j[qx] = fj; best_d[qx] = fd; second_d[qx] = gj >= 0 ? gd : 0.0f;
    }
  }
}

}  // namespace tbm_tc
