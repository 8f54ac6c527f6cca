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
constexpr int ET = 8;          // threads per query in the exact pass
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
// Overflowed queries (0.6 % of the bench scene) are scanned exhaustively -- 1000 x the work of a normal query, which made them 85 % of
// this kernel's time when the 8 threads of the query did it alone with per-lane row loads (ncu: barrier stalls + LSU wavefronts).  They are
// now handled by the WHOLE CTA after the normal queries: 64 candidate rows at a time are staged in shared memory with coalesced 128-bit
// loads (row stride 132 floats: conflict-free), 64 threads run one row each in the reference's term order, then a block-wide top-2 merge.
constexpr int XT = 64;             // candidate rows per staged tile of the exhaustive scan
constexpr int XS = DIM + 4;        // padded row stride (floats)
__global__ void __launch_bounds__(256) k_exact_top2(const float* __restrict__ d, const int* __restrict__ q_row, const int* __restrict__ b_row0,
                                                    const int* __restrict__ b_rows, const int* __restrict__ cand, long long n_q,
                                                    int* __restrict__ best_j, float* __restrict__ best_d, float* __restrict__ second_d,
                                                    unsigned long long* __restrict__ n_exhaustive) {
  __shared__ float s_d[32][ET], s_d2[32][ET];
  __shared__ int s_j[32][ET], s_j2[32][ET];
  __shared__ int s_ovf[32];
  __shared__ int s_novf;
  __shared__ __align__(16) float s_rows[XT * XS];
  __shared__ __align__(16) float s_a[DIM];
  const int ql = threadIdx.x / ET, c = threadIdx.x % ET;
  const long long q0 = (long long)blockIdx.x * 32;
  const long long qi = q0 + ql;
  if (threadIdx.x == 0) s_novf = 0;
  __syncthreads();
  int bj = -1, sj = -1; float bd = 0.0f, sd = 0.0f;
  bool overflowed = false;
  if (qi < n_q) {
    const float* a = d + (size_t)q_row[qi] * DIM;
    const int base = b_row0[qi];
    overflowed = cand[qi * KC] == kOverflow || cand[qi * KC + KC / 2] == kOverflow;
    if (overflowed) {
      if (c == 0) { s_ovf[atomicAdd(&s_novf, 1)] = ql; if (n_exhaustive) atomicAdd(n_exhaustive, 1ull); }
    } else {
      for (int k = c; k < KC; k += ET) {
        const int j = cand[qi * KC + k];
        if (j >= 0) top2_take(bj, bd, sj, sd, j - base, exact_sqdist(a, d + (size_t)j * DIM));
      }
    }
  }
  s_d[ql][c] = bd; s_j[ql][c] = bj; s_d2[ql][c] = sd; s_j2[ql][c] = sj;
  __syncthreads();
  if (c == 0 && qi < n_q && !overflowed) {
    int fj = -1, gj = -1; float fd = 0.0f, gd = 0.0f;
    for (int k = 0; k < ET; ++k) { top2_take(fj, fd, gj, gd, s_j[ql][k], s_d[ql][k]); top2_take(fj, fd, gj, gd, s_j2[ql][k], s_d2[ql][k]); }
    best_j[qi] = fj; best_d[qi] = fd; second_d[qi] = gj >= 0 ? gd : 0.0f;
  }
  // ---- exhaustive scans, one overflowed query of this CTA after the other, all 256 threads
  const int novf = s_novf;
  for (int o = 0; o < novf; ++o) {
    __syncthreads();  // s_rows / s_a / s_d of the previous round consumed
    const long long qx = q0 + s_ovf[o];
    const float* a = d + (size_t)q_row[qx] * DIM;
    const int base = b_row0[qx], nb = b_rows[qx];
    if (threadIdx.x < DIM / 4) reinterpret_cast<float4*>(s_a)[threadIdx.x] = __ldg(reinterpret_cast<const float4*>(a) + threadIdx.x);
    int xbj = -1, xsj = -1; float xbd = 0.0f, xsd = 0.0f;
    for (int r0 = 0; r0 < nb; r0 += XT) {
      __syncthreads();  // the previous tile is consumed (and s_a is written, first time)
      const int rows = nb - r0 < XT ? nb - r0 : XT;
      for (int e = threadIdx.x; e < rows * (DIM / 4); e += 256) {  // coalesced: 32 consecutive threads fetch one 512-byte row
        const int r = e / (DIM / 4), k4 = e % (DIM / 4);
        reinterpret_cast<float4*>(s_rows + r * XS)[k4] = __ldg(reinterpret_cast<const float4*>(d + (size_t)(base + r0 + r) * DIM) + k4);
      }
      __syncthreads();
      if ((int)threadIdx.x < rows) {
        const float4* b4 = reinterpret_cast<const float4*>(s_rows + threadIdx.x * XS);
        const float4* a4 = reinterpret_cast<const float4*>(s_a);
        float sacc = 0.0f;
#pragma unroll 8
        for (int k = 0; k < DIM / 4; ++k) {
          const float4 x = a4[k], y = b4[k];
          float df = __fsub_rn(x.x, y.x); sacc = __fadd_rn(sacc, __fmul_rn(df, df));
          df = __fsub_rn(x.y, y.y); sacc = __fadd_rn(sacc, __fmul_rn(df, df));
          df = __fsub_rn(x.z, y.z); sacc = __fadd_rn(sacc, __fmul_rn(df, df));
          df = __fsub_rn(x.w, y.w); sacc = __fadd_rn(sacc, __fmul_rn(df, df));
        }
        top2_take(xbj, xbd, xsj, xsd, r0 + (int)threadIdx.x, sacc);
      }
    }
    // merge the XT scanners (threads 0 .. XT-1) through the first rows of the per-query staging arrays
    __syncthreads();
    float* m_d = &s_d[0][0]; float* m_d2 = &s_d2[0][0]; int* m_j = &s_j[0][0]; int* m_j2 = &s_j2[0][0];  // 256 entries each, XT used
    if (threadIdx.x < XT) { m_d[threadIdx.x] = xbd; m_j[threadIdx.x] = xbj; m_d2[threadIdx.x] = xsd; m_j2[threadIdx.x] = xsj; }
    __syncthreads();
    if (threadIdx.x == 0) {
      int fj = -1, gj = -1; float fd = 0.0f, gd = 0.0f;
      for (int k = 0; k < XT; ++k) { top2_take(fj, fd, gj, gd, m_j[k], m_d[k]); top2_take(fj, fd, gj, gd, m_j2[k], m_d2[k]); }
      best_j[qx] = fj; best_d[qx] = fd; second_d[qx] = gj >= 0 ? gd : 0.0f;
    }
  }
}

}  // namespace tbm_tc
