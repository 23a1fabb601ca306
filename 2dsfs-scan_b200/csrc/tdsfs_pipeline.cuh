// tdsfs_pipeline.cuh -- EXPERIMENTAL (off unless TDSFS_PIPELINE=1; not yet measured on a GPU): the window scorer split so
// that its background-independent part runs UNDER the count kernel ("Next").
//
//   T = 2 ( sum_bins x ln x  -  sum_SNPs ln b[bin_s]  -  N (ln N - ln B) )
//
//   k3a_window_sums   per window: spectra in shared memory (same tables as k3_score_small), W = sum_bins x ln x, N and the
//                     "every SNP in one bin" flags.  Needs the records only: launched per chunk of rows on a second stream
//                     while the count kernel works on the next chunk (one 8-warp CTA fits beside the count kernel's CTA).
//   k3b_gather_finish after the background is final: G = sum_SNPs ln b[bin_s] by a gather over the window's records, then
//                     the statistic.  The two spectra for which the reference returns exactly 0.0 keep their per-bin form:
//                     a one-bin window is x (ln x - ln b) of that bin; a window with N == B goes to k3_score_large's list.
// Plain fixed-bp scans with a genome-wide background and no per-SNP flags only; everything else takes k3_score_small.
#pragma once
#include "tdsfs_kernels.cuh"

namespace tdsfs {

constexpr int PIPE_MAX_CHUNKS = 8;
constexpr uint8_t PIPE_ONE_BIN_2D = 1, PIPE_ONE_BIN_1A = 2, PIPE_ONE_BIN_1B = 4;  // r_flags between k3a and k3b

struct PipeParams {
  ScoreParams s;
  const double* dxI;            // (m+1) ln(m+1) - m ln m
  int row_lo, row_hi;           // k3a: windows with row_lo < whi <= row_hi (their records are complete)
  unsigned long long* work;     // k3a: hand-out counter of this launch (zeroed by the host)
  int* large_cap_guard;         // unused (layout reserve)
};

// dx[m] = (m+1) ln(m+1) - m ln m, written as ln(m+1) + m log1p(1/m) so that no large terms cancel
__global__ void k_dx_table(double* t, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) t[i] = i ? log((double)i + 1.0) + (double)i * log1p(1.0 / (double)i) : 0.0;
}

// first candidate id whose whi exceeds `row` (whi is non-decreasing over candidate ids)
__device__ __forceinline__ long long first_window_past(const int32_t* whi, long long n, int row) {
  long long lo = 0, hi = n;
  while (lo < hi) {
    const long long mid = (lo + hi) >> 1;
    if (__ldg(whi + mid) <= row) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// One warp per window (<= WCAP SNPs).  2D: open-addressing table, the atomic that bumps a bin returns its old count c and the
// window sum grows by dx[c]; 1D: packed 16-bit bins bumped with predicated reductions, then walked for x ln x (no gathers).
__global__ void __launch_bounds__(SCORE_WARPS * 32, 4) k3a_window_sums(const __grid_constant__ PipeParams q) {
  extern __shared__ __align__(16) uint32_t sm32[];
  const ScoreParams& p = q.s;
  constexpr uint32_t F10 = (1u << KEY_SHIFT) - 1;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gwords = score_group_smem_words(p.n1, p.n2);  // even: every warp's tables start 8-byte aligned
  uint32_t* tab = sm32 + (size_t)warp * gwords;
  uint32_t* h1a = tab + HASH_SLOTS;
  const int nw1 = (p.n1 + 2) / 2, nw2 = (p.n2 + 2) / 2;
  uint32_t* h1b = h1a + nw1;
  const int nw12 = (nw1 + nw2 + 1) & ~1;
  auto clear_table = [&]() {
    uint2* t2 = reinterpret_cast<uint2*>(tab);
#pragma unroll
    for (int i = 0; i < HASH_SLOTS / 64; ++i) t2[lane + i * 32] = make_uint2(EMPTY_KEY, EMPTY_KEY);
  };
  clear_table();
  for (int i = lane; i < nw12; i += 32) h1a[i] = 0;
  __syncwarp();
  const uint32_t last = (uint32_t)p.bins2d - 1;
  const long long id0 = first_window_past(p.whi, p.ncand, q.row_lo), id1 = first_window_past(p.whi, p.ncand, q.row_hi);

  auto grab = [&]() -> long long {
    long long v = 0;
    if (lane == 0) v = id0 + (long long)atomicAdd(q.work, 1ull);
    return __shfl_sync(0xffffffffu, v, 0);
  };
  for (long long id = grab(); id < id1; id = grab()) {
    const int lo = __ldg(p.wlo + id), hi = __ldg(p.whi + id);
    const int cnt = hi - lo;
    if (cnt == 0 || cnt > WCAP) continue;  // empty: flagged by K2; large: k3_score_large
    constexpr int Q = 8;
    double w2 = 0.0;
    uint32_t nn = 0, m2 = 0;  // nn = N2 | N1a << 10 | N1b << 20; m2 = largest old count seen (N2 - 1 iff one 2D bin)
    for (int base = 0; base < cnt; base += Q * 32) {
      uint2 r[Q];
#pragma unroll
      for (int j = 0; j < Q; ++j) {
        const int i = base + j * 32 + lane;
        r[j] = i < cnt ? __ldg(p.rec + lo + i) : make_uint2(0u, 0u);  // default policy: read again by k3b, ideally from L2
      }
#pragma unroll
      for (int j = 0; j < Q; ++j) {
        const uint32_t k = r[j].x;
        if (k != 0 && k != last) {
          uint32_t h = (k * 0x9E3779B1u) >> 22;  // HASH_SLOTS = 2^10
          uint32_t c;
          while (true) {
            uint32_t e = tab[h];
            if (e == EMPTY_KEY) {
              e = atomicCAS(tab + h, EMPTY_KEY, (k << KEY_SHIFT) | 1u);
              if (e == EMPTY_KEY) { c = 0; break; }
            }
            if ((e >> KEY_SHIFT) == k) { c = atomicAdd(tab + h, 1u) & F10; break; }
            h = (h + 1) & (HASH_SLOTS - 1);
          }
          if (c) w2 += __ldg(q.dxI + c);
          nn += 1u;
          m2 = max(m2, c);
        }
        const uint32_t fa = r[j].y & 0xFFFF, fb = r[j].y >> 16;
        bump_half_if(h1a, fa);
        bump_half_if(h1b, fb);
        nn += (fa ? 1u << 10 : 0u) + (fb ? 1u << 20 : 0u);
      }
    }
    __syncwarp();  // every insert of the window is done
    clear_table();
    // 1D: walk the packed bins for x ln x and the largest bin, clearing on the way
    double w1a = 0.0, w1b = 0.0;
    uint32_t m1a = 0, m1b = 0;
    auto walk = [&](uint32_t* h1, int nw, double& w, uint32_t& mx) {
      for (int i = lane; i < nw; i += 32) {
        const uint32_t v = h1[i];
        if (v) {
          h1[i] = 0;
          const uint32_t x0 = v & 0xFFFF, x1 = v >> 16;
          if (x0 > 1) w = fma(u32_to_double(x0), __ldg(p.lnI + x0), w);
          if (x1 > 1) w = fma(u32_to_double(x1), __ldg(p.lnI + x1), w);
          mx = max(mx, max(x0, x1));
        }
      }
    };
    walk(h1a, nw1, w1a, m1a);
    walk(h1b, nw2, w1b, m1b);
    nn = __reduce_add_sync(0xffffffffu, nn);
    m2 = __reduce_max_sync(0xffffffffu, m2);
    m1a = __reduce_max_sync(0xffffffffu, m1a);
    m1b = __reduce_max_sync(0xffffffffu, m1b);
    w2 = warp_sum(w2); w1a = warp_sum(w1a); w1b = warp_sum(w1b);
    const uint32_t N2 = nn & F10, N1a = (nn >> 10) & F10, N1b = nn >> 20;
    if (lane == 0) {
      p.r_T2[id] = w2; p.r_T1a[id] = w1a; p.r_T1b[id] = w1b;
      p.r_n2[id] = (int)N2; p.r_n1a[id] = (int)N1a; p.r_n1b[id] = (int)N1b;
      p.r_count[id] = cnt;
      p.r_flags[id] = (uint8_t)((N2 && m2 + 1 == N2 ? PIPE_ONE_BIN_2D : 0) | (N1a && m1a == N1a ? PIPE_ONE_BIN_1A : 0) |
                                (N1b && m1b == N1b ? PIPE_ONE_BIN_1B : 0));
    }
    __syncwarp();  // the tables are clean before the next window's inserts
  }
}

// One warp per small window: G = sum over the window's SNPs of ln b of their bins, then the three statistics.
__global__ void __launch_bounds__(256) k3b_gather_finish(const __grid_constant__ PipeParams q, int32_t* large, int* nlarge) {
  const ScoreParams& p = q.s;
  const int lane = threadIdx.x & 31;
  const long long nwarp = (long long)gridDim.x * (blockDim.x >> 5);
  const uint32_t last = (uint32_t)p.bins2d - 1;
  const double* Bg = p.B;  // group 0: genome-wide background
  for (long long id = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); id < p.ncand; id += nwarp) {
    const int lo = __ldg(p.wlo + id), hi = __ldg(p.whi + id);
    const int cnt = hi - lo;
    if (cnt == 0 || cnt > WCAP) continue;
    double g2 = 0.0, g1a = 0.0, g1b = 0.0;
    uint32_t k_any = 0, fa_any = 0, fb_any = 0;  // a populated bin of each spectrum (THE bin of a one-bin window)
    constexpr int Q = 4;
    for (int base = 0; base < cnt; base += Q * 32) {
      uint2 r[Q];
      double l2[Q], la[Q], lb[Q];
#pragma unroll
      for (int j = 0; j < Q; ++j) {
        const int i = base + j * 32 + lane;
        r[j] = i < cnt ? __ldcs(p.rec + lo + i) : make_uint2(0u, 0u);
      }
#pragma unroll
      for (int j = 0; j < Q; ++j) {
        const uint32_t k = r[j].x, fa = r[j].y & 0xFFFF, fb = r[j].y >> 16;
        const bool v2 = k != 0 && k != last;
        l2[j] = v2 ? __ldg(p.lb2 + k) : 0.0;
        la[j] = fa ? __ldg(p.lb1a + fa) : 0.0;
        lb[j] = fb ? __ldg(p.lb1b + fb) : 0.0;
        if (v2) k_any = k;
        if (fa) fa_any = fa;
        if (fb) fb_any = fb;
      }
#pragma unroll
      for (int j = 0; j < Q; ++j) { g2 += l2[j]; g1a += la[j]; g1b += lb[j]; }
    }
    g2 = warp_sum(g2); g1a = warp_sum(g1a); g1b = warp_sum(g1b);
    // a populated bin per spectrum from the lowest lane that saw one
    const uint32_t b2m = __ballot_sync(0xffffffffu, k_any != 0), bam = __ballot_sync(0xffffffffu, fa_any != 0),
                   bbm = __ballot_sync(0xffffffffu, fb_any != 0);
    k_any = __shfl_sync(0xffffffffu, k_any, b2m ? __ffs(b2m) - 1 : 0);
    fa_any = __shfl_sync(0xffffffffu, fa_any, bam ? __ffs(bam) - 1 : 0);
    fb_any = __shfl_sync(0xffffffffu, fb_any, bbm ? __ffs(bbm) - 1 : 0);
    const uint8_t one = p.r_flags[id];
    const int Nq = lane == 0 ? p.r_n2[id] : (lane == 1 ? p.r_n1a[id] : (lane == 2 ? p.r_n1b[id] : 0));
    const double Wq = lane == 0 ? p.r_T2[id] : (lane == 1 ? p.r_T1a[id] : (lane == 2 ? p.r_T1b[id] : 0.0));
    const double Gq = lane == 0 ? g2 : (lane == 1 ? g1a : g1b);
    bool none = false, own_bg = false;
    double Tq = 0.0;
    if (lane < 3) {
      double acc = Wq - Gq;
      if (one & (1u << lane)) {  // one populated bin: the per-bin form x (ln x - ln b), exactly as k3_score_small computes it
        const double lbin = lane == 0 ? __ldg(p.lb2 + k_any) : (lane == 1 ? __ldg(p.lb1a + fa_any) : __ldg(p.lb1b + fb_any));
        const double lnx = Nq > 1 ? ln_mult(p, (uint32_t)Nq) : 0.0;
        acc = fma(u32_to_double((uint32_t)Nq), lnx - lbin, 0.0);
      }
      own_bg = Nq > 0 && (double)Nq == __ldg(Bg + lane);  // N == B: possibly the background itself -> per-bin re-score
      Tq = clr_value(p, Nq, acc, Bg, lane, none);
    }
    if (__any_sync(0xffffffffu, own_bg)) {
      if (lane == 0) large[atomicAdd(nlarge, 1)] = (int32_t)id;  // k3_score_large writes every field of this window
      continue;
    }
    const uint32_t nb = __ballot_sync(0xffffffffu, none) & 7u;
    if (lane == 0) {
      p.r_flags[id] = (uint8_t)nb;
      p.r_T2[id] = Tq;
    } else if (lane == 1) {
      p.r_T1a[id] = Tq;
    } else if (lane == 2) {
      p.r_T1b[id] = Tq;
    }
  }
}

}  // namespace tdsfs
