// K1 — plan kernel: the greedy grouping of fuse_tracks_th (tracking.py:652-701) decided from the
// first <=30 tracks of every chunk, step by step, with the reference's operation order
// (explicit round-to-nearest intrinsics, no FMA contraction) so that the discontinuous
// `value < threshold` / `count/30 > 0.8` decisions see the same numbers as numpy does.
//
// One CTA per chunk.  lane = leader track, warps stride over sequences.  Leader-track state
// lives in a per-chunk global scratch block that stays L1/L2 resident.
#pragma once
#include "xt_common.cuh"

struct K1Args {
  const XtChunk* chunks;
  const double* soa;
  double* state;       // [n_chunks][2][cap][CO][32]
  double* hist;        // [n_chunks][2][cap][RH][nS]
  XtPlanPtrs plan;
  XtChunkSummary* summ;
  int32_t cap;         // children capacity
  int32_t RH;          // history rows allocated per sequence
  int32_t bits;        // bits per history row in the window code
};

__device__ __forceinline__ int xt_label(int x, int nS, bool wrap) {
  if (wrap) {
    int v = (int)(int8_t)(x & 0xFF);  // np.arange(..., dtype='int8') wraps, tracking.py:543
    int r = v % nS;
    return r < 0 ? r + nS : r;        // np.mod is non-negative for a positive divisor
  }
  return x % nS;
}

// numpy's pairwise summation (n < 8: plain loop; blocks of 8 accumulators up to 128; recursive
// halving above) applied to f(k), k in [lo, lo+n).  Used where the reference reduces over a
// contiguous axis (sum of weights, and the s2 merge when s2 has one component).
template <typename F>
__device__ __forceinline__ double xt_pairwise_block(F f, int lo, int n) {  // n <= 128
  if (n < 8) {
    double res = 0.0;
    for (int k = 0; k < n; ++k) res = __dadd_rn(res, f(lo + k));
    return res;
  }
  double r[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) r[k] = f(lo + k);
  int i = 8;
  for (; i < n - (n % 8); i += 8) {
#pragma unroll
    for (int k = 0; k < 8; ++k) r[k] = __dadd_rn(r[k], f(lo + i + k));
  }
  double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                         __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
  for (; i < n; ++i) res = __dadd_rn(res, f(lo + i));
  return res;
}

// The recursion "n > 128: split at n/2 rounded down to a multiple of 8" is unrolled into an
// explicit post-order walk with a small value stack (depth <= log2(XT_HARD_CAP/128)+1), so no
// device call stack is needed.
template <typename F>
__device__ double xt_pairwise(F f, int lo, int n) {
  if (n <= 128) return xt_pairwise_block(f, lo, n);
  int s_lo[8], s_n[8], s_state[8];
  double s_val[8];
  int sp = 0;
  s_lo[0] = lo; s_n[0] = n; s_state[0] = 0;
  double ret = 0.0;
  while (sp >= 0) {
    const int clo = s_lo[sp], cn = s_n[sp];
    if (cn <= 128) {
      ret = xt_pairwise_block(f, clo, cn);
      --sp;
      continue;
    }
    int n2 = cn / 2;
    n2 -= n2 % 8;
    if (s_state[sp] == 0) {          // descend into the left half
      s_state[sp] = 1;
      ++sp;
      s_lo[sp] = clo; s_n[sp] = n2; s_state[sp] = 0;
    } else if (s_state[sp] == 1) {   // left done -> keep it, descend into the right half
      s_val[sp] = ret;
      s_state[sp] = 2;
      ++sp;
      s_lo[sp] = clo + n2; s_n[sp] = cn - n2; s_state[sp] = 0;
    } else {                         // both done
      ret = __dadd_rn(s_val[sp], ret);
      --sp;
    }
  }
  return ret;
}

template <int D, int KS>
__global__ void __launch_bounds__(XT_K1_THREADS) k1_plan(const K1Args a, const __grid_constant__ xt_params P) {
  constexpr int CO = D + 2 * KS + 1;  // m[D], s2[KS], s[KS], LP
  constexpr int W = XT_K1_THREADS / 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, tid = threadIdx.x;
  const XtChunk ck = a.chunks[blockIdx.x];
  const int nS = P.nS, nsub = P.nsub, cap = a.cap;
  int K = 1;
  for (int i = 0; i < nsub; ++i) K *= nS;
  const bool wrap = (P.flags & XT_FLAG_INT8_WRAP) != 0;

  extern __shared__ unsigned char k1_smem[];
  // smem carve-up (all sized by cap)
  unsigned long long* codeP = (unsigned long long*)k1_smem;
  unsigned long long* codeC = codeP + cap;
  int* gid = (int*)(codeC + cap);
  int* grank = gid + cap;
  int* gcnt = grank + cap;            // [cap+1]: group sizes -> offsets
  unsigned char* curP = (unsigned char*)(gcnt + cap + 1);
  __shared__ int s_flag;

  XtChunkSummary* sm = &a.summ[blockIdx.x];
  const int nP0 = K * nS;
  if (tid == 0) {
    sm->err = 0;
    sm->need_cap = 0;
    sm->max_nP = nP0;
    sm->max_nC = 0;
    sm->sum_nC = 0;
    sm->sum_nG = 0;
    s_flag = 0;
  }
  const int L = ck.L;
  // steps 2..L-1 expand; steps 2..L-2 fuse.  Work counters for chunks that need no plan:
  if (L < 4) {
    if (tid == 0) {
      long long sc = 0;
      int nC = nP0;
      if (L == 3) { nC = nP0 * K; sc += nC; }
      if (ck.isBL) sc += (long long)nC * K;
      sm->sum_nC = sc;
      sm->max_nC = (L == 3) ? nC : 0;
    }
    return;
  }

  const int Kt = ck.nT < XT_LEADERS ? ck.nT : XT_LEADERS;
  const bool act = lane < Kt;
  const int t = act ? lane : 0;
  const double* Cp = a.soa + ck.xyz_off + t;
  const size_t npad = (size_t)ck.nTpad;

  double* bufP = a.state + (size_t)blockIdx.x * 2 * cap * CO * 32;
  double* bufC = bufP + (size_t)cap * CO * 32;
  double* histP = a.hist + (size_t)blockIdx.x * 2 * cap * a.RH * nS;
  double* histN = histP + (size_t)cap * a.RH * nS;
  const int bits = a.bits;
  const unsigned long long rowmask = (1ull << bits) - 1ull;

#define ST(buf, slot, comp) (buf)[((size_t)(slot) * CO + (comp)) * 32 + lane]

  double l2[KS];
#pragma unroll
  for (int k = 0; k < KS; ++k) l2[k] = P.l2[k];

  // ---- first localisation (tracking.py:478-529) ----
  int nP = nP0;
  for (int c = warp; c < nP; c += W) {
#pragma unroll
    for (int dim = 0; dim < D; ++dim) ST(bufP, c, dim) = Cp[(size_t)(0 * D + dim) * npad];
#pragma unroll
    for (int k = 0; k < KS; ++k) ST(bufP, c, D + k) = __dadd_rn(l2[k], P.dd[c]);
    ST(bufP, c, D + 2 * KS) = __dadd_rn(P.LT[c], P.LF[c]);
  }
  int LhP = nsub + 1;
  for (int c = tid; c < nP; c += XT_K1_THREADS) {
    curP[c] = (unsigned char)(c % nS);
    unsigned long long code = 0;
    int x = c;
    for (int r = 0; r <= nsub; ++r) {
      int dg = x % nS;
      x /= nS;
      code |= (unsigned long long)dg << (bits * r);
      for (int s = 0; s < nS; ++s) histP[((size_t)c * a.RH + r) * nS + s] = (dg == s) ? 1.0 : 0.0;
    }
    codeP[c] = code;
  }
  int hist_dim0_is_nT = 0;  // cur_Bs_cat has a single row until the first fusion
  double th = P.threshold;
  long long sum_nC = 0, sum_nG = 0;
  int max_nP = nP, max_nC = 0;
  __syncthreads();

  for (int step = 2; step <= L - 2; ++step) {
    const int nC = nP * K;
    if (nC > cap) {
      if (tid == 0) {
        sm->err = 2;
        sm->need_cap = nC;
      }
      return;
    }
    sum_nC += nC;
    max_nC = nC > max_nC ? nC : max_nC;
    const int rec = ck.rec0 + (step - 2);
    const int LhC = LhP + nsub;
    const int rows_cmp = LhC < P.frame_len ? LhC : P.frame_len;  // rows kept in the window code
    const bool use_window = LhC > P.frame_len;
    const unsigned long long cmask = (bits * rows_cmp >= 64) ? ~0ull : ((1ull << (bits * rows_cmp)) - 1ull);

    // ---- expansion + Gaussian update on the leader tracks (tracking.py:540-570, :87-98) ----
    double cl[D];
#pragma unroll
    for (int dim = 0; dim < D; ++dim) cl[dim] = Cp[(size_t)((step - 1) * D + dim) * npad];
    for (int c = warp; c < nC; c += W) {
      const int p = c / K, r = c - p * K;
      const int head = r + K * (int)curP[p];
      const double dd = P.dd[head];
      const double LPp = ST(bufP, p, D + 2 * KS);
      double mm[D], s2[KS], q[KS];
#pragma unroll
      for (int dim = 0; dim < D; ++dim) mm[dim] = ST(bufP, p, dim);
#pragma unroll
      for (int k = 0; k < KS; ++k) {
        s2[k] = ST(bufP, p, D + k);
        q[k] = __dadd_rn(l2[k], s2[k]);
      }
      double quad = 0.0, logs = 0.0;
#pragma unroll
      for (int dim = 0; dim < D; ++dim) {
        const int k = (KS == 1) ? 0 : dim;
        const double df = __dsub_rn(cl[dim], mm[dim]);
        const double term = __ddiv_rn(__dmul_rn(df, df), __dmul_rn(2.0, q[k]));
        quad = (dim == 0) ? term : __dadd_rn(quad, term);
        const double nm = __ddiv_rn(__dadd_rn(__dmul_rn(mm[dim], l2[k]), __dmul_rn(cl[dim], s2[k])),
                                    __dadd_rn(l2[k], s2[k]));
        ST(bufC, c, dim) = nm;
      }
      if (KS == 1) {
        logs = __dmul_rn((double)D * -0.5, log(__dmul_rn(XT_TWO_PI, q[0])));
      } else {
#pragma unroll
        for (int k = 0; k < KS; ++k) {
          const double lg = __dmul_rn(-0.5, log(__dmul_rn(XT_TWO_PI, q[k])));
          logs = (k == 0) ? lg : __dadd_rn(logs, lg);
        }
      }
#pragma unroll
      for (int k = 0; k < KS; ++k) {
        const double ns2 = __ddiv_rn(
            __dadd_rn(__dadd_rn(__dmul_rn(dd, l2[k]), __dmul_rn(dd, s2[k])), __dmul_rn(l2[k], s2[k])), q[k]);
        ST(bufC, c, D + k) = ns2;
        ST(bufC, c, D + KS + k) = __dsqrt_rn(ns2);
      }
      const double LC = __dsub_rn(logs, quad);
      double add = __dadd_rn(P.LT[head], LC);
      if (step >= P.min_len) add = __dadd_rn(add, P.Lp_stay[r]);
      ST(bufC, c, D + 2 * KS) = __dadd_rn(LPp, add);
    }
    // window codes of the children: nsub new labels in front of the parent's rows
    for (int c = tid; c < nC; c += XT_K1_THREADS) {
      const int p = c / K;
      unsigned long long code = codeP[p] << (bits * nsub);
      int x = c;
      for (int r = 0; r < nsub; ++r) {
        code |= (unsigned long long)xt_label(x, nS, wrap) << (bits * r);
        x /= nS;
      }
      codeC[c] = code & cmask;
      gid[c] = -1;
    }
    if (nC > P.max_nb_states) th = __dmul_rn(th, 1.2);  // sticky escalation, tracking.py:581-582
    __syncthreads();

    // ---- greedy grouping (tracking.py:667-698) ----
    int nG = 0;
    const double denom = (double)(Kt * KS);
    for (int i = 0; i < nC; ++i) {
      if (gid[i] >= 0) continue;  // uniform
      double mi[D], si[KS];
#pragma unroll
      for (int dim = 0; dim < D; ++dim) mi[dim] = ST(bufC, i, dim);
#pragma unroll
      for (int k = 0; k < KS; ++k) si[k] = ST(bufC, i, D + KS + k);
      const unsigned long long ci = codeC[i];
      for (int j = warp; j < nC; j += W) {
        if (gid[j] >= 0) continue;  // warp-uniform
        const unsigned long long cj = codeC[j];
        double am = 0.0, as = 0.0;
#pragma unroll
        for (int dim = 0; dim < D; ++dim) {
          const double v = fabs(__dsub_rn(ST(bufC, j, dim), mi[dim]));
          am = (dim == 0) ? v : __dadd_rn(am, v);
        }
        am = __ddiv_rn(am, (double)D);
        double sj[KS];
#pragma unroll
        for (int k = 0; k < KS; ++k) {
          sj[k] = ST(bufC, j, D + KS + k);
          const double v = fabs(__dsub_rn(sj[k], si[k]));
          as = (k == 0) ? v : __dadd_rn(as, v);
        }
        as = __ddiv_rn(as, (double)KS);
        int cnt_m = 0, cnt_s = 0;
#pragma unroll
        for (int k = 0; k < KS; ++k) {
          cnt_m += __popc(__ballot_sync(0xffffffffu, act && (__ddiv_rn(am, sj[k]) < th)));
          cnt_s += __popc(__ballot_sync(0xffffffffu, act && (__ddiv_rn(as, sj[k]) < th)));
        }
        const bool m_ok = __ddiv_rn((double)cnt_m, denom) > 0.8;
        const bool s_ok = __ddiv_rn((double)cnt_s, denom) > 0.8;
        const bool same_state = (cj & rowmask) == (ci & rowmask);
        const bool same_win = use_window && (cj == ci);
        if (((m_ok && s_ok && same_state) || same_win) && lane == 0) gid[j] = nG;
      }
      ++nG;
      __syncthreads();
    }
    // every sequence must have been grouped (tracking.py:700-701)
    for (int c = tid; c < nC; c += XT_K1_THREADS)
      if (gid[c] < 0) s_flag = 1;
    for (int g = tid; g <= nC; g += XT_K1_THREADS) gcnt[g] = 0;
    __syncthreads();
    if (s_flag) {
      if (tid == 0) sm->err = 1;
      return;
    }
    // ---- CSR member lists: rank inside the group (ascending child id), sizes, offsets ----
    for (int c = tid; c < nC; c += XT_K1_THREADS) {
      const int g = gid[c];
      int rk = 0;
      for (int c2 = 0; c2 < c; ++c2) rk += (gid[c2] == g);
      grank[c] = rk;
      atomicAdd(&gcnt[g + 1], 1);
    }
    __syncthreads();
    if (tid == 0) {
      for (int g = 0; g < nG; ++g) gcnt[g + 1] += gcnt[g];
    }
    __syncthreads();
    {
      uint16_t* goff = a.plan.goff + (size_t)rec * (a.plan.cap + 1);
      uint32_t* ent = a.plan.ent + (size_t)rec * a.plan.cap;
      uint16_t* pg = a.plan.gid + (size_t)rec * a.plan.cap;
      for (int g = tid; g <= nG; g += XT_K1_THREADS) goff[g] = (uint16_t)gcnt[g];
      for (int c = tid; c < nC; c += XT_K1_THREADS) {
        const int p = c / K, r = c - p * K;
        ent[gcnt[gid[c]] + grank[c]] = xt_pack_ent(p, r + K * (int)curP[p], r);
        pg[c] = (uint16_t)gid[c];
      }
      if (tid == 0) {
        a.plan.hdr[rec].nC = nC;
        a.plan.hdr[rec].nG = nG;
        a.plan.hdr[rec].th = th;
      }
    }
    __syncthreads();  // ent visible to the CTA (read back below through global memory)
    const uint32_t* ent = a.plan.ent + (size_t)rec * a.plan.cap;

    // ---- merge on the leader tracks (tracking.py:723-741), reference summation orders ----
    for (int g = warp; g < nG; g += W) {
      const int o = gcnt[g], n = gcnt[g + 1] - o;
      if (n == 1) {
        const int c = (int)(ent[o] & 0xFFFF) * K + (int)(ent[o] >> 24);
#pragma unroll
        for (int q = 0; q < CO; ++q) ST(bufP, g, q) = ST(bufC, c, q);
        continue;
      }
      auto child = [&](int k) { return (int)(ent[o + k] & 0xFFFF) * K + (int)(ent[o + k] >> 24); };
      double mx = ST(bufC, child(0), D + 2 * KS);
      for (int k = 1; k < n; ++k) mx = fmax(mx, ST(bufC, child(k), D + 2 * KS));
      auto wfun = [&](int k) { return exp(__dsub_rn(ST(bufC, child(k), D + 2 * KS), mx)); };
      const double sw = xt_pairwise(wfun, 0, n);
      double am[D];
#pragma unroll
      for (int dim = 0; dim < D; ++dim) {
        if (D == 1) {
          am[dim] = xt_pairwise([&](int k) { return __dmul_rn(wfun(k), ST(bufC, child(k), dim)); }, 0, n);
        } else {
          double acc = 0.0;
          for (int k = 0; k < n; ++k) {
            const double v = __dmul_rn(wfun(k), ST(bufC, child(k), dim));
            acc = (k == 0) ? v : __dadd_rn(acc, v);
          }
          am[dim] = acc;
        }
        ST(bufP, g, dim) = __ddiv_rn(am[dim], sw);
      }
#pragma unroll
      for (int k2 = 0; k2 < KS; ++k2) {
        double acc;
        if (KS == 1) {
          acc = xt_pairwise([&](int k) { return __dmul_rn(wfun(k), ST(bufC, child(k), D + k2)); }, 0, n);
        } else {
          acc = 0.0;
          for (int k = 0; k < n; ++k) {
            const double v = __dmul_rn(wfun(k), ST(bufC, child(k), D + k2));
            acc = (k == 0) ? v : __dadd_rn(acc, v);
          }
        }
        ST(bufP, g, D + k2) = __ddiv_rn(acc, sw);
      }
      ST(bufP, g, D + 2 * KS) = __dadd_rn(log(sw), mx);
    }
    // ---- history rows of the groups (fit mode: mean one-hot over members and leader tracks,
    //      tracking.py:714-715,735-737), accumulated in numpy's (track, member) order ----
    const int rows_out = rows_cmp;  // truncated to frame_len
    const int Kh = hist_dim0_is_nT ? Kt : 1;
    for (int idx = tid; idx < nG * rows_out * nS; idx += XT_K1_THREADS) {
      const int s = idx % nS, row = (idx / nS) % rows_out, g = idx / (nS * rows_out);
      const int o = gcnt[g], n = gcnt[g + 1] - o;
      auto val = [&](int k) -> double {
        const uint32_t e = ent[o + k];
        const int p = (int)(e & 0xFFFF);
        if (row < nsub) {
          int x = p * K + (int)(e >> 24);
          for (int r = 0; r < row; ++r) x /= nS;
          return xt_label(x, nS, wrap) == s ? 1.0 : 0.0;
        }
        return histP[((size_t)p * a.RH + (row - nsub)) * nS + s];
      };
      double out;
      if (n == 1) {
        out = val(0);
      } else {
        double acc = 0.0;
        bool first = true;
        for (int tt = 0; tt < Kh; ++tt)
          for (int k = 0; k < n; ++k) {
            acc = first ? val(k) : __dadd_rn(acc, val(k));
            first = false;
          }
        out = __ddiv_rn(acc, (double)(Kh * n));
      }
      histN[((size_t)g * a.RH + row) * nS + s] = out;
    }
    __syncthreads();
    // new parents: newest true state, window code (argmax per row, ties -> lowest state)
    {
      uint8_t* pcur = a.plan.curG + (size_t)rec * a.plan.cap;
      for (int g = tid; g < nG; g += XT_K1_THREADS) {
        const uint32_t e = ent[gcnt[g]];
        const int c0 = (int)(e & 0xFFFF) * K + (int)(e >> 24);
        const unsigned char cs = (unsigned char)(c0 % nS);
        unsigned long long code = 0;
        for (int row = 0; row < rows_out; ++row) {
          int best = 0;
          double bv = histN[((size_t)g * a.RH + row) * nS];
          for (int s = 1; s < nS; ++s) {
            const double v = histN[((size_t)g * a.RH + row) * nS + s];
            if (v > bv) { bv = v; best = s; }
          }
          code |= (unsigned long long)best << (bits * row);
        }
        // publish after all reads of curP/codeP of this step are done (next barrier)
        grank[g] = (int)cs;
        ((unsigned long long*)codeC)[g] = code;  // codeC is dead until the next expansion
        pcur[g] = cs;
      }
    }
    __syncthreads();
    for (int g = tid; g < nG; g += XT_K1_THREADS) {
      curP[g] = (unsigned char)grank[g];
      codeP[g] = codeC[g];
    }
    {  // swap history buffers
      double* tmp = histP;
      histP = histN;
      histN = tmp;
    }
    nP = nG;
    LhP = rows_out;
    hist_dim0_is_nT = 1;
    sum_nG += nG;
    max_nP = nP > max_nP ? nP : max_nP;
    __syncthreads();
  }
  if (tid == 0) {
    // last step (no fusion) and the optional end-of-track expansion, for the work counters
    const int nC = nP * K;
    sum_nC += nC;
    max_nC = nC > max_nC ? nC : max_nC;
    if (ck.isBL) sum_nC += (long long)nC * K;
    sm->sum_nC = sum_nC;
    sm->sum_nG = sum_nG;
    sm->max_nP = max_nP;
    sm->max_nC = max_nC;
  }
#undef ST
}
