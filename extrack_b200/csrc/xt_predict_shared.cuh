// K3s — shared plans of the state annotation: predict_Bs(nb_max > 1) (tracking.py:803,860-896) evaluates chunks of
// nb_max tracks, and fuse_tracks_th(do_preds=1) (tracking.py:652-743) decides ONE grouping per chunk and step from
// the chunk's first 30 tracks, each with its own numbers and its own weighted history:
//   m_mask / s_mask : mean over (leader track, LocErr component) of the per-track predicate > 0.8      (:689-691)
//   cur_state_mask  : newest label of the sequence, from track 0                                        (:673-674)
//   state_mask      : the newest frame_len history rows (argmax per row) equal on > 99.9 % of the leaders (:679-681)
// This kernel computes those plans: one warp per chunk walks through the chunk's <= 30 leader tracks in turn in every
// phase (update, votes of the grouping, merge with the per-track weighted history window), with the arithmetic of
// k3_predict (xt_predict.cuh).  All tracks of the chunk, leaders included, are then annotated by k3_predict<.., FOLLOW>,
// which reads the groups from the plan instead of deciding them.  Scalar LocErr / dt models.
#pragma once
#include "xt_common.cuh"
#include "xt_plan.cuh"

struct K3SArgs {
  const XtChunk* chunks;
  const double* soa;
  double* scratch;      // per resident warp: 30 leader blocks + the shared index arrays
  int32_t* splan;       // [n_chunks][splan_stride]: nC[maxL], nG[maxL], then per step goff[cap + 1], order[cap]
  int32_t* err;         // [n_chunks] 0 ok, 1 grouping failure, 2 capacity overflow (need in err_need)
  int32_t* err_need;
  size_t warp_scratch;  // 8-byte units per warp
  size_t splan_stride;  // int32 units per chunk
  int32_t n_chunks;
  int32_t cap, maxL, bits;
  int32_t refine;       // 1: the recursion of the position refinement (get_LC_Km_Ks, refined_localization.py:48-204): no
                        // field-of-view / bleaching term, no initial fraction at the start
  int32_t rev;          // 1: consume the localisations from the last to the first
};

__host__ __device__ inline size_t k3s_splan_stride(int cap, int maxL) { return (size_t)2 * maxL + (size_t)maxL * (2 * cap + 1); }

// scratch of one leader track, in 8-byte units
struct K3SLayout {
  size_t bufP, bufC, histP, histN, codeP, codeC, trk_total;
  size_t gid, order, goff, curP, total;  // shared index arrays (offsets from the warp's base, after the 30 leader blocks)
};
__host__ __device__ inline K3SLayout k3s_layout(int cap, int CO, int fl, int nS) {
  K3SLayout l;
  const size_t capP = (size_t)cap / nS + 1;
  size_t o = 0;
  l.bufP = o;  o += capP * CO;
  l.bufC = o;  o += (size_t)cap * CO;
  l.histP = o; o += capP * fl * nS;
  l.histN = o; o += capP * fl * nS;
  l.codeP = o; o += capP;
  l.codeC = o; o += cap;
  l.trk_total = o;
  o = l.trk_total * XT_LEADERS;
  l.gid = o;   o += (cap + 1) / 2;
  l.order = o; o += (cap + 1) / 2;
  l.goff = o;  o += (cap + 2) / 2;
  l.curP = o;  o += (cap + 1) / 2;
  l.total = (o + 1) & ~(size_t)1;
  return l;
}

template <int D, int KS>
__global__ void __launch_bounds__(256) k3_shared_plan(const K3SArgs a, const __grid_constant__ xt_params P) {
  constexpr int CO = D + 2 * KS + 1;  // m[D], s2[KS], s[KS], LP
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int nS = P.nS, K = P.nS, cap = a.cap, fl = P.frame_len, bits = a.bits;
  const bool wrap = (P.flags & XT_FLAG_INT8_WRAP) != 0;
  const unsigned long long rowmask = (1ull << bits) - 1ull;
  const int capP = cap / nS + 1;
  const K3SLayout lay = k3s_layout(cap, CO, fl, nS);
  double* base = a.scratch + (size_t)(blockIdx.x * nwarps + warp) * a.warp_scratch;
  int* gid = (int*)(base + lay.gid);
  int* order = (int*)(base + lay.order);
  int* goff = (int*)(base + lay.goff);
  int* curP = (int*)(base + lay.curP);
  double l2[KS];
#pragma unroll
  for (int k = 0; k < KS; ++k) l2[k] = P.l2[k];

#define TRK(t) (base + (size_t)(t) * lay.trk_total)
#define BP(T, slot, comp) (T)[lay.bufP + (size_t)(comp) * capP + (slot)]
#define BC(T, slot, comp) (T)[lay.bufC + (size_t)(comp) * cap + (slot)]
#define HP(T, sel) ((T) + ((sel) ? lay.histN : lay.histP))
#define CODEP(T) ((unsigned long long*)((T) + lay.codeP))
#define CODEC(T) ((unsigned long long*)((T) + lay.codeC))

  for (int ci = blockIdx.x * nwarps + warp; ci < a.n_chunks; ci += gridDim.x * nwarps) {
    const XtChunk ck = a.chunks[ci];
    const int L = ck.L;
    int32_t* plan = a.splan + (size_t)ci * a.splan_stride;
    int32_t* planC = plan;
    int32_t* planG = plan + a.maxL;
    int32_t* planL = plan + 2 * a.maxL;  // per step: goff[cap + 1], order[cap]
    if (L < 4) continue;                 // no fusion step
    const int Kt = ck.nT < XT_LEADERS ? ck.nT : XT_LEADERS;
    const size_t npad = (size_t)ck.nTpad;
    // smallest count with (double)count / (Kt*KS) > 0.8  (np.mean(bool) > 0.8, tracking.py:689-691)
    int min_cnt = Kt * KS + 1;
    {
      const double denom = (double)(Kt * KS);
      for (int c = Kt * KS; c >= 0; --c)
        if (__ddiv_rn((double)c, denom) > 0.8) min_cnt = c;
    }
    int errc = 0;
    int hsel = 0;  // which history buffer holds the parents' rows
    const bool rev = a.rev != 0;
#define LROW(j) (rev ? (L - 1 - (j)) : (j))  // localisation consumed j-th
    // ---- first localisation (tracking.py:478-529) ----
    int nP = nS * nS;
    for (int t = 0; t < Kt; ++t) {
      double* T = TRK(t);
      const double* Cp = a.soa + ck.xyz_off + t;
      double* hP = HP(T, hsel);
      for (int c = lane; c < nP; c += 32) {
#pragma unroll
        for (int dim = 0; dim < D; ++dim) BP(T, c, dim) = Cp[(size_t)(LROW(0) * D + dim) * npad];
#pragma unroll
        for (int k = 0; k < KS; ++k) BP(T, c, D + k) = __dadd_rn(l2[k], P.dd[c]);
        BP(T, c, D + 2 * KS) = a.refine ? P.LT[c] : __dadd_rn(P.LT[c], P.LF[c]);
        const int d0 = c % nS, d1 = c / nS;
        CODEP(T)[c] = (unsigned long long)d0 | ((unsigned long long)d1 << bits);
        for (int s = 0; s < nS; ++s) {
          hP[((size_t)c * fl + 0) * nS + s] = (d0 == s) ? 1.0 : 0.0;
          if (fl > 1) hP[((size_t)c * fl + 1) * nS + s] = (d1 == s) ? 1.0 : 0.0;
        }
      }
    }
    for (int c = lane; c < nP; c += 32) curP[c] = c % nS;
    int LhP = 2;
    double th = P.threshold;
    __syncwarp();

    for (int step = 2; step <= L - 2; ++step) {
      const int nC = nP * K;
      if (nC > cap) {
        errc = 2;
        if (lane == 0) atomicMax(&a.err_need[ci], nC);
        break;
      }
      const int LhC = LhP + 1;
      const int rows_cmp = LhC < fl ? LhC : fl;
      const bool use_window = LhC > fl;
      const unsigned long long cmask = (bits * rows_cmp >= 64) ? ~0ull : ((1ull << (bits * rows_cmp)) - 1ull);
      const bool stay = !a.refine && step >= P.min_len;
      // ---- expansion + Gaussian update of every leader track (tracking.py:540-570, :87-98), lane = child ----
      for (int t = 0; t < Kt; ++t) {
        double* T = TRK(t);
        const double* Cp = a.soa + ck.xyz_off + t;
        double cl[D];
#pragma unroll
        for (int dim = 0; dim < D; ++dim) cl[dim] = Cp[(size_t)(LROW(step - 1) * D + dim) * npad];
        for (int c = lane; c < nC; c += 32) {
          const int p = c / K, r = c - p * K;
          const int head = r + K * curP[p];
          const double dd = P.dd[head];
          double s2[KS], q[KS];
#pragma unroll
          for (int k = 0; k < KS; ++k) {
            s2[k] = BP(T, p, D + k);
            q[k] = __dadd_rn(l2[k], s2[k]);
          }
          double quad = 0.0, logs = 0.0;
#pragma unroll
          for (int dim = 0; dim < D; ++dim) {
            const int k = (KS == 1) ? 0 : dim;
            const double mm = BP(T, p, dim);
            const double df = __dsub_rn(cl[dim], mm);
            const double term = __ddiv_rn(__dmul_rn(df, df), __dmul_rn(2.0, q[k]));
            quad = (dim == 0) ? term : __dadd_rn(quad, term);
            BC(T, c, dim) = __ddiv_rn(__dadd_rn(__dmul_rn(mm, l2[k]), __dmul_rn(cl[dim], s2[k])), __dadd_rn(l2[k], s2[k]));
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
            BC(T, c, D + k) = ns2;
            BC(T, c, D + KS + k) = __dsqrt_rn(ns2);
          }
          double add = __dadd_rn(P.LT[head], __dsub_rn(logs, quad));
          if (stay) add = __dadd_rn(add, P.Lp_stay[r]);
          BC(T, c, D + 2 * KS) = __dadd_rn(BP(T, p, D + 2 * KS), add);
          CODEC(T)[c] = ((CODEP(T)[p] << bits) | (unsigned long long)xt_label(c, nS, wrap)) & cmask;
        }
      }
      for (int c = lane; c < nC; c += 32) gid[c] = -1;
      if (nC > P.max_nb_states) th = __dmul_rn(th, 1.2);
      __syncwarp();

      // ---- greedy grouping from the votes of the leader tracks (tracking.py:667-698) ----
      const double th_lo = __dmul_rn(th, 1.0 - 1e-14), th_hi = __dmul_rn(th, 1.0 + 1e-14);
      int nG = 0, off = 0;
      for (int i = 0; i < nC; ++i) {
        if (gid[i] >= 0) continue;  // warp-uniform
        const unsigned long long ci0 = CODEC(TRK(0))[i];
        goff[nG] = off;
        for (int j0 = 0; j0 < nC; j0 += 32) {
          const int j = j0 + lane;
          const bool cand = j < nC && gid[j] < 0;
          // newest label from track 0 (:673-674); the floating-point votes are only needed for those sequences
          const bool cs = cand && ((CODEC(TRK(0))[j] & rowmask) == (ci0 & rowmask));
          int cnt_m = 0, cnt_s = 0, cnt_w = 0;
          if (cand) {
            for (int t = 0; t < Kt; ++t) {
              double* T = TRK(t);
              if (use_window && CODEC(T)[j] == CODEC(T)[i]) ++cnt_w;
              if (!cs) continue;
              double am = 0.0, as = 0.0;
#pragma unroll
              for (int dim = 0; dim < D; ++dim) {
                const double v = fabs(__dsub_rn(BC(T, j, dim), BC(T, i, dim)));
                am = (dim == 0) ? v : __dadd_rn(am, v);
              }
              am = (D == 2) ? __dmul_rn(am, 0.5) : ((D == 1) ? am : __ddiv_rn(am, (double)D));
              double sj[KS];
#pragma unroll
              for (int k = 0; k < KS; ++k) {
                sj[k] = BC(T, j, D + KS + k);
                const double v = fabs(__dsub_rn(sj[k], BC(T, i, D + KS + k)));
                as = (k == 0) ? v : __dadd_rn(as, v);
              }
              as = (KS == 2) ? __dmul_rn(as, 0.5) : ((KS == 1) ? as : __ddiv_rn(as, (double)KS));
#pragma unroll
              for (int k = 0; k < KS; ++k) {
                // fl(x / s) < th decided without a division unless x is within 1e-14 (relative) of th*s
                const double lo = __dmul_rn(th_lo, sj[k]), hi = __dmul_rn(th_hi, sj[k]);
                bool pm = am < lo, ps = as < lo;
                if (!pm && !(am > hi)) pm = __ddiv_rn(am, sj[k]) < th;
                if (!ps && !(as > hi)) ps = __ddiv_rn(as, sj[k]) < th;
                cnt_m += pm;
                cnt_s += ps;
              }
            }
          }
          // state_mask: > 99.9 % of at most 30 leaders = all of them (:679-681)
          const bool ok = cand && ((cs && cnt_m >= min_cnt && cnt_s >= min_cnt) || (use_window && cnt_w == Kt));
          const unsigned m = __ballot_sync(0xffffffffu, ok);
          if (ok) {
            gid[j] = nG;
            order[off + __popc(m & ((1u << lane) - 1u))] = j;
          }
          off += __popc(m);
        }
        if (off == goff[nG]) errc = 1;  // the leader captured nobody, not even itself (:725)
        ++nG;
        __syncwarp();
      }
      if (lane == 0) goff[nG] = off;
      for (int c = lane; c < nC; c += 32)
        if (gid[c] < 0) errc = 1;  // tracking.py:700-701
      if (errc == 0 && nG * K > cap) {
        errc = 2;
        if (lane == 0) atomicMax(&a.err_need[ci], nG * K);
      }
      errc = __reduce_max_sync(0xffffffffu, errc);
      __syncwarp();
      if (errc) break;
      // ---- the plan of this step, for k3_predict<.., FOLLOW> ----
      {
        int32_t* pl = planL + (size_t)step * (2 * cap + 1);
        if (lane == 0) {
          planC[step] = nC;
          planG[step] = nG;
        }
        for (int g = lane; g <= nG; g += 32) pl[g] = goff[g];
        for (int c = lane; c < nC; c += 32) pl[cap + 1 + c] = order[c];
      }
      // ---- merge of every leader track, lane = group (tracking.py:723-741), with the weighted history window ----
      const int rows_out = rows_cmp;
      for (int t = 0; t < Kt; ++t) {
        double* T = TRK(t);
        const double* hP = HP(T, hsel);
        double* hN = HP(T, hsel ^ 1);
        for (int g = lane; g < nG; g += 32) {
          const int o = goff[g], n = goff[g + 1] - o;
          const int c0 = order[o];
          if (n == 1) {
#pragma unroll
            for (int q = 0; q < CO; ++q) BP(T, g, q) = BC(T, c0, q);
            for (int row = 0; row < rows_out; ++row)
              for (int s = 0; s < nS; ++s)
                hN[((size_t)g * fl + row) * nS + s] =
                    (row == 0) ? ((xt_label(c0, nS, wrap) == s) ? 1.0 : 0.0) : hP[((size_t)(c0 / K) * fl + row - 1) * nS + s];
          } else {
            double mx = BC(T, c0, D + 2 * KS);
            for (int k = 1; k < n; ++k) mx = fmax(mx, BC(T, order[o + k], D + 2 * KS));
            double sw = 0.0, am[D], as2[KS];
            for (int k = 0; k < n; ++k) {
              const int c = order[o + k];
              const double w = exp(__dsub_rn(BC(T, c, D + 2 * KS), mx));
              sw = (k == 0) ? w : __dadd_rn(sw, w);
#pragma unroll
              for (int dim = 0; dim < D; ++dim) {
                const double v = __dmul_rn(w, BC(T, c, dim));
                am[dim] = (k == 0) ? v : __dadd_rn(am[dim], v);
              }
#pragma unroll
              for (int k2 = 0; k2 < KS; ++k2) {
                const double v = __dmul_rn(w, BC(T, c, D + k2));
                as2[k2] = (k == 0) ? v : __dadd_rn(as2[k2], v);
              }
            }
            // weighted mean of the members' window rows (tracking.py:733), member order; weights recomputed (same bits)
            for (int row = 0; row < rows_out; ++row)
              for (int s = 0; s < nS; ++s) {
                double acc = 0.0;
                for (int k = 0; k < n; ++k) {
                  const int c = order[o + k];
                  const double w = exp(__dsub_rn(BC(T, c, D + 2 * KS), mx));
                  const double hv = (row == 0) ? ((xt_label(c, nS, wrap) == s) ? 1.0 : 0.0)
                                               : hP[((size_t)(c / K) * fl + row - 1) * nS + s];
                  const double v = __dmul_rn(w, hv);
                  acc = (k == 0) ? v : __dadd_rn(acc, v);
                }
                hN[((size_t)g * fl + row) * nS + s] = __ddiv_rn(acc, sw);
              }
#pragma unroll
            for (int dim = 0; dim < D; ++dim) BP(T, g, dim) = __ddiv_rn(am[dim], sw);
#pragma unroll
            for (int k2 = 0; k2 < KS; ++k2) BP(T, g, D + k2) = __ddiv_rn(as2[k2], sw);
            BP(T, g, D + 2 * KS) = __dadd_rn(log(sw), mx);
          }
          // window code of the merged history (argmax per row, ties -> lowest state)
          unsigned long long code = 0;
          for (int row = 0; row < rows_out; ++row) {
            int best = 0;
            double bv = hN[((size_t)g * fl + row) * nS];
            for (int s = 1; s < nS; ++s) {
              const double v = hN[((size_t)g * fl + row) * nS + s];
              if (v > bv) { bv = v; best = s; }
            }
            code |= (unsigned long long)best << (bits * row);
          }
          CODEP(T)[g] = code;  // (the parents' codes were consumed by the update of this step)
        }
      }
      __syncwarp();
      for (int g = lane; g < nG; g += 32) curP[g] = order[goff[g]] % nS;  // newest true state of the representative (:728)
      hsel ^= 1;
      nP = nG;
      LhP = LhC;
      __syncwarp();
    }
    if (errc && lane == 0) atomicMax(&a.err[ci], errc);
  }
#undef LROW
#undef TRK
#undef BP
#undef BC
#undef HP
#undef CODEP
#undef CODEC
}
