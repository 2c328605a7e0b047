// K2 — replay kernel: every track of a chunk follows the chunk's plan (uniform control flow),
// one thread per track, one warp (32 tracks) per CTA, the live sequences of a track in shared
// memory (slot-major, lane-minor => conflict-free, no barriers needed), FP64 throughout.
//
// Restates, for all tracks: expansion + Gaussian-product update (tracking.py:540-570, :76-98),
// the weighted merge of fuse_tracks_th (:723-741), the end-of-track terms (:613-639) and the
// per-track log-sum-exp of Proba_Cs (:781-786), fused with the first level of the objective's
// sum (:1069).  Children of an expansion are never materialised: a parent slot holds
// (m', u = l2*s2/q, base = LP + LC) and child (p, r) is read as m', u + dd[head],
// base + LT[head] + Lp_stay[r]  (algebraically identical to :88-92; differences ~1 ulp).
#pragma once
#include "xt_common.cuh"

struct K2Args {
  const XtChunk* chunks;
  const XtWork* work;
  const double* soa;
  XtPlanPtrs plan;
  const XtChunkSummary* summ;
  double* gstate;      // global-memory state for chunks whose sequences do not fit shared memory
  double* logp;        // [n_tracks] per-track log P
  double* partial;     // [n_work] per-CTA sums
  int32_t Pcap;        // parent slots per ping-pong buffer (max over chunks of max_nP)
  int32_t n_work;
  double Lsum[XT_MAX_STATES];  // log sum_r exp(L_leave[r + K*state]) (end-of-track expansion folded)
  double ddu[XT_MAX_HEADS];    // VAR_DT: dd per unit time (from 2 D)
  XtAux ax;                    // VAR instantiation only; ax.leave = per-chunk exp-sums [nS] (linear)
};

template <int D, int KS, bool SMEM, bool VAR = false>
__global__ void __launch_bounds__(32) k2_replay(const K2Args a, const __grid_constant__ xt_params P) {
  constexpr int CO = D + KS + 1;  // m[D], s2|u[KS], LP|base
  const int lane = threadIdx.x;
  // SMEM variant: one work item per CTA; global-state variant: resident CTAs stride over items
  for (int wi = blockIdx.x; wi < a.n_work; wi += gridDim.x) {
  const XtWork wk = a.work[wi];
  const XtChunk ck = a.chunks[wk.chunk];

  const int nS = P.nS, nsub = P.nsub;
  int K = 1;
  for (int i = 0; i < nsub; ++i) K *= nS;
  const int t = wk.t0 + lane;
  const bool valid = t < ck.nT;
  const int tt = valid ? t : ck.nT - 1;
  const double* Cp = a.soa + ck.xyz_off + tt;
  const size_t npad = (size_t)ck.nTpad;
  const int L = ck.L;

  extern __shared__ double k2_smem[];
  double* S = SMEM ? (k2_smem + lane) : (a.gstate + (size_t)blockIdx.x * 2 * a.Pcap * CO * 32 + lane);
  const int Pcap = a.Pcap;
#define SA(buf, slot, comp) S[((size_t)((buf)*Pcap + (slot)) * CO + (comp)) * 32]

  double l2[KS];
#pragma unroll
  for (int k = 0; k < KS; ++k) l2[k] = P.l2[k];
  // VAR: row j of the aux block = (sigma components, time-reversed dt) of localisation j
  const bool var_loc = VAR && (P.flags & XT_FLAG_VAR_LOC), var_dt = VAR && (P.flags & XT_FLAG_VAR_DT);
  const double* Ap = VAR ? a.ax.aux + (size_t)(ck.xyz_off / D) * a.ax.R + tt : nullptr;
  const double* Lps = (VAR && a.ax.stay) ? a.ax.stay + (size_t)wk.chunk * K : P.Lp_stay;
  double dtv = 1.0;  // dt of the localisation whose expansion is being consumed
  auto var_row = [&](int j) {
    if (var_loc) {
#pragma unroll
      for (int k = 0; k < KS; ++k) l2[k] = xt_sigma2(P, Ap[(size_t)(j * a.ax.R + k) * npad]);
    }
  };
  auto var_dtrow = [&](int j) { return var_dt ? Ap[(size_t)(j * a.ax.R + a.ax.ka) * npad] : 1.0; };
#define DDV(head) (VAR ? (var_dt ? a.ddu[head] * dtv : P.dd[head]) : P.dd[head])
  if (VAR) {
    var_row(0);
    dtv = var_dtrow(0);
  }

  // ---- first localisation ----
  int nP = K * nS;
  int cur = 0;  // active buffer
  {
    double c0[D];
#pragma unroll
    for (int dim = 0; dim < D; ++dim) c0[dim] = Cp[(size_t)dim * npad];
    for (int c = 0; c < nP; ++c) {
#pragma unroll
      for (int dim = 0; dim < D; ++dim) SA(cur, c, dim) = c0[dim];
#pragma unroll
      for (int k = 0; k < KS; ++k) SA(cur, c, D + k) = l2[k] + DDV(c);
      SA(cur, c, D + KS) = P.LT[c] + P.LF[c];
    }
  }
  const uint8_t* curP = nullptr;  // newest true state of the parents (nullptr: initial c % nS)
  bool implicit = false;          // parent slots hold (m', u, base) of an un-fused last step

  for (int step = 2; step <= L - 1; ++step) {
    double cl[D];
#pragma unroll
    for (int dim = 0; dim < D; ++dim) cl[dim] = Cp[(size_t)((step - 1) * D + dim) * npad];
    if (VAR) {
      var_row(step - 1);
      dtv = var_dtrow(step - 1);
    }
    // phase A: per parent, the part of the update shared by all its children.  Four parents per round:
    // all loads first (the state lives in global memory: independent loads in flight hide its latency)
    for (int p0 = 0; p0 < nP; p0 += 4) {
      double mm4[4][D], s24[4][KS], lp4[4];
#pragma unroll
      for (int u4 = 0; u4 < 4; ++u4) {
        const int p = (p0 + u4 < nP) ? p0 + u4 : p0;
#pragma unroll
        for (int dim = 0; dim < D; ++dim) mm4[u4][dim] = SA(cur, p, dim);
#pragma unroll
        for (int k = 0; k < KS; ++k) s24[u4][k] = SA(cur, p, D + k);
        lp4[u4] = SA(cur, p, D + KS);
      }
#pragma unroll
      for (int u4 = 0; u4 < 4; ++u4) {
        const int p = p0 + u4;
        if (p < nP) {
          double quad = 0.0, logs = 0.0;
          double rq[KS];
#pragma unroll
          for (int k = 0; k < KS; ++k) {
            const double q = l2[k] + s24[u4][k];
            rq[k] = 1.0 / q;
            logs += log(XT_TWO_PI * q);
          }
#pragma unroll
          for (int dim = 0; dim < D; ++dim) {
            const int k = (KS == 1) ? 0 : dim;
            const double df = cl[dim] - mm4[u4][dim];
            quad += df * df * rq[k];
            SA(cur, p, dim) = (mm4[u4][dim] * l2[k] + cl[dim] * s24[u4][k]) * rq[k];
          }
#pragma unroll
          for (int k = 0; k < KS; ++k) SA(cur, p, D + k) = l2[k] * s24[u4][k] * rq[k];
          const double LC = (KS == 1) ? -0.5 * ((double)D * logs + quad) : -0.5 * (logs + quad);
          SA(cur, p, D + KS) = lp4[u4] + LC;
        }
      }
    }
    const bool stay = step >= P.min_len;
    if (step <= L - 2) {
      // phase B: weighted merge of the children into the plan's groups
      const int rec = ck.rec0 + (step - 2);
      const int nG = a.plan.hdr[rec].nG;
      const uint16_t* goff = a.plan.goff + (size_t)rec * (a.plan.cap + 1);
      const uint32_t* ent = a.plan.ent + (size_t)rec * a.plan.cap;
      const int nxt = cur ^ 1;
      int o = 0;
      for (int g = 0; g < nG; ++g) {
        const int o1 = (int)__ldg(&goff[g + 1]);
        const int n = o1 - o;
        if (n == 1) {
          const uint32_t e = __ldg(&ent[o]);
          const int p = (int)(e & 0xFFFF), head = (int)((e >> 16) & 0xFF), r = (int)(e >> 24);
#pragma unroll
          for (int dim = 0; dim < D; ++dim) SA(nxt, g, dim) = SA(cur, p, dim);
#pragma unroll
          for (int k = 0; k < KS; ++k) SA(nxt, g, D + k) = SA(cur, p, D + k) + DDV(head);
          SA(nxt, g, D + KS) = SA(cur, p, D + KS) + (P.LT[head] + (stay ? Lps[r] : 0.0));
        } else {
          double mx = -INFINITY;
          for (int k0 = 0; k0 < n; k0 += 4) {  // four members per round, loads first
            double b4[4], a4[4];
#pragma unroll
            for (int u4 = 0; u4 < 4; ++u4) {
              const uint32_t e = __ldg(&ent[o + (k0 + u4 < n ? k0 + u4 : k0)]);
              const int p = (int)(e & 0xFFFF), head = (int)((e >> 16) & 0xFF), r = (int)(e >> 24);
              b4[u4] = SA(cur, p, D + KS);
              a4[u4] = P.LT[head] + (stay ? Lps[r] : 0.0);
            }
#pragma unroll
            for (int u4 = 0; u4 < 4; ++u4) mx = fmax(mx, b4[u4] + a4[u4]);
          }
          double sw = 0.0, am[D], as[KS];
#pragma unroll
          for (int dim = 0; dim < D; ++dim) am[dim] = 0.0;
#pragma unroll
          for (int k = 0; k < KS; ++k) as[k] = 0.0;
          for (int k = 0; k < n; ++k) {
            const uint32_t e = __ldg(&ent[o + k]);
            const int p = (int)(e & 0xFFFF), head = (int)((e >> 16) & 0xFF), r = (int)(e >> 24);
            const double lp = SA(cur, p, D + KS) + (P.LT[head] + (stay ? Lps[r] : 0.0));
            const double w = exp(lp - mx);
            sw += w;
#pragma unroll
            for (int dim = 0; dim < D; ++dim) am[dim] += w * SA(cur, p, dim);
#pragma unroll
            for (int k2 = 0; k2 < KS; ++k2) as[k2] += w * (SA(cur, p, D + k2) + DDV(head));
          }
          const double rs = 1.0 / sw;
#pragma unroll
          for (int dim = 0; dim < D; ++dim) SA(nxt, g, dim) = am[dim] * rs;
#pragma unroll
          for (int k = 0; k < KS; ++k) SA(nxt, g, D + k) = as[k] * rs;
          SA(nxt, g, D + KS) = log(sw) + mx;
        }
        o = o1;
      }
      cur = nxt;
      nP = nG;
      curP = a.plan.curG + (size_t)rec * a.plan.cap;
    } else {
      implicit = true;
    }
  }

  // ---- end of track: last-localisation term, optional leave expansion (folded into Lsum),
  //      log-sum-exp over the surviving sequences (online max) ----
  double cl[D];
#pragma unroll
  for (int dim = 0; dim < D; ++dim) cl[dim] = Cp[(size_t)((L - 1) * D + dim) * npad];
  if (VAR) var_row(L - 1);  // (dtv still belongs to localisation L-2: the implicit children of step L-1)
  const bool stay_last = (L - 1) >= P.min_len;
  double mx = -INFINITY, acc = 0.0;
  const int Kc = implicit ? K : 1;
  for (int p = 0; p < nP; ++p) {
    const int ps = curP ? (int)__ldg(&curP[p]) : (p % nS);
    for (int r = 0; r < Kc; ++r) {
      double dd = 0.0, lpadd = 0.0;
      int newest = ps;
      if (implicit) {
        const int head = r + K * ps;
        dd = DDV(head);
        lpadd = P.LT[head] + (stay_last ? Lps[r] : 0.0);
        newest = r % nS;
      }
      double term = 0.0;
#pragma unroll
      for (int dim = 0; dim < D; ++dim) {
        const int k = (KS == 1) ? 0 : dim;
        const double q = SA(cur, p, D + k) + dd + l2[k];
        const double df = cl[dim] - SA(cur, p, dim);
        if (KS == 1 && dim > 0) {
          term += -df * df / (2.0 * q);
        } else {
          term += ((KS == 1) ? (double)D : 1.0) * -0.5 * log(XT_TWO_PI * q) - df * df / (2.0 * q);
        }
      }
      double v = SA(cur, p, D + KS) + lpadd + term;
      if (ck.isBL) v += (VAR && a.ax.leave) ? log(a.ax.leave[(size_t)wk.chunk * nS + newest]) : a.Lsum[newest];
      const double nm = fmax(mx, v);
      acc = acc * exp(mx - nm) + exp(v - nm);
      mx = nm;
    }
  }
  double lp = log(acc) + mx;
  if (valid) a.logp[ck.trk_off + t] = lp;
  if (!valid) lp = 0.0;
  // fixed-order warp reduction => bitwise reproducible objective
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) lp += __shfl_down_sync(0xffffffffu, lp, off);
  if (lane == 0) a.partial[wi] = lp;
  }
#undef SA
#undef DDV
}

