// K2 (fast path) — replay kernel in the scaled linear domain.
//
// Same recursion as xt_replay.cuh (the log-domain variant kept for the global-memory fallback),
// restated like the scaled forward algorithm of an HMM: a sequence carries a linear weight W and
// every track carries one log-scale `lnscale`, so that log P(sequence) = log W + lnscale.  Then
//   * the Gaussian-product update (tracking.py:87-98) costs one reciprocal and one exp per parent
//     (no log): W' = W * prod_k q_k^-1/2 * exp(-sum (c-m)^2 / 2q)   [the (2 pi)^-d/2 factors are
//     added analytically at the end of the track];
//   * the merge of fuse_tracks_th (:723-741) needs no exp/log at all: the members' weights ARE
//     the merge weights (w_j / sum w_j), LP_g = log sum exp(LP_j) becomes W_g = sum W_j;
//   * underflow is handled per step and per track by a shift E chosen from
//     key_p = e_p + ln2 * exponent(f_p) so that the largest child weight lands in [1, 2).
// Results agree with the log-domain formulation to ~1e-15 relative per track.
//
// Mapping: a CTA of WPC warps owns 32 tracks (lane = track).  Warp w processes parents
// p = w, w+WPC, ... in the update phase and groups g = w, w+WPC, ... in the merge phase, so the
// same shared-memory tile (2 * Pcap slots of (m[D], s2[KS], W) per track) feeds WPC times more
// warps than one-warp-per-tile; three CTA barriers per step.
#pragma once
#include "xt_common.cuh"
#include "xt_replay.cuh"

struct K2Lin {  // linear-domain tables derived on the host from xt_params (per evaluation)
  double winit[XT_MAX_HEADS];  // exp(LT + LF)
  double tau0[XT_MAX_HEADS];   // exp(LT[head])
  double tau1[XT_MAX_HEADS];   // exp(LT[head] + Lp_stay[r])
  double leave[XT_MAX_STATES]; // sum_r exp(L_leave[r + K*state])
};

#define XT_LN2 0.6931471805599453
#define XT_LN_2PI 1.8378770664093453

// exp() with the polynomial coefficients in the constant bank (operands of DFMA, no UMOV
// traffic) and branch-free two-step scaling (gradual underflow handled, x <= 709.7 assumed to
// matter only up to overflow -> inf).  Taylor degree 13 on |r| <= ln2/2: truncation 4e-18.
__constant__ double c_xt_exp[14] = {1.0,
                                    1.0,
                                    0.5,
                                    1.6666666666666666e-01,
                                    4.1666666666666664e-02,
                                    8.3333333333333332e-03,
                                    1.3888888888888889e-03,
                                    1.9841269841269841e-04,
                                    2.4801587301587302e-05,
                                    2.7557319223985893e-06,
                                    2.7557319223985888e-07,
                                    2.5052108385441720e-08,
                                    2.0876756987868100e-09,
                                    1.6059043836821613e-10};

__device__ __forceinline__ double xt_exp(double x) {
  x = fmax(x, -745.2);
  double t = fma(x, 1.4426950408889634, 6755399441055744.0);
  const int k = __double2loint(t);
  t -= 6755399441055744.0;
  double r = fma(t, -6.93147180369123816490e-01, x);
  r = fma(t, -1.90821492927058770002e-10, r);
  double p = c_xt_exp[13];
#pragma unroll
  for (int i = 12; i >= 0; --i) p = fma(p, r, c_xt_exp[i]);
  const int k1 = k >> 1, k2 = k - k1;
  const double s1 = __hiloint2double((k1 + 1023) << 20, 0);
  const double s2 = __hiloint2double((k2 + 1023) << 20, 0);
  return (p * s1) * s2;
}

// 1/x for normal positive x: hardware seed + two Newton steps (~1 ulp, no slow path)
__device__ __forceinline__ double xt_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  return r;
}

template <int D, int KS>
__device__ __forceinline__ double xt_normfac(const double (&rq)[KS]) {
  // prod over dims of q^-1/2 given rq = 1/q
  if (KS == 1) {
    if (D == 1) return sqrt(rq[0]);
    if (D == 2) return rq[0];
    return rq[0] * sqrt(rq[0]);
  } else {
    double pr = rq[0];
#pragma unroll
    for (int k = 1; k < KS; ++k) pr *= rq[k];
    return sqrt(pr);
  }
}

__device__ __forceinline__ double xt_expo_ln2(double f) {
  const int ex = ((__double2hiint(f) >> 20) & 0x7ff) - 1023;
  return XT_LN2 * (double)ex;
}

template <int D, int KS, int WPC>
__global__ void __launch_bounds__(32 * WPC) k2_replay_lin(const K2Args a, const __grid_constant__ xt_params P,
                                                         const __grid_constant__ K2Lin T) {
  constexpr int CO = D + KS + 1;  // m[D], s2|u[KS], W
  constexpr int SL = CO * 32;     // doubles per slot
  constexpr int IW = (D + KS) * 32;  // offset of the weight inside a slot
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int wi = blockIdx.x;
  const XtWork wk = a.work[wi];
  const XtChunk ck = a.chunks[wk.chunk];
  const int nS = P.nS, nsub = P.nsub;
  int K = 1;
  for (int i = 0; i < nsub; ++i) K *= nS;
  const int t = wk.t0 + lane;
  const bool valid = t < ck.nT;
  const int tt = valid ? t : ck.nT - 1;
  const size_t npad = (size_t)ck.nTpad;
  const double* Cs = a.soa + ck.xyz_off + tt;  // localisation of the current step (advanced per step)
  const size_t cstride = (size_t)D * npad;
  const int L = ck.L;
  const int Pcap = a.Pcap;

  extern __shared__ double k2_smem[];
  double* X = k2_smem + lane;        // parents entering a step / merged groups
  double* Y = X + Pcap * SL;         // parents after the update: (m', u, W')
  double* red = k2_smem + 2 * Pcap * SL + lane;  // [2][WPC][32] cross-warp exchange

  double l2[KS];
#pragma unroll
  for (int k = 0; k < KS; ++k) l2[k] = P.l2[k];

  // ---- first localisation ----
  int nP = K * nS;
  {
    double c0[D];
#pragma unroll
    for (int dim = 0; dim < D; ++dim) c0[dim] = Cs[(size_t)dim * npad];
    for (int c = w; c < nP; c += WPC) {
      double* x = X + c * SL;
#pragma unroll
      for (int dim = 0; dim < D; ++dim) x[dim * 32] = c0[dim];
#pragma unroll
      for (int k = 0; k < KS; ++k) x[(D + k) * 32] = l2[k] + P.dd[c];
      x[IW] = T.winit[c];
    }
  }
  __syncthreads();
  const uint8_t* curP = nullptr;
  bool implicit = false;
  double lnscale = 0.0;
  const int xoff0 = w * SL;

  constexpr int R = 4;  // parents per warp kept in registers across the scale exchange
  double cn[D];         // localisation of the next step, loaded one step ahead (hides DRAM latency)
  Cs += cstride;
#pragma unroll
  for (int dim = 0; dim < D; ++dim) cn[dim] = Cs[(size_t)dim * npad];  // C[1] (exists: L >= 2)

  for (int step = 2; step <= L - 1; ++step) {
    double cl[D];
#pragma unroll
    for (int dim = 0; dim < D; ++dim) cl[dim] = cn[dim];
    Cs += cstride;  // step + 1 <= L  =>  C[step] exists
#pragma unroll
    for (int dim = 0; dim < D; ++dim) cn[dim] = Cs[(size_t)dim * npad];
    double* rd = red + (step & 1) * WPC * 32;
    if (nP <= R * WPC) {
      // ---- update, register path: this warp's <= R parents stay in registers between the
      //      two passes (no smem round trip for e and f, R independent dependency chains) ----
      double e[R], f[R];
      double kmax = -INFINITY;
#pragma unroll
      for (int i = 0; i < R; ++i) {
        const int p = w + i * WPC;
        e[i] = -INFINITY;
        f[i] = 0.0;
        if (p < nP) {
          double* x = X + p * SL;
          double* y = Y + p * SL;
          double rq[KS], s2[KS];
#pragma unroll
          for (int k = 0; k < KS; ++k) {
            s2[k] = x[(D + k) * 32];
            rq[k] = xt_rcp(l2[k] + s2[k]);
          }
          double quad = 0.0;
#pragma unroll
          for (int dim = 0; dim < D; ++dim) {
            const int k = (KS == 1) ? 0 : dim;
            const double mm = x[dim * 32];
            const double df = cl[dim] - mm;
            quad += df * df * rq[k];
            y[dim * 32] = (mm * l2[k] + cl[dim] * s2[k]) * rq[k];
          }
#pragma unroll
          for (int k = 0; k < KS; ++k) y[(D + k) * 32] = l2[k] * s2[k] * rq[k];
          e[i] = -0.5 * quad;
          f[i] = x[IW] * xt_normfac<D, KS>(rq);
          kmax = fmax(kmax, e[i] + xt_expo_ln2(f[i]));
        }
      }
      rd[w * 32] = kmax;
      __syncthreads();
      double E = rd[0];
#pragma unroll
      for (int k = 1; k < WPC; ++k) E = fmax(E, rd[k * 32]);
      E = fmax(E, -1e300);
      lnscale += E;
#pragma unroll
      for (int i = 0; i < R; ++i) {
        const int p = w + i * WPC;
        if (p < nP) Y[p * SL + IW] = f[i] * xt_exp(e[i] - E);
      }
    } else {
      // ---- update, generic path (any number of parents): e and f go through shared memory ----
      double kmax = -INFINITY;
      {
        double* x = X + xoff0;
        double* y = Y + xoff0;
        for (int p = w; p < nP; p += WPC, x += WPC * SL, y += WPC * SL) {
          double rq[KS], s2[KS];
#pragma unroll
          for (int k = 0; k < KS; ++k) {
            s2[k] = x[(D + k) * 32];
            rq[k] = xt_rcp(l2[k] + s2[k]);
          }
          double quad = 0.0;
#pragma unroll
          for (int dim = 0; dim < D; ++dim) {
            const int k = (KS == 1) ? 0 : dim;
            const double mm = x[dim * 32];
            const double df = cl[dim] - mm;
            quad += df * df * rq[k];
            y[dim * 32] = (mm * l2[k] + cl[dim] * s2[k]) * rq[k];
          }
#pragma unroll
          for (int k = 0; k < KS; ++k) y[(D + k) * 32] = l2[k] * s2[k] * rq[k];
          const double e = -0.5 * quad;
          const double f = x[IW] * xt_normfac<D, KS>(rq);
          kmax = fmax(kmax, e + xt_expo_ln2(f));
          y[IW] = e;
          x[IW] = f;
        }
      }
      rd[w * 32] = kmax;
      __syncthreads();
      double E = rd[0];
#pragma unroll
      for (int k = 1; k < WPC; ++k) E = fmax(E, rd[k * 32]);
      E = fmax(E, -1e300);
      lnscale += E;
      {
        double* x = X + xoff0;
        double* y = Y + xoff0;
        for (int p = w; p < nP; p += WPC, x += WPC * SL, y += WPC * SL) y[IW] = x[IW] * xt_exp(y[IW] - E);
      }
    }
    __syncthreads();
    if (step <= L - 2) {
      // ---- merge the children into the plan's groups (weights are the merge weights) ----
      const double* tau = (step >= P.min_len) ? T.tau1 : T.tau0;
      const int rec = ck.rec0 + (step - 2);
      const int nG = a.plan.hdr[rec].nG;
      const unsigned long long* grec = a.plan.grec + (size_t)rec * a.plan.cap;
      unsigned long long gnext = (w < nG) ? __ldg(&grec[w]) : 0ull;
      double* x = X + xoff0;
      for (int g = w; g < nG; g += WPC, x += WPC * SL) {
        const unsigned long long gr = gnext;
        if (g + WPC < nG) gnext = __ldg(&grec[g + WPC]);  // prefetch: no dependent load chain
        const int n = (int)((gr >> 24) & 0xFF);
        const int p0 = (int)(gr & 0xFFFF), h0 = (int)((gr >> 16) & 0xFF);
        const double* y0 = Y + p0 * SL;
        if (n == 1) {
#pragma unroll
          for (int dim = 0; dim < D; ++dim) x[dim * 32] = y0[dim * 32];
#pragma unroll
          for (int k = 0; k < KS; ++k) x[(D + k) * 32] = y0[(D + k) * 32] + P.dd[h0];
          x[IW] = y0[IW] * tau[h0];
        } else {
          const int p1 = (int)((gr >> 32) & 0xFFFF), h1 = (int)((gr >> 48) & 0xFF);
          const double* y1 = Y + p1 * SL;
          const double w0 = y0[IW] * tau[h0], w1 = y1[IW] * tau[h1];
          double sw = w0 + w1, am[D], as[KS];
#pragma unroll
          for (int dim = 0; dim < D; ++dim) am[dim] = w0 * y0[dim * 32] + w1 * y1[dim * 32];
#pragma unroll
          for (int k = 0; k < KS; ++k)
            as[k] = w0 * (y0[(D + k) * 32] + P.dd[h0]) + w1 * (y1[(D + k) * 32] + P.dd[h1]);
          if (n > 2) {  // members beyond the two inlined ones come from the CSR list
            const uint16_t* goff = a.plan.goff + (size_t)rec * (a.plan.cap + 1);
            const uint32_t* ent = a.plan.ent + (size_t)rec * a.plan.cap;
            const int o = (int)__ldg(&goff[g]), o1 = (int)__ldg(&goff[g + 1]);
            for (int k = o + 2; k < o1; ++k) {
              const uint32_t e = __ldg(&ent[k]);
              const int p = (int)(e & 0xFFFF), head = (int)((e >> 16) & 0xFF);
              const double* y = Y + p * SL;
              const double wj = y[IW] * tau[head];
              sw += wj;
#pragma unroll
              for (int dim = 0; dim < D; ++dim) am[dim] += wj * y[dim * 32];
#pragma unroll
              for (int k2 = 0; k2 < KS; ++k2) as[k2] += wj * (y[(D + k2) * 32] + P.dd[head]);
            }
          }
          if (sw > 1e-280) {  // below: 1/sw would overflow (denormal) -> treat as a zero-weight group
            const double rs = xt_rcp(sw);
#pragma unroll
            for (int dim = 0; dim < D; ++dim) x[dim * 32] = am[dim] * rs;
#pragma unroll
            for (int k = 0; k < KS; ++k) x[(D + k) * 32] = as[k] * rs;
          } else {  // every member underflowed: keep finite moments, (near-)zero weight
#pragma unroll
            for (int dim = 0; dim < D; ++dim) x[dim * 32] = y0[dim * 32];
#pragma unroll
            for (int k = 0; k < KS; ++k) x[(D + k) * 32] = y0[(D + k) * 32] + P.dd[h0];
          }
          x[IW] = sw;
        }
      }
      nP = nG;
      curP = a.plan.curG + (size_t)rec * a.plan.cap;
      __syncthreads();
    } else {
      implicit = true;
    }
  }

  // ---- end of track (tracking.py:613-639, :781-786) ----
  double cl[D];  // last localisation C[L-1] (prefetched)
#pragma unroll
  for (int dim = 0; dim < D; ++dim) cl[dim] = cn[dim];
  const double* tau = ((L - 1) >= P.min_len) ? T.tau1 : T.tau0;
  const double* src = implicit ? Y : X;
  const int Kc = implicit ? K : 1;
  double E2 = -INFINITY, acc = 0.0;
  for (int pass = 0; pass < 2; ++pass) {
    double kmax = -INFINITY;
    for (int p = w; p < nP; p += WPC) {
      const double* y = src + p * SL;
      const int ps = curP ? (int)__ldg(&curP[p]) : (p % nS);
      const double Wp = y[IW];
      double df2[D];
#pragma unroll
      for (int dim = 0; dim < D; ++dim) {
        const double df = cl[dim] - y[dim * 32];
        df2[dim] = df * df;
      }
      int newest_r = 0;  // r % nS, maintained incrementally
      for (int r = 0; r < Kc; ++r) {
        double dd = 0.0, fw = Wp;
        int newest = ps;
        if (implicit) {
          const int head = r + K * ps;
          dd = P.dd[head];
          fw *= tau[head];
          newest = newest_r;
        }
        if (++newest_r == nS) newest_r = 0;
        if (ck.isBL) fw *= T.leave[newest];
        double rq[KS];
#pragma unroll
        for (int k = 0; k < KS; ++k) rq[k] = xt_rcp(y[(D + k) * 32] + dd + l2[k]);
        double quad = 0.0;
#pragma unroll
        for (int dim = 0; dim < D; ++dim) quad += df2[dim] * rq[(KS == 1) ? 0 : dim];
        const double e = -0.5 * quad;
        const double f = fw * xt_normfac<D, KS>(rq);
        if (pass == 0) kmax = fmax(kmax, e + xt_expo_ln2(f));
        else acc += f * xt_exp(e - E2);
      }
    }
    double* rd = red + pass * WPC * 32;
    rd[w * 32] = pass == 0 ? kmax : acc;
    __syncthreads();
    if (pass == 0) {
      E2 = rd[0];
#pragma unroll
      for (int k = 1; k < WPC; ++k) E2 = fmax(E2, rd[k * 32]);
      E2 = fmax(E2, -1e300);
    } else {
      acc = rd[0];
#pragma unroll
      for (int k = 1; k < WPC; ++k) acc += rd[k * 32];
    }
  }
  if (w == 0) {
    double lp = lnscale + E2 + log(acc) - (double)(L - 1) * (0.5 * (double)D) * XT_LN_2PI;
    if (valid) a.logp[ck.trk_off + t] = lp;
    if (!valid) lp = 0.0;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) lp += __shfl_down_sync(0xffffffffu, lp, off);
    if (lane == 0) a.partial[wi] = lp;
  }
}
