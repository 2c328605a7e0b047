// K2 (fast path, second generation) — fused merge + update replay kernel with per-sequence
// extended exponents.
//
// Same recursion as xt_replay.cuh (log domain) restated like a scaled forward algorithm:
//   * a sequence carries a linear weight W = Wm * 2^E (Wm in [1, 2) or 0, E a 32-bit integer
//     kept next to the moments), so no per-track scale has to be agreed between warps: one CTA
//     barrier per step, no log/exp pair per merge, exact for any dynamic range;
//   * the Gaussian-product update (tracking.py:87-98) costs one reciprocal and one
//     range-reduced exp polynomial per parent: exp(e) = p(r) * 2^k and k goes to the exponent;
//   * the merge of fuse_tracks_th (tracking.py:723-741) needs no transcendental: the members'
//     weights are the merge weights;
//   * merge and the following update are fused: a warp merges a group from the previous step's
//     parents and immediately updates it with the next localisation, so the state makes one
//     shared-memory round trip per step (ping-pong buffers);
//   * the plan kernel emits one contiguous replay record per step (XtBlobHdr, xt_common.cuh)
//     with the groups already assigned to the replay warps (sorted by size, round-robin);
//     the CTA stages the record of the next step in shared memory while it computes the
//     current one, so the inner loops never wait on global memory.
//
// Mapping: a CTA of WPC warps owns a tile of 32*TPT tracks; a thread carries TPT tracks (lane,
// lane + 32, ...), which gives every instruction stream TPT independent dependency chains and
// amortises the warp-uniform work (record decoding, addresses, loop control) over TPT tracks.
// Warp w processes the groups woff[w]..woff[w+1]-1 of the step's record.  State slot =
// (m[D], u[KS], Wm) as 16-byte vectors [slot][vector][track] (conflict-free 128-bit
// shared-memory accesses) + [slot][track] exponents.
#pragma once
#include "xt_common.cuh"
#include "xt_replay.cuh"
#include "xt_replay_lin.cuh"

#define XT_ZERO_EXP (-(1 << 30))  // exponent of a sequence with zero weight

// build-time variants (A/B timing on the GPU box: python -m extrack_b200.build -DXT_K2_...=0 --out=...)
#ifndef XT_K2_SHORT_CHAIN
#define XT_K2_SHORT_CHAIN 1  // Estrin polynomial + cubic reciprocal step in the update (shorter dependent chains)
#endif
#ifndef XT_K2_SINGLES_X2
#define XT_K2_SINGLES_X2 1  // single-member groups two at a time (one track per thread)
#endif
#ifndef XT_K2_EXP_DEG
#define XT_K2_EXP_DEG 7  // degree of the exp polynomial of the short-chain evaluation (6 | 7)
#endif
#ifndef XT_K2_SINGLES_IL
#define XT_K2_SINGLES_IL 1  // ... with the two updates interleaved stage by stage
#endif

struct K2Tab {  // per-evaluation tables and scalars of the fused replay kernel (built on the host)
  double tau0[XT_MAX_HEADS];      // exp(LT[head])
  double tau1[XT_MAX_HEADS];      // exp(LT[head] + Lp_stay[r])
  double dd[XT_MAX_HEADS];        // xt_params::dd
  double winit[XT_MAX_HEADS];     // exp(LT + LF)
  double leave[XT_MAX_STATES];    // sum_r exp(L_leave[r + K*state])
  double l2[XT_MAX_DIMS];
  double e2[16];                  // 2^(j/16)
  int32_t nS, nsub, K, min_len;
  uint32_t flags;                 // xt_params::flags
  double loc_slope, loc_offset;   // xt_params
  // constants of xt_exp_split: read as constant-bank operands of the FP64 instructions (64-bit
  // immediates would be rebuilt with two moves per use inside the register-bound inner loop)
  double kc[6];
  // ln of the constant c folded into tau0 / tau1 / winit (c = prod over dims of sqrt(l2), scalar-LocErr models): every
  // Gaussian-product update multiplies the weight by prod_dim q^-1/2 >= c^-1, so with c folded in a weight never grows by
  // more than 2x per step and drifts down slowly; the end of the track subtracts (L - 1) * lnc again
  double lnc;
};
#define XT_EXP_CONSTS {23.083120654223414, 6755399441055744.0, -0.04332169878499658, -1.4494042586539372e-18, \
                       4.1666666666666664e-02, 1.6666666666666666e-01}

// global -> shared staging without registers (the issuing thread waits, the CTA barrier publishes)
__device__ __forceinline__ void xt_cp_async16(unsigned dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void xt_cp_async_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// State memory of the replay kernel: shared memory (32-bit shared-window address `a`) or, in the GST
// instantiation (live sequences of a tile do not fit in shared memory), a per-CTA block of global
// memory at `gb` (`a` is then a byte offset).  Plain loads / stores: the CTA barrier between the steps
// orders them.
template <bool GST> __device__ __forceinline__ void xs_ld128(const char* gb, unsigned a, double& x, double& y) {
  if (GST) { const double2 v = *reinterpret_cast<const double2*>(gb + a); x = v.x; y = v.y; } else xt_lds128(a, x, y);
}
template <bool GST> __device__ __forceinline__ double xs_ld64(const char* gb, unsigned a) {
  return GST ? *reinterpret_cast<const double*>(gb + a) : xt_lds64(a);
}
template <bool GST> __device__ __forceinline__ int xs_ld32(const char* gb, unsigned a) {
  return GST ? *reinterpret_cast<const int*>(gb + a) : xt_lds32(a);
}
template <bool GST> __device__ __forceinline__ void xs_st128(char* gb, unsigned a, double x, double y) {
  if (GST) *reinterpret_cast<double2*>(gb + a) = make_double2(x, y); else xt_sts128(a, x, y);
}
template <bool GST> __device__ __forceinline__ void xs_st64(char* gb, unsigned a, double x) {
  if (GST) *reinterpret_cast<double*>(gb + a) = x; else xt_sts64(a, x);
}
template <bool GST> __device__ __forceinline__ void xs_st32(char* gb, unsigned a, int x) {
  if (GST) *reinterpret_cast<int*>(gb + a) = x; else xt_sts32(a, x);
}

// 2^d for -1022 <= d <= 0, +0.0 below (d never exceeds 0 here: exponents relative to a maximum)
__device__ __forceinline__ double xt_pow2_le0(int d) {
  return __hiloint2double(max(d + 1023, 0) << 20, 0);
}

// x <= 0 -> max(x, about -1e6): unsigned min on the high word (the sign bit is set, so a larger
// magnitude is a larger unsigned word); one integer instruction instead of a NaN-aware DSETP/FSEL
// sequence.  Keeps k = round(x / ln 2) far inside the int range of the exponent bookkeeping.
__device__ __forceinline__ double xt_clamp_neg(double x) {
  return __hiloint2double((int)min((unsigned)__double2hiint(x), 0xC12E8480u), __double2loint(x));
}

// exp(x) = p * 2^k for -1e6 <= x <= 0 with p in [0.97, 2): x * 16/ln2 = 16 k + j + rho,
// exp(x) = 2^k * 2^(j/16) * exp(r), |r| <= ln2/32, Taylor degree 7 (truncation 1.2e-18);
// 2^(j/16) from a 16-entry shared-memory table (conflict-free: 16 distinct 8-byte words).
__device__ __forceinline__ double xt_exp_split(double x, unsigned s_e2, int& k, const K2Tab& T) {
  double t = fma(x, T.kc[0], T.kc[1]);
  const int n = __double2loint(t);
  t -= T.kc[1];
  double r = fma(t, T.kc[2], x);
  r = fma(t, T.kc[3], r);
  const double tj = xt_lds64(s_e2 + ((n & 15) << 3));
  k = n >> 4;
  // the three highest coefficients are truncated to their high word (immediate operands):
  // their terms are below 4e-11, so 21 significant bits keep the error under 2e-17
#if XT_K2_SHORT_CHAIN
  // Estrin evaluation: dependency depth 3 instead of 7 (the replay kernel is bound by the latency of its
  // dependent FP64 chains, not by the instruction count)
  const double r2 = r * r;
  const double a0 = r + 1.0;
  const double a1 = fma(r, T.kc[5], 0.5);
  const double a2 = fma(r, 0.00833333283662796, T.kc[4]);
  const double r4 = r2 * r2;
  const double b0 = fma(r2, a1, a0);
#if XT_K2_EXP_DEG == 6
  // degree 6: |r| <= ln2/32, so the dropped r^7/5040 is below 4.4e-16 relative at the ends of the interval (5e-17 rms),
  // and the leaf with two constants (two moves to build the second one in registers at every use) disappears
  const double b1 = fma(r2, 0.00138888880610466, a2);
#else
  const double a3 = fma(r, 0.00019841268658638, 0.00138888880610466);
  const double b1 = fma(r2, a3, a2);
#endif
  return fma(r4, b1, b0) * tj;
#else
  double p = 0.00019841268658638;
  p = fma(p, r, 0.00138888880610466);
  p = fma(p, r, 0.00833333283662796);
  p = fma(p, r, T.kc[4]);
  p = fma(p, r, T.kc[5]);
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  return p * tj;
#endif
}

// 1/x for normal positive x: hardware seed (relative error e <= 2^-20) and one cubic step r0 (1 + e + e^2), error e^3:
// three dependent operations instead of the four of two Newton steps
__device__ __forceinline__ double xt_rcp_fused(double x) {
#if XT_K2_SHORT_CHAIN
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  const double e = fma(-x, r, 1.0);
  const double t = fma(e, e, e);
  return fma(r, t, r);
#else
  return xt_rcp(x);
#endif
}

// v >= 0 (normal or zero) -> mantissa in [1, 2) and exponent base + unbiased exponent of v;
// zero -> (0, XT_ZERO_EXP)
__device__ __forceinline__ void xt_split_exponent(double v, int base, double& mant, int& E) {
  const int hi = __double2hiint(v);
  const int be = (hi >> 20) & 0x7ff;
  mant = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(v));
  E = base + be - 1023;
  if (be == 0) {
    mant = 0.0;
    E = XT_ZERO_EXP;
  }
}

template <int D, int KS>
struct XtSeq {
  double m[D];
  double u[KS];
  double W;
  int E;
};

// State slot = NV 16-byte vectors [vector][track] (components m[D], u[KS], W in this order) at
// va + i*VS, exponents at ea; both addresses already include the lane offset.  Track j of the
// thread sits 32*j lanes further.
template <int D, int KS, int TPT, bool GST = false>
struct XtSlotIO {
  static constexpr int CO = D + KS + 1;
  static constexpr int NV = (CO + 1) / 2;
  static constexpr int VS = 32 * TPT * 16;   // bytes per vector row
  static constexpr int SLOTB = NV * VS;      // bytes per slot
  static constexpr int ESLOT = 32 * TPT * 4; // exponent bytes per slot
  static __device__ __forceinline__ void load(const char* gb, unsigned va, unsigned ea, XtSeq<D, KS> (&s)[TPT]) {
#pragma unroll
    for (int j = 0; j < TPT; ++j) {
      double c[2 * NV];
#pragma unroll
      for (int i = 0; i < NV; ++i) xs_ld128<GST>(gb, va + i * VS + j * 512, c[2 * i], c[2 * i + 1]);
#pragma unroll
      for (int i = 0; i < D; ++i) s[j].m[i] = c[i];
#pragma unroll
      for (int i = 0; i < KS; ++i) s[j].u[i] = c[D + i];
      s[j].W = c[D + KS];
      s[j].E = xs_ld32<GST>(gb, ea + j * 128);
    }
  }
  static __device__ __forceinline__ void store(char* gb, unsigned va, unsigned ea, const XtSeq<D, KS> (&s)[TPT]) {
#pragma unroll
    for (int j = 0; j < TPT; ++j) {
      double c[2 * NV];
#pragma unroll
      for (int i = 0; i < D; ++i) c[i] = s[j].m[i];
#pragma unroll
      for (int i = 0; i < KS; ++i) c[D + i] = s[j].u[i];
      c[D + KS] = s[j].W;
      if (2 * NV > CO) c[CO] = 0.0;
#pragma unroll
      for (int i = 0; i < NV; ++i) xs_st128<GST>(gb, va + i * VS + j * 512, c[2 * i], c[2 * i + 1]);
      xs_st32<GST>(gb, ea + j * 128, s[j].E);
    }
  }
};

// Gaussian-product update of merged sequences (m, s2, W * 2^E) with the localisations cl (one
// per track of the thread, written stage by stage so that the TPT chains interleave); the
// result is the parent record (m', u = l2*s2/q, W' * 2^E') shared by its children.
// (VAR: l2 is per track, l2[j][k]; otherwise one row shared by the thread's tracks)
template <int D, int KS, int TPT, bool VAR = false>
__device__ __forceinline__ void xt_update(XtSeq<D, KS> (&s)[TPT], const double (&cl)[TPT][D],
                                          const double (&l2)[VAR ? TPT : 1][KS], unsigned s_e2, const K2Tab& T) {
  double rq[TPT][KS], e[TPT];
#pragma unroll
  for (int j = 0; j < TPT; ++j)
#pragma unroll
    for (int k = 0; k < KS; ++k) rq[j][k] = xt_rcp_fused(l2[VAR ? j : 0][k] + s[j].u[k]);
#pragma unroll
  for (int j = 0; j < TPT; ++j) {
    double g[KS];
#pragma unroll
    for (int k = 0; k < KS; ++k) {
      g[k] = s[j].u[k] * rq[j][k];
      s[j].u[k] = l2[VAR ? j : 0][k] * g[k];
    }
    if (KS == 1) {
      double q2 = 0.0;
#pragma unroll
      for (int dim = 0; dim < D; ++dim) {
        const double df = cl[j][dim] - s[j].m[dim];
        s[j].m[dim] = fma(df, g[0], s[j].m[dim]);
        q2 = (dim == 0) ? df * df : fma(df, df, q2);
      }
      e[j] = q2 * (-0.5 * rq[j][0]);
    } else {
      double quad = 0.0;
#pragma unroll
      for (int dim = 0; dim < D; ++dim) {
        const double df = cl[j][dim] - s[j].m[dim];
        s[j].m[dim] = fma(df, g[dim], s[j].m[dim]);
        quad = fma(df * df, rq[j][dim], quad);
      }
      e[j] = -0.5 * quad;
    }
  }
  double p[TPT];
  int k2[TPT];
#pragma unroll
  for (int j = 0; j < TPT; ++j) p[j] = xt_exp_split(xt_clamp_neg(e[j]), s_e2, k2[j], T);
  // The weight stays an ordinary double W * 2^E; it is brought back to a mantissa in [1, 2) only when it leaves
  // [2^-512, 2^512) (zero, NaN and negative values take the same branch).  With the per-update constant folded into
  // the tables (K2Tab::lnc) that takes hundreds of steps, so the branch is practically never taken; any member of a
  // later merge is then at most 2^-1022 / 2^-512 below the member it is aligned to before it is flushed.
#pragma unroll
  for (int j = 0; j < TPT; ++j) {
    const double wn = (s[j].W * xt_normfac<D, KS>(rq[j])) * p[j];
    s[j].E += k2[j];
    s[j].W = wn;
    if ((unsigned)(__double2hiint(wn) - 0x1FF00000) >= 0x40000000u) xt_split_exponent(wn, s[j].E, s[j].W, s[j].E);
  }
}

struct K2FArgs {
  const XtChunk* chunks;
  const XtWork* work;   // tiles of 32*TPT tracks
  const double* soa;
  XtPlanPtrs plan;
  double* logp;
  double* partial;
  const XtChunkSummary* summ;  // written by the plan kernel earlier in the same stream
  int* spec_fail;      // set when a chunk needs more parent slots than this launch was sized for
  int32_t Pcap;
  int32_t n_work;      // CTAs of this launch
  int32_t work0;       // first tile of this launch in the work table
  XtAux ax;            // VAR instantiation only (stay / leave tables are per chunk)
  // GST instantiation only: state blocks in global memory, one per resident CTA
  char* gstate;        // [n_sm * 32][gstride] bytes
  unsigned* gslots;    // [n_sm] bitmap of the blocks in use on every SM
  size_t gstride;
};

// shared memory of k2_replay_fused in bytes (host and device agree through these functions)
__host__ __device__ inline int xt_fused_blob16(int Pcap, int K) { return 2 + (Pcap + 1) / 2 + (K * Pcap + 3) / 4; }
// GST instantiation: one staged record + tables in shared memory, the state in global memory
__host__ __device__ inline size_t xt_fused_smem_gst(int Pcap, int K, int H) {
  return (size_t)xt_fused_blob16(Pcap, K) * 16 + (size_t)2 * H * 16 + 128 + 64;
}
__host__ __device__ inline size_t xt_fused_gstride(int D, int KS, int Pcap) {
  const int NV = (D + KS + 1 + 1) / 2;
  return (size_t)2 * Pcap * (NV * 512 + 128);
}
__host__ __device__ inline size_t xt_fused_smem(int D, int KS, int Pcap, int K, int H, int wpc, int tpt, bool var = false) {
  const int NV = (D + KS + 1 + 1) / 2;
  return (var ? 64 : 0)                            // VAR: per-chunk leave sums [nS]
         + (size_t)2 * Pcap * NV * 512 * tpt       // state vectors, ping-pong
         + (size_t)2 * xt_fused_blob16(Pcap, K) * 16  // staged replay records, ping-pong
         + (size_t)2 * H * 16                      // (tau, dd) per head, without / with the stay term
         + 128                                     // 2^(j/16)
         + (size_t)2 * Pcap * 128 * tpt;           // exponents, ping-pong
  // (the end-of-track partial sums, wpc * 384 * tpt bytes, reuse the idle state buffer: Pcap >= 2)
}

template <int D, int KS, int WPC, int TPT, bool VAR = false, bool GST = false>
#ifndef XT_K2_WARPS_T1
#define XT_K2_WARPS_T1 24  // resident warps per SM the register budget is set for (one track per thread)
#endif
__global__ void __launch_bounds__(32 * WPC, (TPT == 1 ? XT_K2_WARPS_T1 : 16) / WPC) k2_replay_fused(const K2FArgs a, const __grid_constant__ K2Tab T) {
  using IO = XtSlotIO<D, KS, TPT, GST>;
  using Seq = XtSeq<D, KS>;
  constexpr int SLOTB = IO::SLOTB, ESLOT = IO::ESLOT;
  constexpr int NT = 32 * WPC;
  static_assert(!GST || (TPT == 1 && !VAR), "the global-state instantiation is built for one track per thread, scalar inputs");
  const int tid = threadIdx.x;
  const int lane = tid & 31, w = tid >> 5;
  const int wi = (int)blockIdx.x + a.work0;
  const XtWork wk = a.work[wi];
  const XtChunk ck = a.chunks[wk.chunk];
  {  // the launch may have been sized speculatively (parent slots of the previous evaluation)
    const XtChunkSummary sm = a.summ[wk.chunk];
    if (sm.err != 0 || sm.max_nP > a.Pcap) {
      if (tid == 0) {
        atomicExch(a.spec_fail, 1);
        a.partial[wi] = 0.0;
      }
      return;
    }
  }
  const int nS = T.nS, K = T.K, H = K * nS;
  const size_t npad = (size_t)ck.nTpad;
  const double* Cs = a.soa + ck.xyz_off;
  int toff[TPT];
  bool valid[TPT];
#pragma unroll
  for (int j = 0; j < TPT; ++j) {
    const int t = wk.t0 + lane + 32 * j;
    valid[j] = t < ck.nT;
    toff[j] = valid[j] ? t : ck.nT - 1;
  }
  const size_t cstride = (size_t)D * npad;
  const int L = ck.L;
  const int Pcap = a.Pcap;
  const int B16 = xt_fused_blob16(Pcap, K);

  extern __shared__ double2 k2f_smem[];
  const unsigned sb = xt_smem_base(k2f_smem);
  const unsigned VB = (unsigned)Pcap * SLOTB, EB = (unsigned)Pcap * ESLOT;
  // shared memory: state vectors [2][Pcap][NV][32*TPT] x 16 B, staged records [2][B16] x 16 B, tables,
  // exponents [2][Pcap][32*TPT] x 4 B.  GST: the state (vectors, then exponents) is a block of global
  // memory and s_vec / s_exp are byte offsets into it; shared memory holds one record and the tables.
  const unsigned s_vec = GST ? lane * 16 : sb + lane * 16;
  const unsigned s_blob = GST ? sb : sb + 2 * VB;
  const unsigned s_tab = s_blob + (GST ? 1 : 2) * B16 * 16;  // [2][H] x 16 B: (tau, dd)
  const unsigned s_e2 = s_tab + 2 * H * 16;                   // [16] x 8 B
  const unsigned s_exp = GST ? 2 * VB + lane * 4 : s_e2 + 128 + lane * 4;
  const unsigned s_leave = s_e2 + 128 + (GST ? 0 : 2 * EB);   // VAR: [nS] x 8 B per-chunk leave sums
  // GST: take one of this SM's state blocks (at most 32 CTAs are resident on an SM, so a bit is free)
  char* gb = nullptr;
  __shared__ int s_slot;
  unsigned smid = 0;
  if (GST) {
    asm("mov.u32 %0, %%smid;" : "=r"(smid));
    if (tid == 0) {
      int b = 0;
      for (;;) {
        const unsigned old = atomicOr(&a.gslots[smid], 1u << b);
        if (!(old & (1u << b))) break;
        b = (b + 1) & 31;
      }
      s_slot = b;
    }
    __syncthreads();
    gb = a.gstate + ((size_t)smid * 32 + s_slot) * a.gstride;
  }

  // stage the first replay record (steps 3..L-1 use records 0..L-4) and the tables; the records are
  // staged by warp 0 alone with cp.async (no registers, no instructions in the other warps)
  const int bstride = xt_blob_stride16(a.plan.cap);
  const uint4* gnext = a.plan.blob + (size_t)ck.rec0 * bstride + lane;
  const int nrec = ck.nrec;
  if (!GST && w == 0 && nrec > 0) {
    for (int i = lane; i < B16; i += 32) xt_cp_async16(s_blob + i * 16, gnext + (i - lane));
    gnext += bstride;
  }
  for (int h = tid; h < 2 * H; h += NT) {
    const int hh = h < H ? h : h - H;
    double tau = h < H ? T.tau0[hh] : T.tau1[hh];
    if (VAR && a.ax.stay && h >= H) tau = T.tau0[hh] * exp(a.ax.stay[(size_t)wk.chunk * K + hh % K]);  // per-chunk p_stay
    xt_sts128(s_tab + h * 16, tau, T.dd[hh]);
  }
  if (tid < 16) xt_sts64(s_e2 + tid * 8, T.e2[tid]);
  if (VAR && tid < nS) xt_sts64(s_leave + tid * 8, a.ax.leave ? a.ax.leave[(size_t)wk.chunk * nS + tid] : T.leave[tid]);

  // l2 of the localisation the next update consumes (VAR: per track, from the aux block; the
  // dd column of the tables then holds dd per unit time and is scaled by the track's dt)
  double l2[VAR ? TPT : 1][KS];
#pragma unroll
  for (int k = 0; k < KS; ++k) l2[0][k] = T.l2[k];
  const bool var_loc = VAR && (T.flags & XT_FLAG_VAR_LOC), var_dt = VAR && (T.flags & XT_FLAG_VAR_DT);
  const double* Ax = VAR ? a.ax.aux + (size_t)(ck.xyz_off / D) * a.ax.R : nullptr;
  const size_t astride = VAR ? (size_t)a.ax.R * npad : 0;
  double l2n[VAR ? TPT : 1][KS], dtc[VAR ? TPT : 1], dtq[VAR ? TPT : 1], dtn[VAR ? TPT : 1];
  auto aux_row = [&](double (&lo)[VAR ? TPT : 1][KS], double (&dto)[VAR ? TPT : 1]) {  // row at Ax
#pragma unroll
    for (int j = 0; j < (VAR ? TPT : 1); ++j) {
#pragma unroll
      for (int k = 0; k < KS; ++k) lo[j][k] = var_loc ? xt_sigma2f(T.flags, T.loc_slope, T.loc_offset, Ax[(size_t)k * npad + toff[j]]) : T.l2[k];
      dto[j] = var_dt ? Ax[(size_t)a.ax.ka * npad + toff[j]] : 1.0;
    }
  };
#define DTQ(j) (VAR ? dtq[VAR ? (j) : 0] : 1.0)
  if (VAR) {
    aux_row(l2, dtq);    // row 0: first localisation
    Ax += astride;
    aux_row(l2n, dtc);   // row 1 (L >= 2)
  }

  double cl[TPT][D], cn[TPT][D];
  double csum[TPT];  // NaN / Inf coordinates anywhere in the track poison the result
#pragma unroll
  for (int j = 0; j < TPT; ++j) {
#pragma unroll
    for (int dim = 0; dim < D; ++dim) cl[j][dim] = Cs[(size_t)dim * npad + toff[j]];  // C[0]
  }
  Cs += cstride;
#pragma unroll
  for (int j = 0; j < TPT; ++j) {
    csum[j] = 0.0;
#pragma unroll
    for (int dim = 0; dim < D; ++dim) {
      cn[j][dim] = Cs[(size_t)dim * npad + toff[j]];  // C[1] (L >= 2)
      csum[j] += cl[j][dim];
    }
  }
  __syncthreads();  // tables visible (the update below reads 2^(j/16))

  // ---- first localisation (tracking.py:478-529) and, if L >= 3, the update of step 2 ----
  int nP = H;
  for (int c = w; c < nP; c += WPC) {
    Seq s[TPT];
    const double ddc = T.dd[c];
#pragma unroll
    for (int j = 0; j < TPT; ++j) {
#pragma unroll
      for (int dim = 0; dim < D; ++dim) s[j].m[dim] = cl[j][dim];
#pragma unroll
      for (int k = 0; k < KS; ++k) s[j].u[k] = l2[VAR ? j : 0][k] + (VAR ? ddc * DTQ(j) : ddc);
      xt_split_exponent(T.winit[c], 0, s[j].W, s[j].E);
    }
    if (L >= 3) xt_update<D, KS, TPT, VAR>(s, cn, VAR ? l2n : l2, s_e2, T);
    IO::store(gb, s_vec + c * SLOTB, s_exp + c * ESLOT, s);
  }
  if (VAR && L >= 3) {  // dtc = dt of localisation 1, l2n/dtn = row 2
    Ax += astride;
    aux_row(l2n, dtn);
  }
  if (L >= 3) {
    Cs += cstride;
#pragma unroll
    for (int j = 0; j < TPT; ++j)
#pragma unroll
      for (int dim = 0; dim < D; ++dim) {
        csum[j] += cn[j][dim];                           // C[1]
        cn[j][dim] = Cs[(size_t)dim * npad + toff[j]];  // C[2]
      }
  }
  if (!GST && w == 0) xt_cp_async_wait();
  __syncthreads();

  // ---- steps 3..L-1: merge by the replay record of step-1, update with C[step-1] ----
  unsigned src_v = s_vec, src_e = s_exp, dst_v = s_vec + VB, dst_e = s_exp + EB;
  for (int step = 3; step <= L - 1; ++step) {
    const int ri = step - 3;
    // prefetch: localisation and replay record of the next step
    Cs += cstride;  // C[step] exists (step <= L-1)
#pragma unroll
    for (int j = 0; j < TPT; ++j)
#pragma unroll
      for (int dim = 0; dim < D; ++dim) {
        cl[j][dim] = cn[j][dim];
        csum[j] += cl[j][dim];
        cn[j][dim] = Cs[(size_t)dim * npad + toff[j]];
      }
    if (VAR) {  // dtq = dt of C[step-2] (children being merged), l2 = of C[step-1] (update), next row prefetched
#pragma unroll
      for (int j = 0; j < (VAR ? TPT : 1); ++j) {
        dtq[j] = dtc[j];
        dtc[j] = dtn[j];
#pragma unroll
        for (int k = 0; k < KS; ++k) l2[j][k] = l2n[j][k];
      }
      Ax += astride;
      aux_row(l2n, dtn);
    }
    const bool more = ri + 1 < nrec;
    if (GST) {  // stage this step's record (any size) with the whole CTA
      const uint4* gr = a.plan.blob + (size_t)(ck.rec0 + ri) * bstride;
      const int n16 = (int)(__ldg(gr).y & 0xFFFFu);  // XtBlobHdr::n16
      for (int i = tid; i < n16; i += NT) xt_sts128u(s_blob + i * 16, __ldg(gr + i));
      __syncthreads();
    } else if (w == 0 && more) {  // record of the next step into the other buffer
      const unsigned nb = s_blob + ((ri + 1) & 1) * (B16 * 16);
      for (int i = lane; i < B16; i += 32) xt_cp_async16(nb + i * 16, gnext + (i - lane));
      gnext += bstride;
    }

    const unsigned rb = s_blob + (GST ? 0 : (ri & 1) * (B16 * 16));
    const int nG = (int)xt_lds16(rb);  // XtBlobHdr::nG
    const unsigned grec = rb + 32;
    const unsigned entb = grec + ((nG + 1) >> 1) * 16;
    const unsigned tab = s_tab + (((step - 1) >= T.min_len) ? H * 16 : 0);
    // group records of this warp: addresses walk [ga, ge); the loop-invariant bases are pinned in
    // registers (the compiler otherwise rematerialises them from uniform registers every iteration)
    unsigned ga = grec + xt_lds16(rb + 8 + 2 * w) * 8;  // XtBlobHdr::woff
    const unsigned ge = grec + xt_lds16(rb + 8 + 2 * (w + 1)) * 8;
    // the first groups of this warp's list have several members, the others exactly one
    int nmw;
    if (WPC <= 4) {
      nmw = (int)xt_lds16(rb + 8 + 2 * (5 + w));  // XtBlobHdr::woff[5 + w]
    } else {
      const int nM = (int)xt_lds16(rb + 6);  // XtBlobHdr::nM, round-robin schedule
      nmw = nM > w ? (nM - w + WPC - 1) / WPC : 0;
    }
    const unsigned gm = ga + (unsigned)nmw * 8;
    unsigned tabp = tab;
    asm volatile("mov.u32 %0, %0;" : "+r"(tabp));
    for (; ga < gm; ga += 8) {
      const uint2 gr = xt_lds64u(ga);
      const unsigned p0o = gr.x & 0x7FF80u;  // p0 * 128
      const unsigned g = gr.x >> 19;
      const unsigned kind = gr.y >> 30;
      Seq G[TPT];
      IO::load(gb, src_v + p0o * (SLOTB / 128), src_e + p0o * (ESLOT / 128), G);
      double tau0, dd0;
      xt_lds128(tabp + (gr.x & 0x7Fu) * 16, tau0, dd0);
      if (kind == 2u) {
        const unsigned p1o = gr.y & 0x7FF80u;
        Seq B[TPT];
        IO::load(gb, src_v + p1o * (SLOTB / 128), src_e + p1o * (ESLOT / 128), B);
        double tau1, dd1;
        xt_lds128(tabp + (gr.y & 0x7Fu) * 16, tau1, dd1);
#pragma unroll
        for (int j = 0; j < TPT; ++j) {
          const int Eg = max(G[j].E, B[j].E);
          const double w0 = (G[j].W * xt_pow2_le0(G[j].E - Eg)) * tau0;
          const double w1 = (B[j].W * xt_pow2_le0(B[j].E - Eg)) * tau1;
          const double sw = w0 + w1;
          const double rs = (sw > 1e-300) ? xt_rcp(sw) : 0.0;
          const double lam = w1 * rs;
#pragma unroll
          for (int dim = 0; dim < D; ++dim) G[j].m[dim] = fma(B[j].m[dim] - G[j].m[dim], lam, G[j].m[dim]);
#pragma unroll
          for (int k = 0; k < KS; ++k) {
            const double ua = G[j].u[k] + (VAR ? dd0 * DTQ(j) : dd0);
            G[j].u[k] = fma((B[j].u[k] + (VAR ? dd1 * DTQ(j) : dd1)) - ua, lam, ua);
          }
          G[j].W = sw;
          G[j].E = Eg;
        }
      } else {
        // member list: two passes (largest exponent, then the weighted sums)
        const unsigned o = gr.y & 0xFFFu;
        const int n = (int)((gr.y >> 12) & 0x1FFFu);
        const unsigned eb = entb + o * 4;
        int Eg[TPT];
#pragma unroll
        for (int j = 0; j < TPT; ++j) Eg[j] = G[j].E;
#pragma unroll 4
        for (int k = 1; k < n; ++k) {
          const unsigned ea = src_e + ((unsigned)xt_lds32(eb + k * 4) & 0xFFFFu) * ESLOT;
#pragma unroll
          for (int j = 0; j < TPT; ++j) Eg[j] = max(Eg[j], xs_ld32<GST>(gb, ea + j * 128));
        }
        double sw[TPT], am[TPT][D], as[TPT][KS];
#pragma unroll
        for (int j = 0; j < TPT; ++j) {
          sw[j] = (G[j].W * xt_pow2_le0(G[j].E - Eg[j])) * tau0;
#pragma unroll
          for (int dim = 0; dim < D; ++dim) am[j][dim] = sw[j] * G[j].m[dim];
#pragma unroll
          for (int k = 0; k < KS; ++k) {
            G[j].u[k] += VAR ? dd0 * DTQ(j) : dd0;
            as[j][k] = sw[j] * G[j].u[k];
          }
        }
#pragma unroll 2
        for (int k = 1; k < n; ++k) {
          const unsigned e = (unsigned)xt_lds32(eb + k * 4);
          const unsigned pm = e & 0xFFFFu;
          Seq M[TPT];
          IO::load(gb, src_v + pm * SLOTB, src_e + pm * ESLOT, M);
          double taum, ddm;
          xt_lds128(tabp + ((e >> 16) & 0xFFu) * 16, taum, ddm);
#pragma unroll
          for (int j = 0; j < TPT; ++j) {
            const double wj = (M[j].W * xt_pow2_le0(M[j].E - Eg[j])) * taum;
            sw[j] += wj;
#pragma unroll
            for (int dim = 0; dim < D; ++dim) am[j][dim] = fma(wj, M[j].m[dim], am[j][dim]);
#pragma unroll
            for (int k2 = 0; k2 < KS; ++k2) as[j][k2] = fma(wj, M[j].u[k2] + (VAR ? ddm * DTQ(j) : ddm), as[j][k2]);
          }
        }
#pragma unroll
        for (int j = 0; j < TPT; ++j) {
          if (sw[j] > 1e-300) {  // otherwise: zero-weight group, keep the first member's moments
            const double rs = xt_rcp(sw[j]);
#pragma unroll
            for (int dim = 0; dim < D; ++dim) G[j].m[dim] = am[j][dim] * rs;
#pragma unroll
            for (int k = 0; k < KS; ++k) G[j].u[k] = as[j][k] * rs;
          }
          G[j].W = sw[j];
          G[j].E = Eg[j];
        }
      }
      xt_update<D, KS, TPT, VAR>(G, cl, l2, s_e2, T);
      IO::store(gb, dst_v + g * SLOTB, dst_e + g * ESLOT, G);
    }
    // single-member groups (three quarters of the groups of a 2-state model): the child itself, no dispatch; two
    // at a time when a thread carries one track (two independent dependency chains per instruction stream)
    if (TPT == 1 && XT_K2_SINGLES_X2) {
      for (; ga + 8 < ge; ga += 16) {
        const uint2 gr0 = xt_lds64u(ga), gr1 = xt_lds64u(ga + 8);
        const unsigned p0o = gr0.x & 0x7FF80u, p1o = gr1.x & 0x7FF80u;
        Seq G0[TPT], G1[TPT];
        IO::load(gb, src_v + p0o * (SLOTB / 128), src_e + p0o * (ESLOT / 128), G0);
        IO::load(gb, src_v + p1o * (SLOTB / 128), src_e + p1o * (ESLOT / 128), G1);
        double tau0, dd0, tau1, dd1;
        xt_lds128(tabp + (gr0.x & 0x7Fu) * 16, tau0, dd0);
        xt_lds128(tabp + (gr1.x & 0x7Fu) * 16, tau1, dd1);
#pragma unroll
        for (int j = 0; j < TPT; ++j) {
          G0[j].W *= tau0;
          G1[j].W *= tau1;
#pragma unroll
          for (int k = 0; k < KS; ++k) {
            G0[j].u[k] += VAR ? dd0 * DTQ(j) : dd0;
            G1[j].u[k] += VAR ? dd1 * DTQ(j) : dd1;
          }
        }
        if (XT_K2_SINGLES_IL && !VAR) {  // both updates stage by stage in one instruction stream (two interleaved chains)
          Seq P2[2] = {G0[0], G1[0]};
          double cl2[2][D];
#pragma unroll
          for (int dim = 0; dim < D; ++dim) cl2[0][dim] = cl2[1][dim] = cl[0][dim];
          double l2s[1][KS];
#pragma unroll
          for (int k = 0; k < KS; ++k) l2s[0][k] = l2[0][k];
          xt_update<D, KS, 2, false>(P2, cl2, l2s, s_e2, T);
          G0[0] = P2[0];
          G1[0] = P2[1];
        } else {
          xt_update<D, KS, TPT, VAR>(G0, cl, l2, s_e2, T);
          xt_update<D, KS, TPT, VAR>(G1, cl, l2, s_e2, T);
        }
        IO::store(gb, dst_v + (gr0.x >> 19) * SLOTB, dst_e + (gr0.x >> 19) * ESLOT, G0);
        IO::store(gb, dst_v + (gr1.x >> 19) * SLOTB, dst_e + (gr1.x >> 19) * ESLOT, G1);
      }
    }
    for (; ga < ge; ga += 8) {
      const uint2 gr = xt_lds64u(ga);
      const unsigned p0o = gr.x & 0x7FF80u;  // p0 * 128
      const unsigned g = gr.x >> 19;
      Seq G[TPT];
      IO::load(gb, src_v + p0o * (SLOTB / 128), src_e + p0o * (ESLOT / 128), G);
      double tau0, dd0;
      xt_lds128(tabp + (gr.x & 0x7Fu) * 16, tau0, dd0);
#pragma unroll
      for (int j = 0; j < TPT; ++j) {
        G[j].W *= tau0;
#pragma unroll
        for (int k = 0; k < KS; ++k) G[j].u[k] += VAR ? dd0 * DTQ(j) : dd0;
      }
      xt_update<D, KS, TPT, VAR>(G, cl, l2, s_e2, T);
      IO::store(gb, dst_v + g * SLOTB, dst_e + g * ESLOT, G);
    }
    nP = nG;
    {  // swap the ping-pong buffers
      unsigned tv = src_v; src_v = dst_v; dst_v = tv;
      unsigned te = src_e; src_e = dst_e; dst_e = te;
    }
    if (!GST && w == 0) xt_cp_async_wait();
    __syncthreads();
  }
  const uint8_t* curP = nullptr;
  if (nrec > 0) curP = a.plan.curG + (size_t)(ck.rec0 + nrec - 1) * a.plan.cap;

  // ---- end of track (tracking.py:613-639, :781-786): last localisation, leave term,
  //      sum over the surviving sequences in extended-exponent arithmetic ----
  const bool implicit = L >= 3;  // slots hold un-fused parents (m', u, W'): children are read on the fly
  const unsigned tab = s_tab + (((L - 1) >= T.min_len) ? H * 16 : 0);
  const int Kc = implicit ? K : 1;
  double acc[TPT];
  int KA[TPT];
#pragma unroll
  for (int j = 0; j < TPT; ++j) {
    acc[j] = 0.0;
    KA[j] = XT_ZERO_EXP;
#pragma unroll
    for (int dim = 0; dim < D; ++dim) csum[j] += cn[j][dim];  // C[L-1]
  }
  for (int p = w; p < nP; p += WPC) {
    Seq S[TPT];
    IO::load(gb, src_v + p * SLOTB, src_e + p * ESLOT, S);
    const int ps = curP ? (int)__ldg(&curP[p]) : (p % nS);
    double df2[TPT][D];
#pragma unroll
    for (int j = 0; j < TPT; ++j)
#pragma unroll
      for (int dim = 0; dim < D; ++dim) {
        const double df = cn[j][dim] - S[j].m[dim];
        df2[j][dim] = df * df;
      }
    int newest_r = 0;  // r % nS, maintained incrementally
    for (int r = 0; r < Kc; ++r) {
      double dd = 0.0, th = 1.0;
      int newest = ps;
      if (implicit) {
        xt_lds128(tab + (r + K * ps) * 16, th, dd);
        newest = newest_r;
      }
      if (++newest_r == nS) newest_r = 0;
      if (ck.isBL) th *= VAR ? xt_lds64(s_leave + newest * 8) : T.leave[newest];
#pragma unroll
      for (int j = 0; j < TPT; ++j) {
        double rq[KS];
#pragma unroll
        for (int k = 0; k < KS; ++k)
          rq[k] = xt_rcp(S[j].u[k] + (VAR ? dd * dtc[VAR ? j : 0] : dd) + (VAR ? l2n[VAR ? j : 0][k] : l2[0][k]));
        double quad = 0.0;
#pragma unroll
        for (int dim = 0; dim < D; ++dim) quad = fma(df2[j][dim], rq[(KS == 1) ? 0 : dim], quad);
        int k2;
        const double pe = xt_exp_split(xt_clamp_neg(-0.5 * quad), s_e2, k2, T);
        const double v = ((S[j].W * th) * xt_normfac<D, KS>(rq)) * pe;
        const int Kv = S[j].E + k2;
        const int Kn = max(KA[j], Kv);
        acc[j] = fma(acc[j], xt_pow2_le0(KA[j] - Kn), v * xt_pow2_le0(Kv - Kn));
        KA[j] = Kn;
      }
    }
  }
  // cross-warp combination through the idle state buffer (every warp is past the last barrier
  // and reads only the current buffer)
  const unsigned s_redA = dst_v - lane * 16;              // [WPC][TPT][32] x 8 B
  const unsigned s_redK = s_redA + WPC * 256 * TPT;       // [WPC][TPT][32] x 4 B
#pragma unroll
  for (int j = 0; j < TPT; ++j) {
    xs_st64<GST>(gb, s_redA + ((w * TPT + j) * 32 + lane) * 8, acc[j]);
    xs_st32<GST>(gb, s_redK + ((w * TPT + j) * 32 + lane) * 4, KA[j]);
  }
  __syncthreads();
  if (w == 0) {
    double lps = 0.0;
#pragma unroll
    for (int j = 0; j < TPT; ++j) {
      int Kn = xs_ld32<GST>(gb, s_redK + (j * 32 + lane) * 4);
#pragma unroll
      for (int k = 1; k < WPC; ++k) Kn = max(Kn, xs_ld32<GST>(gb, s_redK + ((k * TPT + j) * 32 + lane) * 4));
      double tot = 0.0;
#pragma unroll
      for (int k = 0; k < WPC; ++k)
        tot = fma(xs_ld64<GST>(gb, s_redA + ((k * TPT + j) * 32 + lane) * 8),
                  xt_pow2_le0(xs_ld32<GST>(gb, s_redK + ((k * TPT + j) * 32 + lane) * 4) - Kn), tot);
      double lp = XT_LN2 * (double)Kn + log(tot) - (double)(L - 1) * ((0.5 * (double)D) * XT_LN_2PI + T.lnc);
      if (!(fabs(csum[j]) <= 1.7976931348623157e308)) lp = __longlong_as_double(0x7ff8000000000000ll);
      if (valid[j]) {
        a.logp[ck.trk_off + wk.t0 + lane + 32 * j] = lp;
        lps += lp;
      }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) lps += __shfl_down_sync(0xffffffffu, lps, off);
    if (lane == 0) a.partial[wi] = lps;
  }
  if (GST) {  // give the state block back (after warp 0 has read the partial sums parked in it)
    __syncthreads();
    if (tid == 0) atomicAnd(&a.gslots[smid], ~(1u << s_slot));
  }
}
