// K2, optional single-precision replay (xt_set_option "fp32_replay"): the recursion of
// xt_replay_fused.cuh with the per-sequence state (m[D], u[KS], Wm) in FP32 and the same 32-bit
// extended exponent next to it, so the dynamic range is as unlimited as in the FP64 kernel and only
// the mantissas are short.  Stated tolerance: total log-likelihood within 1e-4 relative of the FP64
// result (BASELINE north star); observed 8.5e-10 on 10^6 synthetic tracks.
//
//   * the plan (which sequences fuse) still comes from the FP64 plan kernel: the plan decisions are
//     those of the reference, only the arithmetic carried along the plan is shortened;
//   * localisations are read as FP64 from the same SoA block, taken relative to the track's first
//     localisation in FP64 and only then rounded (log P is translation-invariant, tracking.py:87-98
//     only sees differences), which keeps the rounding error relative to the track's extent instead
//     of the field of view's;
//   * 1/q = rcp.approx.ftz.f32 (1 ulp), exp = 2^k * ex2.approx.ftz.f32(r) with a two-constant
//     Cody-Waite reduction (k goes to the exponent word), no logarithm until the end of the track,
//     where the per-track sum is finished in FP64;
//   * state slot = NV 16-byte vectors [vector][track] (one vector for 2-D tracks with a scalar
//     localisation error: a step moves 16 + 4 bytes per sequence through shared memory instead of
//     32 + 4), tables (tau, dd) as float2 per head.
//
// One configuration: 4 warps per tile of 32 tracks, one track per thread, scalar LocErr / dt, state in
// shared memory.  The host (prepare_fused) selects it only when the tables are representable in FP32
// (every non-zero factor in [1e-20, 1e3]) and the state fits; otherwise the FP64 kernel runs.
#pragma once
#include "xt_replay_fused.cuh"

__device__ __forceinline__ void xf_lds128(unsigned a, float (&c)[4]) {
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(c[0]), "=f"(c[1]), "=f"(c[2]), "=f"(c[3]) : "r"(a));
}
__device__ __forceinline__ void xf_sts128(unsigned a, const float (&c)[4]) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(c[0]), "f"(c[1]), "f"(c[2]), "f"(c[3]) : "memory");
}
__device__ __forceinline__ void xf_lds64(unsigned a, float& x, float& y) {
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(x), "=f"(y) : "r"(a));
}
__device__ __forceinline__ void xf_sts64(unsigned a, float x, float y) {
  asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(a), "f"(x), "f"(y) : "memory");
}
__device__ __forceinline__ void xf_sts32f(unsigned a, float x) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(x) : "memory");
}
__device__ __forceinline__ float xf_lds32f(unsigned a) {
  float x;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x) : "r"(a));
  return x;
}
__device__ __forceinline__ float xf_rcp(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float xf_sqrt(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// 2^d for -126 <= d <= 0, +0 below
__device__ __forceinline__ float xf_pow2_le0(int d) { return __int_as_float(max(d + 127, 0) << 23); }

// exp(x) = p * 2^k for x <= 0 (clamped at -1e6), p in [0.70, 1.42]
__device__ __forceinline__ float xf_exp_split(float x, int& k) {
  x = fmaxf(x, -1.0e6f);
  const float tm = fmaf(x, 1.4426950216293335f, 12582912.0f);  // 1.5 * 2^23: the integer lands in the low bits
  k = __float_as_int(tm) - 0x4B400000;
  const float kf = tm - 12582912.0f;
  float r = fmaf(x, 1.4426950216293335f, -kf);
  r = fmaf(x, 1.9259629911266175e-08f, r);  // log2(e) - (float)log2(e)
  float p;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p) : "f"(r));
  return p;
}

// v >= 0 (normal or zero) -> mantissa in [1, 2) and base + unbiased exponent; zero / denormal -> (0, XT_ZERO_EXP)
__device__ __forceinline__ void xf_split_exponent(float v, int base, float& mant, int& E) {
  const int b = __float_as_int(v);
  const int be = (b >> 23) & 0xff;
  mant = __int_as_float((b & 0x007fffff) | 0x3f800000);
  E = base + be - 127;
  if (be == 0) {
    mant = 0.0f;
    E = XT_ZERO_EXP;
  }
}

template <int D, int KS>
struct XfSeq {
  float m[D];
  float u[KS];
  float W;
  int E;
};

template <int D, int KS>
struct XfSlotIO {
  static constexpr int CO = D + KS + 1;
  static constexpr int NV = (CO + 3) / 4;
  static constexpr int SLOTB = NV * 512;  // bytes per slot (32 tracks)
  static constexpr int ESLOT = 128;       // exponent bytes per slot
  static __device__ __forceinline__ void load(unsigned va, unsigned ea, XfSeq<D, KS>& s) {
    float c[NV][4];
#pragma unroll
    for (int i = 0; i < NV; ++i) xf_lds128(va + i * 512, c[i]);
#pragma unroll
    for (int i = 0; i < D; ++i) s.m[i] = c[i / 4][i % 4];
#pragma unroll
    for (int i = 0; i < KS; ++i) s.u[i] = c[(D + i) / 4][(D + i) % 4];
    s.W = c[(D + KS) / 4][(D + KS) % 4];
    s.E = xt_lds32(ea);
  }
  static __device__ __forceinline__ void store(unsigned va, unsigned ea, const XfSeq<D, KS>& s) {
    float c[NV][4];
#pragma unroll
    for (int i = 0; i < 4 * NV; ++i) c[i / 4][i % 4] = 0.0f;
#pragma unroll
    for (int i = 0; i < D; ++i) c[i / 4][i % 4] = s.m[i];
#pragma unroll
    for (int i = 0; i < KS; ++i) c[(D + i) / 4][(D + i) % 4] = s.u[i];
    c[(D + KS) / 4][(D + KS) % 4] = s.W;
#pragma unroll
    for (int i = 0; i < NV; ++i) xf_sts128(va + i * 512, c[i]);
    xt_sts32(ea, s.E);
  }
};

template <int D, int KS>
__device__ __forceinline__ float xf_normfac(const float (&rq)[KS]) {
  if (KS == 1) {
    if (D == 1) return xf_sqrt(rq[0]);
    if (D == 2) return rq[0];
    return rq[0] * xf_sqrt(rq[0]);
  } else {
    float pr = rq[0];
#pragma unroll
    for (int k = 1; k < KS; ++k) pr *= rq[k];
    return xf_sqrt(pr);
  }
}

// Gaussian-product update (tracking.py:87-98) of a merged sequence with the localisation cl:
// result = parent record (m', u = l2*s2/q, W' * 2^E') shared by its children
template <int D, int KS>
__device__ __forceinline__ void xf_update(XfSeq<D, KS>& s, const float (&cl)[D], const float (&l2)[KS]) {
  float rq[KS], g[KS];
#pragma unroll
  for (int k = 0; k < KS; ++k) {
    rq[k] = xf_rcp(l2[k] + s.u[k]);
    g[k] = s.u[k] * rq[k];
    s.u[k] = l2[k] * g[k];
  }
  float e;
  if (KS == 1) {
    float q2 = 0.0f;
#pragma unroll
    for (int dim = 0; dim < D; ++dim) {
      const float df = cl[dim] - s.m[dim];
      s.m[dim] = fmaf(df, g[0], s.m[dim]);
      q2 = fmaf(df, df, q2);
    }
    e = q2 * (-0.5f * rq[0]);
  } else {
    float quad = 0.0f;
#pragma unroll
    for (int dim = 0; dim < D; ++dim) {
      const float df = cl[dim] - s.m[dim];
      s.m[dim] = fmaf(df, g[dim], s.m[dim]);
      quad = fmaf(df * df, rq[dim], quad);
    }
    e = -0.5f * quad;
  }
  int k2;
  const float p = xf_exp_split(e, k2);
  const float wn = (s.W * xf_normfac<D, KS>(rq)) * p;
  xf_split_exponent(wn, s.E + k2, s.W, s.E);
}

__host__ __device__ inline size_t xt_f32_smem(int D, int KS, int Pcap, int K, int H) {
  const int NV = (D + KS + 1 + 3) / 4;
  return (size_t)2 * Pcap * NV * 512               // state vectors, ping-pong
         + (size_t)2 * xt_fused_blob16(Pcap, K) * 16  // staged replay records, ping-pong
         + (size_t)2 * H * 8                       // (tau, dd) per head, without / with the stay term
         + (size_t)2 * Pcap * 128;                 // exponents, ping-pong
  // (the end-of-track partial sums, 1 KB, reuse the idle state buffer: Pcap >= 2)
}

#ifndef XT_K2F32_CTAS
#define XT_K2F32_CTAS 8
#endif
template <int D, int KS>
__global__ void __launch_bounds__(128, XT_K2F32_CTAS) k2_replay_f32(const K2FArgs a, const __grid_constant__ K2Tab T) {
  using IO = XfSlotIO<D, KS>;
  using Seq = XfSeq<D, KS>;
  constexpr int SLOTB = IO::SLOTB, ESLOT = IO::ESLOT;
  constexpr int WPC = 4, NT = 128;
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
  const int t = wk.t0 + lane;
  const bool valid = t < ck.nT;
  const double* Cs = a.soa + ck.xyz_off + (valid ? t : ck.nT - 1);
  const size_t cstride = (size_t)D * npad;
  const int L = ck.L;
  const int Pcap = a.Pcap;
  const int B16 = xt_fused_blob16(Pcap, K);

  extern __shared__ double2 k2f_smem[];
  const unsigned sb = xt_smem_base(k2f_smem);
  const unsigned VB = (unsigned)Pcap * SLOTB, EB = (unsigned)Pcap * ESLOT;
  const unsigned s_vec = sb + lane * 16;
  const unsigned s_blob = sb + 2 * VB;
  const unsigned s_tab = s_blob + 2 * B16 * 16;        // [2][H] x 8 B: (tau, dd)
  const unsigned s_exp = s_tab + 2 * H * 8 + lane * 4;  // [2][Pcap][32] x 4 B

  const int bstride = xt_blob_stride16(a.plan.cap);
  // replay records are staged by warp 0 alone with cp.async (no registers, no work for the other warps)
  const uint4* gnext = a.plan.blob + (size_t)ck.rec0 * bstride + lane;
  const int nrec = ck.nrec;
  const unsigned B16b = (unsigned)B16 * 16;
  if (w == 0 && nrec > 0) {
    for (int i = lane; i < B16; i += 32) xt_cp_async16(s_blob + i * 16, gnext + (i - lane));
    gnext += bstride;
  }
  for (int h = tid; h < 2 * H; h += NT) {
    const int hh = h < H ? h : h - H;
    xf_sts64(s_tab + h * 8, (float)(h < H ? T.tau0[hh] : T.tau1[hh]), (float)T.dd[hh]);
  }
  float l2[KS];
#pragma unroll
  for (int k = 0; k < KS; ++k) {
    l2[k] = (float)T.l2[k];
    asm volatile("mov.f32 %0, %0;" : "+f"(l2[k]));  // converted once (otherwise rematerialised per group)
  }

  // localisations relative to the first one (FP64 difference, then rounded)
  // NaN / Inf coordinates anywhere in the track poison the result: the differences to the first
  // localisation are then not finite either, and their running FP32 sum records it (a finite offset
  // beyond the FP32 range, > 3e38, counts as not finite on this path)
  double c0[D];
  float cl[D], cn[D], csum = 0.0f;
#pragma unroll
  for (int dim = 0; dim < D; ++dim) {
    c0[dim] = Cs[(size_t)dim * npad];
    cl[dim] = 0.0f;
  }
  Cs += cstride;
#pragma unroll
  for (int dim = 0; dim < D; ++dim) {
    cn[dim] = (float)(Cs[(size_t)dim * npad] - c0[dim]);  // C[1] (L >= 2)
    csum += cn[dim] * 0.0f + (float)(c0[dim] * 0.0);      // (0 or NaN)
  }

  // ---- first localisation (tracking.py:478-529) and, if L >= 3, the update of step 2 ----
  int nP = H;
  for (int c = w; c < nP; c += WPC) {
    Seq s;
    const float ddc = (float)T.dd[c];
#pragma unroll
    for (int dim = 0; dim < D; ++dim) s.m[dim] = 0.0f;
#pragma unroll
    for (int k = 0; k < KS; ++k) s.u[k] = l2[k] + ddc;
    xf_split_exponent((float)T.winit[c], 0, s.W, s.E);
    if (L >= 3) xf_update<D, KS>(s, cn, l2);
    IO::store(s_vec + c * SLOTB, s_exp + c * ESLOT, s);
  }
  if (L >= 3) {
    Cs += cstride;
#pragma unroll
    for (int dim = 0; dim < D; ++dim) {
      cn[dim] = (float)(Cs[(size_t)dim * npad] - c0[dim]);  // C[2]
      csum += cn[dim] * 0.0f;
    }
  }
  if (w == 0) xt_cp_async_wait();
  __syncthreads();

  // ---- steps 3..L-1: merge by the replay record of step-1, update with C[step-1] ----
  // ping-pong buffers: toggled by XOR with the difference of the two addresses
  unsigned src_v = s_vec, src_e = s_exp, dst_v = s_vec + VB, dst_e = s_exp + EB;
  const unsigned xv = src_v ^ dst_v, xe = src_e ^ dst_e, xr = s_blob ^ (s_blob + B16b);
  unsigned rb = s_blob;
  const int last = L - 1;
  for (int step = 3; step <= last; ++step) {
    // prefetch: localisation and replay record of the next step
    Cs += cstride;  // C[step] exists (step <= L-1)
    double cd[D];
#pragma unroll
    for (int dim = 0; dim < D; ++dim) {
      cl[dim] = cn[dim];
      cd[dim] = Cs[(size_t)dim * npad];
    }
    if (w == 0 && step < last) {  // record of the next step into the other buffer
      for (int i = lane; i < B16; i += 32) xt_cp_async16((rb ^ xr) + i * 16, gnext + (i - lane));
      gnext += bstride;
    }

    const int nG = (int)xt_lds16(rb);  // XtBlobHdr::nG
    const unsigned grec = rb + 32;
    const unsigned entb = grec + ((nG + 1) >> 1) * 16;
    unsigned tabp = s_tab + (((step - 1) >= T.min_len) ? H * 8 : 0);
    asm volatile("mov.u32 %0, %0;" : "+r"(tabp));
    unsigned ga = grec + xt_lds16(rb + 8 + 2 * w) * 8;  // XtBlobHdr::woff
    const unsigned ge = grec + xt_lds16(rb + 8 + 2 * (w + 1)) * 8;
    for (; ga < ge; ga += 8) {
      const uint2 gr = xt_lds64u(ga);
      const unsigned p0o = gr.x & 0x7FF80u;  // p0 * 128
      const unsigned g = gr.x >> 19;
      const unsigned kind = gr.y >> 30;
      Seq G;
      IO::load(src_v + p0o * (SLOTB / 128), src_e + p0o, G);
      float tau0, dd0;
      xf_lds64(tabp + (gr.x & 0x7Fu) * 8, tau0, dd0);
      if (kind == 1u) {  // single member: the child itself
        G.W *= tau0;
#pragma unroll
        for (int k = 0; k < KS; ++k) G.u[k] += dd0;
      } else if (kind == 2u) {
        const unsigned p1o = gr.y & 0x7FF80u;
        Seq B;
        IO::load(src_v + p1o * (SLOTB / 128), src_e + p1o, B);
        float tau1, dd1;
        xf_lds64(tabp + (gr.y & 0x7Fu) * 8, tau1, dd1);
        const int Eg = max(G.E, B.E);
        const float w0 = (G.W * xf_pow2_le0(G.E - Eg)) * tau0;
        const float w1 = (B.W * xf_pow2_le0(B.E - Eg)) * tau1;
        const float sw = w0 + w1;
        const float rs = (sw > 1e-36f) ? xf_rcp(sw) : 0.0f;
        const float lam = w1 * rs;
#pragma unroll
        for (int dim = 0; dim < D; ++dim) G.m[dim] = fmaf(B.m[dim] - G.m[dim], lam, G.m[dim]);
#pragma unroll
        for (int k = 0; k < KS; ++k) {
          const float ua = G.u[k] + dd0;
          G.u[k] = fmaf((B.u[k] + dd1) - ua, lam, ua);
        }
        G.W = sw;
        G.E = Eg;
      } else {
        // member list: two passes (largest exponent, then the weighted sums)
        const unsigned o = gr.y & 0xFFFu;
        const int n = (int)((gr.y >> 12) & 0x1FFFu);
        const unsigned eb = entb + o * 4;
        int Eg = G.E;
#pragma unroll 4
        for (int k = 1; k < n; ++k) Eg = max(Eg, xt_lds32(src_e + ((unsigned)xt_lds32(eb + k * 4) & 0xFFFFu) * ESLOT));
        float sw = (G.W * xf_pow2_le0(G.E - Eg)) * tau0, am[D], as[KS];
#pragma unroll
        for (int dim = 0; dim < D; ++dim) am[dim] = sw * G.m[dim];
#pragma unroll
        for (int k = 0; k < KS; ++k) {
          G.u[k] += dd0;
          as[k] = sw * G.u[k];
        }
#pragma unroll 2
        for (int k = 1; k < n; ++k) {
          const unsigned e = (unsigned)xt_lds32(eb + k * 4);
          const unsigned pm = e & 0xFFFFu;
          Seq M;
          IO::load(src_v + pm * SLOTB, src_e + pm * ESLOT, M);
          float taum, ddm;
          xf_lds64(tabp + ((e >> 16) & 0xFFu) * 8, taum, ddm);
          const float wj = (M.W * xf_pow2_le0(M.E - Eg)) * taum;
          sw += wj;
#pragma unroll
          for (int dim = 0; dim < D; ++dim) am[dim] = fmaf(wj, M.m[dim], am[dim]);
#pragma unroll
          for (int k2 = 0; k2 < KS; ++k2) as[k2] = fmaf(wj, M.u[k2] + ddm, as[k2]);
        }
        if (sw > 1e-36f) {  // otherwise: zero-weight group, keep the first member's moments
          const float rs = xf_rcp(sw);
#pragma unroll
          for (int dim = 0; dim < D; ++dim) G.m[dim] = am[dim] * rs;
#pragma unroll
          for (int k = 0; k < KS; ++k) G.u[k] = as[k] * rs;
        }
        G.W = sw;
        G.E = Eg;
      }
      // a merged weight may leave [1, 2): the update renormalises it (xf_split_exponent)
      xf_update<D, KS>(G, cl, l2);
      IO::store(dst_v + g * SLOTB, dst_e + g * ESLOT, G);
    }
    nP = nG;
    src_v ^= xv; dst_v ^= xv;  // swap the ping-pong buffers
    src_e ^= xe; dst_e ^= xe;
    rb ^= xr;
#pragma unroll
    for (int dim = 0; dim < D; ++dim) {
      cn[dim] = (float)(cd[dim] - c0[dim]);
      csum = fmaf(cn[dim], 0.0f, csum);
    }
    if (w == 0) xt_cp_async_wait();
    __syncthreads();
  }
  const uint8_t* curP = nullptr;
  if (nrec > 0) curP = a.plan.curG + (size_t)(ck.rec0 + nrec - 1) * a.plan.cap;

  // ---- end of track (tracking.py:613-639, :781-786): last localisation, leave term,
  //      sum over the surviving sequences in extended-exponent arithmetic ----
  const bool implicit = L >= 3;  // slots hold un-fused parents (m', u, W'): children are read on the fly
  const unsigned tab = s_tab + (((L - 1) >= T.min_len) ? H * 8 : 0);
  const int Kc = implicit ? K : 1;
  float acc = 0.0f;
  int KA = XT_ZERO_EXP;
  for (int p = w; p < nP; p += WPC) {
    Seq S;
    IO::load(src_v + p * SLOTB, src_e + p * ESLOT, S);
    const int ps = curP ? (int)__ldg(&curP[p]) : (p % nS);
    float df2[D];
#pragma unroll
    for (int dim = 0; dim < D; ++dim) {
      const float df = cn[dim] - S.m[dim];
      df2[dim] = df * df;
    }
    int newest_r = 0;  // r % nS, maintained incrementally
    for (int r = 0; r < Kc; ++r) {
      float dd = 0.0f, th = 1.0f;
      int newest = ps;
      if (implicit) {
        xf_lds64(tab + (r + K * ps) * 8, th, dd);
        newest = newest_r;
      }
      if (++newest_r == nS) newest_r = 0;
      if (ck.isBL) th *= (float)T.leave[newest];
      float rq[KS];
#pragma unroll
      for (int k = 0; k < KS; ++k) rq[k] = xf_rcp(S.u[k] + dd + l2[k]);
      float quad = 0.0f;
#pragma unroll
      for (int dim = 0; dim < D; ++dim) quad = fmaf(df2[dim], rq[(KS == 1) ? 0 : dim], quad);
      int k2;
      const float pe = xf_exp_split(-0.5f * quad, k2);
      const float v = ((S.W * th) * xf_normfac<D, KS>(rq)) * pe;
      const int Kv = S.E + k2;
      const int Kn = max(KA, Kv);
      acc = fmaf(acc, xf_pow2_le0(KA - Kn), v * xf_pow2_le0(Kv - Kn));
      KA = Kn;
    }
  }
  // cross-warp combination through the idle state buffer (every warp is past the last barrier
  // and reads only the current buffer)
  const unsigned s_redA = dst_v - lane * 16;   // [WPC][32] x 4 B
  const unsigned s_redK = s_redA + WPC * 128;  // [WPC][32] x 4 B
  xf_sts32f(s_redA + (w * 32 + lane) * 4, acc);
  xt_sts32(s_redK + (w * 32 + lane) * 4, KA);
  __syncthreads();
  if (w == 0) {
    int Kn = xt_lds32(s_redK + lane * 4);
#pragma unroll
    for (int k = 1; k < WPC; ++k) Kn = max(Kn, xt_lds32(s_redK + (k * 32 + lane) * 4));
    double tot = 0.0;  // the four partial sums are combined in FP64
#pragma unroll
    for (int k = 0; k < WPC; ++k)
      tot = fma((double)xf_lds32f(s_redA + (k * 32 + lane) * 4),
                xt_pow2_le0(xt_lds32(s_redK + (k * 32 + lane) * 4) - Kn), tot);
    double lp = XT_LN2 * (double)Kn + log(tot) - (double)(L - 1) * (0.5 * (double)D) * XT_LN_2PI;
    if (csum != 0.0f) lp = __longlong_as_double(0x7ff8000000000000ll);  // csum is 0 or NaN
    double lps = 0.0;
    if (valid) {
      a.logp[ck.trk_off + t] = lp;
      lps = lp;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) lps += __shfl_down_sync(0xffffffffu, lps, off);
    if (lane == 0) a.partial[wi] = lps;
  }
}
