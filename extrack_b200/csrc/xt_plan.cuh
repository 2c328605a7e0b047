// K1 — plan kernel: the greedy grouping of fuse_tracks_th (tracking.py:652-701) decided from the
// first <=30 tracks of every chunk, step by step, with the reference's operation order
// (explicit round-to-nearest intrinsics, no FMA contraction) so that the discontinuous
// `value < threshold` / `count/30 > 0.8` decisions see the same numbers as numpy does.
//
// One CTA per chunk, several CTAs resident per SM (the kernel is latency-bound: the steps of a
// chunk are strictly sequential).  lane = leader track; warps stride over parents / sequences /
// candidate leaders.  Leader-track state lives in a per-chunk global scratch block that stays
// L1/L2 resident.
//
// Grouping, two modes with identical results:
//  * nC <= 64  : every warp evaluates the full row of one *candidate* leader (the next not yet
//                grouped / not yet visited sequences in ascending order); one thread then resolves
//                the candidates greedily in order on 64-bit masks and emits the CSR lists.
//  * nC  > 64  : leaders are visited one at a time and the row is split over all warps.
#pragma once
#include "xt_common.cuh"

#ifdef XT_K1_PROF
#define K1_T(i) do { if (tid == 0) { const long long now__ = clock64(); prof[i] += now__ - tprev; tprev = now__; } } while (0)
#else
#define K1_T(i) do { } while (0)
#endif

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
  int32_t wpc;         // warps per tile of the fused replay kernel (schedule of the replay records)
  int32_t chunk0;      // first position of this launch in `corder`
  const int32_t* corder;  // chunk ids, longest tracks first: chunk id = corder[blockIdx.x + chunk0]
  long long* prof;        // XT_K1_PROF builds: [n_chunks][8] cycles per phase (thread 0)
  // shared-memory scratch (speculative: sized from the previous evaluation; 0 = global scratch).
  // A chunk that needs more parents / children than this reports err = 3 and the host retries with
  // the global-memory scratch.
  int32_t scapP, scapC;
  int32_t want_grec;   // also emit the inline group records of the first-generation replay kernel
  int32_t batch_mode;  // grouping of more than 64 sequences: 1 = batched candidate leaders, 0 = one leader at a time
  int32_t* vflag;      // VERIFY instantiation: [n_chunks] 0 = every decision of the resident plan was reproduced, else the step
                       // (>= 2) at which one changed (or 1: the plan has no verification records for this chunk)
  int32_t lpt;         // replay schedule: 1 = longest-processing-time-first (<= 64 groups, <= 4 replay warps), 0 = round-robin
  int32_t cost[4];     // cost model of the schedule: single-member group, pair, member list (cost[2] + cost[3] * members)
  int32_t cost_w0;     // ... and the per-step extra work of replay warp 0 in the same units
  XtAux ax;            // VAR instantiation only
};

// dynamic shared memory of k1_plan (bytes): per-sequence arrays + the bit rows of the batch-mode
// grouping (grouped mask + one row per warp, (cap + 63) / 64 words each), then the optional scratch
__host__ __device__ inline size_t xt_k1_base(int cap, int nthreads = XT_K1_THREADS) {
  size_t b = (size_t)cap * (8 + 8 + 4 + 4 + 4 + 4 + 1) + 64;
  b = (b + 15) & ~(size_t)15;
  b += (size_t)((cap + 63) / 64) * 8 * (nthreads / 32 + 1);
  return (b + 15) & ~(size_t)15;
}
__host__ __device__ inline size_t xt_k1_smem(int cap, int CO, int RH, int nS, int scapP, int scapC, int varH = 0,
                                             int nthreads = XT_K1_THREADS) {
  size_t b = xt_k1_base(cap, nthreads);
  if (scapC > 0) b += (size_t)(scapP + scapC) * CO * 32 * 8 + (size_t)2 * scapP * RH * nS * 8;
  b += (size_t)varH * 32 * 8;  // VAR: per-lane dd of every head
  return b;
}

__device__ __forceinline__ int xt_label(int x, int nS, bool wrap) {
  if (wrap) {
    int v = (int)(int8_t)(x & 0xFF);  // np.arange(..., dtype='int8') wraps, tracking.py:543
    int r = v % nS;
    return r < 0 ? r + nS : r;        // np.mod is non-negative for a positive divisor
  }
  return x % nS;
}

// NT threads per chunk: 256 (several chunks per SM) or 1024 when the data set has fewer chunks than
// the GPU has SMs (long tracks: the per-step phases then run in a quarter of the rounds)
// SS: the scratch is known at compile time to be in shared memory (scapC > 0): the accesses then compile
// to shared-memory instructions with 32-bit addresses instead of generic loads / stores.
// VERIFY: plan verification instead of plan construction (scalar models, matrix-mode plans).  The leader tracks are
// advanced with the same arithmetic (update, merge) along the RESIDENT plan, and every floating-point decision that plan
// rests on (XtVRec) is re-evaluated with the parameters of this evaluation; codes, histories, member lists and replay
// records are pure functions of the decisions, so they stay valid exactly as long as every decision is reproduced.  The
// replay kernel can therefore run concurrently on the resident records; a chunk whose verification fails is planned
// again from scratch (and replayed again) before the result is used.
template <int D, int KS, bool VAR, int NT, bool SS = false, bool VERIFY = false>
__global__ void __launch_bounds__(NT, NT == XT_K1_THREADS ? XT_K1_MIN_CTAS : 1)
k1_plan(const K1Args a, const __grid_constant__ xt_params P) {
  static_assert(!VERIFY || (!VAR && SS), "verification mode: scalar models with the scratch in shared memory");
  constexpr int CO = D + 2 * KS + 1;  // m[D], s2[KS], s[KS], LP
  constexpr int W = NT / 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, tid = threadIdx.x;
  const int cid = a.corder[(int)blockIdx.x + a.chunk0];
  const XtChunk ck = a.chunks[cid];
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
  uint32_t* ent = (uint32_t*)(grank + cap);  // CSR member entries of the step (mirrored to the plan in global memory)
  int* gcnt = (int*)(ent + cap);      // [cap+1]: group sizes -> offsets
  unsigned char* curP = (unsigned char*)(gcnt + cap + 1);
  __shared__ int s_flag, s_nG, s_left, s_cand[NT / 32];
  __shared__ unsigned long long s_rows[64];  // capture matrix of the matrix-mode grouping
  __shared__ unsigned long long s_todo[64];  // ... and, per row, the sequences its floating-point predicate was evaluated for
  __shared__ __align__(4) unsigned char s_js[(NT / 32) * 72];  // matrix mode: per warp, the sequences a row still has to test
  // batch mode: grouped bit mask and one bit row per candidate (= per warp), BW words each
  const int BW = (cap + 63) / 64;
  unsigned long long* s_grp = (unsigned long long*)(k1_smem + ((((size_t)cap * (8 + 8 + 4 + 4 + 4 + 4 + 1) + 64) + 15) & ~(size_t)15));
  unsigned long long* s_brow = s_grp + BW;

  XtChunkSummary* sm = &a.summ[cid];
  const int nP0 = K * nS;
  if (VERIFY) {
    if (tid == 0) s_flag = 0;
    if (ck.L < 4) {  // no fusion step, nothing to verify
      if (tid == 0) a.vflag[cid] = 0;
      return;
    }
  }
  if (!VERIFY && tid == 0) {
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
  // smallest count with (double)count / (Kt*KS) > 0.8  (np.mean(bool) > 0.8, tracking.py:689-691)
  int min_cnt = Kt * KS + 1;
  {
    const double denom = (double)(Kt * KS);
    for (int c = Kt * KS; c >= 0; --c)
      if (__ddiv_rn((double)c, denom) > 0.8) min_cnt = c;
  }

  // leader-track state and history rows: shared memory when the launch was sized for it (every
  // producer -> consumer hand-off between phases is then a shared-memory round trip instead of
  // an L2 one), else the per-chunk global scratch.  Generic pointers: same code for both.
  const int scapP = a.scapP, scapC = a.scapC;
  double *bufP, *bufC, *histP, *histN;
  if (SS) {
    bufP = (double*)(k1_smem + xt_k1_base(cap, NT));
    bufC = bufP + (size_t)scapP * CO * 32;
    histP = bufC + (size_t)scapC * CO * 32;
    histN = histP + (size_t)scapP * a.RH * nS;
  } else {
    bufP = a.state + (size_t)cid * 2 * cap * CO * 32;
    bufC = bufP + (size_t)cap * CO * 32;
    histP = a.hist + (size_t)cid * 2 * cap * a.RH * nS;
    histN = histP + (size_t)cap * a.RH * nS;
    if (scapC > 0) {
      const size_t o = xt_k1_base(cap, NT);
      bufP = (double*)(k1_smem + o);
      bufC = bufP + (size_t)scapP * CO * 32;
      histP = bufC + (size_t)scapC * CO * 32;
      histN = histP + (size_t)scapP * a.RH * nS;
    }
  }
  if (scapC > 0 && nP0 > scapP) {
    if (tid == 0) {
      if (VERIFY) a.vflag[cid] = 1; else sm->err = 3;
    }
    return;
  }
  const int bits = a.bits;
  const unsigned long long rowmask = (1ull << bits) - 1ull;

#define ST(buf, slot, comp) (buf)[((size_t)(slot) * CO + (comp)) * 32 + lane]
  // children state for the predicate rows: 32-bit shared-window addresses when the scratch is known to be in
  // shared memory (one multiply-add per access instead of 64-bit generic-pointer arithmetic)
  const unsigned cC32 = SS ? xt_smem_base(bufC) + (unsigned)lane * 8u : 0u;
  auto ldC = [&](int slot, int comp) -> double {
    if (SS) return xt_lds64(cC32 + (unsigned)(slot * CO + comp) * 256u);
    return ST(bufC, slot, comp);
  };

  double l2[KS];
#pragma unroll
  for (int k = 0; k < KS; ++k) l2[k] = P.l2[k];

  // VAR: peak-wise LocErr / per-localisation dt of the leader tracks.  Row j of the aux block
  // belongs to localisation j; s_dd[head][lane] holds dd of the current step for every leader.
  const bool var_loc = VAR && (P.flags & XT_FLAG_VAR_LOC), var_dt = VAR && (P.flags & XT_FLAG_VAR_DT);
  const double* Ap = VAR ? a.ax.aux + (size_t)(ck.xyz_off / D) * a.ax.R + t : nullptr;
  const double* Lps = (VAR && a.ax.stay) ? a.ax.stay + (size_t)cid * K : P.Lp_stay;
  double* s_dd = nullptr;
  if (VAR) {
    size_t o = xt_k1_base(cap, NT);
    if (scapC > 0) o += (size_t)(scapP + scapC) * CO * 32 * 8 + (size_t)2 * scapP * a.RH * nS * 8;
    s_dd = (double*)(k1_smem + o);
  }
  auto var_step = [&](int j) {  // call by all threads, followed by a barrier before s_dd is read
    if (var_loc) {
#pragma unroll
      for (int k = 0; k < KS; ++k) l2[k] = xt_sigma2(P, Ap[(size_t)(j * a.ax.R + k) * npad]);
    }
    for (int idx = tid; idx < nP0 * 32; idx += NT) {
      const int h = idx >> 5, ln = idx & 31;
      double v = P.dd[h];
      if (var_dt) v = xt_dd_exact(P, h, Ap[(size_t)(j * a.ax.R + a.ax.ka) * npad - t + (ln < Kt ? ln : 0)]);
      s_dd[idx] = v;
    }
  };
#define DDH(head) (VAR ? s_dd[(head) * 32 + lane] : P.dd[head])
  if (VAR) {
    var_step(0);
    __syncthreads();
  }

  // ---- first localisation (tracking.py:478-529) ----
  int nP = nP0;
  for (int c = warp; c < nP; c += W) {
#pragma unroll
    for (int dim = 0; dim < D; ++dim) ST(bufP, c, dim) = Cp[(size_t)(0 * D + dim) * npad];
#pragma unroll
    for (int k = 0; k < KS; ++k) ST(bufP, c, D + k) = __dadd_rn(l2[k], DDH(c));
    ST(bufP, c, D + 2 * KS) = __dadd_rn(P.LT[c], P.LF[c]);
  }
  int LhP = nsub + 1;
  if (VERIFY) {
    for (int c = tid; c < nP; c += NT) curP[c] = (unsigned char)(c % nS);
  } else
  for (int c = tid; c < nP; c += NT) {
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

#ifdef XT_K1_PROF
  long long prof[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  long long tprev = clock64();
#endif
  double cnext[D];
#pragma unroll
  for (int dim = 0; dim < D; ++dim) cnext[dim] = Cp[(size_t)(1 * D + dim) * npad];  // step 2 consumes localisation 1
  for (int step = 2; step <= L - 2; ++step) {
    const int nC = nP * K;
    if (VERIFY) {  // the resident plan must describe this very step in matrix mode (uniform: same words for every thread)
      const int rec_v = ck.rec0 + (step - 2);
      if (nC > cap || nC > scapC || nC > 64 || a.plan.hdr[rec_v].nC != nC || !a.plan.vok[rec_v] || a.plan.hdr[rec_v].nG > scapP) {
        if (tid == 0) a.vflag[cid] = 1;
        return;
      }
    } else {
    if (nC > cap) {
      if (tid == 0) {
        sm->err = 2;
        sm->need_cap = nC;
      }
      return;
    }
    if (scapC > 0 && nC > scapC) {  // more children than the shared-memory scratch was sized for
      if (tid == 0) sm->err = 3;
      return;
    }
    }
    sum_nC += nC;
    max_nC = nC > max_nC ? nC : max_nC;
    const int rec = ck.rec0 + (step - 2);
    const int LhC = LhP + nsub;
    const int rows_cmp = LhC < P.frame_len ? LhC : P.frame_len;  // rows kept in the window code
    const bool use_window = LhC > P.frame_len;
    const unsigned long long cmask = (bits * rows_cmp >= 64) ? ~0ull : ((1ull << (bits * rows_cmp)) - 1ull);
    uint16_t* goff = a.plan.goff + (size_t)rec * (a.plan.cap + 1);
    uint32_t* gent = a.plan.ent + (size_t)rec * a.plan.cap;
    uint16_t* pgid = a.plan.gid + (size_t)rec * a.plan.cap;

    // ---- expansion + Gaussian update on the leader tracks (tracking.py:540-570, :87-98);
    //      one warp per parent, its K children share q, the new mean and the log term ----
    double cl[D];
#pragma unroll
    for (int dim = 0; dim < D; ++dim) {
      cl[dim] = cnext[dim];
      cnext[dim] = Cp[(size_t)(step * D + dim) * npad];  // localisation of the next step (step <= L - 2: it exists),
    }                                                    // in flight during this step instead of at the head of the next one
    const bool stay = step >= P.min_len;
    if (VAR) {  // (the previous step ended with a barrier: nobody still reads s_dd)
      var_step(step - 1);
      __syncthreads();
    }
    for (int p = warp; p < nP; p += W) {
      const double LPp = ST(bufP, p, D + 2 * KS);
      double mm[D], s2[KS], q[KS], nm[D];
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
        nm[dim] = __ddiv_rn(__dadd_rn(__dmul_rn(mm[dim], l2[k]), __dmul_rn(cl[dim], s2[k])), __dadd_rn(l2[k], s2[k]));
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
      const double LC = __dsub_rn(logs, quad);
      const int hbase = K * (int)curP[p];
      for (int r = 0; r < K; ++r) {
        const int c = p * K + r, head = r + hbase;
        const double dd = DDH(head);
#pragma unroll
        for (int dim = 0; dim < D; ++dim) ST(bufC, c, dim) = nm[dim];
#pragma unroll
        for (int k = 0; k < KS; ++k) {
          const double ns2 = __ddiv_rn(
              __dadd_rn(__dadd_rn(__dmul_rn(dd, l2[k]), __dmul_rn(dd, s2[k])), __dmul_rn(l2[k], s2[k])), q[k]);
          ST(bufC, c, D + k) = ns2;
          ST(bufC, c, D + KS + k) = __dsqrt_rn(ns2);
        }
        double add = __dadd_rn(P.LT[head], LC);
        if (stay) add = __dadd_rn(add, Lps[r]);
        ST(bufC, c, D + 2 * KS) = __dadd_rn(LPp, add);
      }
    }
    // window codes of the children: nsub new labels in front of the parent's rows
    if (!VERIFY)
    for (int c = tid; c < nC; c += NT) {
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
    K1_T(0);

    double th_lo = __dmul_rn(th, 1.0 - 1e-14), th_hi = __dmul_rn(th, 1.0 + 1e-14);
    asm volatile("" : "+d"(th_lo), "+d"(th_hi));  // live in registers (otherwise rebuilt from constants at every use)
    // predicate "leader i captures sequence j" (all lanes of the warp must call it together)
    // floating-point half of the predicate (m_mask and s_mask, tracking.py:689-691)
    auto fp_ok = [&](const double (&mi)[D], const double (&si)[KS], int j) -> bool {
      double am = 0.0, as = 0.0;
#pragma unroll
      for (int dim = 0; dim < D; ++dim) {
        const double v = fabs(__dsub_rn(ST(bufC, j, dim), mi[dim]));
        am = (dim == 0) ? v : __dadd_rn(am, v);
      }
      am = (D == 2) ? __dmul_rn(am, 0.5) : ((D == 1) ? am : __ddiv_rn(am, (double)D));
      double sj[KS];
#pragma unroll
      for (int k = 0; k < KS; ++k) {
        sj[k] = ST(bufC, j, D + KS + k);
        const double v = fabs(__dsub_rn(sj[k], si[k]));
        as = (k == 0) ? v : __dadd_rn(as, v);
      }
      as = (KS == 2) ? __dmul_rn(as, 0.5) : ((KS == 1) ? as : __ddiv_rn(as, (double)KS));
      int cnt_m = 0, cnt_s = 0;
#pragma unroll
      for (int k = 0; k < KS; ++k) {
        // fl(x / s) < th decided without a division unless x is within 1e-14 (relative) of th*s
        const double lo = __dmul_rn(th_lo, sj[k]), hi = __dmul_rn(th_hi, sj[k]);
        bool pm = am < lo, ps = as < lo;
        if (!pm && !(am > hi)) pm = __ddiv_rn(am, sj[k]) < th;
        if (!ps && !(as > hi)) ps = __ddiv_rn(as, sj[k]) < th;
        cnt_m += __popc(__ballot_sync(0xffffffffu, act && pm));
        cnt_s += __popc(__ballot_sync(0xffffffffu, act && ps));
      }
      return cnt_m >= min_cnt && cnt_s >= min_cnt;
    };
    // the same predicate for up to four sequences at once (independent loads and dependency
    // chains: the plan kernel is latency-bound); js[q] for q >= nj repeat a valid index
    auto fp_ok4 = [&](const double (&mi)[D], const double (&si)[KS], const int (&js)[4], bool (&ok)[4]) {
      double am[4], as[4], sj[4][KS];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        am[q] = 0.0;
        as[q] = 0.0;
#pragma unroll
        for (int dim = 0; dim < D; ++dim) {
          const double v = fabs(__dsub_rn(ldC(js[q], dim), mi[dim]));
          am[q] = (dim == 0) ? v : __dadd_rn(am[q], v);
        }
#pragma unroll
        for (int k = 0; k < KS; ++k) {
          sj[q][k] = ldC(js[q], D + KS + k);
          const double v = fabs(__dsub_rn(sj[q][k], si[k]));
          as[q] = (k == 0) ? v : __dadd_rn(as[q], v);
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        am[q] = (D == 2) ? __dmul_rn(am[q], 0.5) : ((D == 1) ? am[q] : __ddiv_rn(am[q], (double)D));
        as[q] = (KS == 2) ? __dmul_rn(as[q], 0.5) : ((KS == 1) ? as[q] : __ddiv_rn(as[q], (double)KS));
      }
      int cnt_m[4] = {0, 0, 0, 0}, cnt_s[4] = {0, 0, 0, 0};
#pragma unroll
      for (int k = 0; k < KS; ++k) {
        bool pm[4], ps[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const double lo = __dmul_rn(th_lo, sj[q][k]), hi = __dmul_rn(th_hi, sj[q][k]);
          pm[q] = am[q] < lo;
          ps[q] = as[q] < lo;
          if (!pm[q] && !(am[q] > hi)) pm[q] = __ddiv_rn(am[q], sj[q][k]) < th;
          if (!ps[q] && !(as[q] > hi)) ps[q] = __ddiv_rn(as[q], sj[q][k]) < th;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          cnt_m[q] += __popc(__ballot_sync(0xffffffffu, act && pm[q]));
          cnt_s[q] += __popc(__ballot_sync(0xffffffffu, act && ps[q]));
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) ok[q] = cnt_m[q] >= min_cnt && cnt_s[q] >= min_cnt;
    };
    // predicate "leader i captures sequence j" (all lanes of the warp must call it together)
    auto pair_ok = [&](const double (&mi)[D], const double (&si)[KS], unsigned long long ci, int j) -> bool {
      const unsigned long long cj = codeC[j];
      if (use_window && cj == ci) return true;             // state_mask, tracking.py:679-681
      if ((cj & rowmask) != (ci & rowmask)) return false;  // cur_state_mask, :673-674
      return fp_ok(mi, si, j);
    };

    int nG = 0;
    if (VERIFY) {
      // ---- verification: the member lists of the resident plan, and every floating-point decision behind them ----
      nG = a.plan.hdr[rec].nG;
      for (int c = tid; c < nC; c += NT) ent[c] = gent[c];
      for (int g = tid; g <= nG; g += NT) gcnt[g] = (int)goff[g];
      const XtVRec* vr = a.plan.vrec + (size_t)rec * XT_VREC_PER_STEP;
      for (int g = warp; g < nG; g += W) {
        const XtVRec v = vr[g];
        const int i = (int)v.lead;
        double mi[D], si[KS];
#pragma unroll
        for (int dim = 0; dim < D; ++dim) mi[dim] = ldC(i, dim);
#pragma unroll
        for (int k = 0; k < KS; ++k) si[k] = ldC(i, D + KS + k);
        const unsigned long long todo = v.cand;
        const int ntodo = __popcll(todo);
        unsigned add_lo = 0u, add_hi = 0u;
        if (ntodo) {
          unsigned char* myjs = s_js + warp * 72;
          const unsigned tlo = (unsigned)todo, thi = (unsigned)(todo >> 32);
          const unsigned below = (1u << lane) - 1u;
          if ((tlo >> lane) & 1u) myjs[__popc(tlo & below)] = (unsigned char)lane;
          if ((thi >> lane) & 1u) myjs[__popc(tlo) + __popc(thi & below)] = (unsigned char)(lane + 32);
          if (lane < 3) myjs[ntodo + lane] = (unsigned char)(__ffsll((long long)todo) - 1);
          __syncwarp();
          for (int r0 = 0; r0 < ntodo; r0 += 4) {
            const unsigned jj = *reinterpret_cast<const unsigned*>(myjs + r0);
            const int js[4] = {(int)(jj & 0xFFu), (int)((jj >> 8) & 0xFFu), (int)((jj >> 16) & 0xFFu), (int)(jj >> 24)};
            bool ok[4];
            fp_ok4(mi, si, js, ok);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const unsigned bit = (r0 + q < ntodo && ok[q]) ? 1u << (js[q] & 31) : 0u;
              if (js[q] & 32) add_hi |= bit; else add_lo |= bit;
            }
          }
          __syncwarp();
        }
        const unsigned long long got = (unsigned long long)add_lo | ((unsigned long long)add_hi << 32);
        if (got != v.exp && lane == 0) s_flag = 1;  // a decision changed: this chunk needs a new plan
      }
      __syncthreads();
      if (s_flag) {
        if (tid == 0) a.vflag[cid] = step;
        return;
      }
      // newest true state of the groups (next step's parents): curP is not read again before the end-of-step barrier
      const uint8_t* pcur_v = a.plan.curG + (size_t)rec * a.plan.cap;
      for (int g = tid; g < nG; g += NT) curP[g] = pcur_v[g];
    } else if (nC <= 64) {
      // ---- matrix mode ----
      // Phase 1 (parallel, no dependency between leaders): row i of the capture matrix for every
      // sequence i, restricted to j >= i.  The greedy loop of the reference visits leaders in
      // ascending order and everything below a leader is already grouped (a leader that captures
      // nothing, not even itself, is the reference's error path, tracking.py:725), so only the
      // upper triangle is ever consulted.  Rows are dealt to the warps in zigzag order (row i has
      // about (nC - i) / 2 floating-point tests).
      // window codes of sequences `lane` and `lane + 32` (never equal to a real code when absent)
      const unsigned long long code_lo = (lane < nC) ? codeC[lane] : ~0ull;
      const unsigned long long code_hi = (lane + 32 < nC) ? codeC[lane + 32] : ~0ull;
      const unsigned long long full = (nC == 64) ? ~0ull : ((1ull << nC) - 1ull);
      for (int base = 0, pass = 0; base < nC; base += W, ++pass) {
        const int cnt = (nC - base) < W ? (nC - base) : W;
        if (warp >= cnt) continue;  // warp-uniform
        const int i = base + ((pass & 1) ? (cnt - 1 - warp) : warp);
        double mi[D], si[KS];
#pragma unroll
        for (int dim = 0; dim < D; ++dim) mi[dim] = ST(bufC, i, dim);
#pragma unroll
        for (int k = 0; k < KS; ++k) si[k] = ST(bufC, i, D + KS + k);
        const unsigned long long ci = codeC[i];
        const unsigned long long upper = full & ~((1ull << i) - 1ull);
        // code tests for all sequences at once (lane = sequence, two per lane)
        const unsigned st_lo = __ballot_sync(0xffffffffu, (code_lo & rowmask) == (ci & rowmask));
        const unsigned st_hi = __ballot_sync(0xffffffffu, (code_hi & rowmask) == (ci & rowmask));
        const unsigned wn_lo = __ballot_sync(0xffffffffu, use_window && code_lo == ci);
        const unsigned wn_hi = __ballot_sync(0xffffffffu, use_window && code_hi == ci);
        const unsigned long long state_eq = (unsigned long long)st_lo | ((unsigned long long)st_hi << 32);
        const unsigned long long win_eq = (unsigned long long)wn_lo | ((unsigned long long)wn_hi << 32);
        unsigned long long row = win_eq & upper;                  // state_mask (:679-681)
        unsigned long long todo = state_eq & ~win_eq & upper;     // need the m/s tests (:689-693)
        // the sequences to test, as a byte list in shared memory (lane l owns bits l and l + 32), padded to a
        // multiple of four with the first one; then four sequences per round
        const int ntodo = __popcll(todo);
        unsigned add_lo = 0u, add_hi = 0u;
        if (ntodo) {
          unsigned char* myjs = s_js + warp * 72;
          const unsigned tlo = (unsigned)todo, thi = (unsigned)(todo >> 32);
          const unsigned below = (1u << lane) - 1u;
          if ((tlo >> lane) & 1u) myjs[__popc(tlo & below)] = (unsigned char)lane;
          if ((thi >> lane) & 1u) myjs[__popc(tlo) + __popc(thi & below)] = (unsigned char)(lane + 32);
          if (lane < 3) myjs[ntodo + lane] = (unsigned char)(__ffsll((long long)todo) - 1);
          __syncwarp();
          for (int r0 = 0; r0 < ntodo; r0 += 4) {
            const unsigned jj = *reinterpret_cast<const unsigned*>(myjs + r0);
            const int js[4] = {(int)(jj & 0xFFu), (int)((jj >> 8) & 0xFFu), (int)((jj >> 16) & 0xFFu), (int)(jj >> 24)};
            bool ok[4];
            fp_ok4(mi, si, js, ok);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const unsigned bit = (r0 + q < ntodo && ok[q]) ? 1u << (js[q] & 31) : 0u;
              if (js[q] & 32) add_hi |= bit; else add_lo |= bit;
            }
          }
          __syncwarp();  // the next row of this warp reuses the list
        }
        row |= (unsigned long long)add_lo | ((unsigned long long)add_hi << 32);
        if (lane == 0) {
          s_rows[i] = row;
          s_todo[i] = todo;
        }
      }
      __syncthreads();
      K1_T(1);
      // Phase 2: greedy resolution on the bit rows (tracking.py:667-698) by one warp (the other
      // warps wait at the barrier and leave the issue slots to the co-resident chunks); lane l
      // emits the CSR entries of sequences l and l + 32 when they are captured.
      if (warp == 0) {
        unsigned long long grouped = 0ull, rem = full;
        int off = 0, ng = 0;
        bool bad = false;
        int jp[2], jr[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int j = lane + 32 * h;
          jp[h] = (j < nC) ? j / K : 0;
          jr[h] = j - jp[h] * K;
        }
        // the sequential part only tracks the masks; every lane remembers the group that captured its
        // two sequences (lane, lane + 32) and emits gid / CSR entries once, after the loop
        int myg[2] = {-1, -1}, myoff[2] = {0, 0};
        unsigned long long mymem[2] = {0ull, 0ull};
        const unsigned long long bit0 = 1ull << lane, bit1 = 1ull << (lane + 32);
        while (rem) {
          const int i = __ffsll((long long)rem) - 1;
          const unsigned long long mem = s_rows[i] & ~grouped;
          if (mem == 0ull) {  // empty group: the reference fails on the zero-size max (:725)
            bad = true;
            rem &= rem - 1ull;
            continue;
          }
          if (mem & bit0) { myg[0] = ng; myoff[0] = off; mymem[0] = mem; }
          if (mem & bit1) { myg[1] = ng; myoff[1] = off; mymem[1] = mem; }
          if (lane == 0) {
            gcnt[ng] = off;
            // verification record: the floating-point decisions this leader's capture rests on
            const unsigned long long cand = s_todo[i] & ~grouped;
            XtVRec* vw = a.plan.vrec + (size_t)rec * XT_VREC_PER_STEP + ng;
            vw->lead = (unsigned long long)i;
            vw->cand = cand;
            vw->exp = mem & cand;
          }
          off += __popcll(mem);
          grouped |= mem;
          rem &= ~mem;
          ++ng;
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int j = lane + 32 * h;
          if (j < nC && myg[h] >= 0) {
            gid[j] = myg[h];
            ent[myoff[h] + __popcll(mymem[h] & ((1ull << j) - 1ull))] = xt_pack_ent(jp[h], jr[h] + K * (int)curP[jp[h]], jr[h]);
          }
        }
        if (lane == 0) {
          gcnt[ng] = off;
          s_nG = ng;
          if (bad || grouped != full) s_flag = 1;  // tracking.py:700-701
        }
      }
      K1_T(2);
      __syncthreads();
      nG = s_nG;
      for (int c = tid; c < nC; c += NT) {
        pgid[c] = (uint16_t)gid[c];
        gent[c] = ent[c];
      }
      for (int g = tid; g <= nG; g += NT) goff[g] = (uint16_t)gcnt[g];
      if (s_flag) {
        if (tid == 0) sm->err = 1;
        return;
      }
    } else {
      // ---- batch mode (nC > 64): the next W not yet grouped sequences are candidate leaders, one
      //      per warp; every warp evaluates the row of its candidate against the sequences that are
      //      still ungrouped (four floating-point tests in flight), then one warp resolves the
      //      candidates in ascending order on the bit rows (a candidate captured by an earlier
      //      leader of the batch is dropped: the reference never visits it as a leader) ----
      const int NW = (nC + 63) >> 6;  // 64-bit words per row
      if (a.batch_mode) {
        for (int wd = tid; wd < NW; wd += NT) s_grp[wd] = 0ull;
        if (tid == 0) { s_nG = 0; s_left = nC; }
        __syncthreads();
        for (;;) {
          if (s_left == 0) break;  // uniform (written before the barrier that ends the previous round)
          // candidate of this warp: the (warp)-th zero bit of the grouped mask (words scanned by lanes)
          int cand = -1;
          {
            int before = 0;  // zero bits in the words below the current 32-word window
            for (int w0 = 0; w0 < NW && cand < 0; w0 += 32) {
              const int wd = w0 + lane;
              unsigned long long z = 0ull;
              if (wd < NW) {
                z = ~s_grp[wd];
                if (wd == NW - 1 && (nC & 63)) z &= (1ull << (nC & 63)) - 1ull;
              }
              const int cnt = __popcll(z);
              int pre = cnt;  // inclusive prefix over the lanes
#pragma unroll
              for (int o2 = 1; o2 < 32; o2 <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, pre, o2);
                if (lane >= o2) pre += v;
              }
              const int lo = before + pre - cnt;  // zero bits before this lane's word
              const bool mine = warp >= lo && warp < lo + cnt;
              int bitpos = -1;
              if (mine) {  // the (warp - lo)-th set bit of z
                unsigned long long zz = z;
                for (int q = 0; q < warp - lo; ++q) zz &= zz - 1ull;
                bitpos = wd * 64 + __ffsll((long long)zz) - 1;
              }
              const unsigned who = __ballot_sync(0xffffffffu, mine);
              if (who) cand = __shfl_sync(0xffffffffu, bitpos, __ffs(who) - 1);
              before += __shfl_sync(0xffffffffu, pre, 31);
            }
          }
          if (lane == 0) s_cand[warp] = cand;
          if (cand >= 0) {
            const int i = cand;
            double mi[D], si[KS];
#pragma unroll
            for (int dim = 0; dim < D; ++dim) mi[dim] = ST(bufC, i, dim);
#pragma unroll
            for (int k = 0; k < KS; ++k) si[k] = ST(bufC, i, D + KS + k);
            const unsigned long long ci = codeC[i];
            unsigned* brow = (unsigned*)&s_brow[warp * BW];
            for (int b = 0; b < 2 * NW; ++b) {
              if (b < (i >> 5)) {  // below the candidate: already grouped (uniform)
                if (lane == 0) brow[b] = 0u;
                continue;
              }
              const int j = 32 * b + lane;
              const bool valid = j < nC && j >= i && !((s_grp[j >> 6] >> (j & 63)) & 1ull);
              const unsigned long long cj = valid ? codeC[j] : ~0ull;
              const unsigned win = __ballot_sync(0xffffffffu, valid && use_window && cj == ci);
              unsigned todo = __ballot_sync(0xffffffffu, valid && (cj & rowmask) == (ci & rowmask)) & ~win;
              unsigned row = win;
              while (todo) {
                int js[4];
                int nj = 0;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  if (todo) {
                    js[q] = 32 * b + __ffs(todo) - 1;
                    todo &= todo - 1u;
                    nj = q + 1;
                  } else {
                    js[q] = js[0];
                  }
                }
                bool ok[4];
                fp_ok4(mi, si, js, ok);
#pragma unroll
                for (int q = 0; q < 4; ++q)
                  if (q < nj && ok[q]) row |= 1u << (js[q] & 31);
              }
              if (lane == 0) brow[b] = row;
            }
          }
          __syncthreads();
          if (warp == 0) {
            int ng = s_nG, left = s_left;
            bool bad = false;
            for (int t2 = 0; t2 < W; ++t2) {
              const int i = s_cand[t2];
              if (i < 0) break;
              if ((s_grp[i >> 6] >> (i & 63)) & 1ull) continue;  // captured by an earlier leader of this batch
              int cnt = 0;
              for (int w0 = 0; w0 < NW; w0 += 32) {
                const int wd = w0 + lane;
                unsigned long long mem = 0ull;
                if (wd < NW) {
                  mem = s_brow[t2 * BW + wd] & ~s_grp[wd];
                  s_grp[wd] |= mem;
                }
                cnt += __popcll(mem);
                unsigned nz = __ballot_sync(0xffffffffu, mem != 0ull);
                while (nz) {  // publish the members word by word: lanes l, l + 32 own bits l, l + 32
                  const int src = __ffs(nz) - 1;
                  nz &= nz - 1u;
                  const unsigned long long m2 = __shfl_sync(0xffffffffu, mem, src);
                  const int base = (w0 + src) * 64;
                  if ((m2 >> lane) & 1ull) gid[base + lane] = ng;
                  if ((m2 >> (lane + 32)) & 1ull) gid[base + lane + 32] = ng;
                }
              }
#pragma unroll
              for (int o2 = 16; o2 > 0; o2 >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o2);
              if (cnt == 0) bad = true;  // empty group: the reference fails on the zero-size max (:725)
              left -= cnt;
              ++ng;
              __syncwarp();
            }
            if (lane == 0) {
              s_nG = ng;
              s_left = bad ? 0 : left;
              if (bad) s_flag = 1;
            }
          }
          __syncthreads();
        }
        nG = s_nG;
      } else
      // ---- split mode (diagnostic option k1_batch = 0): one leader at a time, its row split over the warps ----
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
          if (pair_ok(mi, si, ci, j) && lane == 0) gid[j] = nG;
        }
        ++nG;
        __syncthreads();
      }
      // every sequence must have been grouped (tracking.py:700-701)
      for (int c = tid; c < nC; c += NT)
        if (gid[c] < 0) s_flag = 1;
      for (int g = tid; g <= nC; g += NT) gcnt[g] = 0;
      __syncthreads();
      if (s_flag) {
        if (tid == 0) sm->err = 1;
        return;
      }
      // CSR member lists: rank inside the group (ascending child id), sizes, offsets
      for (int c = tid; c < nC; c += NT) {
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
      for (int g = tid; g <= nG; g += NT) goff[g] = (uint16_t)gcnt[g];
      for (int c = tid; c < nC; c += NT) {
        const int p = c / K, r = c - p * K;
        const uint32_t e = xt_pack_ent(p, r + K * (int)curP[p], r);
        const int pos = gcnt[gid[c]] + grank[c];
        ent[pos] = e;
        gent[pos] = e;
        pgid[c] = (uint16_t)gid[c];
      }
      __syncthreads();  // ent visible to the CTA
    }
    const int rows_out = rows_cmp;  // history rows kept: truncated to frame_len
    int ro_sh = 0;  // rows padded to a power of two: (group, row) of a thread by shifts
    while ((1 << ro_sh) < rows_out) ++ro_sh;
    if (!VERIFY) {
    if (scapC > 0 && nG > scapP) {  // more groups than parent slots in shared memory
      if (tid == 0) sm->err = 3;
      return;
    }
    // The parents' window codes are dead since the update phase: codeP[g] becomes the window code
    // of group g, i.e. of the next step's parent g (history rows OR their argmax bits into it).
    for (int g = tid; g < nG; g += NT) codeP[g] = 0ull;
    __syncthreads();
    K1_T(3);
    // From here to the next barrier the phases are independent of each other (they read ent / gcnt /
    // bufC / histP and write disjoint outputs), so no barrier separates them and the warps overlap:
    // history rows start at thread 0, the replay records at the last thread, the merge on all warps.
    const int rtid = NT - 1 - tid;
    if (tid == 0) {
      a.plan.hdr[rec].nC = nC;
      a.plan.hdr[rec].nG = nG;
      a.plan.hdr[rec].th = th;
      a.plan.vok[rec] = (uint8_t)(nC <= 64);  // matrix mode: the verification records describe the step completely
    }
    // ---- history rows of the groups (fit mode: mean one-hot over members and leader tracks,
    //      tracking.py:714-715,735-737), accumulated in numpy's (member, track) order ----
    const int Kh = hist_dim0_is_nT ? Kt : 1;
    // one thread per (group, row): the nS values of the row, then their argmax (ties -> lowest
    // state) goes straight into the group's window code.  The sums run in numpy's order for the
    // fancy-indexed copy (member outer, track inner; adding to 0.0 first is exact), two states at
    // a time so that the two sequential chains interleave.
#ifdef XT_K1_PROF
    __shared__ unsigned long long s_hmax, s_hslow, s_hmem;
    if (tid == 0) { s_hmax = 0; s_hslow = 0; s_hmem = 0; }
    __syncthreads();
    const long long h_t0 = clock64();
    unsigned h_slow = 0, h_mem = 0, h_slowcyc = 0;
#endif
    for (int idx = tid; (idx >> ro_sh) < nG; idx += NT) {
      const int row = idx & ((1 << ro_sh) - 1), g = idx >> ro_sh;
      if (row >= rows_out) continue;
      const int o = gcnt[g], n = gcnt[g + 1] - o;
      const bool lab_row = row < nsub;
      const double* hrow = histP + (row - nsub) * nS;  // + p * pstride + state
      const int pstride = a.RH * nS;
      int best = 0;
      double bv = 0.0;
      for (int s0 = 0; s0 < nS; s0 += 2) {
        const bool two = s0 + 1 < nS;
        double accA = 0.0, accB = 0.0;
        for (int k = 0; k < n; ++k) {
          const uint32_t e = ent[o + k];
          const int p = (int)(e & 0xFFFF);
          double vA, vB = 0.0;
          if (lab_row) {  // one of the nsub newest rows: the label sits in the child's window code
            const int lab = (int)((codeC[p * K + (int)(e >> 24)] >> (bits * row)) & rowmask);
            vA = lab == s0 ? 1.0 : 0.0;
            vB = lab == s0 + 1 ? 1.0 : 0.0;
          } else {
            const double* hp = hrow + p * pstride + s0;
            vA = hp[0];
            if (two) vB = hp[1];
          }
          if (n == 1) {  // single member: the value itself (no mean)
            accA = vA;
            accB = vB;
            break;
          }
          // x is a multiple of 2^-20 below 2^31  <=>  x + 2^32 is exact.  If the summand and the partial
          // sum both are, every intermediate sum of the Kh sequential additions is exactly
          // representable, so one fused multiply-add gives the same result.
          auto dyadic = [](double x) { return x < 2147483648.0 && __dsub_rn(__dadd_rn(x, 4294967296.0), 4294967296.0) == x; };
          const bool fastA = vA == 0.0 || (dyadic(vA) && dyadic(accA));
          const bool fastB = vB == 0.0 || (dyadic(vB) && dyadic(accB));
#ifdef XT_K1_PROF
          ++h_mem;
          if (!(fastA && fastB)) ++h_slow;
#endif
          if (fastA && fastB) {  // zeros change nothing; dyadic values on dyadic partial sums are exact
            accA = fma(vA, (double)Kh, accA);
            accB = fma(vB, (double)Kh, accB);
          } else {
#ifdef XT_K1_PROF
            const long long ts0 = clock64();
#endif
#pragma unroll 6
            for (int tt = 0; tt < Kh; ++tt) {
              accA = __dadd_rn(accA, vA);
              accB = __dadd_rn(accB, vB);
            }
#ifdef XT_K1_PROF
            h_slowcyc += (unsigned)(clock64() - ts0);
#endif
          }
        }
        double outA = accA, outB = accB;
        if (n > 1) {
          const double den = (double)(Kh * n);
          outA = __ddiv_rn(accA, den);
          outB = __ddiv_rn(accB, den);
        }
        double* hn = histN + ((size_t)g * a.RH + row) * nS + s0;
        hn[0] = outA;
        if (s0 == 0 || outA > bv) {
          bv = outA;
          best = s0;
        }
        if (two) {
          hn[1] = outB;
          if (outB > bv) {
            bv = outB;
            best = s0 + 1;
          }
        }
      }
      if (best) atomicOr(&codeP[g], (unsigned long long)best << (bits * row));
    }
#ifdef XT_K1_PROF
    {
      const unsigned long long dtc = (unsigned long long)(clock64() - h_t0);
      const int myg = tid >> ro_sh;
      const int myn = myg < nG ? gcnt[myg + 1] - gcnt[myg] : 0;
      atomicMax(&s_hmax, (dtc << 32) | ((unsigned long long)(myn & 0xFF) << 24) | ((unsigned long long)(h_slow & 0xFF) << 16) | (unsigned long long)((h_slowcyc >> 4) & 0xFFFF));
    }
    atomicAdd(&s_hslow, (unsigned long long)h_slow);
    atomicAdd(&s_hmem, (unsigned long long)h_mem);
    __syncthreads();
    if (tid == 0) {
      prof[8] += s_hmax >> 32; prof[9] += s_hslow; prof[10] += s_hmem; prof[11] += nG * rows_out;
    }
#endif
    K1_T(4);
    if (a.want_grec) {  // inline group records of the first-generation replay kernel (k2_variant 1)
      unsigned long long* grec = a.plan.grec + (size_t)rec * a.plan.cap;
      for (int g = rtid; g < nG; g += NT) {
        const int o = gcnt[g], n = gcnt[g + 1] - o;
        const unsigned long long e0 = ent[o], e1 = (n > 1) ? ent[o + 1] : 0u;
        grec[g] = (e0 & 0xFFFFFFull) | ((unsigned long long)(n > 255 ? 255 : n) << 24) | ((e1 & 0xFFFFFFull) << 32);
      }
    }

    {  // replay record of this step (XtBlobHdr, xt_common.cuh).  Schedule: the groups sorted by (members
       // descending, group ascending) are dealt to the replay warps - up to 64 groups and 4 replay warps:
       // longest-processing-time-first on a cost model of the replay kernel (the warps of a tile meet at one
       // barrier per step, so the step lasts as long as its most loaded warp); otherwise round-robin.
      uint4* blob = a.plan.blob + (size_t)rec * xt_blob_stride16(a.plan.cap);
      XtBlobHdr* h = (XtBlobHdr*)blob;
      const int wpc = a.wpc;
      const bool lpt = a.lpt && nG <= 64 && wpc <= 4;
      // warp of the schedule: with 16 or more warps one that neither holds history rows (first warps) nor merges (upper half)
      const int rec_warp = (W >= 16 && (nG << ro_sh) <= NT / 4) ? W / 2 - 1 : W - 1;
      unsigned long long* brec = (unsigned long long*)(blob + 2);
      uint8_t* pcur = a.plan.curG + (size_t)rec * a.plan.cap;
      // group record: fields pre-positioned for the replay kernel (xt_common.cuh)
      auto group_record = [&](int g, int o, int n) -> unsigned long long {
        const uint32_t e0 = ent[o];
        const unsigned lo = ((e0 >> 16) & 0x7Fu) | ((e0 & 0xFFFu) << 7) | ((unsigned)g << 19);
        unsigned hi;
        if (n == 1) {
          hi = 1u << 30;
        } else if (n == 2) {
          const uint32_t e1 = ent[o + 1];
          hi = (2u << 30) | ((e1 >> 16) & 0x7Fu) | ((e1 & 0xFFFu) << 7);
        } else {
          hi = (3u << 30) | (unsigned)o | ((unsigned)n << 12);
        }
        return (unsigned long long)lo | ((unsigned long long)hi << 32);
      };
      if (rtid == 0) {
        h->nG = (uint16_t)nG;
        h->nC = (uint16_t)nC;
        h->n16 = (uint16_t)(2 + (nG + 1) / 2 + (nC + 3) / 4);
      }
      if (lpt) {
        // one warp (the last of the CTA; its threads have rtid 0..31): lane l owns groups l and l + 32.  The
        // greedy assignment runs redundantly on every lane (warp-uniform registers), each lane keeps the
        // (warp, position) of its own groups.
        if (warp == rec_warp) {
          const int l = lane;
          int gn[2] = {0, 0}, go[2] = {0, 0};
          int multi = 0;
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            const int g = l + 32 * hh;
            if (g < nG) {
              go[hh] = gcnt[g];
              gn[hh] = gcnt[g + 1] - go[hh];
              const uint32_t e = ent[go[hh]];
              const unsigned char cs = (unsigned char)(((int)(e & 0xFFFF) * K + (int)(e >> 24)) % nS);
              pcur[g] = cs;  // new parent g: newest true state = its first member's (tracking.py:728); curP is
              curP[g] = cs;  // not read again before the barrier that ends the step
              int rank = 0;
              for (int j = 0; j < nG; ++j) {
                const int nj = gcnt[j + 1] - gcnt[j];
                rank += (nj > gn[hh]) || (nj == gn[hh] && j < g);
                if (hh == 0) multi += nj > 1;
              }
              grank[rank] = g;  // (grank is free outside the split-mode grouping)
            }
          }
          __syncwarp();
          // (replay warp 0 also stages the next record and the next localisation's prefetch bookkeeping: a head start for the others)
          int ld0 = a.cost_w0, ld1 = 0, ld2 = 0, ld3 = 0, c0 = 0, c1 = 0, c2 = 0, c3 = 0, m0 = 0, m1 = 0, m2 = 0, m3 = 0;
          int myw[2] = {0, 0}, myp[2] = {0, 0};
          for (int r = 0; r < nG; ++r) {
            const int g = grank[r];
            const int n = gcnt[g + 1] - gcnt[g];
            const int cost = n == 1 ? a.cost[0] : (n == 2 ? a.cost[1] : a.cost[2] + a.cost[3] * n);  // replay cost model
            int w = 0, best = ld0;
            if (wpc > 1 && ld1 < best) { w = 1; best = ld1; }
            if (wpc > 2 && ld2 < best) { w = 2; best = ld2; }
            if (wpc > 3 && ld3 < best) { w = 3; best = ld3; }
            const int pos = w == 0 ? c0 : (w == 1 ? c1 : (w == 2 ? c2 : c3));
            if (g == l) { myw[0] = w; myp[0] = pos; }
            if (g == l + 32) { myw[1] = w; myp[1] = pos; }
            const int mm = n > 1;
            c0 += w == 0; c1 += w == 1; c2 += w == 2; c3 += w == 3;
            ld0 += w == 0 ? cost : 0; ld1 += w == 1 ? cost : 0; ld2 += w == 2 ? cost : 0; ld3 += w == 3 ? cost : 0;
            m0 += (w == 0) & mm; m1 += (w == 1) & mm; m2 += (w == 2) & mm; m3 += (w == 3) & mm;
          }
          const int o1 = c0, o2 = c0 + c1, o3 = o2 + c2;
          if (l == 0) {
            h->nM = (uint16_t)multi;
            h->woff[0] = 0;
            h->woff[1] = (uint16_t)o1;
            h->woff[2] = (uint16_t)(wpc > 1 ? o2 : nG);
            h->woff[3] = (uint16_t)(wpc > 2 ? o3 : nG);
            h->woff[4] = (uint16_t)nG;
            h->woff[5] = (uint16_t)m0;  // multi-member groups at the head of every warp's list (XT_MAX_WPC + 1 - 4 spare entries)
            h->woff[6] = (uint16_t)m1;
            h->woff[7] = (uint16_t)m2;
            h->woff[8] = (uint16_t)m3;
          }
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            const int g = l + 32 * hh;
            if (g < nG) {
              const int base = myw[hh] == 0 ? 0 : (myw[hh] == 1 ? o1 : (myw[hh] == 2 ? o2 : o3));
              brec[base + myp[hh]] = group_record(g, go[hh], gn[hh]);
            }
          }
        }
      } else {
        if (rtid <= XT_MAX_WPC) {
          // groups of warp w: ranks w, w + wpc, ...  => woff[w] = sum_{v<w} ceil((nG - v) / wpc)
          int o = 0;
          for (int v = 0; v < rtid && v < wpc; ++v) o += (nG - v + wpc - 1) / wpc;
          if (wpc > 4 || rtid <= 4) h->woff[rtid] = (uint16_t)o;
        }
        for (int g = rtid; g < nG; g += NT) {
          const int o = gcnt[g], n = gcnt[g + 1] - o;
          {  // new parent g: newest true state = its first member's (tracking.py:728); curP is not read
             // again before the barrier that ends the step
            const uint32_t e = ent[o];
            const unsigned char cs = (unsigned char)(((int)(e & 0xFFFF) * K + (int)(e >> 24)) % nS);
            pcur[g] = cs;
            curP[g] = cs;
          }
          int rank = 0, multi = 0;
          for (int j = 0; j < nG; ++j) {
            const int nj = gcnt[j + 1] - gcnt[j];
            rank += (nj > n) || (nj == n && j < g);
            multi += nj > 1;
          }
          if (rank == nG - 1) {  // one writer: the last group of the schedule
            h->nM = (uint16_t)multi;
            if (wpc <= 4)  // multi-member groups at the head of warp w's list: ranks w, w + wpc, ... below `multi`
              for (int w2 = 0; w2 < 4; ++w2) h->woff[5 + w2] = (uint16_t)(multi > w2 ? (multi - w2 + wpc - 1) / wpc : 0);
          }
          const int wq = rank % wpc, pos = rank / wpc;
          int slot = pos;
          for (int v = 0; v < wq; ++v) slot += (nG - v + wpc - 1) / wpc;
          brec[slot] = group_record(g, o, n);
        }
      }
      uint32_t* bent = (uint32_t*)(blob + 2 + (nG + 1) / 2);
      for (int c = rtid; c < nC; c += NT) bent[c] = ent[c];
    }
    }  // !VERIFY

    K1_T(5);
    // ---- merge on the leader tracks (tracking.py:723-741) ----
    // `LP[:, subgroup]` is an F-ordered fancy-index copy in numpy, so every reduction over the
    // members runs sequentially in ascending member order (checked against numpy 2.3).
    // (warps whose threads hold history rows merge fewer groups: with <= 128 rows the merge runs
    // on the upper half of the CTA)
    const int mw0 = (!VERIFY && (nG << ro_sh) <= NT / 2) ? W / 2 : 0, mW = W - mw0;
    for (int g = warp - mw0; g >= 0 && g < nG; g += mW) {
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
      double sw = 0.0, am[D], as2[KS];
      for (int k0 = 0; k0 < n; k0 += 4) {  // weights of four members at a time (independent exps);
        double w4[4], m4[4][D], s4[4][KS];  // the sums still run in ascending member order
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int c = child(k0 + q < n ? k0 + q : k0);
          w4[q] = exp(__dsub_rn(ST(bufC, c, D + 2 * KS), mx));
#pragma unroll
          for (int dim = 0; dim < D; ++dim) m4[q][dim] = ST(bufC, c, dim);
#pragma unroll
          for (int k2 = 0; k2 < KS; ++k2) s4[q][k2] = ST(bufC, c, D + k2);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (k0 + q < n) {
            const bool first = (k0 + q) == 0;
            sw = first ? w4[q] : __dadd_rn(sw, w4[q]);
#pragma unroll
            for (int dim = 0; dim < D; ++dim) {
              const double v = __dmul_rn(w4[q], m4[q][dim]);
              am[dim] = first ? v : __dadd_rn(am[dim], v);
            }
#pragma unroll
            for (int k2 = 0; k2 < KS; ++k2) {
              const double v = __dmul_rn(w4[q], s4[q][k2]);
              as2[k2] = first ? v : __dadd_rn(as2[k2], v);
            }
          }
        }
      }
#pragma unroll
      for (int dim = 0; dim < D; ++dim) ST(bufP, g, dim) = __ddiv_rn(am[dim], sw);
#pragma unroll
      for (int k2 = 0; k2 < KS; ++k2) ST(bufP, g, D + k2) = __ddiv_rn(as2[k2], sw);
      ST(bufP, g, D + 2 * KS) = __dadd_rn(log(sw), mx);
    }
    K1_T(6);
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
    K1_T(7);
  }
#ifdef XT_K1_PROF
  if (tid == 0 && a.prof) for (int i = 0; i < 12; ++i) a.prof[(size_t)cid * 12 + i] = prof[i];
#endif
  if (VERIFY) {
    if (tid == 0) a.vflag[cid] = 0;
    return;
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
#undef DDH
}
