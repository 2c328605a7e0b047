// Shared device-side definitions of the B200 ExTrack likelihood engine.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "xtrack.h"

#define XT_LEADERS 30      // tracking.py:677  test_chunks = 30
#define XT_HARD_CAP 4096   // most live sequences after an expansion the engine accepts
#define XT_K1_THREADS 256
#define XT_K1_MIN_CTAS 3
#define XT_TWO_PI 6.283185307179586  // 2*np.pi

// One chunk = <= chunk_size tracks of one length bucket (tracking.py:1030-1036).
struct XtChunk {
  int32_t L;         // localisations per track
  int32_t nT;        // tracks in the chunk
  int32_t nTpad;     // nT rounded up to 32 (row pitch of the SoA block)
  int32_t isBL;      // 0 for the longest bucket
  int64_t xyz_off;   // offset (doubles) of the chunk's SoA block [L][d][nTpad]
  int64_t trk_off;   // offset of the chunk's first track in per-track outputs
  int64_t loc_off;   // offset of the chunk's first localisation in per-localisation outputs
  int32_t rec0;      // first plan record of the chunk (records for steps 2..L-2)
  int32_t nrec;      // max(0, L-3)
  int32_t seg;       // uploaded segment the chunk came from
  int32_t seg_t0;    // index of the chunk's first track inside that segment
};

// Per-chunk summary written by the plan kernel, read back by the host once per evaluation.
struct XtChunkSummary {
  int32_t err;        // 0 ok, 1 grouping failure, 2 capacity overflow
  int32_t need_cap;   // on overflow: children capacity that would have been needed
  int32_t max_nP;     // most parent sequences at the start of any step (incl. the initial K*nS)
  int32_t max_nC;     // most children after any expansion
  int64_t sum_nC;     // sum over steps 2..L-1 of nC (+ end-of-track expansion if isBL)
  int64_t sum_nG;     // sum over fused steps of nG
};

// Plan record header of one fusion step.
struct XtRecHdr {
  int32_t nC;   // sequences after the expansion (nB_in)
  int32_t nG;   // groups after the fusion
  double th;    // threshold used at this step
};

// Replay record of one fusion step, consumed by the fused replay kernel (xt_replay_fused.cuh):
// one contiguous blob per record so that a CTA can stage the next step's record in shared
// memory while it computes the current one.  Layout in 16-byte words:
//   words 0-1 : XtBlobHdr
//   words 2.. : nG group records (8 bytes each) in *schedule order*: the groups of replay warp w
//               are entries woff[w] .. woff[w+1]-1 (groups sorted by member count, descending,
//               dealt round-robin to the warps; nM of the nG groups have more than one member)
//   then      : nC member entries (4 bytes each, xt_pack_ent) in CSR order, for groups of > 2 members
// Group record (fields pre-positioned so that the replay kernel gets byte offsets with one mask
// each): lo = head0:7 | p0:12 << 7 | g:12 << 19; hi = kind:2 (1 single, 2 pair, 3 list) << 30 and
//   pair: head1:7 | p1:12 << 7;   list: first member offset:12 | member count:13 << 12.
#define XT_MAX_WPC 8
struct XtBlobHdr {
  uint16_t nG, nC;
  uint16_t n16;   // 16-byte words of this record
  uint16_t nM;    // groups with more than one member: they come first in the schedule order, so replay warp w
                  // owns ceil((nM - w) / wpc) of them at the head of its list and single-member groups after
  uint16_t woff[XT_MAX_WPC + 1];  // <= 4 replay warps: woff[5 + w] = multi-member groups at the head of warp w's list
  uint16_t pad2_[3];
};
static_assert(sizeof(XtBlobHdr) == 32, "XtBlobHdr must be two 16-byte words");
__host__ __device__ inline int xt_blob_stride16(int cap) { return 2 + (cap + 1) / 2 + (cap + 3) / 4; }

// entry of the CSR member list: parent slot | head << 16 | r << 24
__host__ __device__ inline uint32_t xt_pack_ent(int p, int head, int r) {
  return (uint32_t)p | ((uint32_t)head << 16) | ((uint32_t)r << 24);
}

// Verification record of one leader of a matrix-mode fusion step (<= 64 sequences): the sequences whose
// floating-point predicate (m_mask and s_mask, tracking.py:689-691) the greedy loop consulted when it visited this
// leader - the not yet grouped sequences with the leader's newest state that the window test did not already
// capture - and the outcome of every one of them.  The plan kernel in verification mode (k1_plan<.., VERIFY>)
// re-evaluates exactly these predicates with the parameters of a new evaluation: the plan is unchanged if and only if
// every outcome is.
struct XtVRec {
  unsigned long long lead;  // sequence id of the leader
  unsigned long long cand;  // bit j: the pair (leader, j) was decided by the floating-point predicate
  unsigned long long exp;   // its outcome (bit j set: captured)
};
#define XT_VREC_PER_STEP 64

struct XtPlanPtrs {
  XtRecHdr* hdr;     // [nrec_total]
  uint16_t* goff;    // [nrec_total][cap+1]
  uint32_t* ent;     // [nrec_total][cap]   members sorted by group, ascending child id inside a group
  uint8_t* curG;     // [nrec_total][cap]   newest true state of each group's representative
  uint16_t* gid;     // [nrec_total][cap]   group of each incoming child (dump / tests)
  unsigned long long* grec;  // [nrec_total][cap] per group: p0:16|head0:8|n:8 | (p1:16|head1:8)<<32, n capped at 255
  uint4* blob;       // [nrec_total][xt_blob_stride16(cap)] replay records (see XtBlobHdr)
  XtVRec* vrec;      // [nrec_total][XT_VREC_PER_STEP] verification records (matrix-mode steps)
  uint8_t* vok;      // [nrec_total] 1: the step was grouped in matrix mode and vrec describes it completely
  int32_t cap;
};

// Optional per-localisation inputs (xt_upload_aux) and field-of-view tables (xt_set_stay_tables)
// of evaluations with XT_FLAG_VAR_LOC / XT_FLAG_VAR_DT.  Per chunk the block [L][R][nTpad] sits at
// (xyz_off / d) * R: rows 0..ka-1 = sigma components, row ka = dt stored time-reversed, so that
// the row of localisation j holds what the reference reads while it consumes C[j]
// (tracking.py:495,:526,:549,:563,:634).
struct XtAux {
  const double* aux;
  const double* stay;    // [n_chunks | n_tracks][K] Lp_stay, nullptr: xt_params::Lp_stay
  const double* leave;   // [n_chunks | n_tracks][H] L_leave (plan / predict) or [..][nS] exp-sums (replay)
  int32_t R, ka, d;
};

// sigma -> LocErr^2 exactly as the host does (extract_params :926-930, LocErr**2 :457)
__device__ __forceinline__ double xt_sigma2f(uint32_t flags, double slope, double offset, double sg) {
  if (flags & XT_FLAG_LOC_AFFINE) {
    sg = __dadd_rn(__dmul_rn(sg, slope), offset);
    if (sg < 0.000001) sg = 0.000001;  // np.clip(.., 0.000001, inf); NaN stays NaN
  }
  return __dmul_rn(sg, sg);
}
__device__ __forceinline__ double xt_sigma2(const xt_params& P, double sg) {
  return xt_sigma2f(P.flags, P.loc_slope, P.loc_offset, sg);
}

// mean mid-sub-step displacement variance of `head` for the time step dtv, in numpy's operation
// order: ds = sqrt(2 D dt) (:979-982), ds**2, (d2[k+1] + d2[k]) / 2, mean over the sub-steps (:549-553)
__device__ __forceinline__ double xt_dd_exact(const xt_params& P, int head, double dtv) {
  const int nS = P.nS, nsub = P.nsub;
  int x = head;
  double r0 = __dsqrt_rn(__dmul_rn(P.twoD[x % nS], dtv));
  double prev = __dmul_rn(r0, r0), sum = 0.0;
  x /= nS;
  for (int k = 0; k < nsub; ++k) {
    const double r1 = __dsqrt_rn(__dmul_rn(P.twoD[x % nS], dtv));
    const double cur = __dmul_rn(r1, r1);
    x /= nS;
    const double pair = __dmul_rn(__dadd_rn(cur, prev), 0.5);
    sum = (k == 0) ? pair : __dadd_rn(sum, pair);
    prev = cur;
  }
  return nsub == 1 ? sum : __ddiv_rn(sum, (double)nsub);
}

// ---- shared memory through 32-bit shared-window addresses (no generic-pointer arithmetic) ----
__device__ __forceinline__ unsigned xt_smem_base(const void* p) {
  unsigned a = (unsigned)__cvta_generic_to_shared(p);
  asm volatile("mov.u32 %0, %0;" : "+r"(a));  // computed once: keeps the base from being rematerialised
  return a;
}
__device__ __forceinline__ void xt_lds128(unsigned a, double& x, double& y) {
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x), "=d"(y) : "r"(a));
}
__device__ __forceinline__ double xt_lds64(unsigned a) {
  double x;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(x) : "r"(a));
  return x;
}
__device__ __forceinline__ uint2 xt_lds64u(unsigned a) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ int xt_lds32(unsigned a) {
  int x;
  asm volatile("ld.shared.s32 %0, [%1];" : "=r"(x) : "r"(a));
  return x;
}
__device__ __forceinline__ unsigned xt_lds16(unsigned a) {
  unsigned x;
  asm volatile("ld.shared.u16 %0, [%1];" : "=r"(x) : "r"(a));
  return x;
}
__device__ __forceinline__ void xt_sts128(unsigned a, double x, double y) {
  asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a), "d"(x), "d"(y) : "memory");
}
__device__ __forceinline__ void xt_sts128u(unsigned a, uint4 v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void xt_sts64(unsigned a, double x) {
  asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(x) : "memory");
}
__device__ __forceinline__ void xt_sts32(unsigned a, int x) {
  asm volatile("st.shared.s32 [%0], %1;" ::"r"(a), "r"(x) : "memory");
}

struct XtWork {  // one replay CTA: 32 tracks of a chunk
  int32_t chunk;
  int32_t t0;
};

#define XT_CUDA_OK(call)                                                                         \
  do {                                                                                           \
    cudaError_t e__ = (call);                                                                    \
    if (e__ != cudaSuccess) {                                                                    \
      set_error(ctx, std::string(#call) + ": " + cudaGetErrorString(e__));                       \
      return XT_ERR_CUDA;                                                                        \
    }                                                                                            \
  } while (0)
