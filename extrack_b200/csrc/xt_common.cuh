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
//               dealt round-robin to the warps)
//   then      : nC member entries (4 bytes each, xt_pack_ent) in CSR order, for groups of > 2 members
// Group record: lo = p0:12 | head0:8 | g:12; hi = kind:2 (1 single, 2 pair, 3 list) << 30 and
//   pair: p1:12 | head1:8 << 12;   list: first member offset:12 | member count:13 << 12.
#define XT_MAX_WPC 8
struct XtBlobHdr {
  uint16_t nG, nC;
  uint16_t n16;   // 16-byte words of this record
  uint16_t pad_;
  uint16_t woff[XT_MAX_WPC + 1];
  uint16_t pad2_[3];
};
static_assert(sizeof(XtBlobHdr) == 32, "XtBlobHdr must be two 16-byte words");
__host__ __device__ inline int xt_blob_stride16(int cap) { return 2 + (cap + 1) / 2 + (cap + 3) / 4; }

// entry of the CSR member list: parent slot | head << 16 | r << 24
__host__ __device__ inline uint32_t xt_pack_ent(int p, int head, int r) {
  return (uint32_t)p | ((uint32_t)head << 16) | ((uint32_t)r << 24);
}

struct XtPlanPtrs {
  XtRecHdr* hdr;     // [nrec_total]
  uint16_t* goff;    // [nrec_total][cap+1]
  uint32_t* ent;     // [nrec_total][cap]   members sorted by group, ascending child id inside a group
  uint8_t* curG;     // [nrec_total][cap]   newest true state of each group's representative
  uint16_t* gid;     // [nrec_total][cap]   group of each incoming child (dump / tests)
  unsigned long long* grec;  // [nrec_total][cap] per group: p0:16|head0:8|n:8 | (p1:16|head1:8)<<32, n capped at 255
  uint4* blob;       // [nrec_total][xt_blob_stride16(cap)] replay records (see XtBlobHdr)
  int32_t cap;
};

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
