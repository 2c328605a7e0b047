// Host side of libxtrack_b200.so: context, upload + device repack, evaluation driver, C ABI.
// See include/xtrack.h for the contract of every entry point.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "xt_common.cuh"
#include "xt_plan.cuh"
#include "xt_replay.cuh"
#include "xt_replay_lin.cuh"
#include "xt_replay_fused.cuh"
#include "xt_predict.cuh"

struct xt_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  // data
  int d = 0;
  int64_t n_tracks = 0, track_steps = 0, n_locs = 0;
  std::vector<int> seg_chunk0;
  std::vector<int64_t> upload_sig;  // shapes of the resident data set (allocation reuse)
  double* stage[2] = {nullptr, nullptr};
  cudaEvent_t stage_done[2] = {nullptr, nullptr};
  size_t stage_elems = 0;
  std::vector<int64_t> seg_n;
  std::vector<int> seg_L;
  std::vector<XtChunk> chunks;
  std::vector<XtWork> work;
  int nrec_total = 0;
  int maxL = 0;
  double* d_soa = nullptr;
  XtChunk* d_chunks = nullptr;
  XtWork* d_work = nullptr;
  XtWork* d_workf[2] = {nullptr, nullptr};  // fused replay: tiles of 32 / 64 tracks, longest chunks first
  int n_workf[2] = {0, 0};
  double* d_logp = nullptr;
  double* d_partial = nullptr;
  double* d_out = nullptr;
  XtChunkSummary* d_summ = nullptr;
  std::vector<XtChunkSummary> summ;
  XtChunkSummary* h_summ = nullptr;  // pinned
  double* h_out = nullptr;           // pinned
  // plan storage / K1 scratch (sized by cap and the model)
  int cap = 0, RH = 0, nS_alloc = 0, CO1_alloc = 0;
  XtPlanPtrs plan{};
  double* d_state1 = nullptr;
  double* d_hist1 = nullptr;
  // K2 global-state fallback
  double* d_gstate = nullptr;
  size_t gstate_bytes = 0;
  int smem_optin = 0, n_sm = 0;
  bool have_eval = false;
  bool force_global = false;  // test hook: run the log-domain global-memory replay variant
  int k2_wpc = 4;             // warps cooperating on one 32-track tile in the fast replay kernel
  int k2_tpt = 1;             // tracks per thread of the fused replay kernel (tile = 32 * k2_tpt tracks)
  int k2_variant = 0;         // 0: fused merge+update kernel (default), 1: first-generation linear-domain kernel
  xt_params last_p{};
  xt_stats stats{};
  cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
};

static std::string g_create_error;
static void set_error(xt_ctx* ctx, const std::string& s) {
  if (ctx) ctx->err = s; else g_create_error = s;
}

// ------------------------------------------------------------------------------------------
// repack: AoS [n][L][d] (host order) -> per-chunk SoA [L][d][nTpad]
// ------------------------------------------------------------------------------------------
__global__ void k_pack(const double* __restrict__ src, double* __restrict__ soa, const XtChunk* __restrict__ chunks,
                       int chunk0, int chunk_size, int n, int L, int d) {
  // one thread per (row = k*d+dim, track i) with i fastest => coalesced writes
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int rows = L * d;
  if (idx >= (long long)n * rows) return;
  const int i = (int)(idx % n);
  const int row = (int)(idx / n);
  const int c = i / chunk_size, t = i - c * chunk_size;
  const XtChunk ck = chunks[chunk0 + c];
  soa[ck.xyz_off + (size_t)row * ck.nTpad + t] = src[(size_t)i * rows + row];
}

__global__ void k_fp64_peak(double* out, int iters) {
  double a0 = 1.0 + threadIdx.x * 1e-9, a1 = a0 + 1e-3, a2 = a0 + 2e-3, a3 = a0 + 3e-3;
  double a4 = a0 + 4e-3, a5 = a0 + 5e-3, a6 = a0 + 6e-3, a7 = a0 + 7e-3;
  const double b = 0.999999, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
    a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

static void free_plan(xt_ctx* ctx) {
  cudaFree(ctx->plan.hdr); cudaFree(ctx->plan.goff); cudaFree(ctx->plan.ent);
  cudaFree(ctx->plan.curG); cudaFree(ctx->plan.gid); cudaFree(ctx->plan.grec); cudaFree(ctx->plan.blob);
  cudaFree(ctx->d_state1); cudaFree(ctx->d_hist1);
  ctx->plan = XtPlanPtrs{};
  ctx->d_state1 = ctx->d_hist1 = nullptr;
  ctx->cap = 0;
}

static void free_data(xt_ctx* ctx) {
  cudaFree(ctx->d_soa); cudaFree(ctx->d_chunks); cudaFree(ctx->d_work); cudaFree(ctx->d_logp);
  cudaFree(ctx->d_partial); cudaFree(ctx->d_summ); cudaFree(ctx->d_gstate);
  cudaFree(ctx->d_workf[0]); cudaFree(ctx->d_workf[1]);
  ctx->d_workf[0] = ctx->d_workf[1] = nullptr;
  if (ctx->h_summ) cudaFreeHost(ctx->h_summ);
  ctx->d_soa = ctx->d_logp = ctx->d_partial = ctx->d_gstate = nullptr;
  ctx->d_chunks = nullptr; ctx->d_work = nullptr; ctx->d_summ = nullptr; ctx->h_summ = nullptr;
  ctx->gstate_bytes = 0;
  ctx->chunks.clear(); ctx->work.clear(); ctx->summ.clear(); ctx->upload_sig.clear();
  ctx->have_eval = false;
  free_plan(ctx);
}

extern "C" int xt_create(int device, xt_ctx** out) {
  xt_ctx* ctx = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    set_error(nullptr, std::string("no CUDA device: ") + cudaGetErrorString(e));
    return XT_ERR_CUDA;
  }
  if (device < 0 || device >= ndev) {
    set_error(nullptr, "device ordinal out of range");
    return XT_ERR_ARG;
  }
  xt_ctx* c = new xt_ctx();
  c->device = device;
  ctx = c;
  XT_CUDA_OK(cudaSetDevice(device));
  XT_CUDA_OK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  XT_CUDA_OK(cudaMalloc(&c->d_out, sizeof(double)));
  XT_CUDA_OK(cudaMallocHost(&c->h_out, sizeof(double)));
  for (int i = 0; i < 3; ++i) XT_CUDA_OK(cudaEventCreate(&c->ev[i]));
  XT_CUDA_OK(cudaDeviceGetAttribute(&c->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
  XT_CUDA_OK(cudaDeviceGetAttribute(&c->n_sm, cudaDevAttrMultiProcessorCount, device));
  *out = c;
  return XT_OK;
}

extern "C" void xt_destroy(xt_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  free_data(ctx);
  for (int b = 0; b < 2; ++b) {
    cudaFree(ctx->stage[b]);
    if (ctx->stage_done[b]) cudaEventDestroy(ctx->stage_done[b]);
  }
  cudaFree(ctx->d_out);
  if (ctx->h_out) cudaFreeHost(ctx->h_out);
  for (int i = 0; i < 3; ++i) if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
  cudaStreamDestroy(ctx->stream);
  delete ctx;
}

extern "C" const char* xt_last_error(xt_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

extern "C" int xt_host_alloc(void** out, uint64_t bytes) {
  return cudaMallocHost(out, bytes) == cudaSuccess ? XT_OK : XT_ERR_CUDA;
}
extern "C" int xt_host_free(void* p) { return cudaFreeHost(p) == cudaSuccess ? XT_OK : XT_ERR_CUDA; }

extern "C" int xt_upload(xt_ctx* ctx, int32_t n_seg, const int32_t* L, const int64_t* n, const int32_t* isBL,
                         const double* const* xyz, int32_t d, int32_t chunk_size) {
  if (!ctx) return XT_ERR_ARG;
  if (n_seg <= 0 || d < 1 || d > XT_MAX_DIMS || chunk_size < 1) {
    set_error(ctx, "xt_upload: need n_seg >= 1, 1 <= d <= 3, chunk_size >= 1");
    return XT_ERR_ARG;
  }
  for (int s = 0; s < n_seg; ++s) {
    if (L[s] < 2) {
      set_error(ctx, "minimal track length = 2, here track length = " + std::to_string(L[s]));
      return XT_ERR_ARG;
    }
    if (n[s] < 1) {
      set_error(ctx, "xt_upload: empty segment");
      return XT_ERR_ARG;
    }
  }
  XT_CUDA_OK(cudaSetDevice(ctx->device));
  XT_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  // Same shapes as the resident data set (a re-upload of new coordinates, e.g. one objective call
  // per host buffer): keep every allocation, the chunk/work tables and the plan storage.
  std::vector<int64_t> sig;
  sig.push_back(n_seg); sig.push_back(d); sig.push_back(chunk_size);
  for (int s = 0; s < n_seg; ++s) { sig.push_back(L[s]); sig.push_back(n[s]); sig.push_back(isBL[s]); }
  const bool reuse = !ctx->chunks.empty() && sig == ctx->upload_sig;
  int64_t max_seg_elems = 0;
  for (int s = 0; s < n_seg; ++s) max_seg_elems = std::max<int64_t>(max_seg_elems, n[s] * L[s] * d);
  if (!reuse) {
    free_data(ctx);
    ctx->upload_sig = sig;
    ctx->d = d;
    ctx->n_tracks = 0;
    ctx->n_locs = 0;
    ctx->track_steps = 0;
    ctx->maxL = 0;
    int64_t soa_elems = 0;
    int rec = 0;
    ctx->seg_chunk0.assign(n_seg, 0);
    for (int s = 0; s < n_seg; ++s) {
      ctx->seg_chunk0[s] = (int)ctx->chunks.size();
      ctx->maxL = std::max(ctx->maxL, (int)L[s]);
      for (int64_t a = 0; a < n[s]; a += chunk_size) {
        XtChunk ck{};
        ck.L = L[s];
        ck.nT = (int)std::min<int64_t>(chunk_size, n[s] - a);
        ck.nTpad = (ck.nT + 31) & ~31;
        ck.isBL = isBL[s];
        ck.xyz_off = soa_elems;
        ck.trk_off = ctx->n_tracks;
        ck.loc_off = ctx->n_locs;
        ck.rec0 = rec;
        ck.nrec = std::max(0, ck.L - 3);
        ck.seg = s;
        ck.seg_t0 = (int)a;
        rec += ck.nrec;
        soa_elems += (int64_t)ck.L * d * ck.nTpad;
        ctx->n_tracks += ck.nT;
        ctx->n_locs += (int64_t)ck.nT * ck.L;
        ctx->track_steps += (int64_t)ck.nT * (ck.L - 1);
        for (int t0 = 0; t0 < ck.nT; t0 += 32) ctx->work.push_back(XtWork{(int)ctx->chunks.size(), t0});
        ctx->chunks.push_back(ck);
      }
    }
    ctx->nrec_total = rec;
    ctx->seg_n.assign(n, n + n_seg);
    ctx->seg_L.assign(L, L + n_seg);
    const size_t nch = ctx->chunks.size();
    XT_CUDA_OK(cudaMalloc(&ctx->d_soa, sizeof(double) * (size_t)soa_elems));
    XT_CUDA_OK(cudaMemsetAsync(ctx->d_soa, 0, sizeof(double) * (size_t)soa_elems, ctx->stream));  // padding lanes
    XT_CUDA_OK(cudaMalloc(&ctx->d_chunks, sizeof(XtChunk) * nch));
    XT_CUDA_OK(cudaMalloc(&ctx->d_work, sizeof(XtWork) * ctx->work.size()));
    XT_CUDA_OK(cudaMalloc(&ctx->d_logp, sizeof(double) * (size_t)ctx->n_tracks));
    XT_CUDA_OK(cudaMalloc(&ctx->d_partial, sizeof(double) * ctx->work.size()));
    XT_CUDA_OK(cudaMalloc(&ctx->d_summ, sizeof(XtChunkSummary) * nch));
    XT_CUDA_OK(cudaMallocHost(&ctx->h_summ, sizeof(XtChunkSummary) * nch));
    ctx->summ.resize(nch);
    XT_CUDA_OK(cudaMemcpyAsync(ctx->d_chunks, ctx->chunks.data(), sizeof(XtChunk) * nch, cudaMemcpyHostToDevice,
                               ctx->stream));
    XT_CUDA_OK(cudaMemcpyAsync(ctx->d_work, ctx->work.data(), sizeof(XtWork) * ctx->work.size(),
                               cudaMemcpyHostToDevice, ctx->stream));
    // work tables of the fused replay kernel: longest chunks first (the last CTAs of the launch
    // are then the short ones: smaller tail)
    std::vector<int> order(nch);
    for (size_t c = 0; c < nch; ++c) order[c] = (int)c;
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return ctx->chunks[x].L > ctx->chunks[y].L; });
    for (int v = 0; v < 2; ++v) {
      const int tile = 32 << v;
      std::vector<XtWork> wf;
      for (int c : order)
        for (int t0 = 0; t0 < ctx->chunks[c].nT; t0 += tile) wf.push_back(XtWork{c, t0});
      ctx->n_workf[v] = (int)wf.size();
      XT_CUDA_OK(cudaMalloc(&ctx->d_workf[v], sizeof(XtWork) * wf.size()));
      XT_CUDA_OK(cudaMemcpy(ctx->d_workf[v], wf.data(), sizeof(XtWork) * wf.size(), cudaMemcpyHostToDevice));
    }
  }
  ctx->have_eval = false;
  // stage each segment (AoS) and repack on the device; two staging buffers overlap copy and pack
  if ((size_t)max_seg_elems > ctx->stage_elems) {
    for (int b = 0; b < 2; ++b) {
      cudaFree(ctx->stage[b]);
      ctx->stage[b] = nullptr;
      XT_CUDA_OK(cudaMalloc(&ctx->stage[b], sizeof(double) * (size_t)max_seg_elems));
      if (!ctx->stage_done[b]) XT_CUDA_OK(cudaEventCreateWithFlags(&ctx->stage_done[b], cudaEventDisableTiming));
    }
    ctx->stage_elems = (size_t)max_seg_elems;
  }
  for (int s = 0; s < n_seg; ++s) {
    const int b = s & 1;
    const size_t elems = (size_t)n[s] * L[s] * d;
    XT_CUDA_OK(cudaEventSynchronize(ctx->stage_done[b]));
    XT_CUDA_OK(cudaMemcpyAsync(ctx->stage[b], xyz[s], sizeof(double) * elems, cudaMemcpyHostToDevice, ctx->stream));
    const int threads = 256;
    const long long blocks = ((long long)elems + threads - 1) / threads;
    k_pack<<<(unsigned)blocks, threads, 0, ctx->stream>>>(ctx->stage[b], ctx->d_soa, ctx->d_chunks, ctx->seg_chunk0[s],
                                                          chunk_size, (int)n[s], L[s], d);
    XT_CUDA_OK(cudaEventRecord(ctx->stage_done[b], ctx->stream));
  }
  XT_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  XT_CUDA_OK(cudaGetLastError());
  return XT_OK;
}

// ------------------------------------------------------------------------------------------
// evaluation
// ------------------------------------------------------------------------------------------
static int ipow(int b, int e) { int r = 1; for (int i = 0; i < e; ++i) r *= b; return r; }

static int check_params(xt_ctx* ctx, const xt_params* p, int* bits_out) {
  if (!p || p->nS < 1 || p->nS > XT_MAX_STATES || p->nsub < 1 || p->d != ctx->d ||
      (p->n_loc != 1 && p->n_loc != p->d) || p->frame_len < 1) {
    set_error(ctx, "xt_params: inconsistent model (nS, nsub, d, n_loc or frame_len)");
    return XT_ERR_ARG;
  }
  long long heads = 1;
  for (int i = 0; i <= p->nsub; ++i) heads *= p->nS;
  if (heads > XT_MAX_HEADS) {
    set_error(ctx, "xt_params: nS^(nb_substeps+1) exceeds XT_MAX_HEADS");
    return XT_ERR_ARG;
  }
  const int bits = p->nS <= 2 ? 1 : (p->nS <= 4 ? 2 : 3);
  if ((long long)bits * std::max(p->frame_len, p->nsub + 1) > 64) {
    set_error(ctx, "xt_params: frame_len too large for the window code (bits*frame_len must be <= 64)");
    return XT_ERR_ARG;
  }
  *bits_out = bits;
  return XT_OK;
}

static int ensure_plan(xt_ctx* ctx, const xt_params* p, int cap) {
  const int KS = p->n_loc;
  const int CO1 = p->d + 2 * KS + 1;
  const int RH = p->frame_len + p->nsub + 1;
  if (ctx->cap >= cap && ctx->RH >= RH && ctx->nS_alloc >= p->nS && ctx->CO1_alloc >= CO1) return XT_OK;
  cap = std::max(cap, ctx->cap);
  free_plan(ctx);
  const size_t nrec = (size_t)std::max(1, ctx->nrec_total), nch = ctx->chunks.size();
  XT_CUDA_OK(cudaMalloc(&ctx->plan.hdr, sizeof(XtRecHdr) * nrec));
  XT_CUDA_OK(cudaMalloc(&ctx->plan.goff, sizeof(uint16_t) * nrec * (cap + 1)));
  XT_CUDA_OK(cudaMalloc(&ctx->plan.ent, sizeof(uint32_t) * nrec * cap));
  XT_CUDA_OK(cudaMalloc(&ctx->plan.curG, sizeof(uint8_t) * nrec * cap));
  XT_CUDA_OK(cudaMalloc(&ctx->plan.gid, sizeof(uint16_t) * nrec * cap));
  XT_CUDA_OK(cudaMalloc(&ctx->plan.grec, sizeof(unsigned long long) * nrec * cap));
  XT_CUDA_OK(cudaMalloc(&ctx->plan.blob, sizeof(uint4) * nrec * xt_blob_stride16(cap)));
  ctx->plan.cap = cap;
  XT_CUDA_OK(cudaMalloc(&ctx->d_state1, sizeof(double) * nch * 2 * cap * CO1 * 32));
  XT_CUDA_OK(cudaMalloc(&ctx->d_hist1, sizeof(double) * nch * 2 * cap * RH * p->nS));
  ctx->cap = cap;
  ctx->RH = RH;
  ctx->nS_alloc = p->nS;
  ctx->CO1_alloc = CO1;
  return XT_OK;
}

template <int D, int KS>
static cudaError_t launch_k1(xt_ctx* ctx, const K1Args& a, const xt_params& p, size_t smem) {
  auto kern = k1_plan<D, KS>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  kern<<<(unsigned)ctx->chunks.size(), XT_K1_THREADS, smem, ctx->stream>>>(a, p);
  return cudaGetLastError();
}

template <int D, int KS, int WPC>
static cudaError_t launch_k2_lin(xt_ctx* ctx, const K2Args& a, const xt_params& p, const K2Lin& lin, size_t smem, int grid) {
  auto kern = k2_replay_lin<D, KS, WPC>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  kern<<<grid, 32 * WPC, smem, ctx->stream>>>(a, p, lin);
  return cudaGetLastError();
}

template <int D, int KS, int WPC, int TPT>
static cudaError_t launch_k2_fused_w(xt_ctx* ctx, const K2FArgs& a, const K2Tab& tab, size_t smem) {
  auto kern = k2_replay_fused<D, KS, WPC, TPT>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  kern<<<a.n_work, 32 * WPC, smem, ctx->stream>>>(a, tab);
  return cudaGetLastError();
}

template <int D, int KS>
static cudaError_t launch_k2_fused(xt_ctx* ctx, const K2FArgs& a, const K2Tab& tab, size_t smem, int wpc, int tpt) {
  if (tpt == 2) {
    if (wpc == 8) return launch_k2_fused_w<D, KS, 8, 2>(ctx, a, tab, smem);
    if (wpc == 2) return launch_k2_fused_w<D, KS, 2, 2>(ctx, a, tab, smem);
    return launch_k2_fused_w<D, KS, 4, 2>(ctx, a, tab, smem);
  }
  if (wpc == 8) return launch_k2_fused_w<D, KS, 8, 1>(ctx, a, tab, smem);
  if (wpc == 2) return launch_k2_fused_w<D, KS, 2, 1>(ctx, a, tab, smem);
  return launch_k2_fused_w<D, KS, 4, 1>(ctx, a, tab, smem);
}

template <int D, int KS>
static cudaError_t launch_k2(xt_ctx* ctx, const K2Args& a, const xt_params& p, const K2Lin& lin, size_t smem,
                             bool use_smem, int grid, int wpc) {
  if (use_smem) {
    if (wpc == 8) return launch_k2_lin<D, KS, 8>(ctx, a, p, lin, smem, grid);
    if (wpc == 2) return launch_k2_lin<D, KS, 2>(ctx, a, p, lin, smem, grid);
    return launch_k2_lin<D, KS, 4>(ctx, a, p, lin, smem, grid);
  }
  k2_replay<D, KS, false><<<grid, 32, 0, ctx->stream>>>(a, p);
  return cudaGetLastError();
}

#define XT_DISPATCH(D_, KS_, CALL)                                   \
  do {                                                               \
    if (D_ == 1) { CALL(1, 1); }                                     \
    else if (D_ == 2 && KS_ == 1) { CALL(2, 1); }                    \
    else if (D_ == 2) { CALL(2, 2); }                                \
    else if (KS_ == 1) { CALL(3, 1); }                               \
    else { CALL(3, 3); }                                             \
  } while (0)

static int run_plan(xt_ctx* ctx, const xt_params* p, int bits) {
  // run K1, growing the sequence capacity on overflow
  const int K = ipow(p->nS, p->nsub);
  int cap = std::max(ctx->cap, std::max(128, K * K * p->nS));
  for (;;) {
    if (cap > XT_HARD_CAP) {
      set_error(ctx, "more than " + std::to_string(XT_HARD_CAP) + " live state sequences; lower frame_len or raise threshold");
      return XT_ERR_CAPACITY;
    }
    int rc = ensure_plan(ctx, p, cap);
    if (rc) return rc;
    cap = ctx->cap;
    K1Args a{};
    a.chunks = ctx->d_chunks;
    a.soa = ctx->d_soa;
    a.state = ctx->d_state1;
    a.hist = ctx->d_hist1;
    a.plan = ctx->plan;
    a.summ = ctx->d_summ;
    a.cap = cap;
    a.RH = ctx->RH;
    a.bits = bits;
    a.wpc = ctx->k2_wpc;
    const size_t smem = (size_t)cap * (8 + 8 + 4 + 4 + 4 + 1) + 64;
    cudaError_t e = cudaSuccess;
#define CALL_K1(D_, KS_) e = launch_k1<D_, KS_>(ctx, a, *p, smem)
    XT_DISPATCH(p->d, p->n_loc, CALL_K1);
#undef CALL_K1
    XT_CUDA_OK(e);
    ctx->stats.k1_launches++;
    XT_CUDA_OK(cudaMemcpyAsync(ctx->h_summ, ctx->d_summ, sizeof(XtChunkSummary) * ctx->chunks.size(),
                               cudaMemcpyDeviceToHost, ctx->stream));
    XT_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    int need = 0;
    for (size_t c = 0; c < ctx->chunks.size(); ++c) {
      const XtChunkSummary& s = ctx->h_summ[c];
      if (s.err == 1) {
        set_error(ctx, "problem with grouping: a state sequence ended ungrouped in chunk " + std::to_string(c) +
                           " (threshold must be > 0 and the model finite)");
        return XT_ERR_GROUPING;
      }
      if (s.err == 2) need = std::max(need, s.need_cap);
    }
    if (!need) break;
    int ncap = cap;
    while (ncap < need) ncap *= 2;
    cap = ncap;
  }
  std::copy(ctx->h_summ, ctx->h_summ + ctx->chunks.size(), ctx->summ.begin());
  return XT_OK;
}

static int finish_eval(xt_ctx* ctx, const xt_params* p, int n_work, double* d_out, int64_t su, int64_t sg, int maxC) {
  k_reduce<<<1, 1024, 0, ctx->stream>>>(ctx->d_partial, n_work, d_out ? d_out : ctx->d_out);
  XT_CUDA_OK(cudaGetLastError());
  XT_CUDA_OK(cudaEventRecord(ctx->ev[2], ctx->stream));
  ctx->stats.k2_launches = 2;
  ctx->stats.n_tracks = ctx->n_tracks;
  ctx->stats.track_steps = ctx->track_steps;
  ctx->stats.seq_updates = su;
  ctx->stats.seq_groups = sg;
  ctx->stats.max_nB_in = maxC;
  ctx->stats.n_chunks = (int)ctx->chunks.size();
  ctx->last_p = *p;
  ctx->have_eval = true;
  return XT_OK;
}

static int evaluate(xt_ctx* ctx, const xt_params* p, double* d_out, cudaStream_t user_stream) {
  if (!ctx) return XT_ERR_ARG;
  if (ctx->chunks.empty()) {
    set_error(ctx, "no tracks uploaded");
    return XT_ERR_STATE;
  }
  int bits = 0;
  int rc = check_params(ctx, p, &bits);
  if (rc) return rc;
  XT_CUDA_OK(cudaSetDevice(ctx->device));
  (void)user_stream;
  ctx->stats = xt_stats{};
  XT_CUDA_OK(cudaEventRecord(ctx->ev[0], ctx->stream));
  rc = run_plan(ctx, p, bits);
  if (rc) return rc;
  XT_CUDA_OK(cudaEventRecord(ctx->ev[1], ctx->stream));

  // work counters + replay configuration
  int Pmax = 1, maxC = 0;
  int64_t su = 0, sg = 0;
  for (size_t c = 0; c < ctx->chunks.size(); ++c) {
    const XtChunkSummary& s = ctx->summ[c];
    Pmax = std::max(Pmax, s.max_nP);
    maxC = std::max(maxC, s.max_nC);
    su += s.sum_nC * ctx->chunks[c].nT;
    sg += s.sum_nG * ctx->chunks[c].nT;
  }
  const int K = ipow(p->nS, p->nsub);
  maxC = std::max(maxC, K * p->nS);
  const int KS = p->n_loc, CO = p->d + KS + 1;
  K2Args a{};
  a.chunks = ctx->d_chunks;
  a.work = ctx->d_work;
  a.soa = ctx->d_soa;
  a.plan = ctx->plan;
  a.summ = ctx->d_summ;
  a.logp = ctx->d_logp;
  a.partial = ctx->d_partial;
  Pmax = std::max(Pmax, 4);  // the fused kernel parks its end-of-track partial sums in an idle state buffer
  a.Pcap = Pmax;
  a.n_work = (int)ctx->work.size();
  for (int s = 0; s < p->nS; ++s) {
    double mx = -INFINITY;
    for (int r = 0; r < K; ++r) mx = std::max(mx, p->L_leave[r + K * s]);
    double acc = 0;
    for (int r = 0; r < K; ++r) acc += std::exp(p->L_leave[r + K * s] - mx);
    a.Lsum[s] = std::log(acc) + mx;
  }
  // linear-domain tables of the fast replay kernel
  K2Lin lin{};
  for (int h = 0; h < K * p->nS; ++h) {
    lin.winit[h] = std::exp(p->LT[h] + p->LF[h]);
    lin.tau0[h] = std::exp(p->LT[h]);
    lin.tau1[h] = std::exp(p->LT[h] + p->Lp_stay[h % K]);
  }
  for (int s = 0; s < p->nS; ++s) lin.leave[s] = std::exp(a.Lsum[s]);
  const int wpc = ctx->k2_wpc;
  // tracks per thread: two if the state of a 64-track tile fits in shared memory
  int tpt = ctx->k2_tpt;
  if (tpt == 2 && xt_fused_smem(p->d, KS, Pmax, K, K * p->nS, wpc, 2) > (size_t)ctx->smem_optin) tpt = 1;
  const size_t fsmem = xt_fused_smem(p->d, KS, Pmax, K, K * p->nS, wpc, tpt);
  if (ctx->k2_variant == 0 && !ctx->force_global && fsmem <= (size_t)ctx->smem_optin &&
      xt_fused_blob16(Pmax, K) <= 64 * wpc) {
    K2Tab tab{};
    for (int h = 0; h < K * p->nS; ++h) {
      tab.tau0[h] = lin.tau0[h];
      tab.tau1[h] = lin.tau1[h];
      tab.dd[h] = p->dd[h];
      tab.winit[h] = lin.winit[h];
    }
    for (int s = 0; s < p->nS; ++s) tab.leave[s] = lin.leave[s];
    for (int k = 0; k < KS; ++k) tab.l2[k] = p->l2[k];
    for (int j = 0; j < 16; ++j) tab.e2[j] = std::exp2((double)j / 16.0);
    tab.nS = p->nS;
    tab.nsub = p->nsub;
    tab.K = K;
    tab.min_len = p->min_len;
    K2FArgs fa{};
    fa.chunks = ctx->d_chunks;
    fa.work = ctx->d_workf[tpt - 1];
    fa.soa = ctx->d_soa;
    fa.plan = ctx->plan;
    fa.logp = ctx->d_logp;
    fa.partial = ctx->d_partial;
    fa.Pcap = Pmax;
    fa.n_work = ctx->n_workf[tpt - 1];
    cudaError_t ef = cudaSuccess;
#define CALL_K2F(D_, KS_) ef = launch_k2_fused<D_, KS_>(ctx, fa, tab, fsmem, wpc, tpt)
    XT_DISPATCH(p->d, p->n_loc, CALL_K2F);
#undef CALL_K2F
    XT_CUDA_OK(ef);
    return finish_eval(ctx, p, fa.n_work, d_out, su, sg, maxC);
  }
  const size_t state_bytes = (size_t)2 * Pmax * CO * 32 * sizeof(double);
  const size_t smem = state_bytes + (size_t)2 * wpc * 32 * sizeof(double);
  const bool use_smem = smem <= (size_t)ctx->smem_optin && !ctx->force_global;
  int grid = a.n_work;
  if (!use_smem) {
    grid = std::min(a.n_work, ctx->n_sm * 16);
    const size_t need = (size_t)grid * state_bytes;
    if (need > ctx->gstate_bytes) {
      cudaFree(ctx->d_gstate);
      ctx->d_gstate = nullptr;
      ctx->gstate_bytes = 0;
      XT_CUDA_OK(cudaMalloc(&ctx->d_gstate, need));
      ctx->gstate_bytes = need;
    }
    a.gstate = ctx->d_gstate;
  }
  cudaError_t e = cudaSuccess;
#define CALL_K2(D_, KS_) e = launch_k2<D_, KS_>(ctx, a, *p, lin, smem, use_smem, grid, wpc)
  XT_DISPATCH(p->d, p->n_loc, CALL_K2);
#undef CALL_K2
  XT_CUDA_OK(e);
  return finish_eval(ctx, p, a.n_work, d_out, su, sg, maxC);
}

extern "C" int xt_sum_logp(xt_ctx* ctx, const xt_params* p, double* out) {
  int rc = evaluate(ctx, p, nullptr, nullptr);
  if (rc) return rc;
  XT_CUDA_OK(cudaMemcpyAsync(ctx->h_out, ctx->d_out, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  XT_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  *out = *ctx->h_out;
  return XT_OK;
}

extern "C" int xt_sum_logp_async(xt_ctx* ctx, const xt_params* p, double* d_out, void* cuda_stream) {
  // The result lands in d_out in stream order of the context's stream; if the caller passes its
  // own stream it is made to wait for the result.
  int rc = evaluate(ctx, p, d_out, (cudaStream_t)cuda_stream);
  if (rc) return rc;
  if (cuda_stream) XT_CUDA_OK(cudaStreamWaitEvent((cudaStream_t)cuda_stream, ctx->ev[2], 0));
  return XT_OK;
}

extern "C" int xt_set_option(xt_ctx* ctx, const char* name, int value) {
  if (!ctx || !name) return XT_ERR_ARG;
  if (std::strcmp(name, "force_global_replay") == 0) {
    ctx->force_global = value != 0;
    ctx->have_eval = false;
    return XT_OK;
  }
  if (std::strcmp(name, "k2_variant") == 0) {
    if (value != 0 && value != 1) {
      set_error(ctx, "xt_set_option: k2_variant must be 0 (fused) or 1 (first-generation linear)");
      return XT_ERR_ARG;
    }
    ctx->k2_variant = value;
    ctx->have_eval = false;
    return XT_OK;
  }
  if (std::strcmp(name, "k2_tpt") == 0) {
    if (value != 1 && value != 2) {
      set_error(ctx, "xt_set_option: k2_tpt must be 1 or 2");
      return XT_ERR_ARG;
    }
    ctx->k2_tpt = value;
    ctx->have_eval = false;
    return XT_OK;
  }
  if (std::strcmp(name, "k2_wpc") == 0) {
    if (value != 2 && value != 4 && value != 8) {
      set_error(ctx, "xt_set_option: k2_wpc must be 2, 4 or 8");
      return XT_ERR_ARG;
    }
    ctx->k2_wpc = value;
    ctx->have_eval = false;
    return XT_OK;
  }
  set_error(ctx, std::string("xt_set_option: unknown option ") + name);
  return XT_ERR_ARG;
}

extern "C" int xt_get_stats(xt_ctx* ctx, xt_stats* out) {
  if (!ctx || !out) return XT_ERR_ARG;
  if (ctx->have_eval) {
    XT_CUDA_OK(cudaSetDevice(ctx->device));
    XT_CUDA_OK(cudaEventSynchronize(ctx->ev[2]));
    cudaEventElapsedTime(&ctx->stats.ms_plan, ctx->ev[0], ctx->ev[1]);
    cudaEventElapsedTime(&ctx->stats.ms_replay, ctx->ev[1], ctx->ev[2]);
  }
  *out = ctx->stats;
  return XT_OK;
}

extern "C" int xt_chunk_logp(xt_ctx* ctx, int32_t chunk, const xt_params* p, double* out) {
  if (!ctx || chunk < 0 || chunk >= (int)ctx->chunks.size()) {
    set_error(ctx, "xt_chunk_logp: bad chunk index");
    return XT_ERR_ARG;
  }
  if (!ctx->have_eval || std::memcmp(&ctx->last_p, p, sizeof(xt_params)) != 0) {
    int rc = evaluate(ctx, p, nullptr, nullptr);
    if (rc) return rc;
  }
  const XtChunk& ck = ctx->chunks[chunk];
  XT_CUDA_OK(cudaMemcpyAsync(out, ctx->d_logp + ck.trk_off, sizeof(double) * ck.nT, cudaMemcpyDeviceToHost,
                             ctx->stream));
  XT_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  return XT_OK;
}

extern "C" int xt_plan_dump(xt_ctx* ctx, int32_t chunk, int32_t step, int32_t* nB_in, int32_t* nG, int32_t* gid,
                            int32_t cap, double* threshold_used) {
  if (!ctx || !ctx->have_eval || chunk < 0 || chunk >= (int)ctx->chunks.size()) {
    set_error(ctx, "xt_plan_dump: no evaluation yet or bad chunk index");
    return XT_ERR_ARG;
  }
  const XtChunk& ck = ctx->chunks[chunk];
  if (step < 2 || step > ck.L - 2) {
    set_error(ctx, "xt_plan_dump: step must satisfy 2 <= step <= L-2");
    return XT_ERR_ARG;
  }
  const int rec = ck.rec0 + (step - 2);
  XtRecHdr h;
  XT_CUDA_OK(cudaMemcpy(&h, ctx->plan.hdr + rec, sizeof(h), cudaMemcpyDeviceToHost));
  *nB_in = h.nC;
  *nG = h.nG;
  if (threshold_used) *threshold_used = h.th;
  if (cap < h.nC) {
    set_error(ctx, "xt_plan_dump: gid buffer too small");
    return XT_ERR_ARG;
  }
  std::vector<uint16_t> tmp(h.nC);
  XT_CUDA_OK(cudaMemcpy(tmp.data(), ctx->plan.gid + (size_t)rec * ctx->plan.cap, sizeof(uint16_t) * h.nC,
                        cudaMemcpyDeviceToHost));
  for (int i = 0; i < h.nC; ++i) gid[i] = tmp[i];
  return XT_OK;
}

extern "C" int xt_fp64_peak_tflops(xt_ctx* ctx, double* out) {
  if (!ctx) return XT_ERR_ARG;
  XT_CUDA_OK(cudaSetDevice(ctx->device));
  const int blocks = ctx->n_sm * 8, threads = 256, iters = 1 << 15;
  double* buf = nullptr;
  XT_CUDA_OK(cudaMalloc(&buf, sizeof(double) * blocks * threads));
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(a, ctx->stream);
    k_fp64_peak<<<blocks, threads, 0, ctx->stream>>>(buf, iters);
    cudaEventRecord(b, ctx->stream);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    if (rep > 0) best = std::min(best, ms);
  }
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  cudaFree(buf);
  XT_CUDA_OK(cudaGetLastError());
  const double flops = 2.0 * 8.0 * (double)iters * (double)blocks * threads;
  *out = flops / (best * 1e-3) / 1e12;
  return XT_OK;
}

#include "xt_predict_host.inl"
