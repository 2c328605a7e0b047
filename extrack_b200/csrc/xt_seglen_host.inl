// Host side of the segment-length histogram (included by xt_engine.cu).
extern "C" int xt_seglen_hist(xt_ctx* ctx, const xt_params* p, const double* leave_LL, double* hist, int32_t Lmax,
                              int32_t dbg_chunk, double* dbg_LP, int8_t* dbg_Bs, int32_t* n_final) {
  if (!ctx || !hist) return XT_ERR_ARG;
  if (ctx->chunks.empty()) {
    set_error(ctx, "no tracks uploaded");
    return XT_ERR_STATE;
  }
  int bits = 0;
  int rc = check_params(ctx, p, &bits);
  if (rc) return rc;
  if (p->nsub != 1 || is_var(p)) {
    set_error(ctx, "xt_seglen_hist: nb_substeps = 1 and scalar / per-dimension LocErr, scalar dt only");
    return XT_ERR_ARG;
  }
  if (p->max_nb_states < 1) {
    set_error(ctx, "xt_seglen_hist: max_nb_states must be >= 1");
    return XT_ERR_ARG;
  }
  XT_CUDA_OK(cudaSetDevice(ctx->device));
  const int nS = p->nS, KS = p->n_loc, d = p->d;
  const int nch = (int)ctx->chunks.size();
  // live-sequence counts (histograms.py: x nS per step, cut to max_nb_states while step < L-1)
  long long capC = (long long)nS * nS;
  int Lm = 0;
  std::vector<int> nBf(nch);
  for (int c = 0; c < nch; ++c) {
    const XtChunk& ck = ctx->chunks[c];
    Lm = std::max(Lm, ck.L);
    long long nB = (long long)nS * nS;
    for (int step = 2; step <= ck.L - 1; ++step) {
      nB *= nS;
      capC = std::max(capC, nB);
      if (capC > 65535) break;
      if (step < ck.L - 1 && nB > p->max_nb_states) nB = p->max_nb_states;
    }
    nBf[c] = (int)(nB * (ck.isBL ? nS : 1));
  }
  if (Lmax < Lm) {
    set_error(ctx, "xt_seglen_hist: Lmax is smaller than the longest track");
    return XT_ERR_ARG;
  }
  // (max_nb_states itself needs no slots: a step is only cut when it has more children than that)
  int n2 = 4;
  while (n2 < capC && n2 < (1 << 20)) n2 <<= 1;
  const size_t smem = capC <= 65535 ? xt_seg_smem(d, KS, (int)capC, n2, Lmax, nS) : (size_t)-1;
  if (capC > 65535 || n2 > 16 * XT_SEG_THREADS || smem > (size_t)ctx->smem_optin) {
    set_error(ctx, "xt_seglen_hist: " + std::to_string(capC) + " live state sequences per track do not fit in shared memory; lower max_nb_states");
    return XT_ERR_CAPACITY;
  }
  if (n_final && dbg_chunk >= 0 && dbg_chunk < nch) *n_final = nBf[dbg_chunk];
  const int cap = (int)capC;
  // one CTA per chunk (the reference's unit of the > 600 rescale), persistent over the chunk list
  const int grid = (int)std::min<size_t>((size_t)nch, (size_t)ctx->n_sm * std::max<size_t>(1, ((size_t)228 * 1024) / (smem + 2048)));
  double* d_colmax = nullptr;
  uint32_t* d_lat = nullptr;
  double* d_hist = nullptr;
  int32_t* d_flags = nullptr;
  double* d_LP = nullptr;
  int8_t* d_Bs = nullptr;
  int result = XT_OK;
  const size_t hist_n = (size_t)nch * Lmax * nS;
  const bool dbg = dbg_chunk >= 0 && dbg_chunk < nch && (dbg_LP || dbg_Bs);
  const size_t dn = dbg ? (size_t)ctx->chunks[dbg_chunk].nT * nBf[dbg_chunk] : 0;
  auto cleanup = [&]() {
    cudaFree(d_colmax); cudaFree(d_lat); cudaFree(d_hist); cudaFree(d_flags); cudaFree(d_LP); cudaFree(d_Bs);
  };
#define SEG_OK(call)                                                        \
  do {                                                                      \
    cudaError_t e__ = (call);                                               \
    if (e__ != cudaSuccess) {                                               \
      set_error(ctx, std::string(#call) + ": " + cudaGetErrorString(e__));  \
      cleanup();                                                            \
      return XT_ERR_CUDA;                                                   \
    }                                                                       \
  } while (0)
  SEG_OK(cudaMalloc(&d_colmax, sizeof(double) * (size_t)grid * cap));
  SEG_OK(cudaMalloc(&d_lat, sizeof(uint32_t) * (size_t)grid * Lmax * cap));
  SEG_OK(cudaMalloc(&d_hist, sizeof(double) * hist_n));
  SEG_OK(cudaMalloc(&d_flags, sizeof(int32_t) * 2));
  if (dbg && dbg_LP) SEG_OK(cudaMalloc(&d_LP, sizeof(double) * dn));
  if (dbg && dbg_Bs) SEG_OK(cudaMalloc(&d_Bs, dn * ctx->chunks[dbg_chunk].L));
  SEG_OK(cudaMemsetAsync(d_hist, 0, sizeof(double) * hist_n, ctx->stream));
  SEG_OK(cudaMemsetAsync(d_flags, 0, sizeof(int32_t) * 2, ctx->stream));
  K4Args a{};
  a.chunks = ctx->d_chunks;
  a.soa = ctx->d_soa;
  a.n_chunks = nch;
  a.colmax = d_colmax;
  a.cap = cap;
  a.n2 = n2;
  a.Lmax = Lmax;
  a.lattice = d_lat;
  a.hist = d_hist;
  a.flags = d_flags;
  a.corder = ctx->d_corder;
  for (int h = 0; h < nS * nS; ++h) a.leave_LL[h] = leave_LL ? leave_LL[h] : 0.0;
  a.dbg_LP = d_LP;
  a.dbg_Bs = d_Bs;
  a.dbg_chunk = dbg ? dbg_chunk : -1;
  a.dbg_nBf = dbg ? nBf[dbg_chunk] : 0;
  cudaError_t e = cudaSuccess;
  cudaEvent_t e0, e1;
  SEG_OK(cudaEventCreate(&e0));
  SEG_OK(cudaEventCreate(&e1));
  cudaEventRecord(e0, ctx->stream);
  e = xt_launch_k4(a, *p, grid, smem, ctx->stream);
  cudaEventRecord(e1, ctx->stream);
  int32_t flags = 0;
  if (e == cudaSuccess) e = cudaMemcpyAsync(hist, d_hist, sizeof(double) * hist_n, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(&flags, d_flags, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess && d_LP) e = cudaMemcpyAsync(dbg_LP, d_LP, sizeof(double) * dn, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess && d_Bs) e = cudaMemcpyAsync(dbg_Bs, d_Bs, dn * ctx->chunks[dbg_chunk].L, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e == cudaSuccess) cudaEventElapsedTime(&ctx->ms_seglen, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (e != cudaSuccess) {
    set_error(ctx, std::string("xt_seglen_hist: ") + cudaGetErrorString(e));
    result = XT_ERR_CUDA;
  }
  ctx->seglen_rescaled = (flags & 1) != 0;
  cleanup();
#undef SEG_OK
  return result;
}

extern "C" int xt_seglen_last_ms(xt_ctx* ctx, float* ms) {
  if (!ctx || !ms) return XT_ERR_ARG;
  *ms = ctx->ms_seglen;
  return XT_OK;
}
