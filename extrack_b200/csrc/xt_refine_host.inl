// Host driver of the position refinement (included by xt_engine.cu): two passes of the refinement recursion over the
// uploaded buckets (every bucket = one chunk, as get_pos_PDF hands whole buckets to get_LC_Km_Ks,
// refined_localization.py:214-218,:325-328) and their combination.
struct XtRefinePass {
  double* dump = nullptr;
  int64_t* d_off = nullptr;
  int32_t* d_entn = nullptr;
  int32_t* d_splan = nullptr;
  int capD = 0;
};
static void free_refine_pass(XtRefinePass& r) {
  cudaFree(r.dump); cudaFree(r.d_off); cudaFree(r.d_entn); cudaFree(r.d_splan);
  r = XtRefinePass{};
}

static int run_refine_pass(xt_ctx* ctx, const xt_params* p, int bits, int rev, int* cap_io, XtRefinePass* out) {
  const int nS = p->nS, KS = p->n_loc, CO = p->d + 2 * KS + 1, COD = p->d + KS + 2;
  const int nch = (int)ctx->chunks.size(), n_work = (int)ctx->work.size(), maxL = ctx->maxL + 1;
  int cap = *cap_io;
  int32_t* d_err = nullptr;
  double* d_scr = nullptr;
  double* d_scr2 = nullptr;
  int32_t* d_err2 = nullptr;
  int result = XT_OK;
  auto cleanup = [&]() { cudaFree(d_err); cudaFree(d_scr); cudaFree(d_scr2); cudaFree(d_err2); };
  for (;;) {
    free_refine_pass(*out);
    cudaFree(d_err); cudaFree(d_scr); cudaFree(d_scr2); cudaFree(d_err2);
    d_err = d_err2 = nullptr; d_scr = d_scr2 = nullptr;
    if (cap > 1024) {
      set_error(ctx, "position refinement: more than 1024 live state sequences; lower frame_len or raise threshold");
      result = XT_ERR_CAPACITY;
      break;
    }
    // ---- plans of the buckets (first 30 tracks of every bucket) ----
    const K3SLayout sl = k3s_layout(cap, CO, p->frame_len, nS);
    const int sgrid = std::max(1, std::min((nch + 7) / 8, ctx->n_sm * 2));
    K3SArgs sa{};
    sa.chunks = ctx->d_chunks;
    sa.soa = ctx->d_soa;
    sa.warp_scratch = sl.total;
    sa.splan_stride = k3s_splan_stride(cap, maxL);
    sa.n_chunks = nch;
    sa.cap = cap;
    sa.maxL = maxL;
    sa.bits = bits;
    sa.refine = 1;
    sa.rev = rev;
    if (cudaMalloc(&out->d_splan, sizeof(int32_t) * sa.splan_stride * nch) != cudaSuccess ||
        cudaMalloc(&d_scr, sizeof(double) * sl.total * (size_t)sgrid * 8) != cudaSuccess ||
        cudaMalloc(&d_err, sizeof(int32_t) * 2 * (size_t)nch) != cudaSuccess) {
      set_error(ctx, "position refinement: cannot allocate the plan buffers");
      result = XT_ERR_CUDA;
      break;
    }
    cudaMemsetAsync(d_err, 0, sizeof(int32_t) * 2 * (size_t)nch, ctx->stream);
    cudaMemsetAsync(out->d_splan, 0, sizeof(int32_t) * sa.splan_stride * nch, ctx->stream);
    sa.scratch = d_scr;
    sa.splan = out->d_splan;
    sa.err = d_err;
    sa.err_need = d_err + nch;
    cudaError_t e = xt_launch_k3_shared_plan(sa, *p, sgrid, ctx->stream);
    std::vector<int32_t> h_err(2 * (size_t)nch);
    std::vector<int32_t> h_plan((size_t)2 * maxL * nch);
    if (e != cudaSuccess ||
        cudaMemcpyAsync(h_err.data(), d_err, sizeof(int32_t) * 2 * (size_t)nch, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
        cudaMemcpy2DAsync(h_plan.data(), sizeof(int32_t) * 2 * maxL, out->d_splan, sizeof(int32_t) * sa.splan_stride,
                          sizeof(int32_t) * 2 * maxL, nch, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
        cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
      set_error(ctx, std::string("position refinement (plans): ") + cudaGetErrorString(cudaGetLastError()));
      result = XT_ERR_CUDA;
      break;
    }
    int need = 0;
    bool grouping = false;
    for (int i = 0; i < nch; ++i) {
      if (h_err[i] == 1) grouping = true;
      need = std::max(need, h_err[nch + i]);
    }
    if (grouping) {
      set_error(ctx, "problem with grouping: a state sequence ended ungrouped (threshold must be > 0 and the model finite)");
      result = XT_ERR_GROUPING;
      break;
    }
    if (need) {
      while (cap < need) cap *= 2;
      continue;
    }
    // ---- entry capacity: most sequences any stored step holds ----
    int capD = nS * nS;
    for (int c = 0; c < nch; ++c) {
      const int L = ctx->chunks[c].L;
      int last = nS * nS;  // parents of the last (unfused) step
      for (int s = 2; s <= L - 2; ++s) {
        last = h_plan[(size_t)c * 2 * maxL + maxL + s];
        capD = std::max(capD, last);
      }
      if (L >= 3) capD = std::max(capD, last * nS);
    }
    if (capD > cap) {
      cap *= 2;
      continue;
    }
    out->capD = capD;
    std::vector<int64_t> off(nch);
    int64_t tot = 0;
    for (int c = 0; c < nch; ++c) {
      off[c] = tot;
      tot += (int64_t)ctx->chunks[c].nT * (ctx->chunks[c].L - 1) * capD * COD;
    }
    // ---- all tracks along the plans, storing every step ----
    const K3Layout lay = k3_layout(cap, CO, p->frame_len, nS, maxL);
    const int nwarps = XT_K3_WARPS;
    const int grid = std::max(1, std::min(n_work, ctx->n_sm * ctx->k3_ctas_per_sm));
    const size_t warp_units = lay.cold_total + lay.hot_total;
    if (cudaMalloc(&out->dump, sizeof(double) * (size_t)std::max<int64_t>(tot, 1)) != cudaSuccess ||
        cudaMalloc(&out->d_off, sizeof(int64_t) * nch) != cudaSuccess ||
        cudaMalloc(&out->d_entn, sizeof(int32_t) * (size_t)nch * maxL) != cudaSuccess ||
        cudaMalloc(&d_scr2, sizeof(double) * warp_units * (size_t)grid * nwarps) != cudaSuccess ||
        cudaMalloc(&d_err2, sizeof(int32_t) * 2 * (size_t)n_work) != cudaSuccess) {
      set_error(ctx, "position refinement: cannot allocate the per-step storage (" + std::to_string(tot * 8 / (1 << 20)) + " MiB)");
      result = XT_ERR_CUDA;
      break;
    }
    cudaMemcpyAsync(out->d_off, off.data(), sizeof(int64_t) * nch, cudaMemcpyHostToDevice, ctx->stream);
    cudaMemsetAsync(out->d_entn, 0, sizeof(int32_t) * (size_t)nch * maxL, ctx->stream);
    cudaMemsetAsync(d_err2, 0, sizeof(int32_t) * 2 * (size_t)n_work, ctx->stream);
    K3Args a{};
    a.chunks = ctx->d_chunks;
    a.work = ctx->d_work;
    a.soa = ctx->d_soa;
    a.scratch = d_scr2;
    a.pred = out->dump;  // (unused in this mode)
    a.err = d_err2;
    a.err_need = d_err2 + n_work;
    a.n_work = n_work;
    a.cap = cap;
    a.maxL = maxL;
    a.bits = bits;
    a.warp_scratch = warp_units;
    a.hot_smem = 0;
    a.splan = out->d_splan;
    a.splan_stride = sa.splan_stride;
    a.rev = rev;
    a.dump = out->dump;
    a.dump_off = out->d_off;
    a.ent_n = out->d_entn;
    a.capD = capD;
    for (int s = 0; s < nS; ++s) a.LF[s] = p->LF[s * nS];  // log F of state s (head s * nS: oldest state = s)
    e = xt_launch_k3_refine(a, *p, grid, nwarps, 0, ctx->stream);
    std::vector<int32_t> h_err2(2 * (size_t)n_work);
    if (e != cudaSuccess ||
        cudaMemcpyAsync(h_err2.data(), d_err2, sizeof(int32_t) * 2 * (size_t)n_work, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
        cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
      set_error(ctx, std::string("position refinement (recursion): ") + cudaGetErrorString(cudaGetLastError()));
      result = XT_ERR_CUDA;
      break;
    }
    need = 0;
    grouping = false;
    for (int i = 0; i < n_work; ++i) {
      if (h_err2[i] == 1) grouping = true;
      need = std::max(need, h_err2[n_work + i]);
    }
    if (grouping) {
      set_error(ctx, "position refinement: a track could not follow its bucket's plan");
      result = XT_ERR_GROUPING;
      break;
    }
    if (need) {
      while (cap < need) cap *= 2;
      continue;
    }
    break;
  }
  cleanup();
  if (result != XT_OK) free_refine_pass(*out);
  *cap_io = cap;
  return result;
}

extern "C" int xt_refine_positions(xt_ctx* ctx, const xt_params* p_rev, const xt_params* p_fwd, double* const* mu_out,
                                   double* const* sigma_out) {
  if (!ctx || !p_rev || !p_fwd || !mu_out || !sigma_out) return XT_ERR_ARG;
  if (ctx->chunks.empty()) {
    set_error(ctx, "no tracks uploaded");
    return XT_ERR_STATE;
  }
  int bits = 0;
  int rc = check_params(ctx, p_rev, &bits);
  if (!rc) rc = check_params(ctx, p_fwd, &bits);
  if (rc) return rc;
  if (p_rev->nsub != 1 || p_fwd->nsub != 1 || is_var(p_rev) || is_var(p_fwd) || p_rev->nS != p_fwd->nS || p_rev->n_loc != p_fwd->n_loc) {
    set_error(ctx, "xt_refine_positions: nb_substeps must be 1, LocErr scalar or per dimension, both passes of the same model");
    return XT_ERR_UNSUPPORTED;
  }
  for (size_t c = 0; c < ctx->chunks.size(); ++c)
    if (ctx->chunks[c].nT != ctx->seg_n[ctx->chunks[c].seg]) {
      set_error(ctx, "xt_refine_positions: upload every length bucket as one chunk (chunk_size >= its track count)");
      return XT_ERR_ARG;
    }
  XT_CUDA_OK(cudaSetDevice(ctx->device));
  const int nS = p_rev->nS, KS = p_rev->n_loc, d = p_rev->d;
  int cap = std::max(64, nS * nS * nS);
  XtRefinePass r1, r2;
  rc = run_refine_pass(ctx, p_rev, bits, 1, &cap, &r1);
  if (!rc) rc = run_refine_pass(ctx, p_fwd, bits, 0, &cap, &r2);
  double* d_mu = nullptr;
  double* d_sigma = nullptr;
  if (!rc) {
    if (cudaMalloc(&d_mu, sizeof(double) * (size_t)ctx->n_locs * d) != cudaSuccess ||
        cudaMalloc(&d_sigma, sizeof(double) * (size_t)ctx->n_locs) != cudaSuccess) {
      set_error(ctx, "xt_refine_positions: cannot allocate the outputs");
      rc = XT_ERR_CUDA;
    }
  }
  if (!rc) {
    KRArgs a{};
    a.chunks = ctx->d_chunks;
    a.soa = ctx->d_soa;
    a.dump1 = r1.dump; a.dump2 = r2.dump;
    a.dump_off1 = r1.d_off; a.dump_off2 = r2.d_off;
    a.ent_n1 = r1.d_entn; a.ent_n2 = r2.d_entn;
    a.mu = d_mu; a.sigma = d_sigma;
    a.n_chunks = (int)ctx->chunks.size();
    a.maxL = ctx->maxL + 1;
    a.capD1 = r1.capD; a.capD2 = r2.capD;
    for (int k = 0; k < KS; ++k) a.le[k] = std::sqrt(p_rev->l2[k]);
    const long long warps = (long long)ctx->n_locs;
    const unsigned grid = (unsigned)((warps * 32 + 255) / 256);
    cudaError_t e = xt_launch_refine_combine(d, KS, a, grid, ctx->stream);
    if (e != cudaSuccess) {
      set_error(ctx, std::string("xt_refine_positions (combination): ") + cudaGetErrorString(e));
      rc = XT_ERR_CUDA;
    }
  }
  if (!rc) {
    for (size_t s = 0; s < ctx->seg_n.size(); ++s) {
      const XtChunk& c0 = ctx->chunks[ctx->seg_chunk0[s]];
      const size_t cnt = (size_t)ctx->seg_n[s] * ctx->seg_L[s];
      if (cudaMemcpyAsync(mu_out[s], d_mu + (size_t)c0.loc_off * d, sizeof(double) * cnt * d, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
          cudaMemcpyAsync(sigma_out[s], d_sigma + (size_t)c0.loc_off, sizeof(double) * cnt, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) {
        set_error(ctx, "xt_refine_positions: device to host copy failed");
        rc = XT_ERR_CUDA;
        break;
      }
    }
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess && !rc) {
      set_error(ctx, std::string("xt_refine_positions: ") + cudaGetErrorString(cudaGetLastError()));
      rc = XT_ERR_CUDA;
    }
  }
  free_refine_pass(r1);
  free_refine_pass(r2);
  cudaFree(d_mu);
  cudaFree(d_sigma);
  return rc;
}
