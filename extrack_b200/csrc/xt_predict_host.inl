// Host driver of the state-annotation kernel (included by xt_engine.cu).
extern "C" int xt_predict(xt_ctx* ctx, const xt_params* p, double* const* out) {
  if (!ctx) return XT_ERR_ARG;
  if (ctx->chunks.empty()) {
    set_error(ctx, "no tracks uploaded");
    return XT_ERR_STATE;
  }
  int bits = 0;
  int rc = check_params(ctx, p, &bits);
  if (rc) return rc;
  if (p->nsub != 1) {
    set_error(ctx, "xt_predict: nb_substeps must be 1 (predict_Bs forces it, tracking.py:839)");
    return XT_ERR_ARG;
  }
  XT_CUDA_OK(cudaSetDevice(ctx->device));
  const int nS = p->nS, KS = p->n_loc, CO = p->d + 2 * KS + 1;
  const int n_work = (int)ctx->work.size();
  double* d_pred = nullptr;
  int32_t* d_err = nullptr;
  double* d_scratch = nullptr;
  if (cudaMalloc(&d_pred, sizeof(double) * (size_t)ctx->n_locs * nS) != cudaSuccess ||
      cudaMalloc(&d_err, sizeof(int32_t) * 2 * (size_t)n_work) != cudaSuccess) {
    set_error(ctx, std::string("xt_predict: cannot allocate the output buffers: ") + cudaGetErrorString(cudaGetLastError()));
    cudaFree(d_pred);
    cudaFree(d_err);
    return XT_ERR_CUDA;
  }
  std::vector<int32_t> h_err(2 * (size_t)n_work);
  int cap = std::max(ctx->k3_cap0, nS * nS * nS);
  cap += cap & 1;
  int result = XT_OK;
  ctx->k3_launches = 0;
  // nb_max > 1 (the upload's chunk size): one plan per chunk, decided from its first 30 tracks (xt_predict_shared.cuh)
  const bool shared = ctx->k3_shared && (int)ctx->upload_sig[2] > 1;  // option "predict_shared_plans"
  // work items of the current launch: all of them first, then only those whose tracks outgrew the capacity (own-plan
  // mode; the outputs of a track depend on nothing but the track, so the others keep theirs)
  const XtWork* d_work_cur = ctx->d_work;
  XtWork* d_retry = nullptr;
  std::vector<XtWork> retry;
  int n_cur = n_work;
  float ms_total = 0.f;
  // Own-plan mode on a large data set: the first round is launched in a few pieces of consecutive length buckets, and
  // the read-back of a piece (into the caller's pageable arrays: the host thread does the staging and takes the page
  // faults) runs while the kernel works on the next pieces.
  std::vector<int> piece_w0, piece_s0;  // first work item / first segment of every piece (+ one past the end)
  bool copied = false, ran_again = false;
  cudaStream_t s_copy = nullptr;
  std::vector<cudaEvent_t> ev_piece;
  if (!shared && ctx->k3_pieces > 1 && n_work >= 64 * ctx->n_sm) {
    const int n_seg = (int)ctx->seg_n.size();
    std::vector<int> seg_w0(n_seg + 1, n_work);
    for (int i = n_work - 1; i >= 0; --i) seg_w0[ctx->chunks[ctx->work[i].chunk].seg] = i;
    for (int sg = n_seg - 1; sg >= 0; --sg) seg_w0[sg] = std::min(seg_w0[sg], seg_w0[sg + 1]);  // (empty segments)
    double total = 0, acc = 0;
    for (int sg = 0; sg < n_seg; ++sg) total += (double)ctx->seg_n[sg] * ctx->seg_L[sg];
    const int want = std::min(ctx->k3_pieces, n_seg);
    piece_w0.push_back(0);
    piece_s0.push_back(0);
    for (int sg = 0; sg < n_seg; ++sg) {
      acc += (double)ctx->seg_n[sg] * ctx->seg_L[sg];
      if (sg + 1 < n_seg && acc >= total * (double)piece_w0.size() / want && seg_w0[sg + 1] > piece_w0.back()) {
        piece_w0.push_back(seg_w0[sg + 1]);
        piece_s0.push_back(sg + 1);
      }
    }
    piece_w0.push_back(n_work);
    piece_s0.push_back(n_seg);
    if (piece_w0.size() < 3 || cudaStreamCreateWithFlags(&s_copy, cudaStreamNonBlocking) != cudaSuccess) {
      cudaGetLastError();
      piece_w0.clear();
      s_copy = nullptr;
    } else {
      ev_piece.resize(piece_w0.size() - 1, nullptr);
      for (auto& ev : ev_piece) cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    }
  }
  const int nch = (int)ctx->chunks.size();
  int32_t* d_splan = nullptr;
  double* d_sscratch = nullptr;
  int32_t* d_serr = nullptr;
  if (shared && is_var(p)) {
    set_error(ctx, "predict_Bs with nb_max > 1 is implemented for scalar LocErr / dt (per-localisation inputs: use nb_max = 1)");
    cudaFree(d_pred);
    cudaFree(d_err);
    return XT_ERR_UNSUPPORTED;
  }
  for (;;) {
    if (cap > XT_HARD_CAP) {
      set_error(ctx, "more than " + std::to_string(XT_HARD_CAP) + " live state sequences; lower frame_len or raise threshold");
      result = XT_ERR_CAPACITY;
      break;
    }
    const K3Layout lay = k3_layout(cap, CO, p->frame_len, nS, ctx->maxL + 1);
    // hot scratch in shared memory: as many warps per CTA (<= 8) as fit, CTAs per SM accordingly;
    // if not even one warp fits, everything stays in global memory (4 CTAs of 8 warps per SM)
    const size_t hot_bytes = sizeof(double) * lay.hot_total;
    const size_t smem_cap = (size_t)ctx->smem_optin;
    int nwarps = XT_K3_WARPS, ctas_per_sm = ctx->k3_ctas_per_sm, hot_smem = 0;
    if (ctx->k3_hot_smem && hot_bytes <= smem_cap) {
      // warps per CTA (<= 8) that maximise the resident warps per SM under the shared-memory and
      // register limits (the kernel is latency-bound: more resident warps = more throughput)
      hot_smem = 1;
      const int regs = shared ? 128 : xt_k3_regs(*p, 1);  // registers per thread of the kernel that will run
      int best = 0;
      for (int nw = XT_K3_WARPS; nw >= 1; --nw) {
        if (hot_bytes * nw > smem_cap) continue;
        const int by_smem = (int)(((size_t)228 * 1024) / (hot_bytes * nw + 1024));
        const int by_regs = 65536 / (regs * 32 * nw);
        const int ctas = std::max(1, std::min(by_smem, by_regs));
        if (ctas * nw > best) {
          best = ctas * nw;
          nwarps = nw;
          ctas_per_sm = ctas;
        }
      }
    }
    const size_t smem = hot_smem ? hot_bytes * nwarps : 0;
    const int grid = std::min(n_cur, ctx->n_sm * ctas_per_sm);
    const size_t warp_units = lay.cold_total + (hot_smem ? 0 : lay.hot_total);
    const size_t bytes = sizeof(double) * warp_units * (size_t)grid * nwarps;
    cudaFree(d_scratch);
    d_scratch = nullptr;
    if (cudaMalloc(&d_scratch, bytes) != cudaSuccess) {
      set_error(ctx, "xt_predict: cannot allocate scratch");
      result = XT_ERR_CUDA;
      break;
    }
    cudaMemsetAsync(d_err, 0, sizeof(int32_t) * 2 * (size_t)n_work, ctx->stream);
    K3Args a{};
    a.chunks = ctx->d_chunks;
    a.work = d_work_cur;
    a.soa = ctx->d_soa;
    a.scratch = d_scratch;
    a.pred = d_pred;
    a.err = d_err;
    a.err_need = d_err + n_work;
    a.n_work = n_cur;
    a.cap = cap;
    a.maxL = ctx->maxL + 1;
    a.bits = bits;
    a.warp_scratch = warp_units;
    a.hot_smem = hot_smem;
    if (is_var(p)) a.ax = make_aux(ctx, p, 1);
    for (int s = 0; s < nS; ++s) {  // nsub == 1: K = nS
      double mx = -INFINITY;
      for (int r = 0; r < nS; ++r) mx = std::max(mx, p->L_leave[r + nS * s]);
      double acc = 0;
      for (int r = 0; r < nS; ++r) acc += std::exp(p->L_leave[r + nS * s] - mx);
      a.Lsum[s] = std::log(acc) + mx;
    }
    cudaError_t e = cudaSuccess;
    ctx->k3_launches++;
    ctx->k3_cap = cap;
    cudaEventRecord(ctx->ev_k3[0], ctx->stream);
    if (shared) {
      if (cap > 1024) {
        set_error(ctx, "predict_Bs with nb_max > 1: more than 1024 live state sequences; lower frame_len or raise threshold");
        result = XT_ERR_CAPACITY;
        break;
      }
      const K3SLayout sl = k3s_layout(cap, CO, p->frame_len, nS);
      const int sgrid = std::max(1, std::min((nch + 7) / 8, ctx->n_sm * 2));
      K3SArgs sa{};
      sa.chunks = ctx->d_chunks;
      sa.soa = ctx->d_soa;
      sa.warp_scratch = sl.total;
      sa.splan_stride = k3s_splan_stride(cap, a.maxL);
      sa.n_chunks = nch;
      sa.cap = cap;
      sa.maxL = a.maxL;
      sa.bits = bits;
      cudaFree(d_splan); cudaFree(d_sscratch); cudaFree(d_serr);
      d_splan = nullptr; d_sscratch = nullptr; d_serr = nullptr;
      if (cudaMalloc(&d_splan, sizeof(int32_t) * sa.splan_stride * nch) != cudaSuccess ||
          cudaMalloc(&d_sscratch, sizeof(double) * sl.total * (size_t)sgrid * 8) != cudaSuccess ||
          cudaMalloc(&d_serr, sizeof(int32_t) * 2 * (size_t)nch) != cudaSuccess) {
        set_error(ctx, "xt_predict: cannot allocate the shared-plan buffers");
        result = XT_ERR_CUDA;
        break;
      }
      cudaMemsetAsync(d_serr, 0, sizeof(int32_t) * 2 * (size_t)nch, ctx->stream);
      sa.scratch = d_sscratch;
      sa.splan = d_splan;
      sa.err = d_serr;
      sa.err_need = d_serr + nch;
      e = xt_launch_k3_shared_plan(sa, *p, sgrid, ctx->stream);
      std::vector<int32_t> h_serr(2 * (size_t)nch);
      if (e != cudaSuccess || cudaMemcpyAsync(h_serr.data(), d_serr, sizeof(int32_t) * 2 * (size_t)nch, cudaMemcpyDeviceToHost,
                                              ctx->stream) != cudaSuccess ||
          cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
        set_error(ctx, std::string("xt_predict (shared plans): ") + cudaGetErrorString(cudaGetLastError()));
        result = XT_ERR_CUDA;
        break;
      }
      int sneed = 0;
      bool sgrouping = false;
      for (int i = 0; i < nch; ++i) {
        if (h_serr[i] == 1) sgrouping = true;
        sneed = std::max(sneed, h_serr[nch + i]);
      }
      if (sgrouping) {
        set_error(ctx, "problem with grouping: a state sequence ended ungrouped (threshold must be > 0 and the model finite)");
        result = XT_ERR_GROUPING;
        break;
      }
      if (sneed) {
        while (cap < sneed) cap *= 2;
        continue;
      }
      a.splan = d_splan;
      a.splan_stride = sa.splan_stride;
      e = xt_launch_k3_follow(a, *p, grid, nwarps, smem, ctx->stream);
    } else if (!piece_w0.empty() && !copied) {
      const int np = (int)piece_w0.size() - 1;
      for (int g = 0; g < np && e == cudaSuccess; ++g) {
        K3Args ag = a;
        ag.work = ctx->d_work + piece_w0[g];
        ag.n_work = piece_w0[g + 1] - piece_w0[g];
        ag.err = d_err + piece_w0[g];
        ag.err_need = d_err + n_work + piece_w0[g];
        e = xt_launch_k3(ag, *p, std::min(ag.n_work, grid), nwarps, smem, ctx->stream);
        cudaEventRecord(ev_piece[g], ctx->stream);
      }
      ctx->k3_launches += np - 1;
      cudaEventRecord(ctx->ev_k3[1], ctx->stream);
      // The caller's arrays are usually fresh allocations: every 4 KB page faults at its first write, which bounds a
      // pageable read-back to ~2 GB/s when the copy takes the faults.  A few host threads touch the pages, segment by
      // segment, while the kernels run; the copy of a segment starts once its pages are resident.
      const int n_seg_all = (int)ctx->seg_n.size();
      std::atomic<int> touched{0};
      std::thread toucher([&] {
        const unsigned hw = std::thread::hardware_concurrency();
        const int nthr = (int)std::max(1u, std::min(8u, hw / 2));
        for (int sg = 0; sg < n_seg_all; ++sg) {
          char* base = (char*)out[sg];
          const size_t bytes = sizeof(double) * (size_t)ctx->seg_n[sg] * ctx->seg_L[sg] * nS;
          auto touch = [base](size_t b0, size_t b1) {
            for (size_t o = b0; o < b1; o += 4096) *(volatile char*)(base + o) = 0;
            if (b1 > b0) *(volatile char*)(base + b1 - 1) = 0;
          };
          if (bytes >= ((size_t)8 << 20) && nthr > 1) {
            std::vector<std::thread> th;
            const size_t per = ((bytes / nthr) + 4095) & ~(size_t)4095;
            for (int t = 0; t < nthr; ++t) {
              const size_t b0 = std::min(bytes, per * t), b1 = std::min(bytes, per * (t + 1));
              if (b1 > b0) th.emplace_back(touch, b0, b1);
            }
            for (auto& x : th) x.join();
          } else if (bytes) {
            touch(0, bytes);
          }
          touched.store(sg + 1, std::memory_order_release);
        }
      });
      for (int g = 0; g < np && e == cudaSuccess; ++g) {
        cudaStreamWaitEvent(s_copy, ev_piece[g], 0);
        for (int sg = piece_s0[g]; sg < piece_s0[g + 1] && e == cudaSuccess; ++sg) {
          const XtChunk& c0 = ctx->chunks[ctx->seg_chunk0[sg]];
          const size_t cnt = (size_t)ctx->seg_n[sg] * ctx->seg_L[sg] * nS;
          while (touched.load(std::memory_order_acquire) <= sg) std::this_thread::yield();
          if (cnt) e = cudaMemcpyAsync(out[sg], d_pred + (size_t)c0.loc_off * nS, sizeof(double) * cnt, cudaMemcpyDeviceToHost, s_copy);
        }
      }
      toucher.join();
      if (e == cudaSuccess) e = cudaStreamSynchronize(s_copy);
      copied = true;
    } else {
      e = xt_launch_k3(a, *p, grid, nwarps, smem, ctx->stream);
      cudaEventRecord(ctx->ev_k3[1], ctx->stream);
    }
    if (shared) cudaEventRecord(ctx->ev_k3[1], ctx->stream);
    if (e != cudaSuccess || cudaMemcpyAsync(h_err.data(), d_err, sizeof(int32_t) * 2 * (size_t)n_work, cudaMemcpyDeviceToHost,
                                            ctx->stream) != cudaSuccess ||
        cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
      set_error(ctx, std::string("xt_predict: ") + cudaGetErrorString(cudaGetLastError()));
      result = XT_ERR_CUDA;
      break;
    }
    int need = 0;
    bool grouping = false;
    std::vector<XtWork> again;
    for (int i = 0; i < n_cur; ++i) {
      if (h_err[i] == 1) grouping = true;
      need = std::max(need, h_err[n_work + i]);
      if (h_err[i] == 2 && !shared) again.push_back(retry.empty() ? ctx->work[i] : retry[i]);
    }
    {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, ctx->ev_k3[0], ctx->ev_k3[1]);
      ms_total += ms;
    }
    if (grouping) {
      set_error(ctx, "problem with grouping: a state sequence ended ungrouped (threshold must be > 0 and the model finite)");
      result = XT_ERR_GROUPING;
      break;
    }
    if (!need) {
      ctx->ms_predict = ms_total;
      break;
    }
    while (cap < need) cap *= 2;
    ran_again = true;
    if (!shared) {  // only the work items that overflowed run again
      retry.swap(again);
      n_cur = (int)retry.size();
      cudaFree(d_retry);
      d_retry = nullptr;
      if (cudaMalloc(&d_retry, sizeof(XtWork) * retry.size()) != cudaSuccess ||
          cudaMemcpyAsync(d_retry, retry.data(), sizeof(XtWork) * retry.size(), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) {
        set_error(ctx, "xt_predict: cannot stage the work items to run again");
        result = XT_ERR_CUDA;
        break;
      }
      d_work_cur = d_retry;
    }
  }
  if (result == XT_OK && (!copied || ran_again)) {
    for (size_t s = 0; s < ctx->seg_n.size(); ++s) {
      const XtChunk& c0 = ctx->chunks[ctx->seg_chunk0[s]];
      const size_t cnt = (size_t)ctx->seg_n[s] * ctx->seg_L[s] * nS;
      if (cudaMemcpyAsync(out[s], d_pred + (size_t)c0.loc_off * nS, sizeof(double) * cnt, cudaMemcpyDeviceToHost, ctx->stream) !=
          cudaSuccess) {
        set_error(ctx, "xt_predict: device to host copy failed");
        result = XT_ERR_CUDA;
        break;
      }
    }
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess && result == XT_OK) {
      set_error(ctx, "xt_predict: synchronisation failed");
      result = XT_ERR_CUDA;
    }
  }
  for (auto ev : ev_piece)
    if (ev) cudaEventDestroy(ev);
  if (s_copy) cudaStreamDestroy(s_copy);
  cudaFree(d_pred);
  cudaFree(d_err);
  cudaFree(d_scratch);
  cudaFree(d_retry);
  cudaFree(d_splan);
  cudaFree(d_sscratch);
  cudaFree(d_serr);
  return result;
}
