// In-process multi-GPU objective: one xt_multi drives several single-device contexts from the one
// thread that calls the objective (lmfit calls cum_Proba_Cs sequentially from one Python thread,
// tracking.py:1371; the reference's own parallelism is a process pool over chunks, :1061-1063).
//
// Chunks are the atomic unit (the grouping plan is decided per chunk, tracking.py:677-691): the chunk
// list of the whole data set (numbered in upload order, like xt_upload) is dealt to the devices
// longest-processing-time-first on nT * (L - 1); every device uploads only its chunks.  One worker
// thread per device (the launches of one evaluation cost tens of microseconds of host time per device;
// issued from one thread they would serialise), woken per call.  Result: every device returns its
// per-chunk sums and the caller adds them in global chunk order, so the objective has the same bits for
// any number of devices (BFGS finite differences rely on reproducibility, and a fit does not change with
// the GPU count).  No collective is needed in process: the 8-byte-per-chunk read-back replaces the
// all-reduce of the one-process-per-GPU mode.
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>

struct xt_multi {
  std::vector<xt_ctx*> ctx;
  std::vector<int> dev;
  std::string err;
  // chunk layout of the resident data set
  std::vector<std::vector<int>> gchunks;       // per device: global chunk ids, ascending
  std::vector<std::pair<int, int>> where;      // global chunk -> (device, local chunk)
  std::vector<int64_t> dev_steps;              // track-steps per device
  int d = 0, chunk_size = 0;
  // workers (device 0 is served by the calling thread)
  std::vector<std::thread> th;
  std::mutex mu;
  std::condition_variable cv_go, cv_done;
  uint64_t gen = 0;
  int pending = 0;
  bool quit = false;
  std::atomic<uint64_t> gen_atomic{0};  // mirrors of gen / pending / quit for the short spin phases
  std::atomic<int> pending_atomic{0};
  std::atomic<bool> quit_atomic{false};
  std::function<int(int)> job;
  std::vector<int> rc;
};

static std::string g_multi_create_error;

static void xt_multi_worker(xt_multi* m, int g) {
  uint64_t seen = 0;
  for (;;) {
    std::function<int(int)> job;
    {
      // objective calls of a fit follow each other within ~0.1 ms: look at the generation counter for that long before
      // going to sleep on the condition variable (a futex wake-up costs about as much as the GPU work of a small shard)
      const auto t0 = std::chrono::steady_clock::now();
      while (m->gen_atomic.load(std::memory_order_acquire) == seen && !m->quit_atomic.load(std::memory_order_acquire) &&
             std::chrono::steady_clock::now() - t0 < std::chrono::microseconds(200)) {
      }
      std::unique_lock<std::mutex> lk(m->mu);
      m->cv_go.wait(lk, [&] { return m->quit || m->gen != seen; });
      if (m->quit) return;
      seen = m->gen;
      job = m->job;
    }
    const int r = job(g);
    {
      std::lock_guard<std::mutex> lk(m->mu);
      m->rc[g] = r;
      if (--m->pending == 0) m->cv_done.notify_one();
      m->pending_atomic.store(m->pending, std::memory_order_release);
    }
  }
}

// run job(g) for every device g (device 0 on the calling thread); returns the first error
static int xt_multi_run(xt_multi* m, const std::function<int(int)>& job) {
  const int n = (int)m->ctx.size();
  if (n > 1) {
    std::lock_guard<std::mutex> lk(m->mu);
    m->job = job;
    m->pending = n - 1;
    m->pending_atomic.store(n - 1, std::memory_order_release);
    ++m->gen;
    m->gen_atomic.store(m->gen, std::memory_order_release);
  }
  if (n > 1) m->cv_go.notify_all();
  m->rc[0] = job(0);
  if (n > 1) {
    const auto t0 = std::chrono::steady_clock::now();  // the other devices finish within microseconds of this one
    while (m->pending_atomic.load(std::memory_order_acquire) != 0 &&
           std::chrono::steady_clock::now() - t0 < std::chrono::microseconds(200)) {
    }
    std::unique_lock<std::mutex> lk(m->mu);
    m->cv_done.wait(lk, [&] { return m->pending == 0; });
  }
  for (int g = 0; g < n; ++g)
    if (m->rc[g]) {
      m->err = "device " + std::to_string(m->dev[g]) + ": " + m->ctx[g]->err;
      return m->rc[g];
    }
  return XT_OK;
}

extern "C" int xt_multi_create(const int32_t* dev_ids, int32_t n_dev, xt_multi** out) {
  if (!out || n_dev < 1 || n_dev > 64 || !dev_ids) {
    g_multi_create_error = "xt_multi_create: need 1 <= n_dev <= 64 device ordinals";
    return XT_ERR_ARG;
  }
  // (an ordinal may repeat: several contexts on one GPU are logical shards, used by the tests on one-GPU boxes)
  xt_multi* m = new xt_multi();
  for (int i = 0; i < n_dev; ++i) {
    xt_ctx* c = nullptr;
    const int rc = xt_create(dev_ids[i], &c);
    if (rc) {
      g_multi_create_error = xt_last_error(nullptr);
      for (xt_ctx* o : m->ctx) xt_destroy(o);
      delete m;
      return rc;
    }
    m->ctx.push_back(c);
    m->dev.push_back(dev_ids[i]);
  }
  m->rc.assign(n_dev, 0);
  for (int g = 1; g < n_dev; ++g) m->th.emplace_back(xt_multi_worker, m, g);
  *out = m;
  return XT_OK;
}

extern "C" void xt_multi_destroy(xt_multi* m) {
  if (!m) return;
  {
    std::lock_guard<std::mutex> lk(m->mu);
    m->quit = true;
    m->quit_atomic.store(true, std::memory_order_release);
  }
  m->cv_go.notify_all();
  for (std::thread& t : m->th) t.join();
  for (xt_ctx* c : m->ctx) xt_destroy(c);
  delete m;
}

extern "C" const char* xt_multi_last_error(xt_multi* m) { return m ? m->err.c_str() : g_multi_create_error.c_str(); }

extern "C" int xt_multi_n_devices(xt_multi* m) { return m ? (int)m->ctx.size() : 0; }

// Longest-processing-time-first assignment of the chunk list (cost nT * (L - 1); ties: lower chunk id first, lower
// device first) — the same rule as tracking.shard_chunks of the one-process-per-GPU mode.
static void xt_multi_shard(const std::vector<int64_t>& cost, int n_dev, std::vector<std::vector<int>>* owner) {
  std::vector<int> order(cost.size());
  for (size_t i = 0; i < cost.size(); ++i) order[i] = (int)i;
  std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return cost[x] > cost[y]; });
  std::vector<int64_t> load(n_dev, 0);
  owner->assign(n_dev, {});
  for (int i : order) {
    int best = 0;
    for (int g = 1; g < n_dev; ++g)
      if (load[g] < load[best]) best = g;
    (*owner)[best].push_back(i);
    load[best] += cost[i];
  }
  for (auto& v : *owner) std::sort(v.begin(), v.end());
}

struct XtMultiSlice {  // one chunk of the caller's segment list
  int seg;
  int64_t t0;
  int nT;
};

static int xt_multi_layout(xt_multi* m, int32_t n_seg, const int32_t* L, const int64_t* n, int32_t d, int32_t chunk_size,
                           std::vector<XtMultiSlice>* slices) {
  if (n_seg <= 0 || chunk_size < 1 || d < 1 || d > XT_MAX_DIMS) {
    m->err = "xt_multi_upload: need n_seg >= 1, 1 <= d <= 3, chunk_size >= 1";
    return XT_ERR_ARG;
  }
  slices->clear();
  std::vector<int64_t> cost;
  for (int s = 0; s < n_seg; ++s) {
    if (L[s] < 2) {
      m->err = "minimal track length = 2, here track length = " + std::to_string(L[s]);
      return XT_ERR_ARG;
    }
    if (n[s] < 1) {
      m->err = "xt_multi_upload: empty segment";
      return XT_ERR_ARG;
    }
    for (int64_t a = 0; a < n[s]; a += chunk_size) {
      const int nT = (int)std::min<int64_t>(chunk_size, n[s] - a);
      slices->push_back(XtMultiSlice{s, a, nT});
      cost.push_back((int64_t)nT * (L[s] - 1));
    }
  }
  const int n_dev = (int)m->ctx.size();
  xt_multi_shard(cost, n_dev, &m->gchunks);
  m->where.assign(slices->size(), {0, 0});
  m->dev_steps.assign(n_dev, 0);
  for (int g = 0; g < n_dev; ++g)
    for (size_t k = 0; k < m->gchunks[g].size(); ++k) {
      m->where[m->gchunks[g][k]] = {g, (int)k};
      m->dev_steps[g] += cost[m->gchunks[g][k]];
    }
  m->d = d;
  m->chunk_size = chunk_size;
  return XT_OK;
}

extern "C" int xt_multi_upload(xt_multi* m, int32_t n_seg, const int32_t* L, const int64_t* n, const int32_t* isBL,
                               const double* const* xyz, int32_t d, int32_t chunk_size) {
  if (!m) return XT_ERR_ARG;
  std::vector<XtMultiSlice> sl;
  int rc = xt_multi_layout(m, n_seg, L, n, d, chunk_size, &sl);
  if (rc) return rc;
  // every chunk becomes one segment of its device (a pointer into the caller's buffer: no host copy)
  return xt_multi_run(m, [&](int g) -> int {
    const std::vector<int>& mine = m->gchunks[g];
    if (mine.empty()) {
      free_data(m->ctx[g]);
      return XT_OK;
    }
    std::vector<int32_t> Lg, bl;
    std::vector<int64_t> ng;
    std::vector<const double*> pg;
    for (int c : mine) {
      const XtMultiSlice& s = sl[c];
      Lg.push_back(L[s.seg]);
      bl.push_back(isBL[s.seg]);
      ng.push_back(s.nT);
      pg.push_back(xyz[s.seg] + (size_t)s.t0 * L[s.seg] * d);
    }
    return xt_upload(m->ctx[g], (int)mine.size(), Lg.data(), ng.data(), bl.data(), pg.data(), d, chunk_size);
  });
}

extern "C" int xt_multi_upload_aux(xt_multi* m, int32_t n_seg, const int32_t* L, const int64_t* n, int32_t k_sigma,
                                   const double* const* sigma, const double* const* dt) {
  if (!m || m->where.empty()) return XT_ERR_STATE;
  std::vector<XtMultiSlice> sl;
  for (int s = 0; s < n_seg; ++s)
    for (int64_t a = 0; a < n[s]; a += m->chunk_size) sl.push_back(XtMultiSlice{s, a, (int)std::min<int64_t>(m->chunk_size, n[s] - a)});
  if (sl.size() != m->where.size()) {
    m->err = "xt_multi_upload_aux: segments do not match the uploaded tracks";
    return XT_ERR_ARG;
  }
  return xt_multi_run(m, [&](int g) -> int {
    const std::vector<int>& mine = m->gchunks[g];
    if (mine.empty()) return XT_OK;
    std::vector<const double*> sg, tg;
    for (int c : mine) {
      const XtMultiSlice& s = sl[c];
      if (sigma && k_sigma > 0) sg.push_back(sigma[s.seg] + (size_t)s.t0 * L[s.seg] * k_sigma);
      if (dt) tg.push_back(dt[s.seg] + (size_t)s.t0 * L[s.seg]);
    }
    return xt_upload_aux(m->ctx[g], k_sigma, sg.empty() ? nullptr : sg.data(), tg.empty() ? nullptr : tg.data());
  });
}

// per-chunk field-of-view tables (rows in global chunk order), see xt_set_stay_tables
extern "C" int xt_multi_set_stay_tables(xt_multi* m, int32_t K, int32_t H, const double* Lp_stay, const double* L_leave) {
  if (!m || m->where.empty()) return XT_ERR_STATE;
  return xt_multi_run(m, [&](int g) -> int {
    const std::vector<int>& mine = m->gchunks[g];
    if (mine.empty()) return XT_OK;
    if (!Lp_stay || !L_leave) return xt_set_stay_tables(m->ctx[g], 0, 0, 0, nullptr, nullptr);
    std::vector<double> a(mine.size() * (size_t)K), b(mine.size() * (size_t)H);
    for (size_t k = 0; k < mine.size(); ++k) {
      std::copy(Lp_stay + (size_t)mine[k] * K, Lp_stay + (size_t)(mine[k] + 1) * K, a.begin() + k * K);
      std::copy(L_leave + (size_t)mine[k] * H, L_leave + (size_t)(mine[k] + 1) * H, b.begin() + k * H);
    }
    return xt_set_stay_tables(m->ctx[g], 0, K, H, a.data(), b.data());
  });
}

extern "C" int xt_multi_sum_logp(xt_multi* m, const xt_params* p, double* out) {
  if (!m || !out) return XT_ERR_ARG;
  if (m->where.empty()) {
    m->err = "no tracks uploaded";
    return XT_ERR_STATE;
  }
  int rc = xt_multi_run(m, [&](int g) -> int {
    xt_ctx* c = m->ctx[g];
    if (m->gchunks[g].empty()) return XT_OK;
    int r = evaluate(c, p, nullptr, nullptr);
    if (r) return r;
    if (c->csum_fetched) return XT_OK;  // (the verified evaluation already fetched the chunk sums)
    if (cudaMemcpyAsync(c->h_csum, c->d_csum, sizeof(double) * c->chunks.size(), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
        cudaStreamSynchronize(c->stream) != cudaSuccess) {
      c->err = "xt_multi_sum_logp: read-back of the chunk sums failed";
      return XT_ERR_CUDA;
    }
    return XT_OK;
  });
  if (rc) return rc;
  double acc = 0.0;
  for (const auto& w : m->where) acc += m->ctx[w.first]->h_csum[w.second];  // global chunk order
  *out = acc;
  return XT_OK;
}

extern "C" int xt_multi_chunk_logp(xt_multi* m, int32_t chunk, const xt_params* p, double* out) {
  if (!m || chunk < 0 || chunk >= (int)m->where.size()) return XT_ERR_ARG;
  const auto w = m->where[chunk];
  const int rc = xt_chunk_logp(m->ctx[w.first], w.second, p, out);
  if (rc) m->err = m->ctx[w.first]->err;
  return rc;
}

extern "C" int xt_multi_set_option(xt_multi* m, const char* name, int value) {
  if (!m) return XT_ERR_ARG;
  for (xt_ctx* c : m->ctx) {
    const int rc = xt_set_option(c, name, value);
    if (rc) {
      m->err = c->err;
      return rc;
    }
  }
  return XT_OK;
}

// device g's share: chunks and track-steps (load balance diagnostics)
extern "C" int xt_multi_device_load(xt_multi* m, int32_t g, int32_t* device, int32_t* n_chunks, int64_t* track_steps) {
  if (!m || g < 0 || g >= (int)m->ctx.size() || m->gchunks.empty()) return XT_ERR_ARG;
  *device = m->dev[g];
  *n_chunks = (int)m->gchunks[g].size();
  *track_steps = m->dev_steps[g];
  return XT_OK;
}

// aggregated counters of the last evaluation: sums over the devices; times and maxima = max over the devices
extern "C" int xt_multi_get_stats(xt_multi* m, xt_stats* out) {
  if (!m || !out) return XT_ERR_ARG;
  xt_stats acc{};
  bool first = true;
  for (size_t g = 0; g < m->ctx.size(); ++g) {
    if (m->gchunks.empty() || m->gchunks[g].empty()) continue;
    xt_stats s{};
    const int rc = xt_get_stats(m->ctx[g], &s);
    if (rc) {
      m->err = m->ctx[g]->err;
      return rc;
    }
    acc.n_tracks += s.n_tracks;
    acc.track_steps += s.track_steps;
    acc.seq_updates += s.seq_updates;
    acc.seq_groups += s.seq_groups;
    acc.n_chunks += s.n_chunks;
    acc.k1_launches += s.k1_launches;
    acc.k2_launches += s.k2_launches;
    acc.max_nB_in = std::max(acc.max_nB_in, s.max_nB_in);
    acc.ms_plan = std::max(acc.ms_plan, s.ms_plan);
    acc.ms_replay = std::max(acc.ms_replay, s.ms_replay);
    acc.pipelined = s.pipelined;
    acc.fp32 = s.fp32;
    acc.plan_verified = (first || acc.plan_verified) && s.plan_verified;  // 1: every device ran along its resident plan
    first = false;
    acc.replanned += s.replanned;
  }
  *out = acc;
  return XT_OK;
}
