// Translation unit of the fused FP64 replay kernel (all instantiations of k2_replay_fused).
#include "xt_launch.h"
#include "xt_replay_fused.cuh"

template <int D, int KS, int WPC, int TPT, bool VAR, bool GST>
static cudaError_t launch_one(const K2FArgs& a, const K2Tab& tab, size_t smem, cudaStream_t stream) {
  auto kern = k2_replay_fused<D, KS, WPC, TPT, VAR, GST>;
  static unsigned long long smem_ok = 0;
  cudaError_t e = xt_allow_smem(kern, smem, &smem_ok);
  if (e != cudaSuccess) return e;
  kern<<<a.n_work, 32 * WPC, smem, stream>>>(a, tab);
  return cudaGetLastError();
}

template <int D, int KS>
static cudaError_t launch_k2_fused(const K2FArgs& a, const K2Tab& tab, size_t smem, int wpc, int tpt, cudaStream_t stream, bool var) {
  // state in global memory / peak-wise LocErr or per-track dt: one configuration each (4 warps per tile, one track per thread)
  if (a.gstate) return launch_one<D, KS, 4, 1, false, true>(a, tab, smem, stream);
  if (var) return launch_one<D, KS, 4, 1, true, false>(a, tab, smem, stream);
  if (tpt == 2) {
    if (wpc == 8) return launch_one<D, KS, 8, 2, false, false>(a, tab, smem, stream);
    if (wpc == 2) return launch_one<D, KS, 2, 2, false, false>(a, tab, smem, stream);
    return launch_one<D, KS, 4, 2, false, false>(a, tab, smem, stream);
  }
  if (wpc == 8) return launch_one<D, KS, 8, 1, false, false>(a, tab, smem, stream);
  if (wpc == 2) return launch_one<D, KS, 2, 1, false, false>(a, tab, smem, stream);
  return launch_one<D, KS, 4, 1, false, false>(a, tab, smem, stream);
}

cudaError_t xt_launch_k2_fused(int d, int ks, const K2FArgs& a, const K2Tab& tab, size_t smem, int wpc, int tpt,
                               cudaStream_t stream, bool var) {
  cudaError_t e = cudaSuccess;
#define CALL_K2F(D_, KS_) e = launch_k2_fused<D_, KS_>(a, tab, smem, wpc, tpt, stream, var)
  XT_DISPATCH(d, ks, CALL_K2F);
#undef CALL_K2F
  return e;
}
