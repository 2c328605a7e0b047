// Translation unit of the first-generation linear-domain replay kernel (k2_replay_lin, cross-check) and of the
// log-domain kernel with its state in global memory (k2_replay, last resort).
#include "xt_launch.h"
#include "xt_replay.cuh"
#include "xt_replay_lin.cuh"

template <int D, int KS, int WPC>
static cudaError_t launch_k2_lin(const K2Args& a, const xt_params& p, const K2Lin& lin, size_t smem, int grid, cudaStream_t stream) {
  auto kern = k2_replay_lin<D, KS, WPC>;
  static unsigned long long smem_ok = 0;
  cudaError_t e = xt_allow_smem(kern, smem, &smem_ok);
  if (e != cudaSuccess) return e;
  kern<<<grid, 32 * WPC, smem, stream>>>(a, p, lin);
  return cudaGetLastError();
}

template <int D, int KS>
static cudaError_t launch_k2(const K2Args& a, const xt_params& p, const K2Lin& lin, size_t smem, bool use_smem, int grid, int wpc,
                             cudaStream_t stream) {
  if (xt_is_var(&p)) {  // log-domain kernel, state in global memory
    k2_replay<D, KS, false, true><<<grid, 32, 0, stream>>>(a, p);
    return cudaGetLastError();
  }
  if (use_smem) {
    if (wpc == 8) return launch_k2_lin<D, KS, 8>(a, p, lin, smem, grid, stream);
    if (wpc == 2) return launch_k2_lin<D, KS, 2>(a, p, lin, smem, grid, stream);
    return launch_k2_lin<D, KS, 4>(a, p, lin, smem, grid, stream);
  }
  k2_replay<D, KS, false><<<grid, 32, 0, stream>>>(a, p);
  return cudaGetLastError();
}

cudaError_t xt_launch_k2_old(const K2Args& a, const xt_params& p, const K2Lin& lin, size_t smem, bool use_smem, int grid,
                             int wpc, cudaStream_t stream) {
  cudaError_t e = cudaSuccess;
#define CALL_K2(D_, KS_) e = launch_k2<D_, KS_>(a, p, lin, smem, use_smem, grid, wpc, stream)
  XT_DISPATCH(p.d, p.n_loc, CALL_K2);
#undef CALL_K2
  return e;
}
