// Translation unit of the optional single-precision replay kernel (k2_replay_f32).
#include "xt_launch.h"
#include "xt_replay_f32.cuh"

template <int D, int KS>
static cudaError_t launch_f32(const K2FArgs& a, const K2Tab& tab, size_t smem, cudaStream_t stream) {
  auto kern = k2_replay_f32<D, KS>;
  static unsigned long long smem_ok = 0;
  cudaError_t e = xt_allow_smem(kern, smem, &smem_ok);
  if (e != cudaSuccess) return e;
  kern<<<a.n_work, 128, smem, stream>>>(a, tab);
  return cudaGetLastError();
}

cudaError_t xt_launch_k2_f32(int d, int ks, const K2FArgs& a, const K2Tab& tab, size_t smem, cudaStream_t stream) {
  cudaError_t e = cudaSuccess;
#define CALL_F32(D_, KS_) e = launch_f32<D_, KS_>(a, tab, smem, stream)
  XT_DISPATCH(d, ks, CALL_F32);
#undef CALL_F32
  return e;
}
