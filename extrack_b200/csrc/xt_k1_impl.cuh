// Launch dispatch of the plan kernel for one thread count (included by xt_k1.cu / xt_k1w.cu / xt_k1x.cu: one
// translation unit per thread count so that the instantiations compile in parallel).
#pragma once
#include "xt_launch.h"
#include "xt_plan.cuh"

template <int D, int KS, bool VAR, int NT, bool SS, bool VERIFY = false>
static cudaError_t launch_k1_one(const K1Args& a, const xt_params& p, size_t smem, int n_chunks, cudaStream_t stream) {
  auto kern = k1_plan<D, KS, VAR, NT, SS, VERIFY>;
  static unsigned long long smem_ok = 0;
  cudaError_t e = xt_allow_smem(kern, smem, &smem_ok);
  if (e != cudaSuccess) return e;
  kern<<<(unsigned)n_chunks, NT, smem, stream>>>(a, p);
  return cudaGetLastError();
}

template <int NT>
static cudaError_t launch_k1_nt(const K1Args& a, const xt_params& p, size_t smem, int n_chunks, cudaStream_t stream) {
  cudaError_t e = cudaSuccess;
  if (xt_is_var(&p)) {
#define CALL_K1V(D_, KS_) e = launch_k1_one<D_, KS_, true, NT, false>(a, p, smem, n_chunks, stream)
    XT_DISPATCH(p.d, p.n_loc, CALL_K1V);
#undef CALL_K1V
  } else if (a.scapC > 0) {  // scratch in shared memory, known at compile time
#define CALL_K1S(D_, KS_) e = launch_k1_one<D_, KS_, false, NT, true>(a, p, smem, n_chunks, stream)
    XT_DISPATCH(p.d, p.n_loc, CALL_K1S);
#undef CALL_K1S
  } else {
#define CALL_K1(D_, KS_) e = launch_k1_one<D_, KS_, false, NT, false>(a, p, smem, n_chunks, stream)
    XT_DISPATCH(p.d, p.n_loc, CALL_K1);
#undef CALL_K1
  }
  return e;
}

// verification mode (scalar models, scratch in shared memory)
template <int NT>
static cudaError_t launch_k1_verify_nt(const K1Args& a, const xt_params& p, size_t smem, int n_chunks, cudaStream_t stream) {
  cudaError_t e = cudaSuccess;
#define CALL_K1Y(D_, KS_) e = launch_k1_one<D_, KS_, false, NT, true, true>(a, p, smem, n_chunks, stream)
  XT_DISPATCH(p.d, p.n_loc, CALL_K1Y);
#undef CALL_K1Y
  return e;
}
