// Translation unit of the plan kernel (all instantiations of k1_plan).
#include "xt_launch.h"
#include "xt_plan.cuh"

template <int D, int KS, bool VAR, int NT>
static cudaError_t launch_k1_nt(const K1Args& a, const xt_params& p, size_t smem, int n_chunks, cudaStream_t stream) {
  auto kern = k1_plan<D, KS, VAR, NT>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  kern<<<(unsigned)n_chunks, NT, smem, stream>>>(a, p);
  return cudaGetLastError();
}
template <int D, int KS, bool VAR>
static cudaError_t launch_k1(const K1Args& a, const xt_params& p, size_t smem, int n_chunks, cudaStream_t stream, int nthreads) {
  if (nthreads == 1024) return launch_k1_nt<D, KS, VAR, 1024>(a, p, smem, n_chunks, stream);
  if (!VAR && a.scapC > 0) {  // scratch in shared memory, known at compile time
    auto kern = k1_plan<D, KS, false, XT_K1_THREADS, true>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<(unsigned)n_chunks, XT_K1_THREADS, smem, stream>>>(a, p);
    return cudaGetLastError();
  }
  return launch_k1_nt<D, KS, VAR, XT_K1_THREADS>(a, p, smem, n_chunks, stream);
}

cudaError_t xt_launch_k1(const K1Args& a, const xt_params& p, size_t smem, int n_chunks, cudaStream_t stream, int nthreads) {
  cudaError_t e = cudaSuccess;
  if (xt_is_var(&p)) {
#define CALL_K1V(D_, KS_) e = launch_k1<D_, KS_, true>(a, p, smem, n_chunks, stream, nthreads)
    XT_DISPATCH(p.d, p.n_loc, CALL_K1V);
#undef CALL_K1V
  } else {
#define CALL_K1(D_, KS_) e = launch_k1<D_, KS_, false>(a, p, smem, n_chunks, stream, nthreads)
    XT_DISPATCH(p.d, p.n_loc, CALL_K1);
#undef CALL_K1
  }
  return e;
}
