// Translation unit of the segment-length histogram kernel (k4_seglen).
#include "xt_launch.h"
#include "xt_seglen.cuh"

template <int D, int KS, int EPT>
static cudaError_t launch_k4(const K4Args& a, const xt_params& p, int grid, size_t smem, cudaStream_t stream) {
  auto kern = k4_seglen<D, KS, EPT>;
  static unsigned long long smem_ok = 0;
  cudaError_t e = xt_allow_smem(kern, smem, &smem_ok);
  if (e != cudaSuccess) return e;
  kern<<<grid, XT_SEG_THREADS, smem, stream>>>(a, p);
  return cudaGetLastError();
}

// elements per thread of the sort: the smallest of 4 / 8 / 16 that covers a.n2 pairs with XT_SEG_THREADS threads
cudaError_t xt_launch_k4(const K4Args& a, const xt_params& p, int grid, size_t smem, cudaStream_t stream) {
  cudaError_t e = cudaSuccess;
  if (a.n2 > 16 * XT_SEG_THREADS) return cudaErrorInvalidValue;
#define CALL_K4(D_, KS_)                                                                                  \
  e = a.n2 <= 4 * XT_SEG_THREADS ? launch_k4<D_, KS_, 4>(a, p, grid, smem, stream)                        \
                                 : (a.n2 <= 8 * XT_SEG_THREADS ? launch_k4<D_, KS_, 8>(a, p, grid, smem, stream) \
                                                               : launch_k4<D_, KS_, 16>(a, p, grid, smem, stream))
  XT_DISPATCH(p.d, p.n_loc, CALL_K4);
#undef CALL_K4
  return e;
}
