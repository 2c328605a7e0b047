// Translation unit of the segment-length histogram kernel (k4_seglen).
#include "xt_launch.h"
#include "xt_seglen.cuh"

cudaError_t xt_launch_k4(const K4Args& a, const xt_params& p, int grid, size_t smem, cudaStream_t stream) {
  cudaError_t e = cudaSuccess;
#define CALL_K4(D_, KS_)                                                                    \
  do {                                                                                      \
    auto kern = k4_seglen<D_, KS_>;                                                         \
    static unsigned long long smem_ok = 0; e = xt_allow_smem(kern, smem, &smem_ok); \
    if (e == cudaSuccess) {                                                                 \
      kern<<<grid, XT_SEG_THREADS, smem, stream>>>(a, p);                                   \
      e = cudaGetLastError();                                                               \
    }                                                                                       \
  } while (0)
  XT_DISPATCH(p.d, p.n_loc, CALL_K4);
#undef CALL_K4
  return e;
}
