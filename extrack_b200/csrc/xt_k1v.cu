// Translation unit of the plan kernel in verification mode (256 and 512 threads per chunk).
#include "xt_k1_impl.cuh"

cudaError_t xt_launch_k1_verify(const K1Args& a, const xt_params& p, size_t smem, int n_chunks, cudaStream_t stream, int nthreads) {
  if (nthreads == 512) return launch_k1_verify_nt<512>(a, p, smem, n_chunks, stream);
  return launch_k1_verify_nt<XT_K1_THREADS>(a, p, smem, n_chunks, stream);
}
