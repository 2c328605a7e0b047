// Translation unit of the plan kernel, 256 threads per chunk (several chunks per SM).
#include "xt_k1_impl.cuh"

cudaError_t xt_launch_k1_512(const K1Args& a, const xt_params& p, size_t smem, int n_chunks, cudaStream_t stream);
cudaError_t xt_launch_k1_1024(const K1Args& a, const xt_params& p, size_t smem, int n_chunks, cudaStream_t stream);

cudaError_t xt_launch_k1(const K1Args& a, const xt_params& p, size_t smem, int n_chunks, cudaStream_t stream, int nthreads) {
  if (nthreads == 1024) return xt_launch_k1_1024(a, p, smem, n_chunks, stream);
  if (nthreads == 512) return xt_launch_k1_512(a, p, smem, n_chunks, stream);
  return launch_k1_nt<XT_K1_THREADS>(a, p, smem, n_chunks, stream);
}
