// Translation unit of the plan kernel, 1024 threads per chunk (a chunk per SM, many live sequences).
#include "xt_k1_impl.cuh"

cudaError_t xt_launch_k1_1024(const K1Args& a, const xt_params& p, size_t smem, int n_chunks, cudaStream_t stream) {
  return launch_k1_nt<1024>(a, p, smem, n_chunks, stream);
}
