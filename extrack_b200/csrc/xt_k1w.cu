// Translation unit of the plan kernel, 512 threads per chunk (a chunk per SM, <= 64 live sequences: up to 128 registers).
#include "xt_k1_impl.cuh"

cudaError_t xt_launch_k1_512(const K1Args& a, const xt_params& p, size_t smem, int n_chunks, cudaStream_t stream) {
  return launch_k1_nt<512>(a, p, smem, n_chunks, stream);
}
