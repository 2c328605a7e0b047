// Position refinement (extrack/refined_localization.py:207-338): combination of the two passes of the
// refinement recursion (k3_predict<.., FOLLOW, REFINE>, one over the track from its last to its first localisation and
// one in forward time with the transposed transition matrix) with the localisation itself.
//
// For localisation k of a track, get_pos_PDF (:229-296) pairs every sequence of pass 1 that has consumed the
// localisations after k with every sequence of pass 2 that has consumed those before k, if both are in the same state
// at k, and multiplies three Gaussians (:33-43): (mean, std) of either sequence and the localisation with its error.
// position_refinement (:329-337) then takes the weighted mean of the means and of the variances (first LocErr
// component).  The first / last localisation have only one pass next to them (:223-227, :291-294).
// One warp per (track, localisation): lanes stride over the pairs; weights relative to the largest log-weight.
#pragma once
#include "xt_common.cuh"

struct KRArgs {
  const XtChunk* chunks;
  const double* soa;
  const double* dump1;        // pass 1 (reverse time): entries [L - 1][capD][d + KS + 2] per track
  const double* dump2;        // pass 2 (forward time)
  const int64_t* dump_off1;   // per chunk
  const int64_t* dump_off2;
  const int32_t* ent_n1;      // [n_chunks][maxL] sequences per entry
  const int32_t* ent_n2;
  double* mu;                 // [sum over tracks of L][d]
  double* sigma;              // [sum over tracks of L]
  int32_t n_chunks, maxL, capD1, capD2;
  double le[XT_MAX_DIMS];     // localisation error(s), KS of them
};

template <int D, int KS>
__device__ __forceinline__ void xr_prod2(const double (&s1)[KS], const double (&s2)[KS], const double (&mu1)[D], const double (&mu2)[D],
                                         double (&sg)[KS], double (&mu)[D], double& LK) {
  // refined_localization.py:33-37 (component k of the stds applies to dimension k, or to all of them when KS = 1)
  double v[KS];
  LK = 0.0;
#pragma unroll
  for (int k = 0; k < KS; ++k) {
    v[k] = s1[k] * s1[k] + s2[k] * s2[k];
    sg[k] = sqrt((s1[k] * s1[k]) * (s2[k] * s2[k]) / v[k]);
  }
#pragma unroll
  for (int dim = 0; dim < D; ++dim) {
    const int k = (KS == 1) ? 0 : dim;
    mu[dim] = (mu1[dim] * (s2[k] * s2[k]) + mu2[dim] * (s1[k] * s1[k])) / v[k];
    const double df = mu1[dim] - mu2[dim];
    LK += -0.5 * log(XT_TWO_PI * v[k]) - df * df / (2.0 * v[k]);
  }
}

template <int D, int KS>
__global__ void __launch_bounds__(256) k_refine_combine(const KRArgs a) {
  constexpr int COD = D + KS + 2;
  const int lane = threadIdx.x & 31;
  const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  // (chunk, track, localisation) of this warp: chunks are whole buckets here, a short linear search is enough
  int c = 0;
  long long rest = gw;
  for (; c < a.n_chunks; ++c) {
    const long long n = (long long)a.chunks[c].nT * a.chunks[c].L;
    if (rest < n) break;
    rest -= n;
  }
  if (c >= a.n_chunks) return;
  const XtChunk ck = a.chunks[c];
  const int L = ck.L, t = (int)(rest / L), k = (int)(rest % L);
  const int E = L - 1;  // entries per pass
  const double* d1 = a.dump1 + a.dump_off1[c] + (size_t)t * E * a.capD1 * COD;
  const double* d2 = a.dump2 + a.dump_off2[c] + (size_t)t * E * a.capD2 * COD;
  const int32_t* n1v = a.ent_n1 + (size_t)c * a.maxL;
  const int32_t* n2v = a.ent_n2 + (size_t)c * a.maxL;
  double le[KS], ck_[D];
#pragma unroll
  for (int q = 0; q < KS; ++q) le[q] = a.le[q];
#pragma unroll
  for (int dim = 0; dim < D; ++dim) ck_[dim] = a.soa[ck.xyz_off + (size_t)(k * D + dim) * ck.nTpad + t];
  // sequences next to localisation k: pass 1 entry E-1-k (it has consumed L-1 .. k+1), pass 2 entry k-1
  const bool has1 = k <= L - 2, has2 = k >= 1;
  const double* A = has1 ? d1 + (size_t)(E - 1 - k) * a.capD1 * COD : nullptr;
  const double* B = has2 ? d2 + (size_t)(k - 1) * a.capD2 * COD : nullptr;
  const int nA = has1 ? n1v[E - 1 - k] : 1, nB = has2 ? n2v[k - 1] : 1;
  // the reference's first / last localisation use the LAST entry of the one pass (:223, :291)
  const bool edge = !(has1 && has2);
  if (k == 0) { A = d1 + (size_t)(E - 1) * a.capD1 * COD; }
  if (k == L - 1) { B = d2 + (size_t)(E - 1) * a.capD2 * COD; }
  const int nAe = (k == 0) ? n1v[E - 1] : nA, nBe = (k == L - 1) ? n2v[E - 1] : nB;
  const int npair = edge ? (k == 0 ? nAe : nBe) : nA * nB;
  double wmax = -INFINITY;
  for (int pass = 0; pass < 2; ++pass) {
    double sw = 0.0, smu[D], ss2 = 0.0;
#pragma unroll
    for (int dim = 0; dim < D; ++dim) smu[dim] = 0.0;
    for (int pi = lane; pi < npair; pi += 32) {
      double sg[KS], mu[D], LP;
      if (edge) {
        const double* S = (k == 0 ? A : B) + (size_t)pi * COD;
        double s1[KS], m1[D];
#pragma unroll
        for (int dim = 0; dim < D; ++dim) m1[dim] = S[dim];
#pragma unroll
        for (int q = 0; q < KS; ++q) s1[q] = S[D + q];
        double LK;
        xr_prod2<D, KS>(le, s1, ck_, m1, sg, mu, LK);
        LP = S[D + KS] + LK;
      } else {
        const int ia = pi / nB, ib = pi - ia * nB;
        const double* Sa = A + (size_t)ia * COD;
        const double* Sb = B + (size_t)ib * COD;
        if (Sa[D + KS + 1] != Sb[D + KS + 1]) continue;  // different states at k
        double sa[KS], ma[D], sb[KS], mb[D];
#pragma unroll
        for (int dim = 0; dim < D; ++dim) { ma[dim] = Sa[dim]; mb[dim] = Sb[dim]; }
#pragma unroll
        for (int q = 0; q < KS; ++q) { sa[q] = Sa[D + q]; sb[q] = Sb[D + q]; }
        double s12[KS], m12[D], LK1, LK2;
        xr_prod2<D, KS>(sa, le, ma, ck_, s12, m12, LK1);   // prod_3GaussPDF (:39-43)
        xr_prod2<D, KS>(s12, sb, m12, mb, sg, mu, LK2);
        LP = Sa[D + KS] + Sb[D + KS] + (LK1 + LK2);
      }
      if (pass == 0) {
        wmax = fmax(wmax, LP);
      } else {
        const double w = exp(LP - wmax);
        sw += w;
#pragma unroll
        for (int dim = 0; dim < D; ++dim) smu[dim] += w * mu[dim];
        ss2 += w * sg[0] * sg[0];
      }
    }
    if (pass == 0) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) wmax = fmax(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
    } else {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        sw += __shfl_xor_sync(0xffffffffu, sw, o);
        ss2 += __shfl_xor_sync(0xffffffffu, ss2, o);
#pragma unroll
        for (int dim = 0; dim < D; ++dim) smu[dim] += __shfl_xor_sync(0xffffffffu, smu[dim], o);
      }
      if (lane == 0) {
        const size_t row = (size_t)ck.loc_off + (size_t)t * L + k;
#pragma unroll
        for (int dim = 0; dim < D; ++dim) a.mu[row * D + dim] = smu[dim] / sw;
        a.sigma[row] = sqrt(ss2 / sw);
      }
    }
  }
}
