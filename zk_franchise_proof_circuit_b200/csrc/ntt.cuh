// Fr NTT plan (twiddle tables + pass schedule) - see ntt.cu
#pragma once
#include <cuda_runtime.h>
#include "fp.cuh"

namespace zkb {

struct NttPlan {
  int logn = 0;
  Fr *tw_fwd = nullptr;       // omega_N^e,  e < N/2 (Montgomery form)
  Fr *tw_inv = nullptr;       // omega_N^-e
  Fr *coset_scale = nullptr;  // N^-1 * omega_2N^bitrev(p), p < N
  cudaError_t init(int logn, cudaStream_t st);
  void destroy();
  cudaError_t dif(Fr *data, int nvec, size_t vec_stride, bool inverse, bool with_coset_scale, cudaStream_t st) const;
  cudaError_t dit(Fr *data, int nvec, size_t vec_stride, bool inverse, cudaStream_t st) const;
};

cudaError_t ntt_bitrev(Fr *out, const Fr *in, int logn, int nvec, size_t vec_stride, cudaStream_t st);
cudaError_t fr_to_mont(Fr *x, size_t n, cudaStream_t st);
cudaError_t fr_from_mont(Fr *x, size_t n, cudaStream_t st);

}  // namespace zkb
