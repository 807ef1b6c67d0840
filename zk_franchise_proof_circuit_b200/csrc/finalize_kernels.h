// Host-callable launchers of the proof assembly kernels (finalize.cu)
#pragma once
#include <cuda_runtime.h>
#include "finalize.cuh"

namespace zkb {

struct FinalizeParams {
  const XYZZ<Fq> *g1;      // [n][3]: pi_a', pi_b1', pi_c'
  const XYZZ<Fq> *g1h;     // [n]:    pi_h
  const XYZZ<Fq2> *g2;     // [n]:    pi_b'
  const Fr *rs;            // [n][2] canonical r, s
  const Fr *wtns;          // [n][n_wires]
  size_t wtns_stride;
  const Affine<Fq> *alpha1, *beta1, *d1tab;
  const Affine<Fq2> *beta2, *d2tab;
  const XYZZ<Fq> *tconst1; // [3] per-key constant parts of pi_a', pi_b1', pi_c' (NULL = none)
  const XYZZ<Fq2> *tconst2; // [1] of pi_b'
  XYZZ<Fq> *scratch;       // [n][2][15] window tables of the two variable-base products
  uint8_t *out;            // [n][256 + n_public*32]
  uint32_t n_public, n;
};

cudaError_t launch_fixed_tables(Affine<Fq> *d1tab, const Affine<Fq> *delta1, Affine<Fq2> *d2tab, const Affine<Fq2> *delta2,
                                cudaStream_t st);
cudaError_t launch_finalize(const FinalizeParams &P, cudaStream_t st);

}  // namespace zkb
