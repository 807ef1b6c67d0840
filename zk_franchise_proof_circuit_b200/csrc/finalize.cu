// Proof assembly kernels - see finalize.cuh for the formulas
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "finalize_kernels.h"

namespace zkb {

// 4-bit fixed-base tables for delta1 / delta2 (finalize.cuh)
template <class F>
__global__ void k_fixed_table(Affine<F> *tab, const Affine<F> *base) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 64 * 15) return;
  Affine<F> b = *base, a;
  fixed_table_entry<F>(b, i, a);
  tab[i] = a;
}


// one CTA of 4 warps per proof, one warp per independent piece, so the scalar multiplications run on
// four schedulers instead of serialising as divergent lanes.  Two phases keep the affine conversions (one field
// inversion each) off the longest chain, the variable-base products:
//   phase 1   warp 0: A = pi_a' + alpha1 + r*delta1      warp 1: B1 = pi_b1' + beta1 + s*delta1
//             warp 2: -(r*s)*delta1, public signals      warp 3: B = pi_b' + beta2 + s*delta2   (G2)
//             (warp 3 continues with B -> affine; it does not take part in the phase barrier)
//   phase 2   warp 0: s*A      warp 1: r*B1   (four lanes each share every point doubling / addition)
//             warp 2: A -> affine, then T = -(r*s)*delta1 + pi_c' + pi_h
//   then      warp 0: C = s*A + r*B1 + T -> affine
__global__ void __launch_bounds__(128) k_finalize(FinalizeParams P) {
  __shared__ XYZZ<Fq> sh[5];          // A, s*A | B1, r*B1 | T
  __shared__ XYZZ<Fq2> shB;
  const uint32_t p = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t *out = P.out + (size_t)p * (256 + 32 * P.n_public);
  XYZZ<Fq> *scratch = P.scratch + ((size_t)p * 2) * 15;
  const Fr r = P.rs[2 * p], s = P.rs[2 * p + 1];
  // phase 1, the fixed-base products k * delta on the first lane group of each warp (ec_coop.cuh) ...
  const LaneGroup grp = LaneGroup::of(lane);
  if (lane < 4) {
    if (warp == 0) fixed_mul_coop<Fq>(P.d1tab, r.v, &sh[0], grp);
    else if (warp == 1) fixed_mul_coop<Fq>(P.d1tab, s.v, &sh[2], grp);
    else if (warp == 2) {
      Fr rs = (r.to_mont() * s.to_mont()).from_mont();
      fixed_mul_coop<Fq>(P.d1tab, rs.v, &sh[4], grp);
    } else fixed_mul_coop<Fq2>(P.d2tab, s.v, &shB, grp);
  }
  __syncwarp();
  // ... then lane 0 adds the fixed point and the MSM partial sum
  if (lane == 0) {
    if (warp == 0) {
      xyzz_add_affine_ni(&sh[0], P.alpha1);
      xyzz_add_ni(&sh[0], P.g1 + 3 * p);
      if (P.tconst1) xyzz_add_ni(&sh[0], P.tconst1);
    } else if (warp == 1) {
      xyzz_add_affine_ni(&sh[2], P.beta1);
      xyzz_add_ni(&sh[2], P.g1 + 3 * p + 1);
      if (P.tconst1) xyzz_add_ni(&sh[2], P.tconst1 + 1);
    } else if (warp == 2) {
      sh[4].Y = sh[4].Y.neg();                     // -(r s) delta1
      const Fr *w = P.wtns + (size_t)p * P.wtns_stride;
      for (uint32_t i = 0; i < P.n_public; i++) memcpy(out + 256 + 32 * i, w[1 + i].v, 32);
    } else {                     // the G2 point is the longest phase-1 piece: warp 3 goes straight on to its affine form
      xyzz_add_affine_ni(&shB, P.beta2);
      xyzz_add_ni(&shB, P.g2 + p);
      if (P.tconst2) xyzz_add_ni(&shB, P.tconst2);
      Affine<Fq2> b;
      xyzz_to_affine_ni(&shB, &b);
      Fq c[4] = {b.x.a.from_mont(), b.x.b.from_mont(), b.y.a.from_mont(), b.y.b.from_mont()};
      memcpy(out + 64, c, 128);
    }
  }
  if (warp < 3) asm volatile("bar.sync 1, 96;" ::: "memory");   // warps 0-2 only: A, B1 and -(rs)delta1 are ready
  if (warp < 2 && lane < 4) {          // the two variable-base products, four lanes each (finalize.cuh, *_coop)
    if (warp == 0) var_mul_coop(&sh[0], s.v, scratch, &sh[1], grp);
    else var_mul_coop(&sh[2], r.v, scratch + 15, &sh[3], grp);
  }
  if (lane == 0) {
    if (warp == 2) {
      Affine<Fq> a;
      xyzz_to_affine_ni(&sh[0], &a);
      Fq x = a.x.from_mont(), y = a.y.from_mont();
      memcpy(out, x.v, 32);
      memcpy(out + 32, y.v, 32);
      xyzz_add_ni(&sh[4], P.g1 + 3 * p + 2);
      xyzz_add_ni(&sh[4], P.g1h + p);
      if (P.tconst1) xyzz_add_ni(&sh[4], P.tconst1 + 2);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    // C = pi_c' + pi_h + s*A + r*B1 - (r*s)*delta1, accumulated in sh[1]
    xyzz_add_ni(&sh[1], &sh[3]);
    xyzz_add_ni(&sh[1], &sh[4]);
    Affine<Fq> c;
    xyzz_to_affine_ni(&sh[1], &c);
    Fq x = c.x.from_mont(), y = c.y.from_mont();
    memcpy(out + 192, x.v, 32);
    memcpy(out + 224, y.v, 32);
  }
}

cudaError_t launch_fixed_tables(Affine<Fq> *d1tab, const Affine<Fq> *delta1, Affine<Fq2> *d2tab, const Affine<Fq2> *delta2,
                                cudaStream_t st) {
  k_fixed_table<Fq><<<(64 * 15 + 63) / 64, 64, 0, st>>>(d1tab, delta1);
  k_fixed_table<Fq2><<<(64 * 15 + 63) / 64, 64, 0, st>>>(d2tab, delta2);
  return cudaGetLastError();
}

cudaError_t launch_finalize(const FinalizeParams &P, cudaStream_t st) {
  k_finalize<<<P.n, 128, 0, st>>>(P);
  return cudaGetLastError();
}

}  // namespace zkb
