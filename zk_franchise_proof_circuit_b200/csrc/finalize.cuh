// Groth16 proof assembly, host/device code so the formulas can be unit-tested
// without a GPU:
//   A  = pi_a'  + alpha1 + r*delta1
//   B  = pi_b'  + beta2  + s*delta2            (G2)
//   B1 = pi_b1' + beta1  + s*delta1
//   C  = pi_c' + pi_h + s*A + r*B1 - (r*s)*delta1
// r*delta and s*delta use 4-bit fixed-base tables built once per key; s*A and r*B1 are variable-base.
#pragma once
#include "ec.cuh"
#include "ec_coop.cuh"

namespace zkb {

// tab[w * 15 + d - 1] = d * 16^w * P,  w < 64, d = 1..15
template <class F>
ZKB_HD void fixed_table_entry(const Affine<F> &base, uint32_t idx, Affine<F> &out) {
  uint32_t w = idx / 15, d = idx % 15 + 1;
  uint32_t k[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  k[(4 * w) >> 5] = d << ((4 * w) & 31);
  XYZZ<F> p = XYZZ<F>::from_affine(base), r;
  xyzz_mul_ni(&p, k, &r);
  xyzz_to_affine_ni(&r, &out);
}

// NOTE on style: these helpers accumulate straight into *out and never keep two XYZZ temporaries alive
// across an out-of-line call.  nvcc 12.9 was observed to give a caller's live XYZZ local and an inlined
// callee's accumulator the same stack slot in exactly that pattern (pi_a came out as 2*r*delta1); the parity
// tests pin the result either way.
template <class F>
ZKB_HD void fixed_mul(const Affine<F> *tab, const uint32_t k[8], XYZZ<F> *out) {
  *out = XYZZ<F>::infinity();
  for (int w = 0; w < 64; w++) {
    uint32_t d = (k[w >> 3] >> ((w & 7) * 4)) & 15u;
    if (d) xyzz_add_affine_ni(out, tab + w * 15 + d - 1);
  }
}

// variable-base k * P with 4-bit windows; tb = caller-provided scratch of 15 points; out must not alias p
template <class F>
ZKB_HD void var_mul(const XYZZ<F> *p, const uint32_t k[8], XYZZ<F> *tb, XYZZ<F> *out) {
  tb[0] = *p;
  for (int i = 1; i < 15; i++) { tb[i] = tb[i - 1]; xyzz_add_ni(&tb[i], p); }
  *out = XYZZ<F>::infinity();
  for (int w = 63; w >= 0; w--) {
    if (w != 63) for (int j = 0; j < 4; j++) xyzz_dbl_ni(out);
    uint32_t d = (k[w >> 3] >> ((w & 7) * 4)) & 15u;
    if (d) xyzz_add_ni(out, &tb[d - 1]);
  }
}

#if defined(__CUDACC__)
// Scalar products on a lane group (ec_coop.cuh): *out is written by the group's lane 0.
template <class F>
static __device__ __noinline__ void fixed_mul_coop(const Affine<F> *tab, const uint32_t k[8], XYZZ<F> *out, const LaneGroup &g) {
  XYZZ<F> acc = XYZZ<F>::infinity();
  for (int w = 0; w < 64; w++) {
    uint32_t d = (k[w >> 3] >> ((w & 7) * 4)) & 15u;
    if (d) {
      Affine<F> t = tab[w * 15 + d - 1];
      coop_madd(&acc, &t, g);
    }
  }
  if (g.l == 0) *out = acc;
}
// var_mul on a lane group: *p is readable by all four lanes, tb = 15 points of scratch shared by the group
static __device__ __noinline__ void var_mul_coop(const XYZZ<Fq> *p, const uint32_t k[8], XYZZ<Fq> *tb, XYZZ<Fq> *out,
                                                 const LaneGroup &g) {
  XYZZ<Fq> base = *p, acc = base;
  if (g.l == 0) tb[0] = acc;
  for (int i = 1; i < 15; i++) {
    coop_add(&acc, &base, g);
    if (g.l == 0) tb[i] = acc;
  }
  __syncwarp(g.mask);
  acc = XYZZ<Fq>::infinity();
  for (int w = 63; w >= 0; w--) {
    if (w != 63) for (int j = 0; j < 4; j++) coop_dbl(&acc, g);
    uint32_t d = (k[w >> 3] >> ((w & 7) * 4)) & 15u;
    if (d) {
      XYZZ<Fq> t = tb[d - 1];
      coop_add(&acc, &t, g);
    }
  }
  if (g.l == 0) *out = acc;
}
#endif

// out = partial + fixed + k * delta   (A with k = r, B1 with k = s; G2: B with k = s)
template <class F>
ZKB_HD void fin_point(const XYZZ<F> *partial, const Affine<F> *fixed, const Affine<F> *dtab, const uint32_t k[8],
                      XYZZ<F> *out) {
  fixed_mul<F>(dtab, k, out);
  xyzz_add_affine_ni(out, fixed);
  xyzz_add_ni(out, partial);
}

// out = -(r*s) * delta1
ZKB_HD void fin_neg_rs_delta(const Affine<Fq> *d1tab, const Fr &r, const Fr &s, XYZZ<Fq> *out) {
  Fr rs = (r.to_mont() * s.to_mont()).from_mont();
  fixed_mul<Fq>(d1tab, rs.v, out);
  out->Y = out->Y.neg();
}

}  // namespace zkb
