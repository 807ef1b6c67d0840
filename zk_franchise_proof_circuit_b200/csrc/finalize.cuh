// Groth16 proof assembly, host/device code so the formulas can be unit-tested
// without a GPU:
//   A  = pi_a'  + alpha1 + r*delta1
//   B  = pi_b'  + beta2  + s*delta2            (G2)
//   B1 = pi_b1' + beta1  + s*delta1
//   C  = pi_c' + pi_h + s*A + r*B1 - (r*s)*delta1
// r*delta and s*delta use 4-bit fixed-base tables built once per key; s*A and r*B1 are variable-base.
#pragma once
#include "ec.cuh"

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
// ---- lane-cooperative G1 point arithmetic (device only) -------------------------------------------------------
// Lanes 0..3 of a warp hold IDENTICAL copies of every point and run the same control flow; the independent field
// products of one formula level are spread over the lanes (one product each) and broadcast back, so a doubling costs
// 4 product times instead of 9 and a full addition 5 instead of 14.  Same formulas and therefore the same (exact)
// results as XYZZ::dbl / XYZZ::add.
__device__ __forceinline__ Fq bc4(const Fq &x, int src) {
  Fq r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = __shfl_sync(0xFu, x.v[i], src);
  return r;
}
__device__ __forceinline__ Fq sel4(uint32_t l, const Fq &a, const Fq &b, const Fq &c, const Fq &d) {
  return Fq::select(l == 0, a, Fq::select(l == 1, b, Fq::select(l == 2, c, d)));
}
static __device__ __noinline__ void xyzz_dbl_coop(XYZZ<Fq> *p, uint32_t l) {
  if (p->is_inf()) return;
  const Fq U = p->Y.dbl();
  const Fq a1 = Fq::select(l == 0, U, p->X);
  const Fq r1 = a1 * a1;                                  // lane 0: V = U^2, lane 1: X^2
  const Fq V = bc4(r1, 0), A = bc4(r1, 1);
  const Fq r2 = sel4(l, U, p->X, p->ZZ, p->ZZ) * V;       // W = U V | S = X V | ZZ' = V ZZ
  const Fq W = bc4(r2, 0), S = bc4(r2, 1), ZZ3 = bc4(r2, 2);
  const Fq M = A.dbl() + A;
  const Fq r3 = Fq::select(l == 0, M, W) * sel4(l, M, p->Y, p->ZZZ, p->ZZZ);   // M^2 | W Y | ZZZ' = W ZZZ
  const Fq MM = bc4(r3, 0), T1 = bc4(r3, 1), ZZZ3 = bc4(r3, 2);
  const Fq X3 = MM - S.dbl();
  const Fq Y3 = M * (S - X3) - T1;
  p->X = X3; p->Y = Y3; p->ZZ = ZZ3; p->ZZZ = ZZZ3;
}
static __device__ __noinline__ void xyzz_add_coop(XYZZ<Fq> *p, const XYZZ<Fq> *o, uint32_t l) {
  if (o->is_inf()) return;
  if (p->is_inf()) { *p = *o; return; }
  const Fq r1 = sel4(l, p->X, o->X, p->Y, o->Y) * sel4(l, o->ZZ, p->ZZ, o->ZZZ, p->ZZZ);   // U1 | U2 | S1 | S2
  const Fq U1 = bc4(r1, 0), U2 = bc4(r1, 1), S1 = bc4(r1, 2), S2 = bc4(r1, 3);
  const Fq P = U2 - U1, R = S2 - S1;
  if (P.is_zero()) {
    if (R.is_zero()) xyzz_dbl_coop(p, l);
    else *p = XYZZ<Fq>::infinity();
    return;
  }
  const Fq r2 = sel4(l, P, R, p->ZZ, p->ZZZ) * sel4(l, P, R, o->ZZ, o->ZZZ);                 // PP | RR | ZZ1 ZZ2 | ZZZ1 ZZZ2
  const Fq PP = bc4(r2, 0), RR = bc4(r2, 1), ZZ12 = bc4(r2, 2), ZZZ12 = bc4(r2, 3);
  const Fq r3 = sel4(l, P, U1, ZZ12, ZZ12) * PP;                                             // PPP | Q | ZZ'
  const Fq PPP = bc4(r3, 0), Q = bc4(r3, 1), ZZ3 = bc4(r3, 2);
  const Fq r4 = Fq::select(l == 0, S1, ZZZ12) * PPP;                                         // S1 PPP | ZZZ'
  const Fq T = bc4(r4, 0), ZZZ3 = bc4(r4, 1);
  const Fq X3 = RR - PPP - Q.dbl();
  const Fq Y3 = R * (Q - X3) - T;
  p->X = X3; p->Y = Y3; p->ZZ = ZZ3; p->ZZZ = ZZZ3;
}
// mixed addition p += q (q affine), 5 product times instead of 10
static __device__ __noinline__ void xyzz_madd_coop(XYZZ<Fq> *p, const Affine<Fq> *q, uint32_t l) {
  if (q->is_inf()) return;
  if (p->is_inf()) { p->X = q->x; p->Y = q->y; p->ZZ = Fq::one(); p->ZZZ = Fq::one(); return; }
  const Fq r1 = Fq::select(l == 0, q->x, q->y) * Fq::select(l == 0, p->ZZ, p->ZZZ);          // U2 | S2
  const Fq P = bc4(r1, 0) - p->X, R = bc4(r1, 1) - p->Y;
  if (P.is_zero()) {
    if (R.is_zero()) *p = XYZZ<Fq>::dbl_affine(*q);
    else *p = XYZZ<Fq>::infinity();
    return;
  }
  const Fq a2 = Fq::select(l == 0, P, R);
  const Fq r2 = a2 * a2;                                                                      // PP | RR
  const Fq PP = bc4(r2, 0), RR = bc4(r2, 1);
  const Fq r3 = sel4(l, P, p->X, p->ZZ, p->ZZ) * PP;                                          // PPP | Q | ZZ'
  const Fq PPP = bc4(r3, 0), Q = bc4(r3, 1), ZZ3 = bc4(r3, 2);
  const Fq r4 = Fq::select(l == 0, p->Y, p->ZZZ) * PPP;                                       // Y PPP | ZZZ'
  const Fq T = bc4(r4, 0), ZZZ3 = bc4(r4, 1);
  const Fq X3 = RR - PPP - Q.dbl();
  const Fq Y3 = R * (Q - X3) - T;
  p->X = X3; p->Y = Y3; p->ZZ = ZZ3; p->ZZZ = ZZZ3;
}
// fixed_mul on lanes 0..3; *out written by lane 0
static __device__ __noinline__ void fixed_mul_coop(const Affine<Fq> *tab, const uint32_t k[8], XYZZ<Fq> *out, uint32_t l) {
  XYZZ<Fq> acc = XYZZ<Fq>::infinity();
  for (int w = 0; w < 64; w++) {
    uint32_t d = (k[w >> 3] >> ((w & 7) * 4)) & 15u;
    if (d) {
      Affine<Fq> t = tab[w * 15 + d - 1];
      xyzz_madd_coop(&acc, &t, l);
    }
  }
  if (l == 0) *out = acc;
}

// ---- G2: an Fq2 product is three Fq products - one each on lanes 0..2 (identical copies of all values) --------
__device__ __forceinline__ Fq bc3(const Fq &x, int src) {
  Fq r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = __shfl_sync(0x7u, x.v[i], src);
  return r;
}
__device__ __forceinline__ Fq2 fq2_mul_coop(const Fq2 &a, const Fq2 &b, uint32_t l) {   // Karatsuba, as Fq2::operator*
  const Fq x = l == 0 ? a.a : (l == 1 ? a.b : a.a + a.b), y = l == 0 ? b.a : (l == 1 ? b.b : b.a + b.b);
  const Fq t = x * y;
  const Fq t0 = bc3(t, 0), t1 = bc3(t, 1), t2 = bc3(t, 2);
  return {t0 - t1, t2 - t0 - t1};
}
__device__ __forceinline__ Fq2 fq2_sqr_coop(const Fq2 &a, uint32_t l) {                  // complex squaring, as Fq2::sqr
  const Fq x = l == 0 ? a.a + a.b : a.a, y = l == 0 ? a.a - a.b : a.b;
  const Fq t = x * y;
  const Fq t0 = bc3(t, 0), t1 = bc3(t, 1);
  return {t0, t1 + t1};
}
static __device__ __noinline__ void xyzz2_madd_coop(XYZZ<Fq2> *p, const Affine<Fq2> *q, uint32_t l) {
  if (q->is_inf()) return;
  if (p->is_inf()) { p->X = q->x; p->Y = q->y; p->ZZ = Fq2::one(); p->ZZZ = Fq2::one(); return; }
  const Fq2 U2 = fq2_mul_coop(q->x, p->ZZ, l), S2 = fq2_mul_coop(q->y, p->ZZZ, l);
  const Fq2 P = U2 - p->X, R = S2 - p->Y;
  if (P.is_zero()) {
    if (R.is_zero()) *p = XYZZ<Fq2>::dbl_affine(*q);
    else *p = XYZZ<Fq2>::infinity();
    return;
  }
  const Fq2 PP = fq2_sqr_coop(P, l), PPP = fq2_mul_coop(P, PP, l), Q = fq2_mul_coop(p->X, PP, l);
  const Fq2 X3 = fq2_sqr_coop(R, l) - PPP - Q.dbl();
  const Fq2 Y3 = fq2_mul_coop(R, Q - X3, l) - fq2_mul_coop(p->Y, PPP, l);
  p->ZZ = fq2_mul_coop(p->ZZ, PP, l);
  p->ZZZ = fq2_mul_coop(p->ZZZ, PPP, l);
  p->X = X3; p->Y = Y3;
}
static __device__ __noinline__ void fixed_mul2_coop(const Affine<Fq2> *tab, const uint32_t k[8], XYZZ<Fq2> *out, uint32_t l) {
  XYZZ<Fq2> acc = XYZZ<Fq2>::infinity();
  for (int w = 0; w < 64; w++) {
    uint32_t d = (k[w >> 3] >> ((w & 7) * 4)) & 15u;
    if (d) {
      Affine<Fq2> t = tab[w * 15 + d - 1];
      xyzz2_madd_coop(&acc, &t, l);
    }
  }
  if (l == 0) *out = acc;
}

// var_mul on lanes 0..3 (l = lane): *p is readable by all four, tb = 15 points of scratch shared by the group,
// *out is written by lane 0.
static __device__ __noinline__ void var_mul_coop(const XYZZ<Fq> *p, const uint32_t k[8], XYZZ<Fq> *tb, XYZZ<Fq> *out, uint32_t l) {
  XYZZ<Fq> base = *p, acc = base;
  if (l == 0) tb[0] = acc;
  for (int i = 1; i < 15; i++) {
    xyzz_add_coop(&acc, &base, l);
    if (l == 0) tb[i] = acc;
  }
  __syncwarp(0xFu);
  acc = XYZZ<Fq>::infinity();
  for (int w = 63; w >= 0; w--) {
    if (w != 63) for (int j = 0; j < 4; j++) xyzz_dbl_coop(&acc, l);
    uint32_t d = (k[w >> 3] >> ((w & 7) * 4)) & 15u;
    if (d) {
      XYZZ<Fq> t = tb[d - 1];
      xyzz_add_coop(&acc, &t, l);
    }
  }
  if (l == 0) *out = acc;
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
