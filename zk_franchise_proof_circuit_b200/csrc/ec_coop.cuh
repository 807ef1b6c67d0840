// Lane-cooperative point arithmetic for the latency-bound kernels (proof assembly, small bucket reductions).
//
// A group of 4 consecutive lanes holds IDENTICAL copies of every point and runs the same control flow; the
// independent field products of one formula level are spread over the lanes (one product each) and broadcast back
// with shuffles.  G1: a doubling is 4 product times instead of 9, a full addition 5 instead of 14, a mixed addition
// 5 instead of 10.  G2: the three Fq products of an Fq2 product (two for a squaring) run on three lanes, so every Fq2
// product is one product time.  Same formulas as XYZZ::dbl / add / add_affine, hence the same (exact) results.
// Device only.
#pragma once
#include "ec.cuh"

#if defined(__CUDACC__)
namespace zkb {

struct LaneGroup {
  uint32_t l;      // lane within the group, 0..3
  uint32_t l0;     // warp lane of the group's lane 0
  uint32_t mask;   // shuffle mask of the group
  __device__ __forceinline__ static LaneGroup of(uint32_t warp_lane) {
    LaneGroup g;
    g.l = warp_lane & 3u;
    g.l0 = warp_lane & ~3u;
    g.mask = 0xFu << g.l0;
    return g;
  }
};

__device__ __forceinline__ Fq bcg(const LaneGroup &g, const Fq &x, uint32_t src) {
  Fq r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = __shfl_sync(g.mask, x.v[i], g.l0 + src);
  return r;
}
__device__ __forceinline__ Fq sel4(uint32_t l, const Fq &a, const Fq &b, const Fq &c, const Fq &d) {
  return Fq::select(l == 0, a, Fq::select(l == 1, b, Fq::select(l == 2, c, d)));
}

// ---- G1 ------------------------------------------------------------------------------------------------------
static __device__ __noinline__ void coop_dbl(XYZZ<Fq> *p, const LaneGroup &g) {
  if (p->is_inf()) return;
  const uint32_t l = g.l;
  const Fq U = p->Y.dbl();
  const Fq a1 = Fq::select(l == 0, U, p->X);
  const Fq r1 = a1 * a1;                                  // lane 0: V = U^2, lane 1: X^2
  const Fq V = bcg(g, r1, 0), A = bcg(g, r1, 1);
  const Fq r2 = sel4(l, U, p->X, p->ZZ, p->ZZ) * V;       // W = U V | S = X V | ZZ' = V ZZ
  const Fq W = bcg(g, r2, 0), S = bcg(g, r2, 1), ZZ3 = bcg(g, r2, 2);
  const Fq M = A.dbl() + A;
  const Fq r3 = Fq::select(l == 0, M, W) * sel4(l, M, p->Y, p->ZZZ, p->ZZZ);   // M^2 | W Y | ZZZ' = W ZZZ
  const Fq MM = bcg(g, r3, 0), T1 = bcg(g, r3, 1), ZZZ3 = bcg(g, r3, 2);
  const Fq X3 = MM - S.dbl();
  const Fq Y3 = M * (S - X3) - T1;
  p->X = X3; p->Y = Y3; p->ZZ = ZZ3; p->ZZZ = ZZZ3;
}
static __device__ __noinline__ void coop_add(XYZZ<Fq> *p, const XYZZ<Fq> *o, const LaneGroup &g) {
  if (o->is_inf()) return;
  if (p->is_inf()) { *p = *o; return; }
  const uint32_t l = g.l;
  const Fq r1 = sel4(l, p->X, o->X, p->Y, o->Y) * sel4(l, o->ZZ, p->ZZ, o->ZZZ, p->ZZZ);   // U1 | U2 | S1 | S2
  const Fq U1 = bcg(g, r1, 0), U2 = bcg(g, r1, 1), S1 = bcg(g, r1, 2), S2 = bcg(g, r1, 3);
  const Fq P = U2 - U1, R = S2 - S1;
  if (P.is_zero()) {
    if (R.is_zero()) coop_dbl(p, g);
    else *p = XYZZ<Fq>::infinity();
    return;
  }
  const Fq r2 = sel4(l, P, R, p->ZZ, p->ZZZ) * sel4(l, P, R, o->ZZ, o->ZZZ);                 // PP | RR | ZZ1 ZZ2 | ZZZ1 ZZZ2
  const Fq PP = bcg(g, r2, 0), RR = bcg(g, r2, 1), ZZ12 = bcg(g, r2, 2), ZZZ12 = bcg(g, r2, 3);
  const Fq r3 = sel4(l, P, U1, ZZ12, ZZ12) * PP;                                             // PPP | Q | ZZ'
  const Fq PPP = bcg(g, r3, 0), Q = bcg(g, r3, 1), ZZ3 = bcg(g, r3, 2);
  const Fq r4 = Fq::select(l == 0, S1, ZZZ12) * PPP;                                         // S1 PPP | ZZZ'
  const Fq T = bcg(g, r4, 0), ZZZ3 = bcg(g, r4, 1);
  const Fq X3 = RR - PPP - Q.dbl();
  const Fq Y3 = R * (Q - X3) - T;
  p->X = X3; p->Y = Y3; p->ZZ = ZZ3; p->ZZZ = ZZZ3;
}
static __device__ __noinline__ void coop_madd(XYZZ<Fq> *p, const Affine<Fq> *q, const LaneGroup &g) {
  if (q->is_inf()) return;
  if (p->is_inf()) { p->X = q->x; p->Y = q->y; p->ZZ = Fq::one(); p->ZZZ = Fq::one(); return; }
  const uint32_t l = g.l;
  const Fq r1 = Fq::select(l == 0, q->x, q->y) * Fq::select(l == 0, p->ZZ, p->ZZZ);          // U2 | S2
  const Fq P = bcg(g, r1, 0) - p->X, R = bcg(g, r1, 1) - p->Y;
  if (P.is_zero()) {
    if (R.is_zero()) *p = XYZZ<Fq>::dbl_affine(*q);
    else *p = XYZZ<Fq>::infinity();
    return;
  }
  const Fq a2 = Fq::select(l == 0, P, R);
  const Fq r2 = a2 * a2;                                                                      // PP | RR
  const Fq PP = bcg(g, r2, 0), RR = bcg(g, r2, 1);
  const Fq r3 = sel4(l, P, p->X, p->ZZ, p->ZZ) * PP;                                          // PPP | Q | ZZ'
  const Fq PPP = bcg(g, r3, 0), Q = bcg(g, r3, 1), ZZ3 = bcg(g, r3, 2);
  const Fq r4 = Fq::select(l == 0, p->Y, p->ZZZ) * PPP;                                       // Y PPP | ZZZ'
  const Fq T = bcg(g, r4, 0), ZZZ3 = bcg(g, r4, 1);
  const Fq X3 = RR - PPP - Q.dbl();
  const Fq Y3 = R * (Q - X3) - T;
  p->X = X3; p->Y = Y3; p->ZZ = ZZ3; p->ZZZ = ZZZ3;
}

// ---- G2 ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ Fq2 fq2_mul_coop(const Fq2 &a, const Fq2 &b, const LaneGroup &g) {   // Karatsuba, as Fq2::operator*
  const uint32_t l = g.l;
  const Fq x = l == 0 ? a.a : (l == 1 ? a.b : a.a + a.b), y = l == 0 ? b.a : (l == 1 ? b.b : b.a + b.b);
  const Fq t = x * y;
  const Fq t0 = bcg(g, t, 0), t1 = bcg(g, t, 1), t2 = bcg(g, t, 2);
  return {t0 - t1, t2 - t0 - t1};
}
__device__ __forceinline__ Fq2 fq2_sqr_coop(const Fq2 &a, const LaneGroup &g) {                  // complex squaring, as Fq2::sqr
  const uint32_t l = g.l;
  const Fq x = l == 0 ? a.a + a.b : a.a, y = l == 0 ? a.a - a.b : a.b;
  const Fq t = x * y;
  const Fq t0 = bcg(g, t, 0), t1 = bcg(g, t, 1);
  return {t0, t1 + t1};
}
static __device__ __noinline__ void coop_dbl(XYZZ<Fq2> *p, const LaneGroup &g) {
  if (p->is_inf()) return;
  const Fq2 U = p->Y.dbl(), V = fq2_sqr_coop(U, g), W = fq2_mul_coop(U, V, g), S = fq2_mul_coop(p->X, V, g);
  const Fq2 x2 = fq2_sqr_coop(p->X, g), M = x2.dbl() + x2;
  const Fq2 X3 = fq2_sqr_coop(M, g) - S.dbl();
  const Fq2 Y3 = fq2_mul_coop(M, S - X3, g) - fq2_mul_coop(W, p->Y, g);
  p->ZZ = fq2_mul_coop(V, p->ZZ, g);
  p->ZZZ = fq2_mul_coop(W, p->ZZZ, g);
  p->X = X3; p->Y = Y3;
}
static __device__ __noinline__ void coop_add(XYZZ<Fq2> *p, const XYZZ<Fq2> *o, const LaneGroup &g) {
  if (o->is_inf()) return;
  if (p->is_inf()) { *p = *o; return; }
  const Fq2 U1 = fq2_mul_coop(p->X, o->ZZ, g), U2 = fq2_mul_coop(o->X, p->ZZ, g);
  const Fq2 S1 = fq2_mul_coop(p->Y, o->ZZZ, g), S2 = fq2_mul_coop(o->Y, p->ZZZ, g);
  const Fq2 P = U2 - U1, R = S2 - S1;
  if (P.is_zero()) {
    if (R.is_zero()) coop_dbl(p, g);
    else *p = XYZZ<Fq2>::infinity();
    return;
  }
  const Fq2 PP = fq2_sqr_coop(P, g), PPP = fq2_mul_coop(P, PP, g), Q = fq2_mul_coop(U1, PP, g);
  const Fq2 X3 = fq2_sqr_coop(R, g) - PPP - Q.dbl();
  const Fq2 Y3 = fq2_mul_coop(R, Q - X3, g) - fq2_mul_coop(S1, PPP, g);
  p->ZZ = fq2_mul_coop(fq2_mul_coop(p->ZZ, o->ZZ, g), PP, g);
  p->ZZZ = fq2_mul_coop(fq2_mul_coop(p->ZZZ, o->ZZZ, g), PPP, g);
  p->X = X3; p->Y = Y3;
}
static __device__ __noinline__ void coop_madd(XYZZ<Fq2> *p, const Affine<Fq2> *q, const LaneGroup &g) {
  if (q->is_inf()) return;
  if (p->is_inf()) { p->X = q->x; p->Y = q->y; p->ZZ = Fq2::one(); p->ZZZ = Fq2::one(); return; }
  const Fq2 U2 = fq2_mul_coop(q->x, p->ZZ, g), S2 = fq2_mul_coop(q->y, p->ZZZ, g);
  const Fq2 P = U2 - p->X, R = S2 - p->Y;
  if (P.is_zero()) {
    if (R.is_zero()) *p = XYZZ<Fq2>::dbl_affine(*q);
    else *p = XYZZ<Fq2>::infinity();
    return;
  }
  const Fq2 PP = fq2_sqr_coop(P, g), PPP = fq2_mul_coop(P, PP, g), Q = fq2_mul_coop(p->X, PP, g);
  const Fq2 X3 = fq2_sqr_coop(R, g) - PPP - Q.dbl();
  const Fq2 Y3 = fq2_mul_coop(R, Q - X3, g) - fq2_mul_coop(p->Y, PPP, g);
  p->ZZ = fq2_mul_coop(p->ZZ, PP, g);
  p->ZZZ = fq2_mul_coop(p->ZZZ, PPP, g);
  p->X = X3; p->Y = Y3;
}

}  // namespace zkb
#endif
