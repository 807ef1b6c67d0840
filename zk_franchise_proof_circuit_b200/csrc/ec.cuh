// BN254 G1 / G2 group arithmetic for the MSM kernels: Fq2 tower and XYZZ ("extended Jacobian")
// coordinates, generic over the coordinate field (Fq for G1, Fq2 for G2).
//
// XYZZ (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2) needs 8M+2S for a mixed add and has no inversion; it is
// the accumulator form of every bucket.  Bases stay affine (Montgomery form, exactly the bytes of
// zkey sections 5-9).  Both curves have a = 0.
#pragma once
#include "fp.cuh"

namespace zkb {

struct Fq2;
#if defined(__CUDACC__)
// Out-of-line Fq2 product / square whose 3 (2) Fq products are inlined, so they interleave on the IMAD pipe
// (instruction-level parallelism 3) while the call keeps the G2 point formulas small.
template <int TAG = 0> __device__ __noinline__ Fq2 fq2_mul_call(Fq2 a, Fq2 b);
template <int TAG = 0> __device__ __noinline__ Fq2 fq2_sqr_call(Fq2 a);
#endif

struct alignas(16) Fq2 {
  Fq a, b;  // a + b*u, u^2 = -1
  ZKB_HD static Fq2 zero() { return {Fq::zero(), Fq::zero()}; }
  ZKB_HD static Fq2 one() { return {Fq::one(), Fq::zero()}; }
  ZKB_HD bool is_zero() const { return a.is_zero() && b.is_zero(); }
  ZKB_HD bool operator==(const Fq2 &o) const { return a == o.a && b == o.b; }
  ZKB_HD bool operator!=(const Fq2 &o) const { return !(*this == o); }
  ZKB_HD Fq2 operator+(const Fq2 &o) const { return {a + o.a, b + o.b}; }
  ZKB_HD Fq2 operator-(const Fq2 &o) const { return {a - o.a, b - o.b}; }
  ZKB_HD Fq2 neg() const { return {a.neg(), b.neg()}; }
  ZKB_HD static Fq2 select(bool c, const Fq2 &x, const Fq2 &y) { return {Fq::select(c, x.a, y.a), Fq::select(c, x.b, y.b)}; }
  ZKB_HD Fq2 dbl() const { return {a.dbl(), b.dbl()}; }
  ZKB_HD Fq2 operator*(const Fq2 &o) const {  // Karatsuba: 3 Fq products (inlined)
    Fq t0 = a * o.a, t1 = b * o.b;
    Fq t2 = (a + b) * (o.a + o.b);
    return {t0 - t1, t2 - t0 - t1};
  }
  ZKB_HD Fq2 sqr() const {  // complex squaring: 2 Fq products (inlined)
    Fq t = a * b;
    return {(a + b) * (a - b), t + t};
  }
  ZKB_HD Fq2 mulc(const Fq2 &o) const {
#if defined(__CUDA_ARCH__)
    return fq2_mul_call<0>(*this, o);
#else
    return *this * o;
#endif
  }
  ZKB_HD Fq2 sqrc() const {
#if defined(__CUDA_ARCH__)
    return fq2_sqr_call<0>(*this);
#else
    return sqr();
#endif
  }
  ZKB_HD Fq2 inv() const {
    Fq n = (a.sqrc() + b.sqrc()).inv();
    return {a.mulc(n), b.mulc(n).neg()};
  }
};

#if defined(__CUDACC__)
template <int TAG> __device__ __noinline__ Fq2 fq2_mul_call(Fq2 a, Fq2 b) { return a * b; }
template <int TAG> __device__ __noinline__ Fq2 fq2_sqr_call(Fq2 a) { return a.sqr(); }
#endif

template <class F>
struct alignas(16) Affine {
  F x, y;  // (0,0) encodes the point at infinity (not on either curve since b != 0)
  ZKB_HD bool is_inf() const { return x.is_zero() && y.is_zero(); }
  ZKB_HD Affine neg() const { return {x, y.neg()}; }
};

template <class F>
struct alignas(16) XYZZ {
  F X, Y, ZZ, ZZZ;

  ZKB_HD static XYZZ infinity() { return {F::zero(), F::zero(), F::zero(), F::zero()}; }
  ZKB_HD bool is_inf() const { return ZZ.is_zero(); }
  ZKB_HD static XYZZ from_affine(const Affine<F> &p) {
    if (p.is_inf()) return infinity();
    return {p.x, p.y, F::one(), F::one()};
  }
  ZKB_HD XYZZ neg() const { return {X, Y.neg(), ZZ, ZZZ}; }

  // 2*P for affine P (mdbl-2008-s-1)
  ZKB_HD static XYZZ dbl_affine(const Affine<F> &p) {
    if (p.is_inf()) return infinity();
    F U = p.y.dbl(), V = U.sqrc(), W = U.mulc(V), S = p.x.mulc(V);
    F x2 = p.x.sqrc(), M = x2.dbl() + x2;
    F X3 = M.sqrc() - S.dbl();
    F Y3 = M.mulc(S - X3) - W.mulc(p.y);
    return {X3, Y3, V, W};
  }
  // dbl-2008-s-1
  ZKB_HD XYZZ dbl() const {
    if (is_inf()) return *this;
    F U = Y.dbl(), V = U.sqrc(), W = U.mulc(V), S = X.mulc(V);
    F x2 = X.sqrc(), M = x2.dbl() + x2;
    F X3 = M.sqrc() - S.dbl();
    F Y3 = M.mulc(S - X3) - W.mulc(Y);
    return {X3, Y3, V.mulc(ZZ), W.mulc(ZZZ)};
  }
  // this += affine p (madd-2008-s), all special cases handled
  ZKB_HD void add_affine(const Affine<F> &p) {
    if (p.is_inf()) return;
    if (is_inf()) {
      X = p.x; Y = p.y; ZZ = F::one(); ZZZ = F::one();
      return;
    }
    F U2 = p.x.mulc(ZZ), S2 = p.y.mulc(ZZZ);
    F P = U2 - X, R = S2 - Y;
    if (P.is_zero()) {
      if (R.is_zero()) *this = dbl_affine(p);
      else *this = infinity();
      return;
    }
    F PP = P.sqrc(), PPP = P.mulc(PP), Q = X.mulc(PP);
    F X3 = R.sqrc() - PPP - Q.dbl();
    Y = R.mulc(Q - X3) - Y.mulc(PPP);
    X = X3;
    ZZ = ZZ.mulc(PP);
    ZZZ = ZZZ.mulc(PPP);
  }
  // this += o (add-2008-s)
  ZKB_HD void add(const XYZZ &o) {
    if (o.is_inf()) return;
    if (is_inf()) { *this = o; return; }
    F U1 = X.mulc(o.ZZ), U2 = o.X.mulc(ZZ), S1 = Y.mulc(o.ZZZ), S2 = o.Y.mulc(ZZZ);
    F P = U2 - U1, R = S2 - S1;
    if (P.is_zero()) {
      if (R.is_zero()) *this = dbl();
      else *this = infinity();
      return;
    }
    F PP = P.sqrc(), PPP = P.mulc(PP), Q = U1.mulc(PP);
    F X3 = R.sqrc() - PPP - Q.dbl();
    Y = R.mulc(Q - X3) - S1.mulc(PPP);
    X = X3;
    ZZ = ZZ.mulc(o.ZZ).mulc(PP);
    ZZZ = ZZZ.mulc(o.ZZZ).mulc(PPP);
  }
  ZKB_HD Affine<F> to_affine() const {
    if (is_inf()) return {F::zero(), F::zero()};
    F i3 = ZZZ.inv();        // 1/z^3
    F i1 = i3.mulc(ZZ);          // 1/z
    F i2 = i1.sqrc();         // 1/z^2
    return {X.mulc(i2), Y.mulc(i3)};
  }
  // k * this, k = plain 256-bit integer (8 limbs LE); double-and-add from the top bit
  ZKB_HD XYZZ mul(const uint32_t k[8]) const {
    XYZZ r = infinity();
    for (int i = 255; i >= 0; i--) {
      r = r.dbl();
      if ((k[i >> 5] >> (i & 31)) & 1) r.add(*this);
    }
    return r;
  }
};

// Out-of-line variants for the cold kernels (bucket reduction, table build, proof assembly): one
// compiled body per field instead of one per call site keeps code size and build time sane.  The
// hot bucket-accumulation loop uses the inlined members above.
#if defined(__CUDACC__)
#define ZKB_NI __host__ __device__ __noinline__
#else
#define ZKB_NI inline
#endif
template <class F> ZKB_NI void xyzz_add_ni(XYZZ<F> *a, const XYZZ<F> *b) { a->add(*b); }
template <class F> ZKB_NI void xyzz_add_affine_ni(XYZZ<F> *a, const Affine<F> *b) { a->add_affine(*b); }
template <class F> ZKB_NI void xyzz_dbl_ni(XYZZ<F> *a) { *a = a->dbl(); }
template <class F> ZKB_NI void xyzz_to_affine_ni(const XYZZ<F> *a, Affine<F> *out) { *out = a->to_affine(); }
template <class F> ZKB_NI void xyzz_mul_ni(const XYZZ<F> *a, const uint32_t *k, XYZZ<F> *out) {
  XYZZ<F> r = XYZZ<F>::infinity();
  for (int i = 255; i >= 0; i--) {
    xyzz_dbl_ni(&r);
    if ((k[i >> 5] >> (i & 31)) & 1) xyzz_add_ni(&r, a);
  }
  *out = r;
}

typedef Affine<Fq> G1Affine;
typedef Affine<Fq2> G2Affine;
typedef XYZZ<Fq> G1XYZZ;
typedef XYZZ<Fq2> G2XYZZ;

}  // namespace zkb
