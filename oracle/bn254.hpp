// BN254 (alt_bn128) arithmetic for the CPU oracle.  TEST INFRASTRUCTURE - not product code.
//
// The reference's prover arithmetic lives in un-vendored dependencies (go-rapidsnark prover v0.0.9
// -> rapidsnark C++/asm, `go.mod:30`; snarkjs 0.7.0 -> ffjavascript 0.2.59 / wasmcurves 0.2.1,
// `circuit/package-lock.json:112-114,311-313,352-354`).  This header restates the published
// algorithms those libraries implement (Montgomery Fp with R = 2^256, Fq2 = Fq[u]/(u^2+1),
// short-Weierstrass y^2 = x^3 + 3 and its D-type sextic twist y^2 = x^3 + 3/(9+u), optimal-ate
// pairing, radix-2 NTT over Fr with omega_{2^28} = 5^((r-1)/2^28), Pippenger MSM), 4x64-bit limbs
// with unsigned __int128.  Pinned by the reference's own fixtures: proof.json + signals.json verify
// under verification_key.json, and vk_alphabeta_12 is reproduced (tests/test_oracle_groth16.py).
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>
#include <cstdio>
#include <cstdlib>

namespace bn {

typedef unsigned __int128 u128;

struct FpParams {
  uint64_t p[4];
  uint64_t inv;     // -p^-1 mod 2^64
  uint64_t r1[4];   // R mod p
  uint64_t r2[4];   // R^2 mod p
};

static inline int cmp4(const uint64_t *a, const uint64_t *b) {
  for (int i = 3; i >= 0; i--) {
    if (a[i] < b[i]) return -1;
    if (a[i] > b[i]) return 1;
  }
  return 0;
}
static inline uint64_t add4(uint64_t *r, const uint64_t *a, const uint64_t *b) {
  u128 c = 0;
  for (int i = 0; i < 4; i++) {
    c += (u128)a[i] + b[i];
    r[i] = (uint64_t)c;
    c >>= 64;
  }
  return (uint64_t)c;
}
static inline uint64_t sub4(uint64_t *r, const uint64_t *a, const uint64_t *b) {
  uint64_t br = 0;
  for (int i = 0; i < 4; i++) {
    u128 d = (u128)a[i] - b[i] - br;
    r[i] = (uint64_t)d;
    br = (uint64_t)(d >> 64) & 1;
  }
  return br;
}

static inline void make_params(FpParams &P, const uint64_t mod[4]) {
  memcpy(P.p, mod, 32);
  uint64_t x = 1;  // Newton: x = p^-1 mod 2^64
  for (int i = 0; i < 7; i++) x *= 2 - mod[0] * x;
  P.inv = (uint64_t)(0 - x);
  // R mod p and R^2 mod p by repeated doubling of 1
  uint64_t t[4] = {1, 0, 0, 0};
  for (int i = 0; i < 512; i++) {
    uint64_t c = add4(t, t, t);
    if (c || cmp4(t, mod) >= 0) sub4(t, t, mod);
    if (i == 255) memcpy(P.r1, t, 32);
  }
  memcpy(P.r2, t, 32);
}

template <int TAG>
struct Fp {
  uint64_t v[4];
  static FpParams P;

  static Fp zero() { Fp r; memset(r.v, 0, 32); return r; }
  static Fp one() { Fp r; memcpy(r.v, P.r1, 32); return r; }
  bool is_zero() const { return (v[0] | v[1] | v[2] | v[3]) == 0; }
  bool operator==(const Fp &o) const { return memcmp(v, o.v, 32) == 0; }
  bool operator!=(const Fp &o) const { return !(*this == o); }

  Fp operator+(const Fp &o) const {
    Fp r;
    uint64_t c = add4(r.v, v, o.v);
    if (c || cmp4(r.v, P.p) >= 0) sub4(r.v, r.v, P.p);
    return r;
  }
  Fp operator-(const Fp &o) const {
    Fp r;
    if (sub4(r.v, v, o.v)) add4(r.v, r.v, P.p);
    return r;
  }
  Fp neg() const {
    if (is_zero()) return *this;
    Fp r;
    sub4(r.v, P.p, v);
    return r;
  }
  Fp dbl() const { return *this + *this; }

  // Montgomery product a*b/R mod p (CIOS)
  Fp operator*(const Fp &o) const {
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    const uint64_t *p = P.p;
    for (int i = 0; i < 4; i++) {
      u128 c = 0;
      for (int j = 0; j < 4; j++) {
        c += (u128)v[j] * o.v[i] + t[j];
        t[j] = (uint64_t)c;
        c >>= 64;
      }
      c += t[4];
      t[4] = (uint64_t)c;
      t[5] = (uint64_t)(c >> 64);
      uint64_t m = t[0] * P.inv;
      c = (u128)m * p[0] + t[0];
      c >>= 64;
      for (int j = 1; j < 4; j++) {
        c += (u128)m * p[j] + t[j];
        t[j - 1] = (uint64_t)c;
        c >>= 64;
      }
      c += t[4];
      t[3] = (uint64_t)c;
      t[4] = t[5] + (uint64_t)(c >> 64);
    }
    Fp r;
    memcpy(r.v, t, 32);
    if (t[4] || cmp4(r.v, p) >= 0) sub4(r.v, r.v, p);
    return r;
  }
  Fp sqr() const { return *this * *this; }

  // this^e, e = n little-endian 64-bit limbs
  Fp pow(const uint64_t *e, int n) const {
    Fp r = one();
    bool started = false;
    for (int i = n * 64 - 1; i >= 0; i--) {
      if (started) r = r.sqr();
      if ((e[i / 64] >> (i % 64)) & 1) {
        r = started ? r * *this : *this;
        started = true;
      }
    }
    return r;
  }
  Fp inv() const {
    uint64_t e[4], two[4] = {2, 0, 0, 0};
    sub4(e, P.p, two);
    return pow(e, 4);
  }

  // canonical little-endian bytes <-> Montgomery
  static Fp from_raw(const uint64_t raw[4]) {  // value (normal form) -> Montgomery
    Fp a, b;
    memcpy(a.v, raw, 32);
    memcpy(b.v, P.r2, 32);
    return a * b;
  }
  static Fp from_u64(uint64_t x) {
    uint64_t raw[4] = {x, 0, 0, 0};
    return from_raw(raw);
  }
  static Fp from_bytes(const uint8_t *b) {
    uint64_t raw[4];
    memcpy(raw, b, 32);
    return from_raw(raw);
  }
  static Fp from_mont_bytes(const uint8_t *b) {  // bytes already hold the Montgomery form
    Fp r;
    memcpy(r.v, b, 32);
    return r;
  }
  void to_raw(uint64_t raw[4]) const {
    Fp o;
    o.v[0] = 1; o.v[1] = o.v[2] = o.v[3] = 0;
    Fp r = *this * o;
    memcpy(raw, r.v, 32);
  }
  void to_bytes(uint8_t *b) const {
    uint64_t raw[4];
    to_raw(raw);
    memcpy(b, raw, 32);
  }
  void to_mont_bytes(uint8_t *b) const { memcpy(b, v, 32); }
};
template <int TAG> FpParams Fp<TAG>::P;

typedef Fp<0> Fq;
typedef Fp<1> Fr;

static const uint64_t Q_MOD[4] = {0x3c208c16d87cfd47ull, 0x97816a916871ca8dull, 0xb85045b68181585dull,
                                  0x30644e72e131a029ull};
static const uint64_t R_MOD[4] = {0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull,
                                  0x30644e72e131a029ull};

struct Fq2 {
  Fq a, b;  // a + b*u, u^2 = -1
  static Fq2 zero() { return {Fq::zero(), Fq::zero()}; }
  static Fq2 one() { return {Fq::one(), Fq::zero()}; }
  bool is_zero() const { return a.is_zero() && b.is_zero(); }
  bool operator==(const Fq2 &o) const { return a == o.a && b == o.b; }
  bool operator!=(const Fq2 &o) const { return !(*this == o); }
  Fq2 operator+(const Fq2 &o) const { return {a + o.a, b + o.b}; }
  Fq2 operator-(const Fq2 &o) const { return {a - o.a, b - o.b}; }
  Fq2 neg() const { return {a.neg(), b.neg()}; }
  Fq2 dbl() const { return {a.dbl(), b.dbl()}; }
  Fq2 conj() const { return {a, b.neg()}; }
  Fq2 operator*(const Fq2 &o) const {
    Fq t0 = a * o.a, t1 = b * o.b;
    Fq t2 = (a + b) * (o.a + o.b);
    return {t0 - t1, t2 - t0 - t1};
  }
  Fq2 sqr() const {
    Fq t = a * b;
    return {(a + b) * (a - b), t + t};
  }
  Fq2 mul_fq(const Fq &k) const { return {a * k, b * k}; }
  Fq2 inv() const {
    Fq n = (a.sqr() + b.sqr()).inv();
    return {a * n, (b * n).neg()};
  }
  Fq2 mul_xi() const {  // * (9 + u)
    Fq a2 = a.dbl(), a4 = a2.dbl(), a8 = a4.dbl();
    Fq b2 = b.dbl(), b4 = b2.dbl(), b8 = b4.dbl();
    return {a8 + a - b, b8 + b + a};
  }
  Fq2 pow(const uint64_t *e, int n) const {
    Fq2 r = one();
    for (int i = n * 64 - 1; i >= 0; i--) {
      r = r.sqr();
      if ((e[i / 64] >> (i % 64)) & 1) r = r * *this;
    }
    return r;
  }
};

// ---------------------------------------------------------------------------------------------
// Short Weierstrass y^2 = x^3 + b, Jacobian coordinates, generic over the coordinate field.
// ---------------------------------------------------------------------------------------------
template <class F> struct Aff { F x, y; bool inf; };
template <class F> struct Jac {
  F X, Y, Z;
  static Jac infinity() { return {F::one(), F::one(), F::zero()}; }
  bool is_inf() const { return Z.is_zero(); }
  static Jac from_aff(const Aff<F> &a) { return a.inf ? infinity() : Jac{a.x, a.y, F::one()}; }

  Jac dbl() const {
    if (is_inf()) return *this;
    F A = X.sqr(), B = Y.sqr(), C = B.sqr();
    F D = ((X + B).sqr() - A - C).dbl();
    F E = A.dbl() + A, Fv = E.sqr();
    F X3 = Fv - D.dbl();
    F Y3 = E * (D - X3) - C.dbl().dbl().dbl();
    F Z3 = (Y * Z).dbl();
    return {X3, Y3, Z3};
  }
  Jac add_aff(const Aff<F> &q) const {
    if (q.inf) return *this;
    if (is_inf()) return from_aff(q);
    F Z1Z1 = Z.sqr();
    F U2 = q.x * Z1Z1, S2 = q.y * Z * Z1Z1;
    if (U2 == X) {
      if (S2 == Y) return dbl();
      return infinity();
    }
    F H = U2 - X, HH = H.sqr(), I = HH.dbl().dbl(), J = H * I;
    F r = (S2 - Y).dbl(), V = X * I;
    F X3 = r.sqr() - J - V.dbl();
    F Y3 = r * (V - X3) - (Y * J).dbl();
    F Z3 = (Z + H).sqr() - Z1Z1 - HH;
    return {X3, Y3, Z3};
  }
  Jac add(const Jac &q) const {
    if (q.is_inf()) return *this;
    if (is_inf()) return q;
    F Z1Z1 = Z.sqr(), Z2Z2 = q.Z.sqr();
    F U1 = X * Z2Z2, U2 = q.X * Z1Z1;
    F S1 = Y * q.Z * Z2Z2, S2 = q.Y * Z * Z1Z1;
    if (U1 == U2) {
      if (S1 == S2) return dbl();
      return infinity();
    }
    F H = U2 - U1, I = H.dbl().sqr(), J = H * I;
    F r = (S2 - S1).dbl(), V = U1 * I;
    F X3 = r.sqr() - J - V.dbl();
    F Y3 = r * (V - X3) - (S1 * J).dbl();
    F Z3 = ((Z + q.Z).sqr() - Z1Z1 - Z2Z2) * H;
    return {X3, Y3, Z3};
  }
  Jac neg() const { return {X, Y.neg(), Z}; }
  Aff<F> to_aff() const {
    if (is_inf()) return {F::zero(), F::zero(), true};
    F zi = Z.inv(), zi2 = zi.sqr();
    return {X * zi2, Y * zi2 * zi, false};
  }
  // scalar given as canonical 256-bit little-endian limbs
  Jac mul(const uint64_t k[4]) const {
    Jac r = infinity();
    for (int i = 255; i >= 0; i--) {
      r = r.dbl();
      if ((k[i / 64] >> (i % 64)) & 1) r = r.add(*this);
    }
    return r;
  }
};

typedef Aff<Fq> G1A;
typedef Jac<Fq> G1J;
typedef Aff<Fq2> G2A;
typedef Jac<Fq2> G2J;

template <class F>
static void batch_to_affine(const std::vector<Jac<F>> &in, std::vector<Aff<F>> &out) {
  size_t n = in.size();
  out.resize(n);
  std::vector<F> pre(n);
  F acc = F::one();
  for (size_t i = 0; i < n; i++) {
    pre[i] = acc;
    if (!in[i].is_inf()) acc = acc * in[i].Z;
  }
  F inv = acc.inv();
  for (size_t i = n; i-- > 0;) {
    if (in[i].is_inf()) {
      out[i] = {F::zero(), F::zero(), true};
      continue;
    }
    F zi = inv * pre[i];
    inv = inv * in[i].Z;
    F zi2 = zi.sqr();
    out[i] = {in[i].X * zi2, in[i].Y * zi2 * zi, false};
  }
}

// Pippenger MSM: sum scalars[i] * bases[i]; scalars are canonical 32-byte little-endian values.
template <class F>
static Jac<F> msm(const Aff<F> *bases, const uint8_t *scalars, size_t n, int c = 0) {
  if (n == 0) return Jac<F>::infinity();
  if (c == 0) {
    c = 4;
    while ((1ull << (c + 4)) < n && c < 16) c++;
  }
  int nwin = (254 + c - 1) / c;
  std::vector<Jac<F>> wsum(nwin);
#pragma omp parallel for schedule(dynamic, 1)
  for (int w = 0; w < nwin; w++) {
    size_t nb = (size_t)1 << c;
    std::vector<Jac<F>> bucket(nb, Jac<F>::infinity());
    int bit = w * c;
    for (size_t i = 0; i < n; i++) {
      if (bases[i].inf) continue;
      const uint8_t *s = scalars + 32 * i;
      uint64_t limb[5] = {0, 0, 0, 0, 0};
      memcpy(limb, s, 32);
      int li = bit / 64, sh = bit % 64;
      uint64_t d = limb[li] >> sh;
      if (sh + c > 64) d |= limb[li + 1] << (64 - sh);
      d &= nb - 1;
      if (d) bucket[d] = bucket[d].add_aff(bases[i]);
    }
    Jac<F> run = Jac<F>::infinity(), acc = Jac<F>::infinity();
    for (size_t k = nb - 1; k >= 1; k--) {
      run = run.add(bucket[k]);
      acc = acc.add(run);
    }
    wsum[w] = acc;
  }
  Jac<F> r = Jac<F>::infinity();
  for (int w = nwin - 1; w >= 0; w--) {
    for (int k = 0; k < c; k++) r = r.dbl();
    r = r.add(wsum[w]);
  }
  return r;
}

// ---------------------------------------------------------------------------------------------
// Fr NTT (radix-2, in place), omega_{2^28} = 5^((r-1)/2^28) as in snarkjs / ptau
// ---------------------------------------------------------------------------------------------
struct NttDomain {
  int logn;
  size_t n;
  std::vector<Fr> tw;     // tw[k] = w^k, k < n/2
  std::vector<Fr> itw;    // inverse twiddles
  Fr ninv;
  Fr w, winv;
  static Fr root_of_unity(int k) {  // primitive 2^k-th root
    uint64_t e[4], one[4] = {1, 0, 0, 0};
    sub4(e, R_MOD, one);
    // (r-1) >> 28
    for (int i = 0; i < 4; i++) e[i] = (e[i] >> 28) | (i < 3 ? e[i + 1] << 36 : 0);
    Fr g = Fr::from_u64(5).pow(e, 4);
    for (int i = 28; i > k; i--) g = g.sqr();
    return g;
  }
  explicit NttDomain(int logn_) : logn(logn_), n((size_t)1 << logn_) {
    w = root_of_unity(logn);
    winv = w.inv();
    tw.resize(n / 2);
    itw.resize(n / 2);
    Fr a = Fr::one(), b = Fr::one();
    for (size_t i = 0; i < n / 2; i++) {
      tw[i] = a;
      itw[i] = b;
      a = a * w;
      b = b * winv;
    }
    ninv = Fr::from_u64(n).inv();
  }
  void run(Fr *x, bool inverse) const {
    // bit reversal
    for (size_t i = 0; i < n; i++) {
      size_t j = 0;
      for (int b = 0; b < logn; b++) j |= ((i >> b) & 1) << (logn - 1 - b);
      if (i < j) std::swap(x[i], x[j]);
    }
    const std::vector<Fr> &T = inverse ? itw : tw;
    for (size_t len = 2; len <= n; len <<= 1) {
      size_t half = len / 2, step = n / len;
#pragma omp parallel for schedule(static)
      for (size_t s = 0; s < n; s += len) {
        for (size_t k = 0; k < half; k++) {
          Fr u = x[s + k], v = x[s + k + half] * T[k * step];
          x[s + k] = u + v;
          x[s + k + half] = u - v;
        }
      }
    }
    if (inverse) {
#pragma omp parallel for schedule(static)
      for (size_t i = 0; i < n; i++) x[i] = x[i] * ninv;
    }
  }
};

// ---------------------------------------------------------------------------------------------
// Fq6 = Fq2[v]/(v^3 - xi), Fq12 = Fq6[w]/(w^2 - v), pairing
// ---------------------------------------------------------------------------------------------
struct Fq6 {
  Fq2 c0, c1, c2;
  static Fq6 zero() { return {Fq2::zero(), Fq2::zero(), Fq2::zero()}; }
  static Fq6 one() { return {Fq2::one(), Fq2::zero(), Fq2::zero()}; }
  bool operator==(const Fq6 &o) const { return c0 == o.c0 && c1 == o.c1 && c2 == o.c2; }
  Fq6 operator+(const Fq6 &o) const { return {c0 + o.c0, c1 + o.c1, c2 + o.c2}; }
  Fq6 operator-(const Fq6 &o) const { return {c0 - o.c0, c1 - o.c1, c2 - o.c2}; }
  Fq6 neg() const { return {c0.neg(), c1.neg(), c2.neg()}; }
  Fq6 operator*(const Fq6 &o) const {
    Fq2 t0 = c0 * o.c0, t1 = c1 * o.c1, t2 = c2 * o.c2;
    Fq2 r0 = ((c1 + c2) * (o.c1 + o.c2) - t1 - t2).mul_xi() + t0;
    Fq2 r1 = (c0 + c1) * (o.c0 + o.c1) - t0 - t1 + t2.mul_xi();
    Fq2 r2 = (c0 + c2) * (o.c0 + o.c2) - t0 - t2 + t1;
    return {r0, r1, r2};
  }
  Fq6 mul_v() const { return {c2.mul_xi(), c0, c1}; }
  Fq6 inv() const {
    Fq2 t0 = c0.sqr() - (c1 * c2).mul_xi();
    Fq2 t1 = c2.sqr().mul_xi() - c0 * c1;
    Fq2 t2 = c1.sqr() - c0 * c2;
    Fq2 d = (c0 * t0 + (c2 * t1).mul_xi() + (c1 * t2).mul_xi()).inv();
    return {t0 * d, t1 * d, t2 * d};
  }
};

struct Fq12 {
  Fq6 c0, c1;
  static Fq12 one() { return {Fq6::one(), Fq6::zero()}; }
  bool operator==(const Fq12 &o) const { return c0 == o.c0 && c1 == o.c1; }
  Fq12 operator*(const Fq12 &o) const {
    Fq6 t0 = c0 * o.c0, t1 = c1 * o.c1;
    Fq6 r1 = (c0 + c1) * (o.c0 + o.c1) - t0 - t1;
    return {t0 + t1.mul_v(), r1};
  }
  Fq12 sqr() const { return *this * *this; }
  Fq12 conj() const { return {c0, c1.neg()}; }
  Fq12 inv() const {
    Fq6 d = (c0 * c0 - (c1 * c1).mul_v()).inv();
    return {c0 * d, (c1 * d).neg()};
  }
  // coefficient of w^i, i = 0..5  (w^2 = v): even i -> c0.c{i/2}, odd i -> c1.c{i/2}
  Fq2 &coef(int i) {
    Fq6 &h = (i & 1) ? c1 : c0;
    int j = i >> 1;
    return j == 0 ? h.c0 : (j == 1 ? h.c1 : h.c2);
  }
  Fq12 pow(const uint64_t *e, int n) const {
    Fq12 r = one();
    for (int i = n * 64 - 1; i >= 0; i--) {
      r = r.sqr();
      if ((e[i / 64] >> (i % 64)) & 1) r = r * *this;
    }
    return r;
  }
};

struct Pairing {
  Fq2 gamma[6];  // gamma[i] = xi^(i*(q-1)/6)
  Fq2 twist_b;   // 3/xi
  // (q^4 - q^2 + 1)/r, little-endian limbs (computed offline with Python big ints)
  static constexpr int HARD_LIMBS = 12;
  uint64_t hard[HARD_LIMBS];

  Pairing() {
    Fq2 xi = {Fq::from_u64(9), Fq::one()};
    // (q-1)/6
    uint64_t e[4], one[4] = {1, 0, 0, 0};
    sub4(e, Q_MOD, one);
    // divide by 6
    u128 rem = 0;
    for (int i = 3; i >= 0; i--) {
      u128 cur = (rem << 64) | e[i];
      e[i] = (uint64_t)(cur / 6);
      rem = cur % 6;
    }
    Fq2 g = xi.pow(e, 4);
    gamma[0] = Fq2::one();
    for (int i = 1; i < 6; i++) gamma[i] = gamma[i - 1] * g;
    twist_b = Fq2{Fq::from_u64(3), Fq::zero()} * xi.inv();
    static const uint64_t H[HARD_LIMBS] = {
        0xe81bb482ccdf42b1ull, 0x5abf5cc4f49c36d4ull, 0xf1154e7e1da014fdull, 0xdcc7b44c87cdbacfull,
        0xaaa441e3954bcf8aull, 0x6b887d56d5095f23ull, 0x79581e16f3fd90c6ull, 0x3b1b1355d189227dull,
        0x4e529a5861876f6bull, 0x6c0eb522d5b12278ull, 0x331ec15183177fafull, 0x01baaa710b0759adull};
    memcpy(hard, H, sizeof H);
  }

  Fq12 frobenius(Fq12 f) const {
    for (int i = 0; i < 6; i++) f.coef(i) = f.coef(i).conj() * gamma[i];
    return f;
  }
  G2A frob_g2(const G2A &q) const {
    return {q.x.conj() * gamma[2], q.y.conj() * gamma[3], q.inf};
  }
  // line through T (slope lam on the twist) evaluated at P: yP - lam*xP*w + (lam*xT - yT)*w^3
  static Fq12 line(const Fq2 &lam, const G2A &T, const G1A &P) {
    Fq12 l = {Fq6::zero(), Fq6::zero()};
    l.coef(0) = Fq2{P.y, Fq::zero()};
    l.coef(1) = lam.mul_fq(P.x).neg();
    l.coef(3) = lam * T.x - T.y;
    return l;
  }
  static void dbl_step(G2A &T, Fq12 &f, const G1A &P) {
    Fq2 x2 = T.x.sqr();
    Fq2 lam = (x2.dbl() + x2) * T.y.dbl().inv();
    f = f * line(lam, T, P);
    Fq2 x3 = lam.sqr() - T.x.dbl();
    Fq2 y3 = lam * (T.x - x3) - T.y;
    T.x = x3;
    T.y = y3;
  }
  static void add_step(G2A &T, const G2A &Q, Fq12 &f, const G1A &P) {
    Fq2 lam = (Q.y - T.y) * (Q.x - T.x).inv();
    f = f * line(lam, T, P);
    Fq2 x3 = lam.sqr() - T.x - Q.x;
    Fq2 y3 = lam * (T.x - x3) - T.y;
    T.x = x3;
    T.y = y3;
  }
  Fq12 miller(const G1A &P, const G2A &Q) const {
    if (P.inf || Q.inf) return Fq12::one();
    // 6x+2 = 29793968203157093288 = 0x19d797039be763ba8 (65 bits)
    const uint64_t lo = 0x9d797039be763ba8ull;  // low 64 bits; bit 64 is set
    Fq12 f = Fq12::one();
    G2A T = Q;
    for (int i = 63; i >= 0; i--) {
      f = f.sqr();
      dbl_step(T, f, P);
      if ((lo >> i) & 1) add_step(T, Q, f, P);
    }
    G2A Q1 = frob_g2(Q);
    G2A Q2 = frob_g2(Q1);
    Q2.y = Q2.y.neg();
    add_step(T, Q1, f, P);
    add_step(T, Q2, f, P);
    return f;
  }
  Fq12 final_exp(const Fq12 &f) const {
    Fq12 a = f.conj() * f.inv();             // f^(q^6-1)
    a = frobenius(frobenius(a)) * a;         // ^(q^2+1)
    return a.pow(hard, HARD_LIMBS);          // ^((q^4-q^2+1)/r)
  }
};

static inline void init() {
  static bool done = false;
  if (done) return;
  make_params(Fq::P, Q_MOD);
  make_params(Fr::P, R_MOD);
  done = true;
}

static inline G1A g1_gen() { return {Fq::from_u64(1), Fq::from_u64(2), false}; }
static inline G2A g2_gen() {
  static const uint64_t x0[4] = {0x46debd5cd992f6edull, 0x674322d4f75edaddull, 0x426a00665e5c4479ull, 0x1800deef121f1e76ull};
  static const uint64_t x1[4] = {0x97e485b7aef312c2ull, 0xf1aa493335a9e712ull, 0x7260bfb731fb5d25ull, 0x198e9393920d483aull};
  static const uint64_t y0[4] = {0x4ce6cc0166fa7daaull, 0xe3d1e7690c43d37bull, 0x4aab71808dcb408full, 0x12c85ea5db8c6debull};
  static const uint64_t y1[4] = {0x55acdadcd122975bull, 0xbc4b313370b38ef3ull, 0xec9e99ad690c3395ull, 0x090689d0585ff075ull};
  return {Fq2{Fq::from_raw(x0), Fq::from_raw(x1)}, Fq2{Fq::from_raw(y0), Fq::from_raw(y1)}, false};
}

template <class F> static bool on_curve(const Aff<F> &p, const F &b) {
  if (p.inf) return true;
  return p.y.sqr() == p.x.sqr() * p.x + b;
}

}  // namespace bn
