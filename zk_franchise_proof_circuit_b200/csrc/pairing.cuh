// BN254 optimal-ate pairing and the Groth16 verification equation, host/device code (one proof per thread).
//
// The step after the proving path in the reference: `proof.Verify(vkey)` (`zk_census_test.go:122`,
// go-rapidsnark/verifier v0.0.3, `go.mod:32`) / `snarkjs groth16 verify` (`circuit/circuit-compiler.sh:128`):
//     e(-A, B) * e(alpha1, beta2) * e(vk_x, gamma2) * e(C, delta2) == 1,   vk_x = IC_0 + sum_i pub_i * IC_i
// (batch verification on the GPU).  Tower: Fq2 = Fq[u]/(u^2+1), Fq6 = Fq2[v]/(v^3 - xi),
// xi = 9 + u, Fq12 = Fq6[w]/(w^2 - v).  G2 points stay on the twist y^2 = x^3 + 3/xi; the line through
// psi(T), psi(Q) evaluated at P is  yP - (lambda xP) w + (lambda xT - yT) w^3  with lambda the slope on the
// twist, so the Miller loop needs only Fq2 arithmetic for the point steps and one sparse Fq12 element per line.
// Not a hot path: affine steps (one Fq2 inversion each) keep it short and obviously correct.
#pragma once
#include "ec.cuh"

#if defined(__CUDACC__)
#define ZKB_PNI __host__ __device__ __noinline__
#else
#define ZKB_PNI inline
#endif

namespace zkb {

ZKB_HD Fq2 fq2_mul_xi(const Fq2 &x) {  // * (9 + u)
  Fq a2 = x.a.dbl(), a4 = a2.dbl(), a8 = a4.dbl();
  Fq b2 = x.b.dbl(), b4 = b2.dbl(), b8 = b4.dbl();
  return {a8 + x.a - x.b, b8 + x.b + x.a};
}
ZKB_HD Fq2 fq2_conj(const Fq2 &x) { return {x.a, x.b.neg()}; }
ZKB_HD Fq2 fq2_mul_fq(const Fq2 &x, const Fq &k) { return {x.a.mulc(k), x.b.mulc(k)}; }

struct Fq6 {
  Fq2 c0, c1, c2;
  ZKB_HD static Fq6 zero() { return {Fq2::zero(), Fq2::zero(), Fq2::zero()}; }
  ZKB_HD static Fq6 one() { return {Fq2::one(), Fq2::zero(), Fq2::zero()}; }
  ZKB_HD bool operator==(const Fq6 &o) const { return c0 == o.c0 && c1 == o.c1 && c2 == o.c2; }
  ZKB_HD Fq6 operator+(const Fq6 &o) const { return {c0 + o.c0, c1 + o.c1, c2 + o.c2}; }
  ZKB_HD Fq6 operator-(const Fq6 &o) const { return {c0 - o.c0, c1 - o.c1, c2 - o.c2}; }
  ZKB_HD Fq6 neg() const { return {c0.neg(), c1.neg(), c2.neg()}; }
  ZKB_PNI Fq6 mul(const Fq6 &o) const {
    Fq2 t0 = c0.mulc(o.c0), t1 = c1.mulc(o.c1), t2 = c2.mulc(o.c2);
    Fq2 r0 = fq2_mul_xi((c1 + c2).mulc(o.c1 + o.c2) - t1 - t2) + t0;
    Fq2 r1 = (c0 + c1).mulc(o.c0 + o.c1) - t0 - t1 + fq2_mul_xi(t2);
    Fq2 r2 = (c0 + c2).mulc(o.c0 + o.c2) - t0 - t2 + t1;
    return {r0, r1, r2};
  }
  ZKB_HD Fq6 mul_v() const { return {fq2_mul_xi(c2), c0, c1}; }
  ZKB_PNI Fq6 inv() const {
    Fq2 t0 = c0.sqrc() - fq2_mul_xi(c1.mulc(c2));
    Fq2 t1 = fq2_mul_xi(c2.sqrc()) - c0.mulc(c1);
    Fq2 t2 = c1.sqrc() - c0.mulc(c2);
    Fq2 d = (c0.mulc(t0) + fq2_mul_xi(c2.mulc(t1)) + fq2_mul_xi(c1.mulc(t2))).inv();
    return {t0.mulc(d), t1.mulc(d), t2.mulc(d)};
  }
};

struct Fq12 {
  Fq6 c0, c1;
  ZKB_HD static Fq12 one() { return {Fq6::one(), Fq6::zero()}; }
  ZKB_HD bool operator==(const Fq12 &o) const { return c0 == o.c0 && c1 == o.c1; }
  ZKB_PNI Fq12 mul(const Fq12 &o) const {
    Fq6 t0 = c0.mul(o.c0), t1 = c1.mul(o.c1);
    Fq6 r1 = (c0 + c1).mul(o.c0 + o.c1) - t0 - t1;
    return {t0 + t1.mul_v(), r1};
  }
  ZKB_HD Fq12 conj() const { return {c0, c1.neg()}; }
  ZKB_PNI Fq12 inv() const {
    Fq6 d = (c0.mul(c0) - c1.mul(c1).mul_v()).inv();
    return {c0.mul(d), c1.mul(d).neg()};
  }
  // coefficient of w^i, i = 0..5 (w^2 = v): even i -> c0.c{i/2}, odd i -> c1.c{i/2}
  ZKB_HD Fq2 &coef(int i) {
    Fq6 &h = (i & 1) ? c1 : c0;
    int j = i >> 1;
    return j == 0 ? h.c0 : (j == 1 ? h.c1 : h.c2);
  }
};

// constants of the pairing, computed once (host or device) by pairing_consts_init
struct PairingConsts {
  Fq2 gamma[6];   // xi^(i (q-1)/6)
  Fq2 twist_b;    // 3 / xi
};

ZKB_PNI Fq2 fq2_pow(const Fq2 &x, const uint32_t *e, int nlimbs) {
  Fq2 r = Fq2::one();
  for (int i = nlimbs * 32 - 1; i >= 0; i--) {
    r = r.sqrc();
    if ((e[i >> 5] >> (i & 31)) & 1) r = r.mulc(x);
  }
  return r;
}

ZKB_HD void pairing_consts_init(PairingConsts &pc) {
  Fq nine = Fq::zero();
  nine.v[0] = 9;
  Fq2 xi = {nine.to_mont(), Fq::one()};
  // (q - 1) / 6
  uint32_t e[8];
  uint64_t rem = 0;
  for (int i = 7; i >= 0; i--) {
    uint64_t limb = FqParams::mod(i) - (i == 0 ? 1u : 0u);   // q is odd: q - 1 only changes limb 0
    uint64_t cur = (rem << 32) | limb;
    e[i] = (uint32_t)(cur / 6);
    rem = cur % 6;
  }
  Fq2 g = fq2_pow(xi, e, 8);
  pc.gamma[0] = Fq2::one();
  for (int i = 1; i < 6; i++) pc.gamma[i] = pc.gamma[i - 1].mulc(g);
  Fq three = Fq::zero();
  three.v[0] = 3;
  pc.twist_b = Fq2{three.to_mont(), Fq::zero()}.mulc(xi.inv());
}

ZKB_PNI Fq12 fq12_frobenius(Fq12 f, const PairingConsts &pc) {
  for (int i = 0; i < 6; i++) f.coef(i) = fq2_conj(f.coef(i)).mulc(pc.gamma[i]);
  return f;
}

ZKB_HD Fq12 pairing_line(const Fq2 &lam, const G2Affine &T, const G1Affine &P) {
  Fq12 l = {Fq6::zero(), Fq6::zero()};
  l.coef(0) = Fq2{P.y, Fq::zero()};
  l.coef(1) = fq2_mul_fq(lam, P.x).neg();
  l.coef(3) = lam.mulc(T.x) - T.y;
  return l;
}

ZKB_PNI void pairing_dbl_step(G2Affine &T, Fq12 &f, const G1Affine &P) {
  Fq2 x2 = T.x.sqrc();
  Fq2 lam = (x2.dbl() + x2).mulc(T.y.dbl().inv());
  f = f.mul(pairing_line(lam, T, P));
  Fq2 x3 = lam.sqrc() - T.x.dbl();
  Fq2 y3 = lam.mulc(T.x - x3) - T.y;
  T.x = x3;
  T.y = y3;
}
ZKB_PNI void pairing_add_step(G2Affine &T, const G2Affine &Q, Fq12 &f, const G1Affine &P) {
  Fq2 lam = (Q.y - T.y).mulc((Q.x - T.x).inv());
  f = f.mul(pairing_line(lam, T, P));
  Fq2 x3 = lam.sqrc() - T.x - Q.x;
  Fq2 y3 = lam.mulc(T.x - x3) - T.y;
  T.x = x3;
  T.y = y3;
}

// f_{6x+2,Q}(P) * l_{[6x+2]Q, pi(Q)}(P) * l_{., -pi^2(Q)}(P);  6x+2 = 0x19d797039be763ba8
ZKB_PNI Fq12 pairing_miller(const G1Affine &P, const G2Affine &Q, const PairingConsts &pc) {
  if (P.is_inf() || Q.is_inf()) return Fq12::one();
  const uint32_t lo0 = 0xbe763ba8u, lo1 = 0x9d797039u;   // low 64 bits; bit 64 is set
  Fq12 f = Fq12::one();
  G2Affine T = Q;
  for (int i = 63; i >= 0; i--) {
    f = f.mul(f);
    pairing_dbl_step(T, f, P);
    uint32_t bit = i >= 32 ? (lo1 >> (i - 32)) & 1u : (lo0 >> i) & 1u;
    if (bit) pairing_add_step(T, Q, f, P);
  }
  G2Affine Q1 = {fq2_conj(Q.x).mulc(pc.gamma[2]), fq2_conj(Q.y).mulc(pc.gamma[3])};
  G2Affine Q2 = {fq2_conj(Q1.x).mulc(pc.gamma[2]), fq2_conj(Q1.y).mulc(pc.gamma[3]).neg()};
  pairing_add_step(T, Q1, f, P);
  pairing_add_step(T, Q2, f, P);
  return f;
}

// f^((q^12 - 1)/r): easy part by conjugation/Frobenius, hard part (q^4 - q^2 + 1)/r by square-and-multiply
ZKB_PNI Fq12 pairing_final_exp(const Fq12 &f, const PairingConsts &pc) {
  const uint32_t hard[24] = {0xccdf42b1u, 0xe81bb482u, 0xf49c36d4u, 0x5abf5cc4u, 0x1da014fdu, 0xf1154e7eu,
                             0x87cdbacfu, 0xdcc7b44cu, 0x954bcf8au, 0xaaa441e3u, 0xd5095f23u, 0x6b887d56u,
                             0xf3fd90c6u, 0x79581e16u, 0xd189227du, 0x3b1b1355u, 0x61876f6bu, 0x4e529a58u,
                             0xd5b12278u, 0x6c0eb522u, 0x83177fafu, 0x331ec151u, 0x0b0759adu, 0x01baaa71u};
  Fq12 a = f.conj().mul(f.inv());                                   // ^(q^6 - 1)
  a = fq12_frobenius(fq12_frobenius(a, pc), pc).mul(a);             // ^(q^2 + 1)
  Fq12 r = Fq12::one();
  bool started = false;
  for (int i = 24 * 32 - 1; i >= 0; i--) {
    if (started) r = r.mul(r);
    if ((hard[i >> 5] >> (i & 31)) & 1) {
      r = started ? r.mul(a) : a;
      started = true;
    }
  }
  return r;
}

template <class F>
ZKB_HD bool on_curve(const Affine<F> &p, const F &b) {
  if (p.is_inf()) return true;
  return p.y.sqrc() == p.x.sqrc().mulc(p.x) + b;
}

// Verification key in device form (Montgomery coordinates); ab_miller = Miller value of (alpha1, beta2)
struct VerifyingKey {
  G1Affine alpha1;
  G2Affine beta2, gamma2, delta2;
  Fq12 ab_miller;
  uint32_t n_public;
};

// One Groth16 check.  pub: n_public canonical scalars (< r, checked by the caller); proof points in Montgomery form.
// Returns true iff the proof is valid.
ZKB_PNI bool groth16_verify_one(const VerifyingKey &vk, const G1Affine *ic, const Fr *pub, const G1Affine &A,
                               const G2Affine &B, const G1Affine &C, const PairingConsts &pc) {
  Fq three = Fq::zero();
  three.v[0] = 3;
  Fq b1 = three.to_mont();
  if (A.is_inf() || B.is_inf()) return false;
  if (!on_curve(A, b1) || !on_curve(C, b1) || !on_curve(B, pc.twist_b)) return false;
  XYZZ<Fq> vkx = XYZZ<Fq>::from_affine(ic[0]);
  for (uint32_t i = 0; i < vk.n_public; i++) {
    XYZZ<Fq> t, p = XYZZ<Fq>::from_affine(ic[i + 1]);
    xyzz_mul_ni(&p, pub[i].v, &t);
    xyzz_add_ni(&vkx, &t);
  }
  G1Affine vx;
  xyzz_to_affine_ni(&vkx, &vx);
  G1Affine negA = {A.x, A.y.neg()};
  Fq12 f = pairing_miller(negA, B, pc);
  f = f.mul(vk.ab_miller);
  f = f.mul(pairing_miller(vx, vk.gamma2, pc));
  f = f.mul(pairing_miller(C, vk.delta2, pc));
  return pairing_final_exp(f, pc) == Fq12::one();
}

}  // namespace zkb
