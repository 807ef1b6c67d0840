// Host-side check of csrc/fp.cuh + csrc/ec.cuh (their portable code path) against the oracle's
// independent 4x64 implementation.  Test-only: links oracle/bn254.hpp as the checker.
#include "../../zk_franchise_proof_circuit_b200/csrc/ec.cuh"
#include "../../oracle/bn254.hpp"
#include <cstdio>
#include <random>

static std::mt19937_64 rng(12345);
template <class OF> static OF rnd() {
  uint64_t raw[4] = {rng(), rng(), rng(), rng() >> 3};
  while (bn::cmp4(raw, OF::P.p) >= 0) raw[3] >>= 1;
  OF r;
  memcpy(r.v, raw, 32);  // treat as Montgomery representation directly
  return r;
}
template <class ZF, class OF> static ZF cv(const OF &x) { ZF r; memcpy(r.v, x.v, 32); return r; }
template <class ZF, class OF> static bool same(const ZF &a, const OF &b) { return memcmp(a.v, b.v, 32) == 0; }
#define CHECK(c) do { if (!(c)) { printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #c); return 1; } } while (0)

template <class ZF, class OF> static int field_tests() {
  for (int i = 0; i < 2000; i++) {
    OF a = rnd<OF>(), b = rnd<OF>();
    ZF za = cv<ZF>(a), zb = cv<ZF>(b);
    CHECK(same(za * zb, a * b));
    CHECK(same(za + zb, a + b));
    CHECK(same(za - zb, a - b));
    CHECK(same(za.neg(), a.neg()));
    if (i < 20) CHECK(same(za.inv(), a.inv()));
  }
  CHECK(same(ZF::one(), OF::one()));
  CHECK(same(ZF::zero().inv(), OF::zero()));
  return 0;
}

static zkb::Fq2 cv2(const bn::Fq2 &x) { return {cv<zkb::Fq>(x.a), cv<zkb::Fq>(x.b)}; }
static bool same2(const zkb::Fq2 &a, const bn::Fq2 &b) { return same(a.a, b.a) && same(a.b, b.b); }

int main() {
  bn::init();
  if (field_tests<zkb::Fq, bn::Fq>()) return 1;
  if (field_tests<zkb::Fr, bn::Fr>()) return 1;
  for (int i = 0; i < 500; i++) {
    bn::Fq2 a = {rnd<bn::Fq>(), rnd<bn::Fq>()}, b = {rnd<bn::Fq>(), rnd<bn::Fq>()};
    CHECK(same2(cv2(a) * cv2(b), a * b));
    CHECK(same2(cv2(a).sqr(), a.sqr()));
    if (i < 10) CHECK(same2(cv2(a).inv(), a.inv()));
  }
  // G1: random multiples of the generator
  bn::G1A g = bn::g1_gen();
  bn::G2A g2 = bn::g2_gen();
  for (int i = 0; i < 30; i++) {
    uint64_t k1[4] = {rng(), rng(), rng(), rng() >> 4}, k2[4] = {rng(), rng(), 0, 0};
    bn::G1A p = bn::G1J::from_aff(g).mul(k1).to_aff(), q = bn::G1J::from_aff(g).mul(k2).to_aff();
    zkb::G1Affine zp = {cv<zkb::Fq>(p.x), cv<zkb::Fq>(p.y)}, zq = {cv<zkb::Fq>(q.x), cv<zkb::Fq>(q.y)};
    zkb::G1XYZZ acc = zkb::G1XYZZ::from_affine(zp);
    acc.add_affine(zq);
    bn::G1A e = bn::G1J::from_aff(p).add_aff(q).to_aff();
    zkb::G1Affine r = acc.to_affine();
    CHECK(same(r.x, e.x) && same(r.y, e.y));
    // doubling through add_affine, cancellation, general add, scalar mul
    zkb::G1XYZZ d = zkb::G1XYZZ::from_affine(zp);
    d.add_affine(zp);
    bn::G1A e2 = bn::G1J::from_aff(p).dbl().to_aff();
    r = d.to_affine();
    CHECK(same(r.x, e2.x) && same(r.y, e2.y));
    zkb::G1XYZZ c = zkb::G1XYZZ::from_affine(zp);
    c.add_affine(zp.neg());
    CHECK(c.is_inf());
    zkb::G1XYZZ s = acc;
    s.add(d);
    bn::G1A e3 = bn::G1J::from_aff(e).add_aff(e2).to_aff();
    r = s.to_affine();
    CHECK(same(r.x, e3.x) && same(r.y, e3.y));
    zkb::G1XYZZ s2 = d;
    s2.add(d);
    bn::G1A e4 = bn::G1J::from_aff(e2).dbl().to_aff();
    r = s2.to_affine();
    CHECK(same(r.x, e4.x) && same(r.y, e4.y));
    uint32_t k32[8];
    memcpy(k32, k2, 32);
    r = zkb::G1XYZZ::from_affine(zp).mul(k32).to_affine();
    bn::G1A e5 = bn::G1J::from_aff(p).mul(k2).to_aff();
    CHECK(same(r.x, e5.x) && same(r.y, e5.y));
    // G2
    bn::G2A P2 = bn::G2J::from_aff(g2).mul(k1).to_aff(), Q2 = bn::G2J::from_aff(g2).mul(k2).to_aff();
    zkb::G2Affine zP2 = {cv2(P2.x), cv2(P2.y)}, zQ2 = {cv2(Q2.x), cv2(Q2.y)};
    zkb::G2XYZZ a2 = zkb::G2XYZZ::from_affine(zP2);
    a2.add_affine(zQ2);
    bn::G2A E2 = bn::G2J::from_aff(P2).add_aff(Q2).to_aff();
    zkb::G2Affine r2 = a2.to_affine();
    CHECK(same2(r2.x, E2.x) && same2(r2.y, E2.y));
    a2.add_affine(zQ2);
    a2.add(a2);
    bn::G2A E3 = bn::G2J::from_aff(E2).add_aff(Q2).dbl().to_aff();
    r2 = a2.to_affine();
    CHECK(same2(r2.x, E3.x) && same2(r2.y, E3.y));
  }
  printf("OK\n");
  return 0;
}
