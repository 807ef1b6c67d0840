// Host check of csrc/finalize.cuh against the oracle: random "partial sums", alpha/beta/delta = k*G.
#include "../../zk_franchise_proof_circuit_b200/csrc/finalize.cuh"
#include "../../oracle/bn254.hpp"
#include <cstdio>
#include <random>
#include <vector>
static std::mt19937_64 rng(777);
template <class ZF, class OF> static ZF cv(const OF &x) { ZF r; memcpy(r.v, x.v, 32); return r; }
static zkb::Fq2 cv2(const bn::Fq2 &x) { return {cv<zkb::Fq>(x.a), cv<zkb::Fq>(x.b)}; }
static zkb::G1Affine g1(const bn::G1A &p) { return {cv<zkb::Fq>(p.x), cv<zkb::Fq>(p.y)}; }
static zkb::G2Affine g2(const bn::G2A &p) { return {cv2(p.x), cv2(p.y)}; }
static bn::G1A rnd1() { uint64_t k[4] = {rng(), rng(), rng(), rng() >> 4}; return bn::G1J::from_aff(bn::g1_gen()).mul(k).to_aff(); }
static bn::G2A rnd2() { uint64_t k[4] = {rng(), rng(), rng(), rng() >> 4}; return bn::G2J::from_aff(bn::g2_gen()).mul(k).to_aff(); }
#define CHECK(c) do { if (!(c)) { printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #c); return 1; } } while (0)
int main() {
  bn::init();
  bn::G1A alpha = rnd1(), beta1 = rnd1(), delta1 = rnd1();
  bn::G2A beta2 = rnd2(), delta2 = rnd2();
  std::vector<zkb::G1Affine> d1(960);
  std::vector<zkb::G2Affine> d2(960);
  for (uint32_t i = 0; i < 960; i++) { zkb::fixed_table_entry<zkb::Fq>(g1(delta1), i, d1[i]); zkb::fixed_table_entry<zkb::Fq2>(g2(delta2), i, d2[i]); }
  for (int it = 0; it < 3; it++) {
    bn::G1A pa = rnd1(), pb1 = rnd1(), pc = rnd1(), ph = rnd1();
    bn::G2A pb = rnd2();
    uint64_t rr[4] = {rng(), rng(), rng(), rng() >> 4}, ss[4] = {rng(), rng(), rng(), rng() >> 4};
    if (it == 0) { rr[0] = 1234567; rr[1] = rr[2] = rr[3] = 0; ss[0] = 7654321; ss[1] = ss[2] = ss[3] = 0; }
    bn::Fr r = bn::Fr::from_raw(rr), s = bn::Fr::from_raw(ss);
    // oracle
    bn::G1J A = bn::G1J::from_aff(pa).add_aff(alpha).add(bn::G1J::from_aff(delta1).mul(rr));
    bn::G2J B = bn::G2J::from_aff(pb).add_aff(beta2).add(bn::G2J::from_aff(delta2).mul(ss));
    bn::G1J B1 = bn::G1J::from_aff(pb1).add_aff(beta1).add(bn::G1J::from_aff(delta1).mul(ss));
    bn::Fr rs = (r * s).neg();
    uint64_t rsr[4];
    rs.to_raw(rsr);
    bn::G1J C = bn::G1J::from_aff(pc).add_aff(ph).add(A.mul(ss)).add(B1.mul(rr)).add(bn::G1J::from_aff(delta1).mul(rsr));
    bn::G1A eA = A.to_aff(), eC = C.to_aff();
    bn::G2A eB = B.to_aff();
    // product formulas
    zkb::Fr zr, zs;
    memcpy(zr.v, rr, 32);
    memcpy(zs.v, ss, 32);
    zkb::G1XYZZ zA, zB1, t0, t1, t2, tb[15];
    zkb::G2XYZZ zB;
    zkb::G1XYZZ xpa = zkb::G1XYZZ::from_affine(g1(pa)), xpb1 = zkb::G1XYZZ::from_affine(g1(pb1));
    zkb::G2XYZZ xpb = zkb::G2XYZZ::from_affine(g2(pb));
    zkb::G1Affine za = g1(alpha), zb1 = g1(beta1);
    zkb::G2Affine zb2 = g2(beta2);
    zkb::fin_point<zkb::Fq>(&xpa, &za, d1.data(), zr.v, &zA);
    zkb::fin_point<zkb::Fq>(&xpb1, &zb1, d1.data(), zs.v, &zB1);
    zkb::fin_point<zkb::Fq2>(&xpb, &zb2, d2.data(), zs.v, &zB);
    zkb::var_mul<zkb::Fq>(&zA, zs.v, tb, &t0);
    zkb::var_mul<zkb::Fq>(&zB1, zr.v, tb, &t1);
    zkb::fin_neg_rs_delta(d1.data(), zr, zs, &t2);
    zkb::G1XYZZ zC = zkb::G1XYZZ::from_affine(g1(pc));
    zkb::G1XYZZ zh = zkb::G1XYZZ::from_affine(g1(ph));
    zC.add(zh); zC.add(t0); zC.add(t1); zC.add(t2);
    zkb::G1Affine a = zA.to_affine(), c = zC.to_affine();
    zkb::G2Affine b = zB.to_affine();
    CHECK(memcmp(a.x.v, eA.x.v, 32) == 0 && memcmp(a.y.v, eA.y.v, 32) == 0);
    CHECK(memcmp(b.x.a.v, eB.x.a.v, 32) == 0 && memcmp(b.y.b.v, eB.y.b.v, 32) == 0);
    CHECK(memcmp(c.x.v, eC.x.v, 32) == 0 && memcmp(c.y.v, eC.y.v, 32) == 0);
  }
  printf("OK\n");
  return 0;
}
