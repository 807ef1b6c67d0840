// Test-only host build of csrc/pairing.cuh (portable path): Groth16 verification of canonical-form inputs.
#include "../../zk_franchise_proof_circuit_b200/csrc/pairing.cuh"
#include <cstring>
#include <vector>
using namespace zkb;

static Fq fq_from(const uint8_t *b) { Fq x; memcpy(x.v, b, 32); return x.to_mont(); }
static G1Affine g1_from(const uint8_t *b) { return {fq_from(b), fq_from(b + 32)}; }
static G2Affine g2_from(const uint8_t *b) { return {Fq2{fq_from(b), fq_from(b + 32)}, Fq2{fq_from(b + 64), fq_from(b + 96)}}; }

extern "C" int host_groth16_verify(const uint8_t *vk448, const uint8_t *ic, int n_pub, const uint8_t *pub, const uint8_t *proof256) {
  PairingConsts pc;
  pairing_consts_init(pc);
  VerifyingKey vk;
  vk.alpha1 = g1_from(vk448);
  vk.beta2 = g2_from(vk448 + 64);
  vk.gamma2 = g2_from(vk448 + 192);
  vk.delta2 = g2_from(vk448 + 320);
  vk.n_public = (uint32_t)n_pub;
  vk.ab_miller = pairing_miller(vk.alpha1, vk.beta2, pc);
  std::vector<G1Affine> icv(n_pub + 1);
  for (int i = 0; i <= n_pub; i++) icv[i] = g1_from(ic + 64 * i);
  std::vector<Fr> pv(n_pub);
  for (int i = 0; i < n_pub; i++) memcpy(pv[i].v, pub + 32 * i, 32);
  return groth16_verify_one(vk, icv.data(), pv.data(), g1_from(proof256), g2_from(proof256 + 64), g1_from(proof256 + 192), pc) ? 1 : 0;
}
