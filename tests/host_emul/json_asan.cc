// AddressSanitizer check of the decimal formatter / JSON helpers of csrc/json_io.cc (u256_to_dec wrote one
// byte past its stack buffer for values >= 10^72, i.e. for almost every field element).
#include <cstdio>
#include <cstring>
#include <string>
#include "../../zk_franchise_proof_circuit_b200/csrc/json_io.cc"

int main() {
  using namespace zkb;
  const uint32_t R1[8] = {0xf0000000u, 0x43e1f593u, 0x79b97091u, 0x2833e848u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
  const uint32_t Q1[8] = {0xd87cfd46u, 0x3c208c16u, 0x6871ca8du, 0x97816a91u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
  uint32_t ALL[8], ZERO[8] = {0}, ONE[8] = {1};
  memset(ALL, 0xff, 32);
  const char *want[5] = {"21888242871839275222246405745257275088548364400416034343698204186575808495616",
                         "21888242871839275222246405745257275088696311157297823662689037894645226208582",
                         "115792089237316195423570985008687907853269984665640564039457584007913129639935", "0", "1"};
  const uint32_t *in[5] = {R1, Q1, ALL, ZERO, ONE};
  for (int i = 0; i < 5; i++) {
    std::string s = u256_to_dec(in[i]);
    if (s != want[i]) { printf("mismatch %d: %s\n", i, s.c_str()); return 1; }
  }
  uint8_t proof[256];
  for (int i = 0; i < 256; i++) proof[i] = (uint8_t)(i * 37 + 11);
  for (int k = 0; k < 8; k++) proof[32 * k + 31] &= 0x2f;
  std::string pj = proof_to_json(proof, true), sj = publics_to_json(proof, 8);
  if (pj.find("pi_a") == std::string::npos || sj[0] != '[') return 2;
  puts("json_asan ok");
  return 0;
}
