// Reader for the proving key the reference passes to `prover.Prove` (`zk_census_test.go:81,89`) and
// `groth16.fullProve` (`ts_inputs/src/example.ts:361`): snarkjs binfile "zkey", version 1, sections
// 1 (protocol) 2 (groth16 header) 3 (IC) 4 (coefficients) 5 (A) 6 (B1) 7 (B2) 8 (C) 9 (H) 10 (contributions);
// layout per SURVEY.md 8a G1.  Points are affine with Montgomery coordinates (infinity = all zero) and are
// uploaded byte-for-byte; coefficients are value * R^2 mod r so that one Montgomery product with a canonical
// witness value yields a Montgomery-form term.
#include "zkey.h"
#include <cstring>

namespace zkb {

static const uint32_t QMOD[8] = {0xd87cfd47u, 0x3c208c16u, 0x6871ca8du, 0x97816a91u,
                                 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
static const uint32_t RMOD[8] = {0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u,
                                 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};

static uint32_t rd32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
static uint64_t rd64(const uint8_t *p) { uint64_t v; memcpy(&v, p, 8); return v; }

bool parse_zkey(const uint8_t *buf, size_t len, ZkeyView &z, std::string &err) {
  if (len < 12 || memcmp(buf, "zkey", 4) != 0) { err = "zkey: bad magic"; return false; }
  if (rd32(buf + 4) != 1) { err = "zkey: unsupported version"; return false; }
  uint32_t nsec = rd32(buf + 8);
  const uint8_t *sec[11] = {nullptr};
  uint64_t sz[11] = {0};
  size_t p = 12;
  for (uint32_t i = 0; i < nsec; i++) {
    if (p + 12 > len) { err = "zkey: truncated section header"; return false; }
    uint32_t id = rd32(buf + p);
    uint64_t s = rd64(buf + p + 4);
    p += 12;
    if (s > len - p) { err = "zkey: truncated section"; return false; }
    if (id >= 1 && id <= 10) { sec[id] = buf + p; sz[id] = s; }
    p += s;
  }
  for (int id = 1; id <= 9; id++)
    if (!sec[id]) { err = "zkey: missing section " + std::to_string(id); return false; }
  if (sz[1] < 4 || rd32(sec[1]) != 1) { err = "zkey: protocol is not groth16"; return false; }
  const uint8_t *h = sec[2];
  if (sz[2] < 4 + 32 + 4 + 32 + 12 + 64 * 3 + 128 * 3) { err = "zkey: short header"; return false; }
  if (rd32(h) != 32 || memcmp(h + 4, QMOD, 32) != 0) { err = "zkey: base field is not bn128"; return false; }
  if (rd32(h + 36) != 32 || memcmp(h + 40, RMOD, 32) != 0) { err = "zkey: scalar field is not bn128"; return false; }
  z.n_vars = rd32(h + 72);
  z.n_public = rd32(h + 76);
  z.domain = rd32(h + 80);
  z.power = 0;
  while ((1u << z.power) < z.domain) z.power++;
  if ((1u << z.power) != z.domain) { err = "zkey: domain size is not a power of two"; return false; }
  const uint8_t *q = h + 84;
  z.alpha1 = q; q += 64;
  z.beta1 = q; q += 64;
  z.beta2 = q; q += 128;
  z.gamma2 = q; q += 128;
  z.delta1 = q; q += 64;
  z.delta2 = q;
  if (z.n_public + 1 > z.n_vars) { err = "zkey: nPublic >= nVars"; return false; }
  if (sz[3] != (uint64_t)(z.n_public + 1) * 64) { err = "zkey: IC section size"; return false; }
  z.ic = sec[3];
  if (sz[4] < 4) { err = "zkey: coefficient section"; return false; }
  z.n_coefs = rd32(sec[4]);
  if (sz[4] != 4 + (uint64_t)z.n_coefs * 44) { err = "zkey: coefficient section size"; return false; }
  z.coefs = sec[4] + 4;
  if (sz[5] != (uint64_t)z.n_vars * 64 || sz[6] != (uint64_t)z.n_vars * 64 || sz[7] != (uint64_t)z.n_vars * 128 ||
      sz[8] != (uint64_t)(z.n_vars - z.n_public - 1) * 64 || sz[9] != (uint64_t)z.domain * 64) {
    err = "zkey: point section sizes";
    return false;
  }
  z.a = sec[5]; z.b1 = sec[6]; z.b2 = sec[7]; z.c = sec[8]; z.h = sec[9];
  for (uint32_t i = 0; i < z.n_coefs; i++) {
    const uint8_t *c = z.coefs + (size_t)i * 44;
    if (rd32(c) > 1 || rd32(c + 4) >= z.domain || rd32(c + 8) >= z.n_vars) { err = "zkey: coefficient out of range"; return false; }
  }
  return true;
}

void build_csr(const ZkeyView &z, uint32_t matrix, CoefCsr &out) {
  out.row_ptr.assign(z.domain + 1, 0);
  for (uint32_t i = 0; i < z.n_coefs; i++) {
    const uint8_t *c = z.coefs + (size_t)i * 44;
    if (rd32(c) == matrix) out.row_ptr[rd32(c + 4) + 1]++;
  }
  for (uint32_t r = 0; r < z.domain; r++) out.row_ptr[r + 1] += out.row_ptr[r];
  uint32_t nnz = out.row_ptr[z.domain];
  out.wire.resize(nnz);
  out.value.resize((size_t)nnz * 32);
  std::vector<uint32_t> cur(out.row_ptr.begin(), out.row_ptr.end() - 1);
  for (uint32_t i = 0; i < z.n_coefs; i++) {
    const uint8_t *c = z.coefs + (size_t)i * 44;
    if (rd32(c) != matrix) continue;
    uint32_t k = cur[rd32(c + 4)]++;
    out.wire[k] = rd32(c + 8);
    memcpy(out.value.data() + (size_t)k * 32, c + 12, 32);
  }
}

}  // namespace zkb
