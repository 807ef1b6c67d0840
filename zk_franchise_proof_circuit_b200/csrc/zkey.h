// snarkjs .zkey (Groth16, bn128) reader - see zkey.cc
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace zkb {

struct ZkeyCoef { uint32_t matrix, row, wire; const uint8_t *value; };  // value: 32 B, coefficient * R^2 mod r

struct ZkeyView {
  uint32_t n_vars = 0, n_public = 0, domain = 0, power = 0;
  const uint8_t *alpha1 = nullptr, *beta1 = nullptr, *beta2 = nullptr, *gamma2 = nullptr, *delta1 = nullptr,
                *delta2 = nullptr;          // affine, Montgomery coordinates, as stored
  const uint8_t *ic = nullptr;              // (n_public + 1) x 64
  const uint8_t *coefs = nullptr;           // section 4 payload after the count
  uint32_t n_coefs = 0;
  const uint8_t *a = nullptr, *b1 = nullptr, *b2 = nullptr, *c = nullptr, *h = nullptr;   // sections 5..9
};

// Borrows `buf` (must outlive the view).  Returns false + err on malformed / non-bn128 / non-groth16 keys.
bool parse_zkey(const uint8_t *buf, size_t len, ZkeyView &z, std::string &err);

// CSR of the A and B matrices by row: for row r the entries [row_ptr[r], row_ptr[r+1]) of `wire` / `value`
struct CoefCsr {
  std::vector<uint32_t> row_ptr;    // domain + 1
  std::vector<uint32_t> wire;
  std::vector<uint8_t> value;       // 32 B each (still * R^2)
};
void build_csr(const ZkeyView &z, uint32_t matrix, CoefCsr &out);

}  // namespace zkb
