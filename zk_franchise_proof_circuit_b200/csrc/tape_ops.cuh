// Semantics of one tape operation (wasm_symexec.h) on BN254 scalar-field values in Montgomery form - host/device, so
// the same code is unit-tested on a machine without a GPU (tests/host_emul/tape_host.cc) and run by tape_eval.cu.
// Each case restates the circom field runtime function it stands for (the wasm's Fr_* functions, circom 2.1.5):
//   comparisons treat values above (r - 1) / 2 as negative; bit operations work on the canonical value, mask the
//   result to 254 bits and reduce it once (Fr_adjustBinResult); shifts by a "negative" amount shift the other way,
//   shifts by 254 or more give 0; logical operations and Fr_isTrue test for non-zero; 0 has inverse 0; integer
//   division, remainder and the exponent of a power read their operands as canonical values in [0, r).
#pragma once
#include "fp.cuh"
#include "wasm_symexec.h"

namespace zkb {

struct U256 { uint32_t w[8]; };

ZKB_HD U256 fr_canon(const Fr &x) {
  Fr c = x.from_mont();
  U256 r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.w[i] = c.v[i];
  return r;
}
ZKB_HD Fr fr_from_canon(const U256 &x) {
  Fr c;
#pragma unroll
  for (int i = 0; i < 8; i++) c.v[i] = x.w[i];
  return c.to_mont();
}
ZKB_HD int u256_cmp(const U256 &a, const U256 &b) {
  for (int i = 7; i >= 0; i--) {
    if (a.w[i] < b.w[i]) return -1;
    if (a.w[i] > b.w[i]) return 1;
  }
  return 0;
}
ZKB_HD U256 fr_half() {          // (r - 1) / 2
  U256 h;
  uint32_t carry = 0;
  for (int i = 7; i >= 0; i--) {
    uint32_t m = FrParams::mod(i) - (i == 0 ? 1u : 0u);     // r is odd: r - 1 only changes limb 0
    h.w[i] = (m >> 1) | (carry << 31);
    carry = m & 1u;
  }
  return h;
}
ZKB_HD bool fr_is_negative(const U256 &a) { return u256_cmp(a, fr_half()) > 0; }
// signed comparison of canonical values: -1 / 0 / 1
ZKB_HD int fr_signed_cmp(const U256 &a, const U256 &b) {
  const bool na = fr_is_negative(a), nb = fr_is_negative(b);
  if (na != nb) return na ? -1 : 1;
  return u256_cmp(a, b);
}
// mask to 254 bits, subtract r once if needed
ZKB_HD U256 fr_adjust_bin(U256 x) {
  x.w[7] &= 0x3fffffffu;
  U256 m;
  for (int i = 0; i < 8; i++) m.w[i] = FrParams::mod(i);
  if (u256_cmp(x, m) >= 0) {
    uint64_t br = 0;
    for (int i = 0; i < 8; i++) {
      uint64_t d = (uint64_t)x.w[i] - m.w[i] - br;
      x.w[i] = (uint32_t)d;
      br = (d >> 32) & 1;
    }
  }
  return x;
}
ZKB_HD U256 u256_shr(const U256 &a, uint32_t n) {          // n < 256
  U256 r;
  const uint32_t ws = n >> 5, bs = n & 31;
  for (int i = 0; i < 8; i++) {
    const uint32_t lo = (uint32_t)i + ws < 8 ? a.w[i + ws] : 0u, hi = (uint32_t)i + ws + 1 < 8 ? a.w[i + ws + 1] : 0u;
    r.w[i] = bs ? (lo >> bs) | (hi << (32 - bs)) : lo;
  }
  return r;
}
ZKB_HD U256 u256_shl(const U256 &a, uint32_t n) {          // n < 256; bits above 2^256 are dropped (masked to 254 anyway)
  U256 r;
  const uint32_t ws = n >> 5, bs = n & 31;
  for (int i = 7; i >= 0; i--) {
    const uint32_t hi = (uint32_t)i >= ws ? a.w[i - ws] : 0u, lo = (uint32_t)i >= ws + 1 ? a.w[i - ws - 1] : 0u;
    r.w[i] = bs ? (hi << bs) | (lo >> (32 - bs)) : hi;
  }
  return r;
}
// shift amount semantics shared by Fr_shr / Fr_shl: `right` = the operation asked for
ZKB_HD Fr fr_shift(const Fr &a, const Fr &b, bool right) {
  U256 bc = fr_canon(b);
  bool r = right;
  if (fr_is_negative(bc)) {                                // negative amount: shift the other way by |b|
    U256 m;
    uint64_t br = 0;
    for (int i = 0; i < 8; i++) {
      uint64_t d = (uint64_t)FrParams::mod(i) - bc.w[i] - br;
      m.w[i] = (uint32_t)d;
      br = (d >> 32) & 1;
    }
    bc = m;
    r = !r;
  }
  bool big = bc.w[0] >= 254;
  for (int i = 1; i < 8; i++) big = big || bc.w[i] != 0;
  if (big) return Fr::zero();
  const U256 ac = fr_canon(a);
  return fr_from_canon(r ? u256_shr(ac, bc.w[0]) : fr_adjust_bin(u256_shl(ac, bc.w[0])));
}
ZKB_HD Fr fr_bool(bool x) { return x ? Fr::one() : Fr::zero(); }

// quotient and remainder of canonical values, b != 0: restoring division, one bit per step from the top set bit of a
// (r < b < 2^254, so r << 1 stays inside 256 bits).  Rare in witness programs (range reductions of `<--` hints).
ZKB_HD void u256_divmod(const U256 &a, const U256 &b, U256 &q, U256 &r) {
  for (int i = 0; i < 8; i++) q.w[i] = r.w[i] = 0;
  int top = 255;
  while (top >= 0 && !((a.w[top >> 5] >> (top & 31)) & 1u)) top--;
  for (int i = top; i >= 0; i--) {
    uint32_t carry = (a.w[i >> 5] >> (i & 31)) & 1u;
    for (int k = 0; k < 8; k++) {
      const uint32_t nc = r.w[k] >> 31;
      r.w[k] = (r.w[k] << 1) | carry;
      carry = nc;
    }
    if (u256_cmp(r, b) >= 0) {
      uint64_t br = 0;
      for (int k = 0; k < 8; k++) {
        const uint64_t d = (uint64_t)r.w[k] - b.w[k] - br;
        r.w[k] = (uint32_t)d;
        br = (d >> 32) & 1;
      }
      q.w[i >> 5] |= 1u << (i & 31);
    }
  }
}

// result of `op`; asserts report through *failed
ZKB_HD Fr tape_apply(uint8_t op, const Fr &a, const Fr &b, const Fr &c, bool *failed) {
  switch (op) {
    case T_ADD: return a + b;
    case T_SUB: return a - b;
    case T_MUL: return a * b;
    case T_NEG: return a.neg();
    case T_DIV: return a * b.inv();
    case T_INV: return a.inv();
    case T_EQ: return fr_bool(a == b);
    case T_NEQ: return fr_bool(a != b);
    case T_LT: return fr_bool(fr_signed_cmp(fr_canon(a), fr_canon(b)) < 0);
    case T_GT: return fr_bool(fr_signed_cmp(fr_canon(a), fr_canon(b)) > 0);
    case T_LEQ: return fr_bool(fr_signed_cmp(fr_canon(a), fr_canon(b)) <= 0);
    case T_GEQ: return fr_bool(fr_signed_cmp(fr_canon(a), fr_canon(b)) >= 0);
    case T_LAND: return fr_bool(!a.is_zero() && !b.is_zero());
    case T_LOR: return fr_bool(!a.is_zero() || !b.is_zero());
    case T_LNOT: return fr_bool(a.is_zero());
    case T_SHR: return fr_shift(a, b, true);
    case T_SHL: return fr_shift(a, b, false);
    case T_BAND: case T_BOR: case T_BXOR: {
      const U256 x = fr_canon(a), y = fr_canon(b);
      U256 r;
      for (int i = 0; i < 8; i++) r.w[i] = op == T_BAND ? (x.w[i] & y.w[i]) : op == T_BOR ? (x.w[i] | y.w[i]) : (x.w[i] ^ y.w[i]);
      return fr_from_canon(fr_adjust_bin(r));
    }
    case T_BNOT: {
      U256 x = fr_canon(a);
      for (int i = 0; i < 8; i++) x.w[i] = ~x.w[i];
      return fr_from_canon(fr_adjust_bin(x));
    }
    case T_IDIV: case T_MOD: {
      if (b.is_zero()) { *failed = true; return Fr::zero(); }
      U256 q, r;
      u256_divmod(fr_canon(a), fr_canon(b), q, r);
      return fr_from_canon(op == T_IDIV ? q : r);
    }
    case T_POW: {
      const U256 e = fr_canon(b);
      return a.pow(e.w);
    }
    case T_ISTRUE: return fr_bool(!a.is_zero());
    case T_SELECT: return c.is_zero() ? b : a;
    case T_ASSERT_TRUE: if (a.is_zero()) *failed = true; return Fr::zero();
    case T_ASSERT_FALSE: if (!a.is_zero()) *failed = true; return Fr::zero();
    default: return a;                                       // T_COPY
  }
}

}  // namespace zkb
