// JSON shapes of the reference's artifacts:
//   inputs_example.json  - 12 keys, decimal strings, arrays flattened in order (`internal/inputs.go:14-31`)
//   proof.json           - {"pi_a":[x,y,"1"],"pi_b":[[x0,x1],[y0,y1],["1","0"]],"pi_c":[x,y,"1"]} compact, as
//                          the Go path re-marshals it (`zk_census_test.go:93-98`); snarkjs adds protocol/curve
//   signals.json         - JSON array of decimal strings
#include "json_io.h"
#include <cstring>

namespace zkb {

static const uint32_t RMOD[8] = {0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u,
                                 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};

static const uint32_t QMOD[8] = {0xd87cfd47u, 0x3c208c16u, 0x6871ca8du, 0x97816a91u,
                                 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};

static bool geq_m(const uint32_t v[9], const uint32_t *M) {
  if (v[8]) return true;
  for (int i = 7; i >= 0; i--) {
    if (v[i] > M[i]) return true;
    if (v[i] < M[i]) return false;
  }
  return true;
}
static void sub_m(uint32_t v[9], const uint32_t *M) {
  uint64_t br = 0;
  for (int i = 0; i < 8; i++) {
    uint64_t d = (uint64_t)v[i] - M[i] - br;
    v[i] = (uint32_t)d;
    br = (d >> 32) & 1;
  }
  v[8] -= (uint32_t)br;
}

static thread_local bool g_reduced = false;   // set when a parsed value was >= the modulus (or negative)
static bool dec_to_mod(const char *s, size_t len, uint32_t out[8], const uint32_t *RMOD);
bool dec_to_fr(const char *s, size_t len, uint32_t out[8]) { return dec_to_mod(s, len, out, RMOD); }
bool dec_to_fq(const char *s, size_t len, uint32_t out[8]) { return dec_to_mod(s, len, out, QMOD); }
bool json_value_was_reduced() { return g_reduced; }
void json_reset_reduced() { g_reduced = false; }

static bool dec_to_mod(const char *s, size_t len, uint32_t out[8], const uint32_t *RMOD) {
  uint32_t v[9] = {0};
  bool neg = false;
  size_t i = 0;
  if (i < len && s[i] == '-') { neg = true; i++; }
  if (i >= len) return false;
  for (; i < len; i++) {
    if (s[i] < '0' || s[i] > '9') return false;
    // v = v * 10 + d, then reduce (v < r before => v*10+9 < 10r + 9: at most 10 subtractions)
    uint64_t c = (uint64_t)(s[i] - '0');
    for (int k = 0; k < 9; k++) {
      c += (uint64_t)v[k] * 10;
      v[k] = (uint32_t)c;
      c >>= 32;
    }
    while (geq_m(v, RMOD)) { sub_m(v, RMOD); g_reduced = true; }
  }
  if (neg) {
    g_reduced = true;
    bool zero = true;
    for (int k = 0; k < 8; k++) zero = zero && v[k] == 0;
    if (!zero) {
      uint64_t br = 0;
      for (int k = 0; k < 8; k++) {
        uint64_t d = (uint64_t)RMOD[k] - v[k] - br;
        v[k] = (uint32_t)d;
        br = (d >> 32) & 1;
      }
    }
  }
  memcpy(out, v, 32);
  return true;
}

std::string u256_to_dec(const uint32_t val[8]) {
  uint32_t t[8];
  memcpy(t, val, 32);
  char buf[96];   // 2^256 - 1 has 78 digits; 9 passes of 9 digits each write up to 81 before the zeros are trimmed
  int n = 0;
  for (;;) {
    bool zero = true;
    for (int i = 0; i < 8; i++) zero = zero && t[i] == 0;
    if (zero) break;
    uint64_t rem = 0;
    for (int i = 7; i >= 0; i--) {
      uint64_t cur = (rem << 32) | t[i];
      t[i] = (uint32_t)(cur / 1000000000u);
      rem = cur % 1000000000u;
    }
    for (int k = 0; k < 9; k++) { buf[n++] = (char)('0' + rem % 10); rem /= 10; }
  }
  while (n > 1 && buf[n - 1] == '0') n--;
  if (n == 0) buf[n++] = '0';
  std::string s(buf, n);
  return std::string(s.rbegin(), s.rend());
}

namespace {
struct P {
  const char *s;
  size_t i, n;
  void ws() { while (i < n && (s[i] == ' ' || s[i] == '\n' || s[i] == '\t' || s[i] == '\r')) i++; }
  bool eat(char c) { ws(); if (i < n && s[i] == c) { i++; return true; } return false; }
  bool peek(char c) { ws(); return i < n && s[i] == c; }
  bool str(std::string &out) {
    ws();
    if (i >= n || s[i] != '"') return false;
    i++;
    size_t b = i;
    while (i < n && s[i] != '"') { if (s[i] == '\\') i++; i++; }
    if (i >= n) return false;
    out.assign(s + b, i - b);
    i++;
    return true;
  }
  // string or bare number token
  bool scalar(std::string &out) {
    ws();
    if (i < n && s[i] == '"') return str(out);
    size_t b = i;
    while (i < n && ((s[i] >= '0' && s[i] <= '9') || s[i] == '-')) i++;
    if (i == b) return false;
    out.assign(s + b, i - b);
    return true;
  }
};

// strict = false: a non-decimal scalar (e.g. "groth16") marks the whole value as non-numeric instead of failing
bool flatten(P &p, std::vector<uint32_t> &out, int depth, bool base_field = false, bool strict = true,
             bool *non_numeric = nullptr) {
  if (depth > 8) return false;
  if (p.eat('[')) {
    if (p.eat(']')) return true;
    do { if (!flatten(p, out, depth + 1, base_field, strict, non_numeric)) return false; } while (p.eat(','));
    return p.eat(']');
  }
  std::string tok;
  if (!p.scalar(tok)) return false;
  uint32_t v[8];
  bool ok = base_field ? dec_to_fq(tok.data(), tok.size(), v) : dec_to_fr(tok.data(), tok.size(), v);
  if (!ok) {
    if (strict) return false;
    if (non_numeric) *non_numeric = true;
    return true;
  }
  out.insert(out.end(), v, v + 8);
  return true;
}
}  // namespace

bool parse_inputs_json(const char *s, size_t len, std::map<std::string, std::vector<uint32_t>> &out, std::string &err) {
  P p{s, 0, len};
  if (!p.eat('{')) { err = "inputs: expected object"; return false; }
  if (p.eat('}')) return true;
  do {
    std::string key;
    if (!p.str(key) || !p.eat(':')) { err = "inputs: bad key"; return false; }
    std::vector<uint32_t> vals;
    if (!flatten(p, vals, 0)) { err = "inputs: bad value for " + key; return false; }
    out[key] = std::move(vals);
  } while (p.eat(','));
  if (!p.eat('}')) { err = "inputs: expected }"; return false; }
  return true;
}

// Generic {"key": number | "decimal" | nested arrays of those | "text"} reader used for verification_key.json
// and proof.json: numeric values are flattened and reduced mod q (base_field) or mod r; text values are skipped.
bool parse_json_numbers(const char *s, size_t len, bool base_field, std::map<std::string, std::vector<uint32_t>> &out,
                        std::string &err) {
  P p{s, 0, len};
  if (!p.eat('{')) { err = "expected a JSON object"; return false; }
  if (p.eat('}')) return true;
  do {
    std::string key;
    if (!p.str(key) || !p.eat(':')) { err = "bad key"; return false; }
    std::vector<uint32_t> vals;
    bool non_numeric = false;
    if (!flatten(p, vals, 0, base_field, false, &non_numeric)) { err = "bad value for " + key; return false; }
    if (!non_numeric) out[key] = std::move(vals);
  } while (p.eat(','));
  if (!p.eat('}')) { err = "expected }"; return false; }
  return true;
}

bool parse_json_array(const char *s, size_t len, std::vector<uint32_t> &out, std::string &err) {
  P p{s, 0, len};
  if (!p.peek('[') || !flatten(p, out, 0)) { err = "expected a JSON array of decimal strings"; return false; }
  return true;
}

static std::string dec32(const uint8_t *b) {
  uint32_t v[8];
  memcpy(v, b, 32);
  return u256_to_dec(v);
}

std::string proof_to_json(const uint8_t *pf, bool snarkjs_style) {
  std::string s = "{\"pi_a\":[\"" + dec32(pf) + "\",\"" + dec32(pf + 32) + "\",\"1\"],\"pi_b\":[[\"" + dec32(pf + 64) +
                  "\",\"" + dec32(pf + 96) + "\"],[\"" + dec32(pf + 128) + "\",\"" + dec32(pf + 160) +
                  "\"],[\"1\",\"0\"]],\"pi_c\":[\"" + dec32(pf + 192) + "\",\"" + dec32(pf + 224) + "\",\"1\"]";
  if (snarkjs_style) s += ",\"protocol\":\"groth16\",\"curve\":\"bn128\"";
  return s + "}";
}

std::string publics_to_json(const uint8_t *pub, uint32_t n) {
  std::string s = "[";
  for (uint32_t i = 0; i < n; i++) {
    if (i) s += ",";
    s += "\"" + dec32(pub + 32 * i) + "\"";
  }
  return s + "]";
}

}  // namespace zkb
