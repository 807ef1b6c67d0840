// Reads what the GPU witness kernel needs out of a circom-2 witness-calculator wasm, the same file
// the reference hands to its wasm runtime (`zk_census_test.go:83,89`, `ts_inputs/src/example.ts:360`):
//   * witness size / input size / prime (exported getters)
//   * the witness->signal table (data segment of 4 * nWitness bytes)
//   * the input-name hashmap (FNV-1a-64 of the signal name -> first signal, length)
//   * circomlib's Poseidon constants C/S/M/P for t = 3,4,5, recovered from the bodies of the
//     generated POSEIDON_C/S/M/P functions and the Fr constant table they copy from
// Nothing is executed: this is a structural read of an MVP wasm.  A wasm that
// is not the census circuit is rejected with ZKB_UNSUPPORTED_CIRCUIT by the caller.
#include "wasm_circuit.h"
#include <cstring>
#include <map>

namespace zkb {

namespace {

struct Rd {
  const uint8_t *b;
  size_t p, end;
  bool ok = true;
  uint8_t u8() { if (p >= end) { ok = false; return 0; } return b[p++]; }
  uint64_t leb() {
    uint64_t r = 0;
    int s = 0;
    for (;;) {
      uint8_t x = u8();
      r |= (uint64_t)(x & 0x7f) << s;
      s += 7;
      if (!(x & 0x80) || !ok || s > 70) return r;
    }
  }
  int64_t sleb() {
    int64_t r = 0;
    int s = 0;
    for (;;) {
      uint8_t x = u8();
      r |= (int64_t)(x & 0x7f) << s;
      s += 7;
      if (!(x & 0x80) || !ok || s > 70) {
        if ((x & 0x40) && s < 64) r -= (int64_t)1 << s;
        return r;
      }
    }
  }
  std::string str() {
    size_t l = (size_t)leb();
    if (p + l > end) { ok = false; return ""; }
    std::string s((const char *)b + p, l);
    p += l;
    return s;
  }
};

struct Func { size_t start, end; };
struct Ins { uint8_t op; int64_t imm; };

// decode the instruction stream of one function (only what circom emits)
bool decode(const uint8_t *b, const Func &f, std::vector<Ins> &out) {
  Rd r{b, f.start, f.end};
  // skip locals
  uint64_t nl = r.leb();
  for (uint64_t i = 0; i < nl; i++) { r.leb(); r.u8(); }
  while (r.p < r.end && r.ok) {
    uint8_t op = r.u8();
    int64_t imm = 0;
    if (op == 0x02 || op == 0x03 || op == 0x04) r.u8();
    else if (op == 0x0c || op == 0x0d || op == 0x10 || (op >= 0x20 && op <= 0x24)) imm = (int64_t)r.leb();
    else if (op == 0x0e) { uint64_t n = r.leb(); for (uint64_t i = 0; i <= n; i++) r.leb(); }
    else if (op == 0x11) { r.leb(); r.leb(); }
    else if (op >= 0x28 && op <= 0x3e) { r.leb(); imm = (int64_t)r.leb(); }
    else if (op == 0x3f || op == 0x40) r.u8();
    else if (op == 0x41 || op == 0x42) imm = r.sleb();
    else if (op == 0x43) r.p += 4;
    else if (op == 0x44) r.p += 8;
    out.push_back({op, imm});
  }
  return r.ok;
}

uint64_t fnv1a64(const std::string &s) {
  uint64_t h = 0xCBF29CE484222325ull;
  for (unsigned char c : s) { h ^= c; h *= 0x100000001B3ull; }
  return h;
}

}  // namespace

bool parse_circom_wasm(const uint8_t *buf, size_t len, WasmCircuit &out, std::string &err) {
  if (len < 8 || memcmp(buf, "\0asm\x01\0\0\0", 8) != 0) { err = "not a wasm v1 module"; return false; }
  Rd r{buf, 8, len};
  uint32_t n_imports = 0;
  std::vector<Func> funcs;
  std::map<std::string, uint32_t> exports;
  std::map<uint32_t, std::string> names;
  struct Seg { uint32_t off; const uint8_t *p; size_t len; };
  std::vector<Seg> segs;
  while (r.p < len && r.ok) {
    uint8_t id = r.u8();
    size_t size = (size_t)r.leb();
    size_t end = r.p + size;
    if (end > len) { err = "truncated section"; return false; }
    Rd s{buf, r.p, end};
    if (id == 2) {
      uint64_t n = s.leb();
      for (uint64_t i = 0; i < n; i++) { s.str(); s.str(); if (s.u8() != 0) { err = "non-function import"; return false; } s.leb(); }
      n_imports = (uint32_t)n;
    } else if (id == 7) {
      uint64_t n = s.leb();
      for (uint64_t i = 0; i < n; i++) { std::string nm = s.str(); uint8_t kind = s.u8(); uint32_t idx = (uint32_t)s.leb(); if (kind == 0) exports[nm] = idx; }
    } else if (id == 10) {
      uint64_t n = s.leb();
      for (uint64_t i = 0; i < n; i++) { size_t bs = (size_t)s.leb(); funcs.push_back({s.p, s.p + bs}); s.p += bs; }
    } else if (id == 11) {
      uint64_t n = s.leb();
      for (uint64_t i = 0; i < n; i++) {
        if (s.leb() != 0 || s.u8() != 0x41) { err = "unsupported data segment"; return false; }
        int64_t off = s.sleb();
        if (s.u8() != 0x0b) { err = "unsupported data segment"; return false; }
        size_t l = (size_t)s.leb();
        if (s.p + l > end) { err = "truncated data segment"; return false; }
        segs.push_back({(uint32_t)off, buf + s.p, l});
        s.p += l;
      }
    } else if (id == 0) {
      if (s.str() == "name") {
        while (s.p < end && s.ok) {
          uint8_t sub = s.u8();
          size_t ssz = (size_t)s.leb(), send = s.p + ssz;
          if (sub == 1) {
            uint64_t n = s.leb();
            for (uint64_t i = 0; i < n; i++) { uint32_t fi = (uint32_t)s.leb(); names[fi] = s.str(); }
          }
          s.p = send;
        }
      }
    }
    if (!s.ok) { err = "malformed section"; return false; }
    r.p = end;
  }
  auto getter = [&](const char *nm, int64_t &v) -> bool {
    auto it = exports.find(nm);
    if (it == exports.end() || it->second < n_imports || it->second - n_imports >= funcs.size()) return false;
    std::vector<Ins> ins;
    if (!decode(buf, funcs[it->second - n_imports], ins)) return false;
    for (auto &i : ins) if (i.op == 0x41) { v = i.imm; return true; }
    return false;
  };
  int64_t nwit, ninp, n32;
  if (!getter("getWitnessSize", nwit) || !getter("getInputSize", ninp) || !getter("getFieldNumLen32", n32)) {
    err = "not a circom 2 witness calculator (getters missing)";
    return false;
  }
  if (n32 != 8) { err = "field is not 8 x 32 bits"; return false; }
  out.n_wires = (uint32_t)nwit;
  out.n_inputs = (uint32_t)ninp;
  // linear-memory image
  size_t top = 0;
  for (auto &s : segs) top = std::max(top, (size_t)s.off + s.len);
  std::vector<uint8_t> mem(top, 0);
  const Seg *wm = nullptr, *hm = nullptr;
  for (auto &s : segs) {
    memcpy(mem.data() + s.off, s.p, s.len);
    if (s.len == (size_t)4 * out.n_wires) wm = &s;
    if (s.len == 4096) hm = &s;
  }
  if (!wm || !hm) { err = "witness map / input hashmap segments not found"; return false; }
  out.witness_map.resize(out.n_wires);
  memcpy(out.witness_map.data(), wm->p, wm->len);
  for (uint32_t i = 1; i < out.n_wires; i++)
    if (out.witness_map[i] <= out.witness_map[i - 1]) { err = "witness map not increasing"; return false; }
  out.hashmap.resize(256);
  for (int i = 0; i < 256; i++) {
    memcpy(&out.hashmap[i].hash, hm->p + 16 * i, 8);
    memcpy(&out.hashmap[i].pos, hm->p + 16 * i + 8, 4);
    memcpy(&out.hashmap[i].size, hm->p + 16 * i + 12, 4);
  }
  // prime: getRawPrime copies from a fixed address; check that r appears in the image's constant block
  static const uint32_t RMOD[8] = {0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u,
                                   0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
  bool prime_ok = false;
  for (size_t a = 0; a + 32 <= mem.size() && a < 8192 && !prime_ok; a += 8)
    prime_ok = memcmp(mem.data() + a, RMOD, 32) == 0;
  if (!prime_ok) { err = "wasm prime is not the BN254 scalar field"; return false; }

  // Poseidon tables
  uint32_t fr_eq = ~0u, fr_copy = ~0u;
  std::map<char, uint32_t> pfun;
  for (auto &kv : names) {
    if (kv.second == "Fr_eq") fr_eq = kv.first;
    else if (kv.second == "Fr_copy") fr_copy = kv.first;
    else if (kv.second.rfind("POSEIDON_", 0) == 0 && kv.second.size() > 10 && kv.second[10] == '_') pfun[kv.second[9]] = kv.first;
  }
  if (fr_eq == ~0u || fr_copy == ~0u || pfun.size() != 4) { err = "POSEIDON_* functions not found (name section needed)"; return false; }
  auto fr_value = [&](uint32_t addr, PoseidonConst &c) -> bool {
    if ((size_t)addr + 40 > mem.size()) return false;
    int32_t sv;
    uint32_t fl;
    memcpy(&sv, mem.data() + addr, 4);
    memcpy(&fl, mem.data() + addr + 4, 4);
    if (fl & 0x80000000u) {
      memcpy(c.v, mem.data() + addr + 8, 32);
      c.form = (fl & 0x40000000u) ? 1 : 0;
    } else {
      // short value: canonical small integer (negative = r - |v|, handled on device as neg)
      memset(c.v, 0, 32);
      c.form = sv < 0 ? 2 : 0;
      c.v[0] = (uint32_t)(sv < 0 ? -(int64_t)sv : sv);
    }
    return true;
  };
  for (int ti = 0; ti < 3; ti++) out.poseidon[ti] = PoseidonTables();
  const char keys[4] = {'C', 'S', 'M', 'P'};
  for (int ki = 0; ki < 4; ki++) {
    uint32_t fi = pfun[keys[ki]];
    if (fi < n_imports || fi - n_imports >= funcs.size()) { err = "bad function index"; return false; }
    std::vector<Ins> ins;
    if (!decode(buf, funcs[fi - n_imports], ins)) { err = "cannot decode POSEIDON function"; return false; }
    int64_t cur_t = -1;
    for (size_t i = 0; i < ins.size(); i++) {
      if (ins[i].op != 0x10) continue;
      if ((uint32_t)ins[i].imm == fr_eq && i >= 1 && ins[i - 1].op == 0x41) {
        PoseidonConst c;
        if (!fr_value((uint32_t)ins[i - 1].imm, c)) { err = "bad constant address"; return false; }
        // t is a small integer constant; recover its value
        if (c.form == 1) { err = "unexpected Montgomery-only t constant"; return false; }
        cur_t = c.v[0];
      } else if ((uint32_t)ins[i].imm == fr_copy && cur_t >= 3 && cur_t <= 5 && i >= 6 && ins[i - 1].op == 0x41 &&
                 ins[i - 2].op == 0x6a && ins[i - 3].op == 0x20 && ins[i - 4].op == 0x6c && ins[i - 5].op == 0x41 &&
                 ins[i - 5].imm == 40 && ins[i - 6].op == 0x41) {
        size_t k = (size_t)ins[i - 6].imm - 1;
        std::vector<PoseidonConst> &v = out.poseidon[cur_t - 3].tab[ki];
        if (v.size() <= k) v.resize(k + 1);
        if (!fr_value((uint32_t)ins[i - 1].imm, v[k])) { err = "bad constant address"; return false; }
      }
    }
  }
  static const int RP[3] = {57, 56, 60};
  for (int ti = 0; ti < 3; ti++) {
    int t = 3 + ti;
    const PoseidonTables &pt = out.poseidon[ti];
    if ((int)pt.tab[0].size() != 8 * t + RP[ti] || (int)pt.tab[1].size() != RP[ti] * (2 * t - 1) ||
        (int)pt.tab[2].size() != t * t || (int)pt.tab[3].size() != t * t) {
      err = "Poseidon table sizes do not match circomlib's t=" + std::to_string(t);
      return false;
    }
  }
  return true;
}

bool wasm_input_lookup(const WasmCircuit &w, const std::string &name, uint32_t &pos, uint32_t &size) {
  uint64_t h = fnv1a64(name);
  uint32_t slot = (uint32_t)(h & 0xff);
  for (int i = 0; i < 256; i++) {
    const WasmInputSlot &s = w.hashmap[slot];
    if (s.hash == h) { pos = s.pos; size = s.size; return true; }
    if (s.hash == 0 && s.size == 0) return false;
    slot = (slot + 1) & 0xff;
  }
  return false;
}

}  // namespace zkb
