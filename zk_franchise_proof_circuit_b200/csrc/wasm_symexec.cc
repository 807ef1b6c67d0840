// Generic witness-program extraction from a circom-2 witness calculator.
//
// The reference hands `circuit.wasm` to a wasm runtime and drives it with the circom_runtime protocol
// (`zk_census_test.go:83,89`; `ts_inputs/src/example.ts:360`).  Here the same file is executed ONCE,
// at load time, by a small MVP-wasm interpreter in which the inputs are symbols:
//   * integer and control state (component tables, input counters, loop variables, the Fr stack pointer) runs
//     concretely, exactly as in a wasm runtime;
//   * every 40-byte field-element record in linear memory is either concrete (its bytes are valid) or symbolic (a
//     value id in a shadow map).  Calls of the circom field runtime (Fr_add, Fr_mul, Fr_lt, Fr_shr, ...) whose
//     operands are all concrete are executed by interpreting the wasm's own code; calls with a symbolic operand are
//     not executed but appended to a tape (TapeOp) that defines a new symbolic value;
//   * Fr_isTrue of a symbolic value yields a tagged integer.  An `if` on such a tag is either a circuit assert (one
//     arm reports an error through the runtime imports: recorded as T_ASSERT_*, SURVEY 8b "Errors") or a two-armed
//     conditional assignment (`inv <-- in != 0 ? 1/in : 0`): both arms are executed against a journal and the field
//     records they wrote are merged with T_SELECT.  Any other use of a tag (arithmetic, addressing, loop bounds) is
//     outside the supported subset and reported as such.
// The tape is then pruned to what the witness needs, ordered by dependency level and given reusable value slots;
// tape_eval.cu runs it for a batch of proofs (one warp per proof, the lanes share each level).
// Nothing here is census-specific: the hand-written census kernel stays the fast path for that program.
#include "wasm_symexec.h"
#include <algorithm>
#include <array>
#include <cstring>
#include <map>
#include <unordered_map>
#include "fp.cuh"

namespace zkb {

uint64_t fnv1a64_name(const std::string &s) {
  uint64_t h = 0xCBF29CE484222325ull;
  for (unsigned char c : s) { h ^= c; h *= 0x100000001B3ull; }
  return h;
}

namespace {

struct Unsupported { std::string why; };
struct ProbeDone {};

struct FuncType { std::vector<uint8_t> params, results; };
struct Ins { uint8_t op; uint32_t aux; int64_t imm; };       // aux: matching end (block/loop/if) or memarg offset
struct Func {
  uint32_t type = 0;
  size_t start = 0, end = 0;
  std::vector<uint8_t> local_types;      // params + locals
  std::vector<Ins> code;
  bool decoded = false;
};
struct Val { uint64_t v; uint32_t tag; };                   // tag: 0 = plain, else ((value ref) << 1 | negated) + 2
struct CtrlFrame { uint32_t open; uint32_t height; };
struct Frame {
  std::vector<Val> locals, stack;
  std::vector<CtrlFrame> ctrl;
};

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
    if (!ok || p > end || l > end - p) { ok = false; return ""; }
    std::string s((const char *)b + p, l);
    p += l;
    return s;
  }
};

enum RtKind { RT_NONE = 0, RT_BIN, RT_UN, RT_COPY, RT_COPYN, RT_ISTRUE, RT_TOINT };
struct RtInfo { RtKind kind = RT_NONE; uint8_t op = 0; };

struct RecState { bool is_sym; uint32_t ref; uint8_t bytes[40]; };
struct RawOld { uint32_t addr; uint8_t width; uint64_t old; };
struct Journal {
  std::map<uint32_t, RecState> rec_old;     // first-touch state of field records
  std::vector<RawOld> raw_old;              // every raw store, in order
};

static const uint32_t RMOD[8] = {0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u,
                                 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};

// census.circom at 160 levels takes 2.7e8 interpreted instructions; a program that needs 100x that is not a witness
// calculator this extractor can serve (and a hostile wasm must not hang zkb_load_circuit)
static const uint64_t kMaxSteps = 30000000000ull;

struct Machine {
  const uint8_t *bin = nullptr;
  size_t bin_len = 0;
  std::vector<FuncType> types;
  uint32_t n_imports = 0;
  std::vector<std::string> import_names;
  std::vector<uint32_t> import_types;
  std::vector<Func> funcs;                 // defined functions (index = func index - n_imports)
  std::vector<uint32_t> table;
  std::map<std::string, uint32_t> exports;
  std::map<uint32_t, std::string> names;
  std::vector<uint8_t> mem;
  std::vector<RtInfo> rt;                  // per function index
  std::vector<uint8_t> is_run_fn;          // per function index: name ends in "_run"

  // symbolic state
  std::unordered_map<uint32_t, uint32_t> sym;       // record address -> operand ref
  std::vector<uint8_t> symbits;                     // one bit per 8 bytes of memory covered by a symbolic record
  std::vector<Journal> journals;
  std::vector<TapeOp> ssa;                          // dst = value id << 1
  uint32_t next_value = 0;
  std::vector<uint32_t> consts;
  std::map<std::array<uint32_t, 8>, uint32_t> const_index;
  bool inputs_marked = false, probing = false;
  uint32_t probe_src = 0;
  std::vector<uint32_t> input_signal_addr;          // record address of every input element, in signal order
  uint64_t steps = 0;
  uint32_t n_asserts = 0, n_selects = 0;
  int depth = 0;

  // ---- module ------------------------------------------------------------------------------------
  bool parse(std::string &err) {
    if (bin_len < 8 || memcmp(bin, "\0asm\x01\0\0\0", 8) != 0) { err = "not a wasm v1 module"; return false; }
    Rd r{bin, 8, bin_len};
    std::vector<uint32_t> ftypes;
    uint32_t mem_pages = 0;
    struct Seg { uint32_t off; const uint8_t *p; size_t len; };
    std::vector<Seg> segs;
    while (r.p < bin_len && r.ok) {
      uint8_t id = r.u8();
      const size_t size = (size_t)r.leb();
      if (!r.ok || size > bin_len - r.p) { err = "truncated section"; return false; }
      const size_t end = r.p + size;
      Rd s{bin, r.p, end};
      // every vector of a section has at least one byte per element: a count above the section size is damage
      auto count = [&]() -> uint64_t { uint64_t n = s.leb(); if (n > size) { s.ok = false; return 0; } return n; };
      if (id == 1) {
        uint64_t n = count();
        for (uint64_t i = 0; i < n; i++) {
          if (s.u8() != 0x60) { err = "bad type section"; return false; }
          FuncType t;
          uint64_t np = count();
          for (uint64_t k = 0; k < np; k++) t.params.push_back(s.u8());
          uint64_t nr = count();
          for (uint64_t k = 0; k < nr; k++) t.results.push_back(s.u8());
          types.push_back(t);
        }
      } else if (id == 2) {
        uint64_t n = count();
        for (uint64_t i = 0; i < n; i++) {
          s.str();
          std::string nm = s.str();
          if (s.u8() != 0) { err = "non-function import"; return false; }
          import_types.push_back((uint32_t)s.leb());
          import_names.push_back(nm);
        }
        n_imports = (uint32_t)n;
      } else if (id == 3) {
        uint64_t n = count();
        for (uint64_t i = 0; i < n; i++) ftypes.push_back((uint32_t)s.leb());
      } else if (id == 5) {
        uint64_t n = s.leb();
        if (n != 1) { err = "expected one memory"; return false; }
        uint8_t fl = s.u8();
        mem_pages = (uint32_t)s.leb();
        if (fl & 1) s.leb();
      } else if (id == 7) {
        uint64_t n = count();
        for (uint64_t i = 0; i < n; i++) {
          std::string nm = s.str();
          uint8_t kind = s.u8();
          uint32_t idx = (uint32_t)s.leb();
          if (kind == 0) exports[nm] = idx;
        }
      } else if (id == 9) {
        uint64_t n = count();
        for (uint64_t i = 0; i < n; i++) {
          if (s.leb() != 0 || s.u8() != 0x41) { err = "unsupported element segment"; return false; }
          int64_t off = s.sleb();
          if (s.u8() != 0x0b) { err = "unsupported element segment"; return false; }
          uint64_t cnt = count();
          if (!s.ok || off < 0 || off > (1 << 20) || cnt > (1u << 20)) { err = "implausible element segment"; return false; }
          if (table.size() < (size_t)off + cnt) table.resize((size_t)off + cnt, ~0u);
          for (uint64_t k = 0; k < cnt; k++) table[(size_t)off + k] = (uint32_t)s.leb();
        }
      } else if (id == 10) {
        uint64_t n = count();
        if (n != ftypes.size()) { err = "function / code section mismatch"; return false; }
        for (uint64_t i = 0; i < n; i++) {
          size_t bs = (size_t)s.leb();
          if (!s.ok || bs > end - s.p) { err = "function body outside the code section"; return false; }
          Func f;
          f.type = ftypes[i];
          f.start = s.p;
          f.end = s.p + bs;
          funcs.push_back(f);
          s.p += bs;
        }
      } else if (id == 11) {
        uint64_t n = count();
        for (uint64_t i = 0; i < n; i++) {
          if (s.leb() != 0 || s.u8() != 0x41) { err = "unsupported data segment"; return false; }
          int64_t off = s.sleb();
          if (s.u8() != 0x0b) { err = "unsupported data segment"; return false; }
          size_t l = (size_t)s.leb();
          if (!s.ok || l > end - s.p) { err = "truncated data segment"; return false; }
          segs.push_back({(uint32_t)off, bin + s.p, l});
          s.p += l;
        }
      } else if (id == 0) {
        if (s.str() == "name") {
          while (s.p < end && s.ok) {
            uint8_t sub = s.u8();
            size_t ssz = (size_t)s.leb(), send = s.p + ssz;
            if (ssz > end - s.p) { s.ok = false; break; }
            if (sub == 1) {
              uint64_t n = count();
              for (uint64_t i = 0; i < n && s.ok; i++) { uint32_t fi = (uint32_t)s.leb(); names[fi] = s.str(); }
            }
            s.p = send;
          }
        }
      }
      if (!s.ok) { err = "malformed section"; return false; }
      r.p = end;
    }
    for (uint32_t t : import_types) if (t >= types.size()) { err = "import with an undefined type"; return false; }
    for (auto &f : funcs) if (f.type >= types.size()) { err = "function with an undefined type"; return false; }
    if (!mem_pages || mem_pages > 16384) { err = "memory size"; return false; }
    mem.assign((size_t)mem_pages << 16, 0);
    symbits.assign(mem.size() / 64 + 8, 0);
    for (auto &sg : segs) {
      if ((size_t)sg.off + sg.len > mem.size()) { err = "data segment outside memory"; return false; }
      memcpy(mem.data() + sg.off, sg.p, sg.len);
    }
    // the circom field runtime, by name
    static const struct { const char *name; RtKind kind; uint8_t op; } RT[] = {
        {"Fr_add", RT_BIN, T_ADD}, {"Fr_sub", RT_BIN, T_SUB}, {"Fr_mul", RT_BIN, T_MUL}, {"Fr_div", RT_BIN, T_DIV},
        {"Fr_eq", RT_BIN, T_EQ}, {"Fr_neq", RT_BIN, T_NEQ}, {"Fr_lt", RT_BIN, T_LT}, {"Fr_gt", RT_BIN, T_GT},
        {"Fr_leq", RT_BIN, T_LEQ}, {"Fr_geq", RT_BIN, T_GEQ}, {"Fr_land", RT_BIN, T_LAND}, {"Fr_lor", RT_BIN, T_LOR},
        {"Fr_shr", RT_BIN, T_SHR}, {"Fr_shl", RT_BIN, T_SHL}, {"Fr_band", RT_BIN, T_BAND}, {"Fr_bor", RT_BIN, T_BOR},
        {"Fr_bxor", RT_BIN, T_BXOR}, {"Fr_idiv", RT_BIN, T_IDIV}, {"Fr_mod", RT_BIN, T_MOD}, {"Fr_pow", RT_BIN, T_POW},
        {"Fr_neg", RT_UN, T_NEG}, {"Fr_inv", RT_UN, T_INV}, {"Fr_lnot", RT_UN, T_LNOT}, {"Fr_bnot", RT_UN, T_BNOT},
        {"Fr_copy", RT_COPY, 0}, {"Fr_copyn", RT_COPYN, 0}, {"Fr_isTrue", RT_ISTRUE, 0}, {"Fr_toInt", RT_TOINT, 0}};
    rt.assign(n_imports + funcs.size(), RtInfo());
    is_run_fn.assign(n_imports + funcs.size(), 0);
    int found = 0;
    for (auto &kv : names) {
      if (kv.first >= rt.size()) continue;
      for (auto &e : RT)
        if (kv.second == e.name) { rt[kv.first].kind = e.kind; rt[kv.first].op = e.op; found++; }
      const std::string &nm = kv.second;
      if (nm.size() > 4 && nm.compare(nm.size() - 4, 4, "_run") == 0) is_run_fn[kv.first] = 1;
    }
    if (found < 20) { err = "circom field runtime not found by name (the wasm has no name section?)"; return false; }
    return true;
  }

  bool decode(Func &f, std::string &err) {
    Rd r{bin, f.start, f.end};
    const FuncType &t = types[f.type];
    f.local_types = t.params;
    uint64_t nl = r.leb();
    for (uint64_t i = 0; i < nl; i++) {
      uint64_t cnt = r.leb();
      uint8_t ty = r.u8();
      if (f.local_types.size() + cnt > 100000) { err = "too many locals"; return false; }
      f.local_types.insert(f.local_types.end(), (size_t)cnt, ty);
    }
    std::vector<uint32_t> open;
    while (r.p < r.end && r.ok) {
      Ins in{r.u8(), 0, 0};
      const uint8_t op = in.op;
      if (op == 0x02 || op == 0x03 || op == 0x04) {
        if (r.u8() != 0x40) { err = "non-void block type"; return false; }
        in.imm = -1;
        open.push_back((uint32_t)f.code.size());
      } else if (op == 0x05) {
        if (open.empty() || f.code[open.back()].op != 0x04) { err = "else without if"; return false; }
        f.code[open.back()].imm = (int64_t)f.code.size();
      } else if (op == 0x0b) {
        if (!open.empty()) {
          f.code[open.back()].aux = (uint32_t)f.code.size();
          open.pop_back();
        }
      } else if (op == 0x0c || op == 0x0d || op == 0x10 || (op >= 0x20 && op <= 0x22)) {
        in.imm = (int64_t)r.leb();
      } else if (op == 0x11) {
        in.imm = (int64_t)r.leb();
        r.leb();
      } else if (op >= 0x28 && op <= 0x3e) {
        r.leb();
        in.aux = (uint32_t)r.leb();
      } else if (op == 0x3f || op == 0x40) {
        r.u8();
      } else if (op == 0x41 || op == 0x42) {
        in.imm = r.sleb();
      } else if (op == 0x0e || op == 0x1b || op == 0x23 || op == 0x24 || op == 0x43 || op == 0x44 || op >= 0xb2) {
        err = "unsupported opcode " + std::to_string(op);
        return false;
      }
      f.code.push_back(in);
    }
    if (!r.ok || !open.empty()) { err = "malformed function body"; return false; }
    f.decoded = true;
    return true;
  }

  // ---- constants and records ----------------------------------------------------------------------
  uint32_t const_ref(const uint32_t v[8]) {
    std::array<uint32_t, 8> k;
    memcpy(k.data(), v, 32);
    auto it = const_index.find(k);
    if (it != const_index.end()) return (it->second << 1) | 1u;
    uint32_t idx = (uint32_t)(consts.size() / 8);
    consts.insert(consts.end(), v, v + 8);
    const_index[k] = idx;
    return (idx << 1) | 1u;
  }
  void check_addr(uint64_t a, size_t n) const {
    if (a + n > mem.size()) throw Unsupported{"memory access out of bounds"};
  }
  // canonical value of the concrete record at addr
  void record_value(uint32_t addr, uint32_t out[8]) const {
    check_addr(addr, 40);
    int32_t sv;
    uint32_t fl;
    memcpy(&sv, mem.data() + addr, 4);
    memcpy(&fl, mem.data() + addr + 4, 4);
    if (fl & 0x80000000u) {
      Fr x;
      memcpy(x.v, mem.data() + addr + 8, 32);
      if (fl & 0x40000000u) x = x.from_mont();
      memcpy(out, x.v, 32);
    } else if (sv >= 0) {
      memset(out, 0, 32);
      out[0] = (uint32_t)sv;
    } else {
      uint64_t sub = (uint64_t)(-(int64_t)sv);                  // r - |v|, |v| < 2^32
      for (int i = 0; i < 8; i++) {
        const uint64_t cur = RMOD[i];
        if (cur >= sub) { out[i] = (uint32_t)(cur - sub); sub = 0; }
        else { out[i] = (uint32_t)(cur + (1ull << 32) - sub); sub = 1; }
      }
    }
  }
  bool is_sym(uint32_t addr) const { return sym.find(addr) != sym.end(); }
  uint32_t operand_ref(uint32_t addr) {
    auto it = sym.find(addr);
    if (it != sym.end()) return it->second;
    uint32_t v[8];
    record_value(addr, v);
    return const_ref(v);
  }
  // symbits: one bit per 8-byte word that belongs to a symbolic record (records are 8-byte aligned)
  void set_bits(uint32_t addr, bool on) {
    if (addr & 7u) throw Unsupported{"field record not 8-byte aligned"};
    for (uint32_t w = addr >> 3; w < (addr >> 3) + 5; w++) {
      if (on) symbits[w >> 3] |= (uint8_t)(1u << (w & 7));
      else symbits[w >> 3] &= (uint8_t)~(1u << (w & 7));
    }
  }
  bool word_is_sym(uint64_t addr, int width) const {
    for (uint64_t w = addr >> 3; w <= (addr + width - 1) >> 3; w++)
      if (symbits[w >> 3] & (1u << (w & 7))) return true;
    return false;
  }
  void sym_set(uint32_t addr, uint32_t ref) {
    if (!sym.count(addr)) set_bits(addr, true);
    sym[addr] = ref;
  }
  void sym_erase(uint32_t addr) {
    if (sym.erase(addr)) set_bits(addr, false);
  }
  void touch_record(uint32_t addr) {
    if (journals.empty()) return;
    Journal &j = journals.back();
    if (j.rec_old.count(addr)) return;
    RecState s;
    auto it = sym.find(addr);
    s.is_sym = it != sym.end();
    s.ref = s.is_sym ? it->second : 0;
    check_addr(addr, 40);
    memcpy(s.bytes, mem.data() + addr, 40);
    j.rec_old[addr] = s;
  }
  void set_sym(uint32_t addr, uint32_t ref) {
    check_addr(addr, 40);
    touch_record(addr);
    sym_set(addr, ref);
  }
  void clear_sym(uint32_t addr) {
    touch_record(addr);
    sym_erase(addr);
  }
  RecState state_of(uint32_t addr) {
    RecState s;
    auto it = sym.find(addr);
    s.is_sym = it != sym.end();
    s.ref = s.is_sym ? it->second : 0;
    memcpy(s.bytes, mem.data() + addr, 40);
    return s;
  }
  void restore(uint32_t addr, const RecState &s) {
    memcpy(mem.data() + addr, s.bytes, 40);
    if (s.is_sym) sym_set(addr, s.ref); else sym_erase(addr);
  }
  uint32_t emit(uint8_t op, uint32_t a, uint32_t b = 0, uint32_t c = 0) {
    uint32_t id = next_value++;
    ssa.push_back({op, id << 1, a, b, c});
    return id << 1;
  }

  // ---- memory access from interpreted code --------------------------------------------------------
  uint64_t load(uint64_t addr, int width) {
    check_addr(addr, width);
    if (word_is_sym(addr, width)) throw Unsupported{"integer read of a field element that depends on an input"};
    uint64_t v = 0;
    memcpy(&v, mem.data() + addr, width);
    return v;
  }
  void store(uint64_t addr, int width, uint64_t v) {
    check_addr(addr, width);
    if (word_is_sym(addr, width)) throw Unsupported{"integer write into a field element that depends on an input"};
    if (!journals.empty()) {
      uint64_t old = 0;
      memcpy(&old, mem.data() + addr, width);
      journals.back().raw_old.push_back({(uint32_t)addr, (uint8_t)width, old});
    }
    memcpy(mem.data() + addr, &v, width);
  }

  // ---- the field runtime ------------------------------------------------------------------------------
  // returns true when the call was handled symbolically (result pushed if any)
  bool intercept(uint32_t fi, std::vector<Val> &args, Val &result, bool &has_result) {
    const RtInfo &info = rt[fi];
    has_result = false;
    for (auto &a : args)
      if (a.tag) throw Unsupported{"a value derived from a signal is passed to a function as an integer"};
    static const size_t need[] = {0, 3, 2, 2, 3, 1, 1};      // RT_NONE, RT_BIN, RT_UN, RT_COPY, RT_COPYN, RT_ISTRUE, RT_TOINT
    if (args.size() < need[info.kind]) throw Unsupported{"a function of the field runtime has an unexpected signature: " + names[fi]};
    switch (info.kind) {
      case RT_BIN: {
        uint32_t dst = (uint32_t)args[0].v, a = (uint32_t)args[1].v, b = (uint32_t)args[2].v;
        if (!is_sym(a) && !is_sym(b)) { clear_sym(dst); return false; }
        uint32_t ra = operand_ref(a), rb = operand_ref(b);
        set_sym(dst, emit(info.op, ra, rb));
        return true;
      }
      case RT_UN: {
        uint32_t dst = (uint32_t)args[0].v, a = (uint32_t)args[1].v;
        if (!is_sym(a)) { clear_sym(dst); return false; }
        set_sym(dst, emit(info.op, operand_ref(a)));
        return true;
      }
      case RT_COPY: {
        uint32_t dst = (uint32_t)args[0].v, src = (uint32_t)args[1].v;
        if (probing) { probe_src = src; throw ProbeDone{}; }
        copy_record(dst, src);
        return true;
      }
      case RT_COPYN: {
        uint32_t dst = (uint32_t)args[0].v, src = (uint32_t)args[1].v, n = (uint32_t)args[2].v;
        for (uint32_t i = 0; i < n; i++) copy_record(dst + 40 * i, src + 40 * i);
        return true;
      }
      case RT_ISTRUE: {
        uint32_t a = (uint32_t)args[0].v;
        if (!is_sym(a)) return false;
        uint32_t bref = emit(T_ISTRUE, operand_ref(a));
        result = {0, (bref << 1) + 2};
        has_result = true;
        return true;
      }
      case RT_TOINT: {
        if (is_sym((uint32_t)args[0].v)) throw Unsupported{"a signal value is used as an integer (array index or loop bound)"};
        return false;
      }
      default: return false;
    }
  }
  void copy_record(uint32_t dst, uint32_t src) {
    check_addr(dst, 40);
    check_addr(src, 40);
    auto it = sym.find(src);
    if (it != sym.end()) { set_sym(dst, it->second); return; }
    clear_sym(dst);
    if (!journals.empty()) {
      // the bytes are restored from rec_old on rollback
    }
    memmove(mem.data() + dst, mem.data() + src, 40);
  }

  // ---- interpreter ------------------------------------------------------------------------------------
  Func &func(uint32_t fi) {
    Func &f = funcs[fi - n_imports];
    if (!f.decoded) {
      std::string err;
      if (!decode(f, err)) throw Unsupported{err};
    }
    return f;
  }
  bool arm_reports_error(const Func &f, uint32_t lo, uint32_t hi) const {
    for (uint32_t i = lo; i < hi; i++)
      if (f.code[i].op == 0x10 && (uint32_t)f.code[i].imm < n_imports) return true;
    return false;
  }
  void mark_inputs() {
    for (size_t i = 0; i < input_signal_addr.size(); i++) set_sym(input_signal_addr[i], (uint32_t)i << 1);
    inputs_marked = true;
  }

  // calls function fi; returns its result (if any)
  bool call(uint32_t fi, std::vector<Val> &args, Val &result) {
    if (fi < n_imports) {
      const std::string &nm = import_names[fi];
      if (nm == "exceptionHandler") throw Unsupported{"the program raised exception " + std::to_string((uint32_t)args[0].v) + " on constant data"};
      return false;   // printErrorMessage, writeBufferMessage, showSharedRWMemory: no effect on the witness
    }
    if (fi - n_imports >= funcs.size()) throw Unsupported{"call of an undefined function"};
    if (rt[fi].kind != RT_NONE) {
      bool has = false;
      if (intercept(fi, args, result, has)) return has;
    }
    if (is_run_fn[fi] && !inputs_marked && !probing) mark_inputs();
    if (++depth > 2000) throw Unsupported{"call depth"};
    Func &f = func(fi);
    const FuncType &t = types[f.type];
    Frame fr;
    fr.locals.assign(f.local_types.size(), Val{0, 0});
    for (size_t i = 0; i < args.size(); i++) fr.locals[i] = args[i];
    int ex = run(f, fr, 0, (uint32_t)f.code.size(), 0);
    depth--;
    (void)ex;
    if (!t.results.empty()) {
      if (fr.stack.empty()) throw Unsupported{"function returned no value"};
      result = fr.stack.back();
      return true;
    }
    return false;
  }

  static uint32_t tag_of(uint32_t bool_ref, bool neg) { return ((bool_ref << 1) | (neg ? 1u : 0u)) + 2; }

  // executes f.code[pc .. end) in frame fr.  base_ctrl: control frames below this index belong to the caller of this
  // range (an arm of a symbolic if may not branch out of them).  returns 0 = fell off the end / branched to `end`,
  // 1 = function return.
  int run(Func &f, Frame &fr, uint32_t pc, uint32_t end, size_t base_ctrl) {
    auto &st = fr.stack;
    auto pop = [&]() -> Val {
      if (st.empty()) throw Unsupported{"operand stack underflow"};
      Val v = st.back();
      st.pop_back();
      return v;
    };
    auto popv = [&]() -> uint64_t {
      Val v = pop();
      if (v.tag) throw Unsupported{"a value derived from a signal is used in integer arithmetic"};
      return v.v;
    };
    auto push = [&](uint64_t v) { st.push_back(Val{v, 0}); };
    auto branch = [&](uint32_t d, uint32_t &pcr) -> int {     // 0 = continue at pcr, 1 = return, 2 = left the range
      if (d >= fr.ctrl.size()) return 1;
      const size_t idx = fr.ctrl.size() - 1 - d;
      if (idx < base_ctrl) throw Unsupported{"a branch leaves a conditional whose condition depends on a signal"};
      const CtrlFrame cf = fr.ctrl[idx];
      st.resize(cf.height);
      if (f.code[cf.open].op == 0x03) { fr.ctrl.resize(idx + 1); pcr = cf.open + 1; }
      else { fr.ctrl.resize(idx); pcr = f.code[cf.open].aux + 1; }
      return 0;
    };
    while (pc < end) {
      const Ins &in = f.code[pc];
      if (++steps > kMaxSteps) throw Unsupported{"the program did not finish within the interpreter's step budget"};
      switch (in.op) {
        case 0x00: throw Unsupported{"unreachable executed"};
        case 0x01: pc++; break;
        case 0x02: case 0x03:
          fr.ctrl.push_back({pc, (uint32_t)st.size()});
          pc++;
          break;
        case 0x04: {
          Val c = pop();
          const uint32_t else_idx = in.imm >= 0 ? (uint32_t)in.imm : in.aux, end_idx = in.aux;
          if (!c.tag) {
            fr.ctrl.push_back({pc, (uint32_t)st.size()});
            if (c.v & 0xffffffffu) pc++;
            else if (in.imm >= 0) pc = else_idx + 1;
            else { fr.ctrl.pop_back(); pc = end_idx + 1; }
            break;
          }
          // ---- condition depends on a signal ----
          const uint32_t t = c.tag - 2, bref = t >> 1;
          const bool neg = t & 1;
          const bool then_err = arm_reports_error(f, pc + 1, else_idx);
          const bool else_err = in.imm >= 0 && arm_reports_error(f, else_idx + 1, end_idx);
          if (then_err != else_err) {
            // circuit assert: the error arm is never taken for a valid input; record the requirement
            const bool need_true = else_err;                 // effective condition must be true to avoid the else arm
            const bool bool_true = need_true != neg;
            ssa.push_back({(uint8_t)(bool_true ? T_ASSERT_TRUE : T_ASSERT_FALSE), 0, bref, 0, 0});
            n_asserts++;
            fr.ctrl.push_back({pc, (uint32_t)st.size()});
            if (else_err) pc++;                               // run the then arm
            else if (in.imm >= 0) pc = else_idx + 1;          // run the else arm
            else { fr.ctrl.pop_back(); pc = end_idx + 1; }
            break;
          }
          if (then_err && else_err) throw Unsupported{"both arms of a signal-dependent conditional report an error"};
          // two-armed conditional assignment: run both arms against journals, merge the field records with selects
          const size_t h = st.size();
          const std::vector<Val> locals0 = fr.locals;
          const size_t nctrl = fr.ctrl.size();
          journals.emplace_back();
          fr.ctrl.push_back({pc, (uint32_t)h});
          if (run(f, fr, pc + 1, else_idx, nctrl) == 1) throw Unsupported{"return inside a signal-dependent conditional"};
          fr.ctrl.resize(nctrl);
          st.resize(h);
          Journal j1 = std::move(journals.back());
          journals.pop_back();
          std::map<uint32_t, RecState> then_state;
          for (auto &kv : j1.rec_old) then_state[kv.first] = state_of(kv.first);
          std::vector<std::pair<RawOld, uint64_t>> then_raw;
          for (auto &ro : j1.raw_old) {
            uint64_t now = 0;
            memcpy(&now, mem.data() + ro.addr, ro.width);
            then_raw.push_back({ro, now});
          }
          const std::vector<Val> locals1 = fr.locals;
          for (size_t i = j1.raw_old.size(); i-- > 0;) memcpy(mem.data() + j1.raw_old[i].addr, &j1.raw_old[i].old, j1.raw_old[i].width);
          for (auto &kv : j1.rec_old) restore(kv.first, kv.second);
          fr.locals = locals0;
          journals.emplace_back();
          if (in.imm >= 0) {
            fr.ctrl.push_back({pc, (uint32_t)h});
            if (run(f, fr, else_idx + 1, end_idx, nctrl) == 1) throw Unsupported{"return inside a signal-dependent conditional"};
            fr.ctrl.resize(nctrl);
            st.resize(h);
          }
          Journal j2 = std::move(journals.back());
          journals.pop_back();
          // integer state must not depend on the arm taken
          for (size_t i = 0; i < fr.locals.size(); i++)
            if (fr.locals[i].v != locals1[i].v || fr.locals[i].tag != locals1[i].tag)
              throw Unsupported{"an integer variable is assigned differently in the arms of a signal-dependent conditional"};
          std::map<uint32_t, uint64_t> raw_final;   // (addr, width) keyed by addr: bytes after the then arm
          for (auto &tr : then_raw) {
            uint64_t now = 0;
            memcpy(&now, mem.data() + tr.first.addr, tr.first.width);
            // inside a record that both arms handle as a record? then it is merged below
            bool in_rec = false;
            for (auto &kv : j1.rec_old) if (tr.first.addr >= kv.first && tr.first.addr < kv.first + 40) in_rec = true;
            if (!in_rec && now != tr.second)
              throw Unsupported{"integer memory is written differently in the arms of a signal-dependent conditional"};
          }
          for (auto &ro : j2.raw_old) {
            bool in_rec = false;
            for (auto &kv : j2.rec_old) if (ro.addr >= kv.first && ro.addr < kv.first + 40) in_rec = true;
            if (in_rec) continue;
            bool in_then = false;
            for (auto &tr : then_raw) if (tr.first.addr == ro.addr) in_then = true;
            uint64_t now = 0;
            memcpy(&now, mem.data() + ro.addr, ro.width);
            if (!in_then && now != ro.old)
              throw Unsupported{"integer memory is written in only one arm of a signal-dependent conditional"};
          }
          // merge the records
          std::map<uint32_t, RecState> pre;
          for (auto &kv : j1.rec_old) pre[kv.first] = kv.second;
          for (auto &kv : j2.rec_old) if (!pre.count(kv.first)) pre[kv.first] = kv.second;
          for (auto &kv : pre) {
            const uint32_t addr = kv.first;
            const RecState tv = then_state.count(addr) ? then_state[addr] : kv.second;
            const RecState ev = state_of(addr);
            auto ref_of = [&](const RecState &s) -> uint32_t {
              if (s.is_sym) return s.ref;
              uint8_t keep[40];
              memcpy(keep, mem.data() + addr, 40);
              memcpy(mem.data() + addr, s.bytes, 40);
              uint32_t v[8];
              record_value(addr, v);
              memcpy(mem.data() + addr, keep, 40);
              return const_ref(v);
            };
            if (!tv.is_sym && !ev.is_sym && memcmp(tv.bytes, ev.bytes, 40) == 0) continue;
            if (tv.is_sym && ev.is_sym && tv.ref == ev.ref) continue;
            const uint32_t rt_ = ref_of(tv), re_ = ref_of(ev);
            const uint32_t sel = neg ? emit(T_SELECT, re_, rt_, bref) : emit(T_SELECT, rt_, re_, bref);
            n_selects++;
            sym_set(addr, sel);
          }
          // the enclosing journal (if any) must learn what this conditional touched
          if (!journals.empty()) {
            Journal &up = journals.back();
            for (auto &kv : pre) if (!up.rec_old.count(kv.first)) up.rec_old[kv.first] = kv.second;
            for (auto &ro : j1.raw_old) up.raw_old.push_back(ro);
            for (auto &ro : j2.raw_old) up.raw_old.push_back(ro);
          }
          pc = end_idx + 1;
          break;
        }
        case 0x05: {                                         // reached the end of a then arm: skip the else arm
          if (fr.ctrl.empty()) throw Unsupported{"else outside if"};
          const CtrlFrame cf = fr.ctrl.back();
          fr.ctrl.pop_back();
          pc = f.code[cf.open].aux + 1;
          break;
        }
        case 0x0b:
          if (fr.ctrl.empty()) return 0;                     // function end
          fr.ctrl.pop_back();
          pc++;
          break;
        case 0x0c: {
          uint32_t npc = pc;
          int k = branch((uint32_t)in.imm, npc);
          if (k == 1) return 1;
          pc = npc;
          break;
        }
        case 0x0d: {
          Val c = pop();
          if (c.tag) throw Unsupported{"a loop or branch condition depends on a signal value"};
          if (c.v & 0xffffffffu) {
            uint32_t npc = pc;
            int k = branch((uint32_t)in.imm, npc);
            if (k == 1) return 1;
            pc = npc;
          } else pc++;
          break;
        }
        case 0x0f: return 1;
        case 0x10: case 0x11: {
          uint32_t fi;
          if (in.op == 0x10) fi = (uint32_t)in.imm;
          else {
            uint64_t idx = popv() & 0xffffffffu;
            if (idx >= table.size() || table[idx] == ~0u) throw Unsupported{"call_indirect outside the table"};
            fi = table[idx];
          }
          if (fi >= n_imports && fi - n_imports >= funcs.size()) throw Unsupported{"call of an undefined function"};
          const FuncType &t = fi < n_imports ? types[import_types[fi]] : types[funcs[fi - n_imports].type];
          std::vector<Val> args(t.params.size());
          for (size_t i = args.size(); i-- > 0;) args[i] = pop();
          Val res{0, 0};
          if (call(fi, args, res)) st.push_back(res);
          else if (!t.results.empty()) st.push_back(Val{0, 0});
          pc++;
          break;
        }
        case 0x1a: pop(); pc++; break;
        case 0x20: st.push_back(fr.locals.at((size_t)in.imm)); pc++; break;
        case 0x21: fr.locals.at((size_t)in.imm) = pop(); pc++; break;
        case 0x22: fr.locals.at((size_t)in.imm) = st.empty() ? throw Unsupported{"stack underflow"} : st.back(); pc++; break;
        case 0x28: push(load((popv() & 0xffffffffu) + in.aux, 4)); pc++; break;
        case 0x29: push(load((popv() & 0xffffffffu) + in.aux, 8)); pc++; break;
        case 0x2c: push((uint32_t)(int32_t)(int8_t)load((popv() & 0xffffffffu) + in.aux, 1)); pc++; break;
        case 0x2d: push(load((popv() & 0xffffffffu) + in.aux, 1)); pc++; break;
        case 0x2e: push((uint32_t)(int32_t)(int16_t)load((popv() & 0xffffffffu) + in.aux, 2)); pc++; break;
        case 0x2f: push(load((popv() & 0xffffffffu) + in.aux, 2)); pc++; break;
        case 0x30: push((uint64_t)(int64_t)(int8_t)load((popv() & 0xffffffffu) + in.aux, 1)); pc++; break;
        case 0x31: push(load((popv() & 0xffffffffu) + in.aux, 1)); pc++; break;
        case 0x32: push((uint64_t)(int64_t)(int16_t)load((popv() & 0xffffffffu) + in.aux, 2)); pc++; break;
        case 0x33: push(load((popv() & 0xffffffffu) + in.aux, 2)); pc++; break;
        case 0x34: push((uint64_t)(int64_t)(int32_t)load((popv() & 0xffffffffu) + in.aux, 4)); pc++; break;
        case 0x35: push(load((popv() & 0xffffffffu) + in.aux, 4)); pc++; break;
        case 0x36: { uint64_t v = popv(); store((popv() & 0xffffffffu) + in.aux, 4, v); pc++; break; }
        case 0x37: { uint64_t v = popv(); store((popv() & 0xffffffffu) + in.aux, 8, v); pc++; break; }
        case 0x3a: case 0x3c: { uint64_t v = popv(); store((popv() & 0xffffffffu) + in.aux, 1, v); pc++; break; }
        case 0x3b: case 0x3d: { uint64_t v = popv(); store((popv() & 0xffffffffu) + in.aux, 2, v); pc++; break; }
        case 0x3e: { uint64_t v = popv(); store((popv() & 0xffffffffu) + in.aux, 4, v); pc++; break; }
        case 0x3f: push((uint32_t)(mem.size() >> 16)); pc++; break;
        case 0x40: {
          uint64_t n = popv() & 0xffffffffu, old = mem.size() >> 16;
          if (old + n > 32768) push(0xffffffffu);
          else { mem.resize((size_t)(old + n) << 16, 0); symbits.resize(mem.size() / 64 + 8, 0); push((uint32_t)old); }
          pc++;
          break;
        }
        case 0x41: push((uint32_t)in.imm); pc++; break;
        case 0x42: push((uint64_t)in.imm); pc++; break;
        case 0x45: {
          Val a = pop();
          if (a.tag) { uint32_t t = a.tag - 2; st.push_back(Val{0, ((t ^ 1u)) + 2}); }
          else push((uint32_t)((a.v & 0xffffffffu) == 0));
          pc++;
          break;
        }
        case 0x50: push((uint32_t)(popv() == 0)); pc++; break;
        default: {
          const uint8_t op = in.op;
          if (op >= 0x46 && op <= 0x4f) {                    // i32 comparisons
            uint32_t b = (uint32_t)popv(), a = (uint32_t)popv();
            int32_t sa = (int32_t)a, sb = (int32_t)b;
            uint32_t r;
            switch (op) {
              case 0x46: r = a == b; break; case 0x47: r = a != b; break;
              case 0x48: r = sa < sb; break; case 0x49: r = a < b; break;
              case 0x4a: r = sa > sb; break; case 0x4b: r = a > b; break;
              case 0x4c: r = sa <= sb; break; case 0x4d: r = a <= b; break;
              case 0x4e: r = sa >= sb; break; default: r = a >= b; break;
            }
            push(r);
          } else if (op >= 0x51 && op <= 0x5a) {             // i64 comparisons
            uint64_t b = popv(), a = popv();
            int64_t sa = (int64_t)a, sb = (int64_t)b;
            uint32_t r;
            switch (op) {
              case 0x51: r = a == b; break; case 0x52: r = a != b; break;
              case 0x53: r = sa < sb; break; case 0x54: r = a < b; break;
              case 0x55: r = sa > sb; break; case 0x56: r = a > b; break;
              case 0x57: r = sa <= sb; break; case 0x58: r = a <= b; break;
              case 0x59: r = sa >= sb; break; default: r = a >= b; break;
            }
            push(r);
          } else if (op >= 0x67 && op <= 0x78) {             // i32 arithmetic
            if (op <= 0x69) {
              uint32_t a = (uint32_t)popv();
              push(op == 0x67 ? (a ? (uint32_t)__builtin_clz(a) : 32u) : op == 0x68 ? (a ? (uint32_t)__builtin_ctz(a) : 32u)
                                                                                      : (uint32_t)__builtin_popcount(a));
            } else {
              uint32_t b = (uint32_t)popv(), a = (uint32_t)popv(), r;
              switch (op) {
                case 0x6a: r = a + b; break; case 0x6b: r = a - b; break; case 0x6c: r = a * b; break;
                case 0x6d: if (!b) throw Unsupported{"division by zero"}; r = (uint32_t)((int32_t)a / (int32_t)b); break;
                case 0x6e: if (!b) throw Unsupported{"division by zero"}; r = a / b; break;
                case 0x6f: if (!b) throw Unsupported{"division by zero"}; r = (uint32_t)((int32_t)a % (int32_t)b); break;
                case 0x70: if (!b) throw Unsupported{"division by zero"}; r = a % b; break;
                case 0x71: r = a & b; break; case 0x72: r = a | b; break; case 0x73: r = a ^ b; break;
                case 0x74: r = a << (b & 31); break;
                case 0x75: r = (uint32_t)((int32_t)a >> (b & 31)); break;
                case 0x76: r = a >> (b & 31); break;
                case 0x77: r = (a << (b & 31)) | (a >> ((32 - (b & 31)) & 31)); break;
                default: r = (a >> (b & 31)) | (a << ((32 - (b & 31)) & 31)); break;
              }
              push(r);
            }
          } else if (op >= 0x79 && op <= 0x8a) {             // i64 arithmetic
            if (op <= 0x7b) {
              uint64_t a = popv();
              push(op == 0x79 ? (a ? (uint64_t)__builtin_clzll(a) : 64u) : op == 0x7a ? (a ? (uint64_t)__builtin_ctzll(a) : 64u)
                                                                                        : (uint64_t)__builtin_popcountll(a));
            } else {
              uint64_t b = popv(), a = popv(), r;
              switch (op) {
                case 0x7c: r = a + b; break; case 0x7d: r = a - b; break; case 0x7e: r = a * b; break;
                case 0x7f: if (!b) throw Unsupported{"division by zero"}; r = (uint64_t)((int64_t)a / (int64_t)b); break;
                case 0x80: if (!b) throw Unsupported{"division by zero"}; r = a / b; break;
                case 0x81: if (!b) throw Unsupported{"division by zero"}; r = (uint64_t)((int64_t)a % (int64_t)b); break;
                case 0x82: if (!b) throw Unsupported{"division by zero"}; r = a % b; break;
                case 0x83: r = a & b; break; case 0x84: r = a | b; break; case 0x85: r = a ^ b; break;
                case 0x86: r = a << (b & 63); break;
                case 0x87: r = (uint64_t)((int64_t)a >> (b & 63)); break;
                case 0x88: r = a >> (b & 63); break;
                case 0x89: r = (a << (b & 63)) | (a >> ((64 - (b & 63)) & 63)); break;
                default: r = (a >> (b & 63)) | (a << ((64 - (b & 63)) & 63)); break;
              }
              push(r);
            }
          } else if (op == 0xa7) push((uint32_t)popv());
          else if (op == 0xac) push((uint64_t)(int64_t)(int32_t)(uint32_t)popv());
          else if (op == 0xad) push((uint64_t)(uint32_t)popv());
          else throw Unsupported{"unsupported opcode " + std::to_string(op)};
          pc++;
        }
      }
    }
    return 0;
  }

  bool call_export(const char *name, std::vector<uint64_t> a, uint64_t *res = nullptr) {
    auto it = exports.find(name);
    if (it == exports.end()) throw Unsupported{std::string("export not found: ") + name};
    std::vector<Val> args;
    for (auto x : a) args.push_back(Val{x, 0});
    Val r{0, 0};
    bool has = call(it->second, args, r);
    if (has && r.tag) throw Unsupported{"an export returned a signal-dependent value"};
    if (res) *res = r.v;
    return has;
  }
};

}  // namespace

bool build_witness_program(const uint8_t *wasm, size_t len, WitnessProgram &out, std::string &err) {
  Machine m;
  m.bin = wasm;
  m.bin_len = len;
  if (!m.parse(err)) return false;
  try {
    uint64_t v = 0;
    m.call_export("getFieldNumLen32", {}, &v);
    if (v != 8) { err = "field is not 8 x 32 bits"; return false; }
    m.call_export("getWitnessSize", {}, &v);
    out.n_wires = (uint32_t)v;
    m.call_export("getInputSize", {}, &v);
    out.n_inputs = (uint32_t)v;
    if (!out.n_wires || !out.n_inputs || out.n_wires > (1u << 26)) { err = "implausible witness / input size"; return false; }
    // prime
    m.call_export("getRawPrime", {});
    for (int j = 0; j < 8; j++) {
      m.call_export("readSharedRWMemory", {(uint64_t)j}, &v);
      if ((uint32_t)v != RMOD[j]) { err = "wasm prime is not the BN254 scalar field"; return false; }
    }
    m.call_export("init", {0});
    // signal addresses of the wires: getWitness(w) = Fr_copy(shared, &signal[map[w]]) ...
    std::vector<uint32_t> wire_addr(out.n_wires);
    m.probing = true;
    for (uint32_t w = 0; w < out.n_wires; w++) {
      try { m.call_export("getWitness", {w}); throw Unsupported{"getWitness does not copy a signal"}; }
      catch (ProbeDone &) { wire_addr[w] = m.probe_src; }
      m.depth = 0;
    }
    m.probing = false;
    if (out.n_wires < 2 || wire_addr[1] <= wire_addr[0] || (wire_addr[1] - wire_addr[0]) % 40) { err = "unexpected signal layout"; return false; }
    const uint32_t sig0 = wire_addr[0];                        // signal 0 (the constant 1)
    // inputs: the 256-slot name hashmap ({hash u64, first signal u32, size u32}) sits in a 4096-byte data segment
    {
      Rd r{wasm, 8, len};
      const uint8_t *hm = nullptr;
      while (r.p < len && r.ok) {
        uint8_t id = r.u8();
        size_t size = (size_t)r.leb(), end = r.p + size;
        if (id == 11) {
          Rd s{wasm, r.p, end};
          uint64_t n = s.leb();
          for (uint64_t i = 0; i < n; i++) {
            s.leb(); s.u8(); s.sleb(); s.u8();
            size_t l = (size_t)s.leb();
            if (l == 4096) hm = wasm + s.p;
            s.p += l;
          }
        }
        r.p = end;
      }
      if (!hm) { err = "input-name hashmap not found"; return false; }
      uint32_t total = 0, first = ~0u;
      for (int i = 0; i < 256; i++) {
        WitnessProgram::Input in;
        memcpy(&in.hash, hm + 16 * i, 8);
        memcpy(&in.pos, hm + 16 * i + 8, 4);
        memcpy(&in.size, hm + 16 * i + 12, 4);
        if (in.hash == 0 && in.size == 0) continue;
        out.inputs.push_back(in);
        total += in.size;
        first = std::min(first, in.pos);
      }
      if (total != out.n_inputs) { err = "input hashmap does not add up to getInputSize"; return false; }
      std::sort(out.inputs.begin(), out.inputs.end(), [](const WitnessProgram::Input &a, const WitnessProgram::Input &b) { return a.pos < b.pos; });
      uint32_t expect = first;
      for (auto &in : out.inputs) {
        if (in.pos != expect) { err = "input signals are not contiguous"; return false; }
        expect += in.size;
      }
      out.first_input_signal = first;
    }
    m.input_signal_addr.resize(out.n_inputs);
    for (uint32_t i = 0; i < out.n_inputs; i++) m.input_signal_addr[i] = sig0 + 40 * (out.first_input_signal + i);
    m.next_value = out.n_inputs;
    // drive the circom_runtime protocol with zeros; the last setInputSignal runs the main component, at which point
    // (first call of a *_run function) every input record becomes a symbol
    for (auto &in : out.inputs)
      for (uint32_t k = 0; k < in.size; k++) {
        for (int j = 0; j < 8; j++) m.call_export("writeSharedRWMemory", {(uint64_t)j, 0});
        m.call_export("setInputSignal", {in.hash >> 32, in.hash & 0xffffffffu, k});
      }
    if (!m.inputs_marked) { err = "the main component never ran (no *_run function was called)"; return false; }
    // ---- wires -----------------------------------------------------------------------------------
    std::vector<uint32_t> wire_ref(out.n_wires);
    for (uint32_t w = 0; w < out.n_wires; w++) wire_ref[w] = m.operand_ref(wire_addr[w]);
    // ---- prune: keep what the wires and the asserts need ------------------------------------------
    const uint32_t nv = m.next_value;
    std::vector<int32_t> def(nv, -1);                          // value id -> ssa index
    for (size_t i = 0; i < m.ssa.size(); i++)
      if (m.ssa[i].op != T_ASSERT_TRUE && m.ssa[i].op != T_ASSERT_FALSE) def[m.ssa[i].dst >> 1] = (int32_t)i;
    std::vector<uint8_t> live(m.ssa.size(), 0);
    std::vector<uint32_t> work;
    auto need = [&](uint32_t ref) {
      if (ref & 1u) return;
      uint32_t id = ref >> 1;
      if (id < out.n_inputs) return;
      int32_t d = def[id];
      if (d >= 0 && !live[d]) { live[d] = 1; work.push_back((uint32_t)d); }
    };
    for (uint32_t w = 0; w < out.n_wires; w++) need(wire_ref[w]);
    for (size_t i = 0; i < m.ssa.size(); i++)
      if (m.ssa[i].op == T_ASSERT_TRUE || m.ssa[i].op == T_ASSERT_FALSE) { live[i] = 1; work.push_back((uint32_t)i); }
    while (!work.empty()) {
      const TapeOp &o = m.ssa[work.back()];
      work.pop_back();
      need(o.a);
      if (o.op != T_NEG && o.op != T_INV && o.op != T_LNOT && o.op != T_BNOT && o.op != T_ISTRUE && o.op != T_ASSERT_TRUE &&
          o.op != T_ASSERT_FALSE && o.op != T_COPY) need(o.b);
      if (o.op == T_SELECT) need(o.c);
    }
    // ---- levels -------------------------------------------------------------------------------------
    std::vector<uint32_t> level_of_value(nv, 0);               // inputs: level 0 (available before level 0 runs)
    std::vector<uint32_t> op_level(m.ssa.size(), 0);
    uint32_t n_levels = 0;
    auto lvl = [&](uint32_t ref) -> uint32_t { return (ref & 1u) ? 0u : level_of_value[ref >> 1]; };
    std::vector<uint32_t> kept;
    for (size_t i = 0; i < m.ssa.size(); i++) {
      if (!live[i]) continue;
      const TapeOp &o = m.ssa[i];
      uint32_t l = lvl(o.a);
      const bool unary = o.op == T_NEG || o.op == T_INV || o.op == T_LNOT || o.op == T_BNOT || o.op == T_ISTRUE ||
                         o.op == T_ASSERT_TRUE || o.op == T_ASSERT_FALSE || o.op == T_COPY;
      if (!unary) l = std::max(l, lvl(o.b));
      if (o.op == T_SELECT) l = std::max(l, lvl(o.c));
      op_level[i] = l;                                          // runs in level l, result available from level l + 1
      if (o.op != T_ASSERT_TRUE && o.op != T_ASSERT_FALSE) level_of_value[o.dst >> 1] = l + 1;
      n_levels = std::max(n_levels, l + 1);
      kept.push_back((uint32_t)i);
    }
    std::stable_sort(kept.begin(), kept.end(), [&](uint32_t x, uint32_t y) {
      if (op_level[x] != op_level[y]) return op_level[x] < op_level[y];
      return m.ssa[x].op < m.ssa[y].op;                         // equal opcodes side by side: less lane divergence
    });
    // ---- slots: a value's slot is reused once the level of its last reader has completed --------------
    std::vector<uint32_t> last_use(nv, 0);
    std::vector<uint8_t> pinned(nv, 0);
    for (uint32_t w = 0; w < out.n_wires; w++)
      if (!(wire_ref[w] & 1u)) pinned[wire_ref[w] >> 1] = 1;
    auto use = [&](uint32_t ref, uint32_t l) { if (!(ref & 1u)) last_use[ref >> 1] = std::max(last_use[ref >> 1], l); };
    for (uint32_t i : kept) {
      const TapeOp &o = m.ssa[i];
      use(o.a, op_level[i]); use(o.b, op_level[i]); use(o.c, op_level[i]);
    }
    std::vector<uint32_t> slot_of(nv, ~0u);
    for (uint32_t i = 0; i < out.n_inputs; i++) slot_of[i] = i;
    uint32_t n_slots = out.n_inputs;
    std::vector<uint32_t> free_slots;
    std::vector<std::vector<uint32_t>> release(n_levels + 2);    // slots that become free after level l
    out.tape.clear();
    out.level_start.clear();
    uint32_t cur_level = ~0u;
    auto sref = [&](uint32_t ref) -> uint32_t { return (ref & 1u) ? ref : (slot_of[ref >> 1] << 1); };
    for (uint32_t i : kept) {
      const uint32_t l = op_level[i];
      while (cur_level != l) {
        cur_level = cur_level == ~0u ? 0 : cur_level + 1;
        out.level_start.push_back((uint32_t)out.tape.size());
        if (cur_level >= 1) for (uint32_t s : release[cur_level - 1]) free_slots.push_back(s);
      }
      TapeOp o = m.ssa[i];
      const bool is_assert = o.op == T_ASSERT_TRUE || o.op == T_ASSERT_FALSE;
      TapeOp t{o.op, 0, sref(o.a), 0, 0};
      const bool unary = o.op == T_NEG || o.op == T_INV || o.op == T_LNOT || o.op == T_BNOT || o.op == T_ISTRUE || is_assert || o.op == T_COPY;
      if (!unary) t.b = sref(o.b);
      if (o.op == T_SELECT) t.c = sref(o.c);
      if (!is_assert) {
        const uint32_t id = o.dst >> 1;
        uint32_t s;
        if (!free_slots.empty()) { s = free_slots.back(); free_slots.pop_back(); }
        else s = n_slots++;
        slot_of[id] = s;
        t.dst = s << 1;
        if (!pinned[id]) release[std::max(last_use[id], l + 1)].push_back(s);   // a value nobody reads is released after the next level
      }
      out.tape.push_back(t);
    }
    out.level_start.push_back((uint32_t)out.tape.size());
    out.n_slots = n_slots;
    out.wire_ref.resize(out.n_wires);
    for (uint32_t w = 0; w < out.n_wires; w++) {
      uint32_t r = wire_ref[w];
      if (!(r & 1u) && slot_of[r >> 1] == ~0u) { err = "internal: wire value without a slot"; return false; }
      out.wire_ref[w] = sref(r);
    }
    out.consts = m.consts;
    out.interpreted_instructions = m.steps;
    out.n_values = nv;
    out.n_asserts = m.n_asserts;
    out.n_selects = m.n_selects;
  } catch (Unsupported &u) {
    err = "circuit outside the supported subset of the generic witness extractor: " + u.why;
    return false;
  } catch (ProbeDone &) {
    err = "internal: stray probe";
    return false;
  } catch (std::exception &e) {
    err = std::string("witness extraction failed: ") + e.what();
    return false;
  }
  return true;
}

}  // namespace zkb
