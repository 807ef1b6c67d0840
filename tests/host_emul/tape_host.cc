// Host build of the generic witness path (csrc/wasm_symexec.cc + csrc/tape_ops.cuh): extracts the witness program from
// a circom wasm and evaluates it on the CPU with the SAME operation semantics tape_eval.cu runs per lane on the GPU.
// Test infrastructure for machines without a GPU; the product evaluates tapes on the device only.
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include "../../zk_franchise_proof_circuit_b200/csrc/wasm_symexec.cc"
#include "../../zk_franchise_proof_circuit_b200/csrc/tape_ops.cuh"

using namespace zkb;

static WitnessProgram g_prog;
static bool g_have = false;

extern "C" {

// info[8] = n_inputs, n_wires, n_slots, tape length, levels, consts, asserts, selects
int tape_host_build(const uint8_t *wasm, size_t len, uint32_t *info, char *err, size_t errmax) {
  std::string e;
  g_prog = WitnessProgram();
  g_have = build_witness_program(wasm, len, g_prog, e);
  if (!g_have) { snprintf(err, errmax, "%s", e.c_str()); return 1; }
  info[0] = g_prog.n_inputs; info[1] = g_prog.n_wires; info[2] = g_prog.n_slots; info[3] = (uint32_t)g_prog.tape.size();
  info[4] = (uint32_t)g_prog.level_start.size() - 1; info[5] = (uint32_t)(g_prog.consts.size() / 8);
  info[6] = g_prog.n_asserts; info[7] = g_prog.n_selects;
  return 0;
}

// position of the input named `name` in the flat input vector (-1 if unknown); *size = its element count
int tape_host_input(const char *name, uint32_t *size) {
  const uint64_t h = fnv1a64_name(name);
  for (auto &in : g_prog.inputs)
    if (in.hash == h) { *size = in.size; return (int)(in.pos - g_prog.first_input_signal); }
  return -1;
}

// inputs: n_inputs x 32 B canonical; witness: n_wires x 32 B canonical.  returns 0, or 4 when an assert failed
int tape_host_eval(const uint8_t *inputs, uint8_t *witness) {
  if (!g_have) return 1;
  const WitnessProgram &p = g_prog;
  std::vector<Fr> slot(p.n_slots, Fr::zero()), cst(p.consts.size() / 8);
  for (size_t i = 0; i < cst.size(); i++) { Fr c; memcpy(c.v, &p.consts[8 * i], 32); cst[i] = c.to_mont(); }
  for (uint32_t i = 0; i < p.n_inputs; i++) { Fr x; memcpy(x.v, inputs + 32 * i, 32); slot[i] = x.to_mont(); }
  auto get = [&](uint32_t ref) -> Fr { return (ref & 1u) ? cst[ref >> 1] : slot[ref >> 1]; };
  bool failed = false;
  // level by level, results of a level written after all its operands were read (as the GPU lanes do)
  for (size_t l = 0; l + 1 < p.level_start.size(); l++) {
    std::vector<Fr> res(p.level_start[l + 1] - p.level_start[l]);
    for (uint32_t i = p.level_start[l]; i < p.level_start[l + 1]; i++) {
      const TapeOp &o = p.tape[i];
      res[i - p.level_start[l]] = tape_apply(o.op, get(o.a), get(o.b), get(o.c), &failed);
    }
    for (uint32_t i = p.level_start[l]; i < p.level_start[l + 1]; i++) {
      const TapeOp &o = p.tape[i];
      if (o.op != T_ASSERT_TRUE && o.op != T_ASSERT_FALSE) slot[o.dst >> 1] = res[i - p.level_start[l]];
    }
  }
  for (uint32_t w = 0; w < p.n_wires; w++) { Fr x = get(p.wire_ref[w]).from_mont(); memcpy(witness + 32 * w, x.v, 32); }
  return failed ? 4 : 0;
}

}  // extern "C"

// ---- the wasm run CONCRETELY by the interpreter (no symbols): an oracle for programs that have no native build ------
// Drives the circom_runtime protocol (init, writeSharedRWMemory + setInputSignal per input element, getWitness +
// readSharedRWMemory per wire) on the program tape_host_build was last given.  Every Fr_* call executes the wasm's own
// code.  returns 0, 4 when the program raised exception 4 (assert / constraint check), 3 for any other trap.
extern "C" int tape_host_wasm_witness(const uint8_t *wasm, size_t len, const uint8_t *inputs, uint8_t *witness) {
  if (!g_have) return 1;
  Machine m;
  m.bin = wasm;
  m.bin_len = len;
  std::string err;
  if (!m.parse(err)) return 2;
  m.inputs_marked = true;                               // nothing becomes a symbol
  try {
    m.call_export("init", {0});
    uint32_t flat = 0;
    for (auto &in : g_prog.inputs)
      for (uint32_t k = 0; k < in.size; k++, flat++) {
        for (int j = 0; j < 8; j++) {
          uint32_t w;
          memcpy(&w, inputs + 32 * flat + 4 * j, 4);
          m.call_export("writeSharedRWMemory", {(uint64_t)j, w});
        }
        m.call_export("setInputSignal", {in.hash >> 32, in.hash & 0xffffffffu, k});
      }
    for (uint32_t w = 0; w < g_prog.n_wires; w++) {
      m.call_export("getWitness", {w});
      for (int j = 0; j < 8; j++) {
        uint64_t v = 0;
        m.call_export("readSharedRWMemory", {(uint64_t)j}, &v);
        const uint32_t v32 = (uint32_t)v;
        memcpy(witness + 32 * w + 4 * j, &v32, 4);
      }
    }
  } catch (Unsupported &u) {
    return u.why.find("exception 4") != std::string::npos ? 4 : 3;
  }
  return 0;
}

// ---- the reference wasm's OWN field runtime as the oracle of tape_ops.cuh ------------------------------------------
// The census program only uses part of the operation set.  Every operation of tape_ops.cuh is therefore also compared
// with the wasm's Fr_* function of the same name, executed concretely by the interpreter of wasm_symexec.cc on the
// same operands (the mechanism the extractor uses for constant sub-expressions).
static Machine *g_rt = nullptr;
static std::vector<uint8_t> g_rt_wasm;

extern "C" {

int tape_host_rt_load(const uint8_t *wasm, size_t len) {
  delete g_rt;
  g_rt_wasm.assign(wasm, wasm + len);
  g_rt = new Machine();
  g_rt->bin = g_rt_wasm.data();
  g_rt->bin_len = g_rt_wasm.size();
  std::string err;
  if (!g_rt->parse(err)) { delete g_rt; g_rt = nullptr; return 1; }
  return 0;
}

// out = Fr_<name>(a, b) as the wasm computes it; a, b, out canonical 32-byte little-endian.  unary: b ignored.
int tape_host_rt_op(const char *name, int unary, const uint8_t *a32, const uint8_t *b32, uint8_t *out32) {
  if (!g_rt) return 1;
  Machine &m = *g_rt;
  uint32_t fi = ~0u;
  for (auto &kv : m.names) if (kv.second == name) fi = kv.first;
  if (fi == ~0u) return 2;
  const uint32_t base = (uint32_t)m.mem.size() - 4096, pa = base, pb = base + 64, pd = base + 128;
  auto put = [&](uint32_t addr, const uint8_t *v) {
    uint32_t hdr[2] = {0, 0x80000000u};              // long form, not Montgomery
    memcpy(m.mem.data() + addr, hdr, 8);
    memcpy(m.mem.data() + addr + 8, v, 32);
  };
  put(pa, a32);
  put(pb, b32);
  memset(m.mem.data() + pd, 0, 40);
  try {
    std::vector<Val> args;
    args.push_back(Val{pd, 0});
    args.push_back(Val{pa, 0});
    if (!unary) args.push_back(Val{pb, 0});
    Val res{0, 0};
    m.depth = 0;
    m.call(fi, args, res);
    uint32_t v[8];
    m.record_value(pd, v);
    memcpy(out32, v, 32);
  } catch (Unsupported &) {
    return 3;
  }
  return 0;
}

// the same operation through tape_ops.cuh (operands converted to Montgomery form and back, as the evaluators do)
int tape_host_apply(int op, const uint8_t *a32, const uint8_t *b32, const uint8_t *c32, uint8_t *out32) {
  Fr a, b, c;
  memcpy(a.v, a32, 32); memcpy(b.v, b32, 32); memcpy(c.v, c32, 32);
  bool failed = false;
  Fr r = tape_apply((uint8_t)op, a.to_mont(), b.to_mont(), c.to_mont(), &failed).from_mont();
  memcpy(out32, r.v, 32);
  return failed ? 4 : 0;
}

}  // extern "C"
