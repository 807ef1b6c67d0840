// Generic witness-program extraction from a circom-2 witness-calculator wasm - see wasm_symexec.cc
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace zkb {

// Straight-line program over field elements ("tape").  Operand references: (index << 1) | 1 = constant pool entry,
// (index << 1) | 0 = value slot of the proof being evaluated.  Slots [0, n_inputs) hold the inputs in signal order.
enum TapeOpcode : uint8_t {
  T_ADD = 0, T_SUB, T_MUL, T_NEG, T_DIV, T_INV,
  T_EQ, T_NEQ, T_LT, T_GT, T_LEQ, T_GEQ, T_LAND, T_LOR, T_LNOT,
  T_SHR, T_SHL, T_BAND, T_BOR, T_BXOR, T_BNOT,
  T_ISTRUE,        // dst = (a != 0)
  T_SELECT,        // dst = (c != 0) ? a : b
  T_ASSERT_TRUE,   // status 4 unless a != 0   (the wasm's exceptionHandler(4) path)
  T_ASSERT_FALSE,  // status 4 unless a == 0
  T_COPY,          // dst = a (only used to materialise a wire)
  T_IDIV,          // dst = floor(a / b) on the canonical values (circom's integer division); status 4 when b == 0
  T_MOD,           // dst = a mod b on the canonical values (circom's `%`); status 4 when b == 0
  T_POW,           // dst = a ^ b, b read as its canonical value (circom's `**`)
  T_NOPS
};
struct TapeOp {
  uint8_t op;
  uint32_t dst, a, b, c;
};

struct WitnessProgram {
  uint32_t n_inputs = 0;                 // field elements the circuit takes (getInputSize)
  uint32_t n_wires = 0;                  // witness length (getWitnessSize)
  uint32_t n_slots = 0;                  // value slots per proof after slot reuse (inputs first)
  uint32_t first_input_signal = 0;       // smallest signal number of an input
  std::vector<TapeOp> tape;              // ordered by dependency level
  std::vector<uint32_t> level_start;     // tape index of the first op of each level, plus tape.size() at the end
  std::vector<uint32_t> consts;          // constant pool, 8 words each, canonical little-endian values
  std::vector<uint32_t> wire_ref;        // per wire: operand reference of its value
  struct Input { uint64_t hash; uint32_t pos, size; };   // FNV-1a-64(name), first signal number, element count
  std::vector<Input> inputs;
  // statistics
  uint64_t interpreted_instructions = 0, n_values = 0;
  uint32_t n_asserts = 0, n_selects = 0;
};

// Runs the wasm's own code once with symbolic inputs.  Integer / control state is executed concretely, calls of the
// circom field runtime (Fr_add, Fr_mul, ...) whose operands depend on an input are recorded instead of executed.
// Returns false + err for programs outside the supported subset (control flow or addressing that depends on a
// signal value, other than circuit asserts and two-armed conditionals that only assign field elements).
bool build_witness_program(const uint8_t *wasm, size_t len, WitnessProgram &out, std::string &err);

uint64_t fnv1a64_name(const std::string &s);

}  // namespace zkb
