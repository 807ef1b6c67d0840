// Structural reader for circom-2 witness-calculator wasm files - see wasm_circuit.cc
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace zkb {

struct PoseidonConst {
  uint32_t v[8];
  int form;   // 0 = canonical value, 1 = Montgomery (R = 2^256), 2 = canonical magnitude of a negative value
};
struct PoseidonTables {
  std::vector<PoseidonConst> tab[4];   // C, S, M, P
};
struct WasmInputSlot { uint64_t hash; uint32_t pos, size; };

struct WasmCircuit {
  uint32_t n_wires = 0, n_inputs = 0;
  std::vector<uint32_t> witness_map;      // wire -> circom signal
  std::vector<WasmInputSlot> hashmap;     // 256 slots, FNV-1a-64(name) & 0xff, linear probing
  PoseidonTables poseidon[3];             // t = 3, 4, 5
};

bool parse_circom_wasm(const uint8_t *buf, size_t len, WasmCircuit &out, std::string &err);
bool wasm_input_lookup(const WasmCircuit &w, const std::string &name, uint32_t &pos, uint32_t &size);

}  // namespace zkb
