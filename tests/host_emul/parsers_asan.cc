// AddressSanitizer harness for the host-side parsers of caller-supplied data: the JSON parsers of csrc/json_io.cc
// (inputs.json, proof.json / verification_key.json numbers, public.json arrays), the census-wasm recogniser
// (csrc/wasm_circuit.cc) and the .zkey reader (csrc/zkey.cc) on a document read from a file.  The document is held in
// an exact-size heap buffer without a terminator: any read past its end is a report.
//     g++ -fsanitize=address parsers_asan.cc csrc/json_io.cc csrc/wasm_circuit.cc csrc/zkey.cc;  parsers_asan json|wasm|zkey <file>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>
#include "../../zk_franchise_proof_circuit_b200/csrc/json_io.h"
#include "../../zk_franchise_proof_circuit_b200/csrc/wasm_circuit.h"
#include "../../zk_franchise_proof_circuit_b200/csrc/zkey.h"

int main(int argc, char **argv) {
  if (argc < 3) return 2;
  FILE *f = fopen(argv[2], "rb");
  if (!f) return 2;
  fseek(f, 0, SEEK_END);
  long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  char *buf = new char[n ? n : 1];
  if (fread(buf, 1, n, f) != (size_t)n) return 2;
  fclose(f);
  std::string err;
  int ok = 0;
  if (!strcmp(argv[1], "json")) {
    std::map<std::string, std::vector<uint32_t>> m, m2;
    std::vector<uint32_t> arr;
    ok += zkb::parse_inputs_json(buf, n, m, err);
    ok += 2 * zkb::parse_json_numbers(buf, n, true, m2, err);
    ok += 4 * zkb::parse_json_array(buf, n, arr, err);
  } else if (!strcmp(argv[1], "wasm")) {
    zkb::WasmCircuit w;
    ok = zkb::parse_circom_wasm((const uint8_t *)buf, n, w, err);
    if (ok) { uint32_t p, s; zkb::wasm_input_lookup(w, "censusSiblings", p, s); }
  } else {
    zkb::ZkeyView z;
    ok = zkb::parse_zkey((const uint8_t *)buf, (size_t)n, z, err);
    if (ok) {                                       // what the loader reads: both matrices, the last byte of every section
      zkb::CoefCsr A, B;
      zkb::build_csr(z, 0, A);
      zkb::build_csr(z, 1, B);
      volatile uint8_t s = 0;
      s += z.a[(size_t)z.n_vars * 64 - 1]; s += z.b1[(size_t)z.n_vars * 64 - 1]; s += z.b2[(size_t)z.n_vars * 128 - 1];
      if (z.n_vars > z.n_public + 1) s += z.c[(size_t)(z.n_vars - z.n_public - 1) * 64 - 1];
      s += z.h[(size_t)z.domain * 64 - 1]; s += z.ic[(size_t)(z.n_public + 1) * 64 - 1];
    }
  }
  delete[] buf;
  printf("%d\n", ok);
  return 0;
}
