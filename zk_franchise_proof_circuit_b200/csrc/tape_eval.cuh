// Device side of the generic witness path: evaluates a WitnessProgram (wasm_symexec.h) for a batch
// of proofs - see tape_eval.cu
#pragma once
#include <cuda_runtime.h>
#include "fp.cuh"
#include "wasm_symexec.h"

namespace zkb {

struct TapeDev {
  TapeOp *tape = nullptr;
  uint32_t *level_start = nullptr;   // n_levels + 1
  Fr *consts = nullptr;              // Montgomery form
  uint32_t *wire_ref = nullptr;
  uint32_t n_levels = 0, n_slots = 0, n_inputs = 0, n_wires = 0, n_ops = 0, n_consts = 0;
  cudaError_t upload(const WitnessProgram &p, cudaStream_t st);
  void free_all();
};

// inputs: [n][n_inputs] canonical; slots: [n][n_slots] workspace; wtns: [n][n_wires] canonical; status[p] = 4 when an
// assert of the circuit failed (max-combined with what is there)
cudaError_t tape_eval(const TapeDev &T, const Fr *inputs, Fr *slots, Fr *wtns, int *status, uint32_t n, cudaStream_t st);

}  // namespace zkb
