// Generic witness evaluation on the GPU: the straight-line program that wasm_symexec.cc extracts
// from a circom wasm, run for a batch of proofs.  One warp per proof: the program is ordered by dependency level and
// the 32 lanes share the operations of a level (consecutive lanes take consecutive operations, which the extractor
// sorted by opcode), with one __syncwarp per level.  A proof's values live in its own slot array in HBM (slots are
// reused once their last reader's level has run, so ~1 slot per wire instead of one per intermediate value); values
// are kept in Montgomery form, constants are shared by all proofs.  Operation semantics: tape_ops.cuh (the same code
// the CPU test harness runs).  A witness program is latency-bound (a Poseidon chain is ~85 k levels for the census
// circuit): throughput comes from the number of proofs in flight, not from a single proof.
#include "tape_eval.cuh"
#include "tape_ops.cuh"
#include <vector>

namespace zkb {

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return e_; } while (0)

__device__ __forceinline__ Fr tape_ld(const Fr *p) {
  const uint4 *q = reinterpret_cast<const uint4 *>(p);
  const uint4 a = q[0], b = q[1];
  Fr r;
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
__device__ __forceinline__ void tape_st(Fr *p, const Fr &x) {
  uint4 *q = reinterpret_cast<uint4 *>(p);
  q[0] = make_uint4(x.v[0], x.v[1], x.v[2], x.v[3]);
  q[1] = make_uint4(x.v[4], x.v[5], x.v[6], x.v[7]);
}

__global__ void k_tape_consts(Fr *c, uint32_t n) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) c[i] = c[i].to_mont();
}

__global__ void __launch_bounds__(128) k_tape_eval(TapeDev T, const Fr *__restrict__ inputs, Fr *slots_all, int *status,
                                                   uint32_t n) {
  const uint32_t p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (p >= n) return;                                   // whole warps leave together
  Fr *slots = slots_all + (size_t)p * T.n_slots;
  for (uint32_t i = lane; i < T.n_inputs; i += 32) tape_st(slots + i, tape_ld(inputs + (size_t)p * T.n_inputs + i).to_mont());
  __syncwarp();
  bool failed = false;
  auto exec = [&](const TapeOp &o) {
    const Fr a = (o.a & 1u) ? tape_ld(T.consts + (o.a >> 1)) : tape_ld(slots + (o.a >> 1));
    Fr b = a, c = a;
    if ((o.op <= T_BXOR && o.op != T_NEG && o.op != T_INV && o.op != T_LNOT) || o.op >= T_IDIV)
      b = (o.b & 1u) ? tape_ld(T.consts + (o.b >> 1)) : tape_ld(slots + (o.b >> 1));
    if (o.op == T_SELECT) {
      b = (o.b & 1u) ? tape_ld(T.consts + (o.b >> 1)) : tape_ld(slots + (o.b >> 1));
      c = (o.c & 1u) ? tape_ld(T.consts + (o.c >> 1)) : tape_ld(slots + (o.c >> 1));
    }
    Fr r;
    // the three operations that make up almost all of a circuit stay out of the big switch
    if (o.op == T_MUL) r = a * b;
    else if (o.op == T_ADD) r = a + b;
    else if (o.op == T_SUB) r = a - b;
    else r = tape_apply(o.op, a, b, c, &failed);
    if (o.op != T_ASSERT_TRUE && o.op != T_ASSERT_FALSE) tape_st(slots + (o.dst >> 1), r);
  };
  // The operations of level l + 1 (and the level table two ahead) are fetched while level l computes: a level is a
  // chain tape entry -> operands -> product -> store, and the first link does not depend on the previous level.
  uint32_t beg = T.level_start[0], end = T.level_start[1];
  TapeOp cur = T.tape[beg + lane];                       // (the tape is padded by 64 entries: no clamping needed)
  for (uint32_t l = 0; l < T.n_levels; l++) {
    const uint32_t nend = T.level_start[min(l + 2, T.n_levels)];
    const TapeOp nxt = T.tape[end + lane];
    if (beg + lane < end) exec(cur);
    for (uint32_t i = beg + lane + 32; i < end; i += 32) exec(T.tape[i]);     // levels wider than a warp (rare)
    __syncwarp();
    cur = nxt;
    beg = end;
    end = nend;
  }
  if (__any_sync(0xffffffffu, failed) && lane == 0) atomicMax(status + p, 4);
}

// thread = (wire, proof)
__global__ void __launch_bounds__(256) k_tape_gather(TapeDev T, const Fr *slots_all, Fr *wtns) {
  const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= T.n_wires) return;
  const uint32_t ref = T.wire_ref[w];
  const Fr v = (ref & 1u) ? tape_ld(T.consts + (ref >> 1)) : tape_ld(slots_all + (size_t)blockIdx.y * T.n_slots + (ref >> 1));
  tape_st(wtns + (size_t)blockIdx.y * T.n_wires + w, v.from_mont());
}

cudaError_t TapeDev::upload(const WitnessProgram &p, cudaStream_t st) {
  n_levels = (uint32_t)p.level_start.size() - 1;
  n_slots = p.n_slots; n_inputs = p.n_inputs; n_wires = p.n_wires;
  n_ops = (uint32_t)p.tape.size();
  n_consts = (uint32_t)(p.consts.size() / 8);
  CK(cudaMalloc(&tape, (p.tape.size() + 64) * sizeof(TapeOp)));          // the evaluator prefetches up to a warp ahead
  CK(cudaMemsetAsync(tape, 0, (p.tape.size() + 64) * sizeof(TapeOp), st));
  CK(cudaMalloc(&level_start, p.level_start.size() * 4));
  CK(cudaMalloc(&consts, (p.consts.size() + 8) * 4));
  CK(cudaMalloc(&wire_ref, p.wire_ref.size() * 4));
  CK(cudaMemcpyAsync(tape, p.tape.data(), p.tape.size() * sizeof(TapeOp), cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(level_start, p.level_start.data(), p.level_start.size() * 4, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(consts, p.consts.data(), p.consts.size() * 4, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(wire_ref, p.wire_ref.data(), p.wire_ref.size() * 4, cudaMemcpyHostToDevice, st));
  if (n_consts) k_tape_consts<<<(n_consts + 127) / 128, 128, 0, st>>>(consts, n_consts);
  CK(cudaStreamSynchronize(st));
  return cudaGetLastError();
}
void TapeDev::free_all() {
  cudaFree(tape); cudaFree(level_start); cudaFree(consts); cudaFree(wire_ref);
  tape = nullptr; level_start = nullptr; consts = nullptr; wire_ref = nullptr;
}

cudaError_t tape_eval(const TapeDev &T, const Fr *inputs, Fr *slots, Fr *wtns, int *status, uint32_t n, cudaStream_t st) {
  if (!n) return cudaSuccess;
  k_tape_eval<<<(n * 32 + 127) / 128, 128, 0, st>>>(T, inputs, slots, status, n);
  k_tape_gather<<<dim3((T.n_wires + 255) / 256, n), 256, 0, st>>>(T, slots, wtns);
  return cudaGetLastError();
}

}  // namespace zkb
