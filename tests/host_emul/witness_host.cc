// Test-only host build of csrc/census_witness.cuh (portable code path), driven from pytest through
// ctypes and compared wire-for-wire with the reference wasm (oracle/_ref).  Not part of the product.
#include "../../zk_franchise_proof_circuit_b200/csrc/census_witness.cuh"
#include <vector>
#include <cstring>
using namespace zkb;

extern "C" {
// consts: concatenation for t = 3,4,5 of C|S|M|P (Montgomery, 8 x u32 each); lens[12] their lengths.
// sig2wire: int32[n_signals].  inputs: canonical (2n+12) x 32 B.  out: n_wires x 32 B (pre-filled by caller
// when skip_const).  Returns status (0 ok, 4 assert failed, -1 layout mismatch).
int host_witness(const uint32_t *consts, const uint32_t *lens, const int32_t *sig2wire, uint32_t n_signals,
                 uint32_t n_levels_plus1, const uint8_t *inputs, uint8_t *out, int skip_const) {
  CensusLayout L;
  if (!census_layout_build(L, n_levels_plus1)) return -1;
  if (L.n_signals != n_signals) return -1;
  uint32_t off = 0;
  for (int i = 0; i < 3; i++) {
    L.pex[i].c_off = off; off += lens[4 * i];
    L.pex[i].s_off = off; off += lens[4 * i + 1];
    L.pex[i].m_off = off; off += lens[4 * i + 2];
    L.pex[i].p_off = off; off += lens[4 * i + 3];
  }
  std::vector<Fr> stage(L.n_signals, Fr::zero());
  WitnessEnv e;
  e.L = &L;
  e.consts = reinterpret_cast<const Fr *>(consts);
  e.stage = stage.data();
  e.status = 0;
  const Fr *in = reinterpret_cast<const Fr *>(inputs);
  Fr zero2[2] = {Fr::zero(), Fr::zero()};
  Fr h00 = poseidon_ex<3>(e, 0, zero2, false);
  Fr z3[3] = {Fr::zero(), Fr::zero(), Fr::one()};
  Fr h001 = poseidon_ex<4>(e, 0, z3, false);
  for (int task = 0; task < WITNESS_TASKS; task++) census_witness_task(e, task, in, h00, h001, skip_const != 0);
  // gather: `out` holds the template on entry when skip_const (as the device kernel reads it)
  Fr *o = reinterpret_cast<Fr *>(out);
  for (uint32_t s = 0, w = 0; s < n_signals; s++) {
    if (sig2wire[s] < 0) continue;
    w = (uint32_t)sig2wire[s];
    o[w] = witness_gather_one(stage.data(), s, skip_const ? o : nullptr, w);
  }
  return e.status;
}
uint32_t host_layout_signals(uint32_t n_levels_plus1) {
  CensusLayout L;
  if (!census_layout_build(L, n_levels_plus1)) return 0;
  return L.n_signals;
}
}
