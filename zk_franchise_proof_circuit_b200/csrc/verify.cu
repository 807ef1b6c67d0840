// Groth16 verification on the GPU, one proof per thread (pairing.cuh): the product side of
// `proof.Verify(vkey)` (`zk_census_test.go:122`) and of `snarkjs groth16 verify`.
// Inputs are the reference's JSON documents: verification_key.json, the public-signals array and proof.json.
#include <cuda_runtime.h>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>
#include "common.h"
#include "json_io.h"
#include "pairing.cuh"

namespace zkb {

#define CKR(x, what) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return cuda_fail(e_, what); } while (0)

__global__ void k_pairing_consts(PairingConsts *pc) {
  if (threadIdx.x || blockIdx.x) return;
  pairing_consts_init(*pc);
}

// canonical -> Montgomery, then the Miller value of (alpha1, beta2) once per key
__global__ void k_vkey_prepare(VerifyingKey *vk, G1Affine *ic, uint32_t n_ic, const PairingConsts *pc) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_ic) { ic[i].x = ic[i].x.to_mont(); ic[i].y = ic[i].y.to_mont(); }
  if (i == 0) {
    VerifyingKey k = *vk;
    Fq *f = reinterpret_cast<Fq *>(&k);                       // alpha1 (2) beta2 (4) gamma2 (4) delta2 (4) coordinates
    for (int j = 0; j < 14; j++) f[j] = f[j].to_mont();
    k.ab_miller = pairing_miller(k.alpha1, k.beta2, *pc);
    *vk = k;
  }
}

// proofs: n x 8 canonical Fq (A.x A.y B.x0 B.x1 B.y0 B.y1 C.x C.y); pubs: n x n_public canonical Fr
__global__ void __launch_bounds__(32) k_verify(const VerifyingKey *vk, const G1Affine *ic, const PairingConsts *pc,
                                                const Fq *proofs, const Fr *pubs, int *ok, uint32_t n) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const Fq *p = proofs + (size_t)i * 8;
  G1Affine A = {p[0].to_mont(), p[1].to_mont()}, C = {p[6].to_mont(), p[7].to_mont()};
  G2Affine B = {Fq2{p[2].to_mont(), p[3].to_mont()}, Fq2{p[4].to_mont(), p[5].to_mont()}};
  bool good = groth16_verify_one(*vk, ic, pubs + (size_t)i * vk->n_public, A, B, C, *pc);
  if (good) {
    // B must lie in the order-r subgroup of the twist (its cofactor is not 1): [r]B == O
    uint32_t r[8];
    for (int k = 0; k < 8; k++) r[k] = FrParams::mod(k);
    XYZZ<Fq2> b = XYZZ<Fq2>::from_affine(B), t;
    xyzz_mul_ni(&b, r, &t);
    good = t.is_inf();
  }
  ok[i] = good ? 1 : 0;
}

struct VKeyDev {
  VerifyingKey *vk = nullptr;
  G1Affine *ic = nullptr;
  uint32_t n_public = 0;
};

// Device objects are cached per GPU: a verify runs on the calling thread's current device (whatever the last prove on
// this thread bound, or the default device) and never touches another GPU's allocations.
static std::mutex g_mu;
static constexpr int MAX_DEV = 64;
static constexpr size_t MAX_VKEYS = 16;             // per device; the oldest entry is evicted
static PairingConsts *g_pc_dev[MAX_DEV] = {};
struct VKeyEntry { std::string text; VKeyDev dv; };
static std::vector<VKeyEntry> g_vkeys_dev[MAX_DEV]; // keyed by the vkey JSON text

static int get_vkey(int dev, const char *json, size_t len, VKeyDev &out, PairingConsts *&g_pc) {
  std::string key(json, len);
  auto &g_vkeys = g_vkeys_dev[dev];
  g_pc = g_pc_dev[dev];
  for (auto &e : g_vkeys)
    if (e.text == key) { out = e.dv; return ZKB_OK; }
  std::map<std::string, std::vector<uint32_t>> m;
  std::string err;
  json_reset_reduced();
  if (!parse_json_numbers(json, len, true, m, err)) { set_error("vkey: " + err); return ZKB_ERROR; }
  if (json_value_was_reduced()) { set_error("vkey: coordinate not in [0, q)"); return ZKB_ERROR; }
  auto need = [&](const char *k, size_t words) -> const uint32_t * {
    auto f = m.find(k);
    return (f == m.end() || f->second.size() != words) ? nullptr : f->second.data();
  };
  const uint32_t *a = need("vk_alpha_1", 24), *b = need("vk_beta_2", 48), *g = need("vk_gamma_2", 48),
                 *d = need("vk_delta_2", 48), *np = need("nPublic", 8);
  if (!a || !b || !g || !d || !np) { set_error("vkey: missing vk_alpha_1 / vk_beta_2 / vk_gamma_2 / vk_delta_2 / nPublic"); return ZKB_ERROR; }
  uint32_t n_public = np[0];
  auto icv = m.find("IC");
  if (icv == m.end() || icv->second.size() != (size_t)(n_public + 1) * 24) { set_error("vkey: IC has the wrong length"); return ZKB_ERROR; }
  if (!g_pc) {
    CKR(cudaMalloc(&g_pc, sizeof(PairingConsts)), "alloc");
    k_pairing_consts<<<1, 1>>>(g_pc);
    g_pc_dev[dev] = g_pc;
  }
  VerifyingKey h;
  memset(&h, 0, sizeof h);
  memcpy(&h.alpha1, a, 64);                 // x, y (the third projective coordinate "1" is dropped)
  memcpy(&h.beta2, b, 128);
  memcpy(&h.gamma2, g, 128);
  memcpy(&h.delta2, d, 128);
  h.n_public = n_public;
  std::vector<uint8_t> ic((size_t)(n_public + 1) * 64);
  for (uint32_t i = 0; i <= n_public; i++) memcpy(ic.data() + 64 * i, icv->second.data() + 24 * i, 64);
  VKeyDev dv;
  dv.n_public = n_public;
  CKR(cudaMalloc(&dv.vk, sizeof(VerifyingKey)), "alloc");
  CKR(cudaMalloc(&dv.ic, ic.size()), "alloc");
  CKR(cudaMemcpy(dv.vk, &h, sizeof h, cudaMemcpyHostToDevice), "h2d");
  CKR(cudaMemcpy(dv.ic, ic.data(), ic.size(), cudaMemcpyHostToDevice), "h2d");
  k_vkey_prepare<<<(n_public + 1 + 31) / 32, 32>>>(dv.vk, dv.ic, n_public + 1, g_pc);
  CKR(cudaDeviceSynchronize(), "vkey prepare");
  if (g_vkeys.size() >= MAX_VKEYS) {
    cudaFree(g_vkeys.front().dv.vk);
    cudaFree(g_vkeys.front().dv.ic);
    g_vkeys.erase(g_vkeys.begin());
  }
  g_vkeys.push_back({key, dv});
  out = dv;
  return ZKB_OK;
}

// proofs: n x 64 words (8 canonical Fq), pubs: n x n_public x 8 words (canonical Fr) -> ok[i]
static int verify_device(const VKeyDev &vk, const PairingConsts *g_pc, const uint32_t *proofs, const uint32_t *pubs, int n,
                         int *ok) {
  Fq *dp = nullptr;
  Fr *dq = nullptr;
  int *dok = nullptr;
  const size_t pw = (size_t)n * 64, qw = (size_t)n * vk.n_public * 8;
  CKR(cudaMalloc(&dp, pw * 4), "alloc");
  CKR(cudaMalloc(&dq, qw * 4 + 32), "alloc");
  CKR(cudaMalloc(&dok, (size_t)n * 4), "alloc");
  CKR(cudaMemcpy(dp, proofs, pw * 4, cudaMemcpyHostToDevice), "h2d");
  CKR(cudaMemcpy(dq, pubs, qw * 4, cudaMemcpyHostToDevice), "h2d");
  k_verify<<<(n + 31) / 32, 32>>>(vk.vk, vk.ic, g_pc, dp, dq, dok, (uint32_t)n);
  CKR(cudaGetLastError(), "verify launch");
  CKR(cudaMemcpy(ok, dok, (size_t)n * 4, cudaMemcpyDeviceToHost), "d2h");
  cudaFree(dp); cudaFree(dq); cudaFree(dok);
  return ZKB_OK;
}

}  // namespace zkb

using namespace zkb;

extern "C" {

// Verifies n (public signals, proof) pairs under one verification key on the GPU.  ok[i] = 1 valid, 0 invalid
// (including malformed documents, coordinates outside [0,q), public signals outside [0,r)).  Returns ZKB_OK unless
// the key itself is unusable or no GPU is present.
int zkb_verify_batch(const char *vkey_json, size_t vkey_len, int n, const char *const *publics_json,
                     const size_t *publics_len, const char *const *proofs_json, const size_t *proofs_len, int *ok) {
  if (require_device()) return ZKB_ERROR;
  std::lock_guard<std::mutex> g(g_mu);
  int dev = 0;
  CKR(cudaGetDevice(&dev), "get device");
  if (dev < 0 || dev >= MAX_DEV) { set_error("verify: device index out of range"); return ZKB_ERROR; }
  VKeyDev vk;
  PairingConsts *g_pc = nullptr;
  int rc = get_vkey(dev, vkey_json, vkey_len, vk, g_pc);
  if (rc) return rc;
  if (n <= 0) return ZKB_OK;
  std::vector<uint32_t> proofs((size_t)n * 64, 0), pubs((size_t)n * vk.n_public * 8, 0);
  std::vector<int> bad(n, 0);
  for (int i = 0; i < n; i++) {
    std::map<std::string, std::vector<uint32_t>> m;
    std::string err;
    json_reset_reduced();
    std::vector<uint32_t> pv;
    if (!parse_json_numbers(proofs_json[i], proofs_len[i], true, m, err) || !m.count("pi_a") || !m.count("pi_b") ||
        !m.count("pi_c") || m["pi_a"].size() != 24 || m["pi_b"].size() != 48 || m["pi_c"].size() != 24 ||
        !parse_json_array(publics_json[i], publics_len[i], pv, err) || pv.size() != (size_t)vk.n_public * 8 ||
        json_value_was_reduced()) {
      bad[i] = 1;
      continue;
    }
    uint32_t *p = proofs.data() + (size_t)i * 64;
    memcpy(p, m["pi_a"].data(), 64);
    memcpy(p + 16, m["pi_b"].data(), 128);
    memcpy(p + 48, m["pi_c"].data(), 64);
    memcpy(pubs.data() + (size_t)i * vk.n_public * 8, pv.data(), pv.size() * 4);
  }
  rc = verify_device(vk, g_pc, proofs.data(), pubs.data(), n, ok);
  if (rc) return rc;
  for (int i = 0; i < n; i++)
    if (bad[i]) ok[i] = 0;
  return ZKB_OK;
}

// Same check on binary results (what zkb_batch_get_results returns): proofs256 = n x 256 B canonical LE coordinates
// (A.x A.y B.x.c0 B.x.c1 B.y.c0 B.y.c1 C.x C.y), publics = n x nPublic x 32 B canonical LE.  Values outside
// [0, q) / [0, r) make the proof invalid (ok[i] = 0).
int zkb_verify_batch_bin(const char *vkey_json, size_t vkey_len, int n, const void *publics, const void *proofs256, int *ok) {
  if (require_device()) return ZKB_ERROR;
  if (!vkey_json || (n > 0 && (!publics || !proofs256 || !ok))) { set_error("null argument"); return ZKB_ERROR; }
  std::lock_guard<std::mutex> g(g_mu);
  int dev = 0;
  CKR(cudaGetDevice(&dev), "get device");
  if (dev < 0 || dev >= MAX_DEV) { set_error("verify: device index out of range"); return ZKB_ERROR; }
  VKeyDev vk;
  PairingConsts *g_pc = nullptr;
  int rc = get_vkey(dev, vkey_json, vkey_len, vk, g_pc);
  if (rc) return rc;
  if (n <= 0) return ZKB_OK;
  static const uint32_t QMOD[8] = {0xd87cfd47u, 0x3c208c16u, 0x6871ca8du, 0x97816a91u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
  static const uint32_t RMOD[8] = {0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
  auto below = [](const uint32_t *v, const uint32_t *m) {
    for (int k = 7; k >= 0; k--) if (v[k] != m[k]) return v[k] < m[k];
    return false;
  };
  const uint32_t *pw = (const uint32_t *)proofs256, *qw = (const uint32_t *)publics;
  std::vector<int> bad(n, 0);
  for (int i = 0; i < n; i++) {
    for (int k = 0; k < 8; k++) if (!below(pw + (size_t)i * 64 + 8 * k, QMOD)) bad[i] = 1;
    for (uint32_t k = 0; k < vk.n_public; k++) if (!below(qw + ((size_t)i * vk.n_public + k) * 8, RMOD)) bad[i] = 1;
  }
  rc = verify_device(vk, g_pc, pw, qw, n, ok);
  if (rc) return rc;
  for (int i = 0; i < n; i++)
    if (bad[i]) ok[i] = 0;
  return ZKB_OK;
}

// prover.ParseProof(...).Verify(vkey): 0 = valid, ZKB_INVALID_PROOF = the proof does not verify
int zkb_verify(const char *vkey_json, size_t vkey_len, const char *public_json, size_t public_len,
               const char *proof_json, size_t proof_len) {
  int ok = 0;
  const char *pubs[1] = {public_json}, *proofs[1] = {proof_json};
  size_t pl[1] = {public_len}, fl[1] = {proof_len};
  int rc = zkb_verify_batch(vkey_json, vkey_len, 1, pubs, pl, proofs, fl, &ok);
  if (rc) return rc;
  if (!ok) { set_error("proof does not verify"); return ZKB_INVALID_PROOF; }
  return ZKB_OK;
}

}  // extern "C"
