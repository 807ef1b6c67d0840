// Groth16 verification on the GPU, one proof per thread (pairing.cuh): the product side of
// `proof.Verify(vkey)` (`zk_census_test.go:122`) and of `snarkjs groth16 verify`.
// Inputs are the reference's JSON documents: verification_key.json, the public-signals array and proof.json.
#include <cuda_runtime.h>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>
#include <cstdlib>
#include <cerrno>
#include <sys/random.h>
#include "common.h"
#include "json_io.h"
#include "pairing.cuh"
#include "zkey.h"

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

// ---- random-linear-combination batch check ---------------------------------------------------------
// For random 128-bit rho_i:   prod_i e(-rho_i A_i, B_i) * e(alpha, beta)^(sum rho_i) * e(sum rho_i vkx_i, gamma)
//                             * e(sum rho_i C_i, delta) == 1
// holds for a batch of valid proofs and fails with probability 1 - 2^-128 if any proof is invalid: one Miller loop
// per proof plus two Miller loops and ONE final exponentiation per batch, instead of three Miller loops and a final
// exponentiation per proof.  A failing batch falls back to the per-proof kernel to name the invalid proofs.
struct RlcItem {
  Fq12 f;
  XYZZ<Fq> sx, sc;
};
__device__ __noinline__ void rlc_combine(RlcItem *a, const RlcItem *b) {
  a->f = a->f.mul(b->f);
  xyzz_add_ni(&a->sx, &b->sx);
  xyzz_add_ni(&a->sc, &b->sc);
}
// thread = proof: structural checks, vk_x, the proof's Miller value and its scaled G1 terms
__global__ void __launch_bounds__(32) k_verify_rlc_prep(const VerifyingKey *vk, const G1Affine *ic, const PairingConsts *pc,
                                                         const Fq *proofs, const Fr *pubs, const uint32_t *rho, RlcItem *items,
                                                         int *valid, uint32_t n) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const Fq *p = proofs + (size_t)i * 8;
  G1Affine A = {p[0].to_mont(), p[1].to_mont()}, C = {p[6].to_mont(), p[7].to_mont()};
  G2Affine B = {Fq2{p[2].to_mont(), p[3].to_mont()}, Fq2{p[4].to_mont(), p[5].to_mont()}};
  Fq three = Fq::zero();
  three.v[0] = 3;
  const Fq b1 = three.to_mont();
  bool good = !A.is_inf() && !B.is_inf() && on_curve(A, b1) && on_curve(C, b1) && on_curve(B, pc->twist_b);
  if (good) {                                       // B in the order-r subgroup of the twist
    uint32_t r[8];
    for (int k = 0; k < 8; k++) r[k] = FrParams::mod(k);
    XYZZ<Fq2> b = XYZZ<Fq2>::from_affine(B), t;
    xyzz_mul_ni(&b, r, &t);
    good = t.is_inf();
  }
  RlcItem it;
  it.f = Fq12::one();
  it.sx = XYZZ<Fq>::infinity();
  it.sc = XYZZ<Fq>::infinity();
  valid[i] = good ? 1 : 0;
  if (good) {
    uint32_t k[8] = {rho[4 * i], rho[4 * i + 1], rho[4 * i + 2], rho[4 * i + 3], 0, 0, 0, 0};
    XYZZ<Fq> vkx = XYZZ<Fq>::from_affine(ic[0]);
    for (uint32_t j = 0; j < vk->n_public; j++) {
      XYZZ<Fq> t, q = XYZZ<Fq>::from_affine(ic[j + 1]);
      xyzz_mul_ni(&q, pubs[(size_t)i * vk->n_public + j].v, &t);
      xyzz_add_ni(&vkx, &t);
    }
    xyzz_mul_ni(&vkx, k, &it.sx);
    XYZZ<Fq> c = XYZZ<Fq>::from_affine(C), a = XYZZ<Fq>::from_affine(A), ra;
    xyzz_mul_ni(&c, k, &it.sc);
    xyzz_mul_ni(&a, k, &ra);
    G1Affine rA;
    xyzz_to_affine_ni(&ra, &rA);
    rA.y = rA.y.neg();
    it.f = pairing_miller(rA, B, *pc);
  }
  items[i] = it;
}
// items[i] <- items[i] o items[i + stride] for i = 2 * stride * t
__global__ void __launch_bounds__(32) k_verify_rlc_reduce(RlcItem *items, uint32_t n, uint32_t stride) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t i = (uint64_t)t * 2 * stride;
  if (i + stride >= n) return;
  rlc_combine(items + i, items + i + stride);
}
// one thread: the three batch-level pairings and the final exponentiation.  rho_sum: 8 words (sum of the rho_i)
__global__ void k_verify_rlc_final(const VerifyingKey *vk, const PairingConsts *pc, const RlcItem *items, const uint32_t *rho_sum,
                                   int *ok) {
  if (threadIdx.x || blockIdx.x) return;
  Fq12 f = items[0].f;
  Fq12 ab = Fq12::one();
  for (int i = 255; i >= 0; i--) {
    ab = ab.mul(ab);
    if ((rho_sum[i >> 5] >> (i & 31)) & 1) ab = ab.mul(vk->ab_miller);
  }
  f = f.mul(ab);
  G1Affine sx, sc;
  xyzz_to_affine_ni(&items[0].sx, &sx);
  xyzz_to_affine_ni(&items[0].sc, &sc);
  f = f.mul(pairing_miller(sx, vk->gamma2, *pc));
  f = f.mul(pairing_miller(sc, vk->delta2, *pc));
  *ok = pairing_final_exp(f, *pc) == Fq12::one() ? 1 : 0;
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
static bool os_random(void *dst, size_t n) {
  uint8_t *d = (uint8_t *)dst;
  while (n) {
    ssize_t k = getrandom(d, n, 0);
    if (k < 0) { if (errno == EINTR) continue; return false; }
    d += k;
    n -= (size_t)k;
  }
  return true;
}

static int verify_device(const VKeyDev &vk, const PairingConsts *g_pc, const uint32_t *proofs, const uint32_t *pubs, int n,
                         int *ok) {
  Fq *dp = nullptr;
  Fr *dq = nullptr;
  int *dok = nullptr;
  const size_t pw = (size_t)n * 64, qw = (size_t)n * vk.n_public * 8;
  CKR(cudaMalloc(&dp, pw * 4), "alloc");
  CKR(cudaMalloc(&dq, qw * 4 + 32), "alloc");
  CKR(cudaMalloc(&dok, (size_t)n * 4 + 4), "alloc");
  CKR(cudaMemcpy(dp, proofs, pw * 4, cudaMemcpyHostToDevice), "h2d");
  CKR(cudaMemcpy(dq, pubs, qw * 4, cudaMemcpyHostToDevice), "h2d");
  static const uint32_t rlc_min = []() { const char *e = getenv("ZKB_VERIFY_RLC_MIN"); return e ? (uint32_t)atoi(e) : 8u; }();
  bool done = false;
  if (rlc_min && (uint32_t)n >= rlc_min) {
    // random-linear-combination batch check; per-proof checks only if it fails
    std::vector<uint32_t> rho((size_t)n * 4);
    if (!os_random(rho.data(), rho.size() * 4)) { set_error("getrandom failed"); return ZKB_ERROR; }
    uint32_t sum[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    uint32_t *drho = nullptr;
    RlcItem *items = nullptr;
    CKR(cudaMalloc(&drho, rho.size() * 4 + 32), "alloc");
    CKR(cudaMalloc(&items, (size_t)n * sizeof(RlcItem)), "alloc");
    for (int i = 0; i < n; i++) rho[4 * (size_t)i] |= 1u;                  // never zero
    CKR(cudaMemcpy(drho, rho.data(), rho.size() * 4, cudaMemcpyHostToDevice), "h2d");
    k_verify_rlc_prep<<<(n + 31) / 32, 32>>>(vk.vk, vk.ic, g_pc, dp, dq, drho, items, dok, (uint32_t)n);
    CKR(cudaMemcpy(ok, dok, (size_t)n * 4, cudaMemcpyDeviceToHost), "d2h");      // structural validity
    for (int i = 0; i < n; i++) {
      if (!ok[i]) continue;
      uint64_t carry = 0;
      for (int k = 0; k < 8; k++) {
        carry += (uint64_t)sum[k] + (k < 4 ? rho[4 * (size_t)i + k] : 0u);
        sum[k] = (uint32_t)carry;
        carry >>= 32;
      }
    }
    for (uint32_t stride = 1; stride < (uint32_t)n; stride <<= 1) {
      const uint32_t threads = ((uint32_t)n + 2 * stride - 1) / (2 * stride);
      k_verify_rlc_reduce<<<(threads + 31) / 32, 32>>>(items, (uint32_t)n, stride);
    }
    uint32_t *dsum = drho;                                                 // reuse: the prep kernel is done with rho
    CKR(cudaMemcpy(dsum, sum, 32, cudaMemcpyHostToDevice), "h2d");
    k_verify_rlc_final<<<1, 1>>>(vk.vk, g_pc, items, dsum, dok + n);
    int batch_ok = 0;
    CKR(cudaGetLastError(), "verify launch");
    CKR(cudaMemcpy(&batch_ok, dok + n, 4, cudaMemcpyDeviceToHost), "d2h");
    cudaFree(drho); cudaFree(items);
    done = batch_ok == 1;                                                  // ok[] already holds 1 for every proof in the product
  }
  if (!done) {
    k_verify<<<(n + 31) / 32, 32>>>(vk.vk, vk.ic, g_pc, dp, dq, dok, (uint32_t)n);
    CKR(cudaGetLastError(), "verify launch");
    CKR(cudaMemcpy(ok, dok, (size_t)n * 4, cudaMemcpyDeviceToHost), "d2h");
  }
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

// `snarkjs zkey export verificationkey` (circuit/circuit-compiler.sh:128-134): verification_key.json of a proving key,
// byte for byte the text snarkjs writes (JSON.stringify(vk, null, 1)): protocol, curve, nPublic, vk_alpha_1, vk_beta_2,
// vk_gamma_2, vk_delta_2, vk_alphabeta_12, IC.  vk_alphabeta_12 is e(alpha1, beta2) as ffjavascript computes it - the
// reduced pairing raised to 2x(6x^2+3x+1), x = 4965661367192848881 - as a [2][3][2] tower
// T[i][j][k] <-> u^k v^j w^i.  Host arithmetic (one pairing); *out_len: capacity in, bytes written (or needed) out.
int zkb_export_vkey(const void *zkey, size_t zkey_len, char *out, size_t *out_len) {
  if (!zkey || !out_len) { set_error("null argument"); return ZKB_ERROR; }
  ZkeyView z;
  std::string err;
  if (!parse_zkey((const uint8_t *)zkey, zkey_len, z, err)) { set_error(err); return ZKB_ERROR; }
  auto fq = [](const uint8_t *p) { Fq x; memcpy(x.v, p, 32); return x; };
  auto dec = [](const Fq &m) { Fq c = m.from_mont(); return "\"" + u256_to_dec(c.v) + "\""; };
  auto ind = [](int n) { return std::string((size_t)n, ' '); };
  auto g1 = [&](const uint8_t *p, int d) {
    Fq x = fq(p), y = fq(p + 32);
    const bool inf = x.is_zero() && y.is_zero();
    return "[\n" + ind(d + 1) + (inf ? "\"0\"" : dec(x)) + ",\n" + ind(d + 1) + (inf ? "\"1\"" : dec(y)) + ",\n" + ind(d + 1) +
           (inf ? "\"0\"" : "\"1\"") + "\n" + ind(d) + "]";
  };
  auto pair2 = [&](const std::string &a, const std::string &b, int d) {
    return "[\n" + ind(d + 1) + a + ",\n" + ind(d + 1) + b + "\n" + ind(d) + "]";
  };
  auto g2 = [&](const uint8_t *p, int d) {
    Fq c[4] = {fq(p), fq(p + 32), fq(p + 64), fq(p + 96)};
    bool inf = true;
    for (auto &x : c) inf = inf && x.is_zero();
    if (inf) return "[\n" + ind(d + 1) + pair2("\"0\"", "\"0\"", d + 1) + ",\n" + ind(d + 1) + pair2("\"1\"", "\"0\"", d + 1) + ",\n" +
                    ind(d + 1) + pair2("\"0\"", "\"0\"", d + 1) + "\n" + ind(d) + "]";
    return "[\n" + ind(d + 1) + pair2(dec(c[0]), dec(c[1]), d + 1) + ",\n" + ind(d + 1) + pair2(dec(c[2]), dec(c[3]), d + 1) + ",\n" +
           ind(d + 1) + pair2("\"1\"", "\"0\"", d + 1) + "\n" + ind(d) + "]";
  };
  // e(alpha1, beta2)^(2x(6x^2+3x+1))
  static PairingConsts pc;
  static bool pc_ready = false;
  {
    static std::mutex pmu;
    std::lock_guard<std::mutex> g(pmu);
    if (!pc_ready) { pairing_consts_init(pc); pc_ready = true; }
  }
  G1Affine a1 = {fq(z.alpha1), fq(z.alpha1 + 32)};
  G2Affine b2 = {Fq2{fq(z.beta2), fq(z.beta2 + 32)}, Fq2{fq(z.beta2 + 64), fq(z.beta2 + 96)}};
  Fq12 e = pairing_final_exp(pairing_miller(a1, b2, pc), pc);
  static const uint32_t EXTRA[6] = {0x3ddedaf4u, 0x2e5d4e22u, 0xd9d9e38du, 0x1ea96b02u, 0x15e307c8u, 0x3bec47dfu};
  Fq12 t = Fq12::one();
  for (int i = 6 * 32 - 1; i >= 0; i--) {
    t = t.mul(t);
    if ((EXTRA[i >> 5] >> (i & 31)) & 1) t = t.mul(e);
  }
  std::string ab = "[\n";
  for (int i = 0; i < 2; i++) {
    const Fq6 &h = i ? t.c1 : t.c0;
    const Fq2 *cf[3] = {&h.c0, &h.c1, &h.c2};
    ab += ind(2) + "[\n";
    for (int j = 0; j < 3; j++)
      ab += ind(3) + pair2(dec(cf[j]->a), dec(cf[j]->b), 3) + (j < 2 ? ",\n" : "\n");
    ab += ind(2) + "]" + (i == 0 ? ",\n" : "\n");
  }
  ab += ind(1) + "]";
  std::string js = "{\n \"protocol\": \"groth16\",\n \"curve\": \"bn128\",\n \"nPublic\": " + std::to_string(z.n_public) + ",\n";
  js += " \"vk_alpha_1\": " + g1(z.alpha1, 1) + ",\n";
  js += " \"vk_beta_2\": " + g2(z.beta2, 1) + ",\n";
  js += " \"vk_gamma_2\": " + g2(z.gamma2, 1) + ",\n";
  js += " \"vk_delta_2\": " + g2(z.delta2, 1) + ",\n";
  js += " \"vk_alphabeta_12\": " + ab + ",\n \"IC\": [\n";
  for (uint32_t i = 0; i <= z.n_public; i++) js += ind(2) + g1(z.ic + (size_t)i * 64, 2) + (i < z.n_public ? ",\n" : "\n");
  js += " ]\n}";
  const size_t need = js.size() + 1;
  if (!out || *out_len < need) { *out_len = need; return ZKB_SHORT_BUFFER; }
  memcpy(out, js.c_str(), need);
  *out_len = js.size();
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
