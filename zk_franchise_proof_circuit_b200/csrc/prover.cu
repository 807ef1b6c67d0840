// Census Groth16 prover: circuit loading, the batched GPU pipeline and the C ABI
// (include/zkcensus_b200.h).  Replaces what runs behind `prover.Prove(zkey, wasm, inputs)`
// (`zk_census_test.go:89`) / `groth16.fullProve` (`ts_inputs/src/example.ts:358-362`):
//
//   inputs.json --host parse--> 2n+12 canonical Fr values per proof --H2D-->
//   k_witness        census witness, 5 tasks per proof (hash chains on 8 lanes each), then k_witness_gather (W1-W7)
//   k_build_abc      A_T, B_T from the CSR coefficients, C_T = A_T o B_T   (G2)
//   NTT              iNTT -> coset shift -> NTT on 3 vectors per proof     (G3)
//   k_join           h = a*b - c, to canonical form                        (G4)
//   MSM              A, B1, C (G1) + B2 (G2) over the witness, H over h    (G5)
//   k_finalize       blinding with r,s and affine conversion               (G6)
//   --D2H--> 256 B proof + nPublic*32 B public signals --host format--> proof.json / public.json (G7)
//
// Everything heavy is device code; the host only parses/prints decimal strings.  No CPU fallback.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <map>
#include <mutex>
#include <sys/random.h>
#include <sys/stat.h>
#include <cerrno>
#include <memory>
#include <atomic>
#include <thread>
#include <algorithm>
#include <cstdlib>
#include "common.h"
#include "ntt.cuh"
#include "msm.cuh"
#include "census_witness.cuh"
#include "finalize_kernels.h"
#include "wasm_circuit.h"
#include "wasm_symexec.h"
#include "tape_eval.cuh"
#include "zkey.h"
#include "json_io.h"

namespace zkb {

#define CKR(x, what) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return cuda_fail(e_, what); } while (0)

// ---------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------
// Poseidon constants as found in the wasm -> Montgomery form
__global__ void k_consts_to_mont(Fr *c, const int *form, uint32_t n) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fr v = c[i];
  if (form[i] == 0) v = v.to_mont();
  else if (form[i] == 2) v = v.to_mont().neg();
  c[i] = v;
}

__global__ void k_hash_consts(CensusLayout L, const Fr *consts, Fr *hc) {
  if (threadIdx.x || blockIdx.x) return;
  WitnessEnv e;
  e.L = &L; e.consts = consts; e.stage = nullptr; e.status = 0;
  Fr z2[2] = {Fr::zero(), Fr::zero()};
  hc[0] = poseidon_ex<3>(e, 0, z2, false);
  Fr z3[3] = {Fr::zero(), Fr::zero(), Fr::one()};
  hc[1] = poseidon_ex<4>(e, 0, z3, false);
}

// The census witness program (census_witness.cuh) writing each proof's staging array.  The two hash-chain tasks
// (blockIdx.y < 2) run on COOP_LANES = 8 lanes per proof that share every Poseidon permutation; the three side tasks
// use one lane of the group.  inputs: [n][n_inputs] canonical; stage: [n][n_signals], zeroed
__global__ void __launch_bounds__(32) k_witness(CensusLayout L, const Fr *consts, const Fr *hc, const Fr *inputs,
                                                 Fr *stage, int *status, uint32_t n, int skip_const) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x, p = t / COOP_LANES, li = t % COOP_LANES;
  const int task = blockIdx.y;
  if (p >= n) return;                          // whole lane groups leave together
  const bool chain = task < 2;
  if (!chain && li) return;
  WitnessEnv e;
  e.L = &L;
  e.consts = consts;
  e.stage = stage + (size_t)p * L.n_signals;
  e.status = 0;
  e.coop = chain;
  e.li = chain ? li : 0;
  e.glane0 = threadIdx.x & ~(COOP_LANES - 1);
  e.gmask = ((1u << COOP_LANES) - 1) << e.glane0;
  census_witness_task(e, task, inputs + (size_t)p * L.n_inputs, hc[0], hc[1], skip_const != 0);
  if (e.status && e.li == 0) atomicMax(status + p, e.status);
}

// thread = (wire, proof): wire value from its staged signal, or from the template when the signal was not produced
__global__ void __launch_bounds__(256) k_witness_gather(const Fr *stage, size_t n_signals, const uint32_t *wmap,
                                                        const Fr *tmpl, Fr *wtns, uint32_t n_wires) {
  uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= n_wires) return;
  wtns[(size_t)blockIdx.y * n_wires + w] = witness_gather_one(stage + (size_t)blockIdx.y * n_signals, wmap[w], tmpl, w);
}

// batched Poseidon hash with the circuit's constants: in [n][T-1] canonical -> out [n] canonical
template <int T>
__global__ void k_poseidon_batch(CensusLayout L, const Fr *consts, const Fr *in, Fr *out, uint32_t n) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  WitnessEnv e;
  e.L = &L; e.consts = consts; e.stage = nullptr; e.status = 0;
  Fr x[T - 1];
  for (int j = 0; j < T - 1; j++) x[j] = in[(size_t)i * (T - 1) + j].to_mont();
  out[i] = poseidon_ex<T>(e, 0, x, false).from_mont();
}

__device__ __forceinline__ Fr ldg_fr8(const Fr *p);
__device__ __forceinline__ void stg_fr8(Fr *p, const Fr &x);

// d[b][i] = w[b][i] - tmpl[i] mod r.  ~91 % of the census wires are identical in every proof (the
// Poseidon2(0,0) blocks below the leaf level, the oldKey = 0 sub-circuits), so d is zero there and the four
// witness MSMs only see the wires that differ:  sum_i w_i P_i = sum_i tmpl_i P_i + sum_i d_i P_i, with the first
// sum computed once per key.  Decided per wire on the data, exact for any tree depth.
__global__ void k_witness_diff(const Fr *wtns, size_t wtns_stride, const Fr *tmpl, Fr *d, size_t d_stride, uint32_t n) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fr w = ldg_fr8(wtns + (size_t)blockIdx.y * wtns_stride + i);
  stg_fr8(d + (size_t)blockIdx.y * d_stride + i, tmpl ? w - ldg_fr8(tmpl + i) : w);
}

__device__ __forceinline__ Fr ldg_fr8(const Fr *p) {
  const uint4 *q = reinterpret_cast<const uint4 *>(p);
  uint4 a = __ldg(q), b = __ldg(q + 1);
  Fr r;
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
__device__ __forceinline__ void stg_fr8(Fr *p, const Fr &x) {
  uint4 *q = reinterpret_cast<uint4 *>(p);
  q[0] = make_uint4(x.v[0], x.v[1], x.v[2], x.v[3]);
  q[1] = make_uint4(x.v[4], x.v[5], x.v[6], x.v[7]);
}

// thread = (row, proof): a = sum coefA * w, b = sum coefB * w, c = a * b  (snarkjs buildABC1)
// Rows are visited in `row_order` (descending number of terms, fixed per key): the 32 rows of a warp then have
// (nearly) the same trip count.  In natural order ncu showed 6.7 of 32 threads active per instruction (census rows
// have 1 .. 319 terms, p99 = 60, and the last third of the domain is empty).
struct CsrDev { const uint32_t *row_ptr, *wire; const Fr *value; };
__global__ void __launch_bounds__(128) k_build_abc(CsrDev A, CsrDev B, const uint32_t *__restrict__ row_order,
                                                   const Fr *wtns, size_t wtns_stride, Fr *abc, uint32_t domain,
                                                   uint32_t n_long_rows) {
  uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x + n_long_rows;   // the first n_long_rows go to k_build_abc_long
  if (tid >= domain) return;
  const uint32_t row = row_order[tid];
  const Fr *w = wtns + (size_t)blockIdx.y * wtns_stride;
  Fr a = Fr::zero(), b = Fr::zero();
  for (uint32_t k = A.row_ptr[row], e = A.row_ptr[row + 1]; k < e; k++) a = a + ldg_fr8(A.value + k) * ldg_fr8(w + A.wire[k]);
  for (uint32_t k = B.row_ptr[row], e = B.row_ptr[row + 1]; k < e; k++) b = b + ldg_fr8(B.value + k) * ldg_fr8(w + B.wire[k]);
  Fr *out = abc + (size_t)blockIdx.y * 3 * domain;
  stg_fr8(out + row, a);
  stg_fr8(out + domain + row, b);
  stg_fr8(out + 2 * (size_t)domain + row, a * b);
}

// The rows with more than ABC_LONG_ROW terms (the head of row_order), one warp each: lanes stride the terms, the 32
// partial sums are folded with shuffles.  A single thread walking a 300-term row was the whole kernel's tail (0.6 ms
// for one proof).
static constexpr uint32_t ABC_LONG_ROW = 48;
__device__ __forceinline__ Fr warp_sum_fr(Fr x) {
  for (int d = 16; d >= 1; d >>= 1) {
    Fr o;
#pragma unroll
    for (int l = 0; l < 8; l++) o.v[l] = __shfl_down_sync(0xffffffffu, x.v[l], d);
    x = x + o;
  }
  return x;
}
__global__ void __launch_bounds__(128) k_build_abc_long(CsrDev A, CsrDev B, const uint32_t *__restrict__ row_order,
                                                        const Fr *wtns, size_t wtns_stride, Fr *abc, uint32_t domain,
                                                        uint32_t n_long_rows) {
  const uint32_t wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (wid >= n_long_rows) return;
  const uint32_t row = row_order[wid];
  const Fr *w = wtns + (size_t)blockIdx.y * wtns_stride;
  Fr a = Fr::zero(), b = Fr::zero();
  for (uint32_t k = A.row_ptr[row] + lane, e = A.row_ptr[row + 1]; k < e; k += 32) a = a + ldg_fr8(A.value + k) * ldg_fr8(w + A.wire[k]);
  for (uint32_t k = B.row_ptr[row] + lane, e = B.row_ptr[row + 1]; k < e; k += 32) b = b + ldg_fr8(B.value + k) * ldg_fr8(w + B.wire[k]);
  a = warp_sum_fr(a);
  b = warp_sum_fr(b);
  if (lane == 0) {
    Fr *out = abc + (size_t)blockIdx.y * 3 * domain;
    stg_fr8(out + row, a);
    stg_fr8(out + domain + row, b);
    stg_fr8(out + 2 * (size_t)domain + row, a * b);
  }
}

// h = a * b - c on the odd coset, back to canonical form (snarkjs joinABC)
__global__ void __launch_bounds__(256) k_join(const Fr *abc, Fr *h, uint32_t domain) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= domain) return;
  const Fr *in = abc + (size_t)blockIdx.y * 3 * domain;
  Fr a = ldg_fr8(in + i), b = ldg_fr8(in + domain + i), c = ldg_fr8(in + 2 * (size_t)domain + i);
  stg_fr8(h + (size_t)blockIdx.y * domain + i, (a * b - c).from_mont());
}

// ---- key sharded over GPUs (one large proof, MSMs split by point range) ---------------------------
// Every rank sums its own point ranges; rank r > 0 then writes its five partial sums into slot r of rank 0's exchange
// buffer (a CUDA IPC / peer mapping: the stores travel over NVLink) and releases the epoch flag; rank 0 waits for the
// flags, adds the partial sums to its own and goes on to the proof assembly.  No NCCL, no host round trip.
static constexpr int SHARD_MAX_RANKS = 16;
struct alignas(256) ShardSlot {
  XYZZ<Fq> g1[4];      // pi_a', pi_b1', pi_c', pi_h partial sums
  XYZZ<Fq2> g2;        // pi_b' partial sum
  uint32_t flag;
};
__global__ void k_shard_publish(ShardSlot *slot, const XYZZ<Fq> *g1, const XYZZ<Fq> *g1h, const XYZZ<Fq2> *g2, uint32_t epoch) {
  const uint32_t t = threadIdx.x;                       // 256 threads: 4 x 32 + 64 words
  uint32_t *dst = reinterpret_cast<uint32_t *>(slot);
  if (t < 96) dst[t] = reinterpret_cast<const uint32_t *>(g1)[t];
  else if (t < 128) dst[t] = reinterpret_cast<const uint32_t *>(g1h)[t - 96];
  else if (t < 192) dst[t] = reinterpret_cast<const uint32_t *>(g2)[t - 128];
  __threadfence_system();
  __syncthreads();
  if (t == 0) asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(&slot->flag), "r"(epoch) : "memory");
}
__global__ void k_shard_combine(ShardSlot *slots, int nranks, uint32_t epoch, XYZZ<Fq> *g1, XYZZ<Fq> *g1h, XYZZ<Fq2> *g2,
                                int *status) {
  const uint32_t t = threadIdx.x;                       // 5 threads: one per partial sum
  __shared__ int bad;
  if (t == 0) {
    bad = 0;
    const long long t0 = clock64();
    for (int r = 1; r < nranks && !bad; r++) {
      uint32_t f;
      do {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(f) : "l"(&slots[r].flag) : "memory");
        if (f != epoch && clock64() - t0 > 8000000000LL) { bad = r; break; }
      } while (f != epoch);
    }
    if (bad) atomicMax(status, 7);                      // ZKB shard timeout marker
  }
  __syncthreads();
  if (bad || t >= 5) return;
  for (int r = 1; r < nranks; r++) {
    if (t < 3) { XYZZ<Fq> p = slots[r].g1[t]; xyzz_add_ni(g1 + t, &p); }
    else if (t == 3) { XYZZ<Fq> p = slots[r].g1[3]; xyzz_add_ni(g1h, &p); }
    else { XYZZ<Fq2> p = slots[r].g2; xyzz_add_ni(g2, &p); }
  }
}

// Witness of a sharded proof: rank r uploads only values [n r / G, n (r+1) / G) from the host, then every rank reads
// the other ranks' slices out of their exchange buffers over NVLink (CUDA IPC mappings) - instead of G processes each
// pushing the whole witness (134 MB at 2^22 constraints) through the host memory system at the same time.
struct WitnessPeers { const Fr *slice[SHARD_MAX_RANKS]; const uint32_t *flag[SHARD_MAX_RANKS]; };
__global__ void k_wt_set_flag(uint32_t *flag, uint32_t epoch) {
  __threadfence_system();
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(epoch) : "memory");
}
__global__ void k_wt_wait(WitnessPeers P, int nranks, uint32_t epoch, int *status) {
  const int r = threadIdx.x;
  if (r >= nranks) return;
  const long long t0 = clock64();
  uint32_t f;
  do {
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(f) : "l"(P.flag[r]) : "memory");
    if (f != epoch && clock64() - t0 > 8000000000LL) { atomicMax(status, 7); return; }
  } while (f != epoch);
}
// wtns[i] = slice of the rank that owns i (plain loads: the data was published with a system-scope release)
__global__ void __launch_bounds__(256) k_wt_gather(WitnessPeers P, int nranks, uint32_t n, Fr *wtns) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int r = (int)(((uint64_t)i * nranks) / n);
  while ((uint32_t)(((uint64_t)n * r) / nranks) > i) r--;
  while ((uint32_t)(((uint64_t)n * (r + 1)) / nranks) <= i) r++;
  const uint32_t lo = (uint32_t)(((uint64_t)n * r) / nranks);
  const uint4 *src = reinterpret_cast<const uint4 *>(P.slice[r] + (i - lo));
  uint4 a, b;
  asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w) : "l"(src));
  asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(src + 1));
  uint4 *dst = reinterpret_cast<uint4 *>(wtns + i);
  dst[0] = a;
  dst[1] = b;
}

// ---------------------------------------------------------------------------------------------
// host objects
// ---------------------------------------------------------------------------------------------
static std::atomic<uint64_t> g_launches{0};   // kernels of this library launched by the proving pipeline

struct Ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
};

static const char *INPUT_NAMES[12] = {"electionId", "nullifier", "voteHash", "sikRoot", "censusRoot", "voteWeight",
                                      "availableWeight", "address", "password", "signature", "censusSiblings",
                                      "sikSiblings"};

// Blinding scalars r, s come from the kernel CSPRNG (getrandom(2)), buffered; uniform in [0, r) by rejection.
// (snarkjs uses crypto.getRandomValues, rapidsnark libsodium's randombytes_buf: both are OS-CSPRNG backed.)
struct OsRandom {
  uint8_t buf[4096];
  size_t pos = sizeof(buf);
  bool fill(void *dst, size_t n) {
    uint8_t *d = (uint8_t *)dst;
    while (n) {
      if (pos == sizeof(buf)) {
        size_t got = 0;
        while (got < sizeof(buf)) {
          ssize_t k = getrandom(buf + got, sizeof(buf) - got, 0);
          if (k < 0) { if (errno == EINTR) continue; return false; }
          got += (size_t)k;
        }
        pos = 0;
      }
      size_t take = sizeof(buf) - pos < n ? sizeof(buf) - pos : n;
      memcpy(d, buf + pos, take);
      pos += take; d += take; n -= take;
    }
    return true;
  }
};

// Per-chunk working set.  Chunks alternate between lanes (each with its own stream), so the latency-bound kernels
// of one chunk (witness, bucket reduction, proof assembly, scans) overlap the IMAD-bound accumulation of the other.
struct Lane {
  cudaStream_t st = nullptr;
  cudaEvent_t done = nullptr;
  // small batches (latency): the witness-side MSMs run on two side streams beside the H pipeline
  cudaStream_t sW = nullptr, s2 = nullptr;
  cudaEvent_t e_fork = nullptr, e_sortW = nullptr, e_W = nullptr, e_2 = nullptr;
  Fr *abc = nullptr, *hs = nullptr, *dw = nullptr, *stage = nullptr;
  MsmSort sortW, sortH;
  MsmAffineWs affH;               // batched-affine pair tree of the H MSM (batch shape only)
  MsmWork<Fq> work1, workH;
  MsmWork<Fq2> work2;
  XYZZ<Fq> *g1out = nullptr, *g1raw = nullptr;   // g1raw / g2raw: per point-range sums before the fold (large keys)
  XYZZ<Fq2> *g2out = nullptr, *g2raw = nullptr;
  XYZZ<Fq> *fin_scratch = nullptr;
  void free_all() {
    cudaFree(abc); cudaFree(hs); cudaFree(dw); cudaFree(g1out); cudaFree(g2out); cudaFree(fin_scratch); cudaFree(stage);
    cudaFree(g1raw); cudaFree(g2raw);
    abc = hs = dw = stage = nullptr; g1out = fin_scratch = g1raw = nullptr; g2out = g2raw = nullptr;
    if (sortW.counts) sortW.free_all();
    if (sortH.counts) sortH.free_all();
    if (work1.buckets) work1.free_all();
    if (workH.buckets) workH.free_all();
    if (work2.buckets) work2.free_all();
    if (affH.pa) affH.free_all();
  }
};
static constexpr int MAX_LANES = 8;

struct Circuit {
  Ctx *ctx = nullptr;
  uint32_t n_vars = 0, n_public = 0, domain = 0, power = 0;
  CensusLayout L;
  uint32_t in_pos[12], in_size[12];
  // device constants
  Fr *consts = nullptr, *hc = nullptr, *tmpl = nullptr;
  uint32_t *wmap = nullptr;        // wire -> circom signal (the wasm's witness table)
  uint32_t *csr_buf = nullptr, *row_order = nullptr;
  uint32_t n_long_rows = 0;        // head of row_order: rows with more than ABC_LONG_ROW terms (one warp each)
  Fr *csr_val = nullptr;
  CsrDev csrA, csrB;
  NttPlan ntt;
  MsmTable<Fq> tabA, tabB1, tabC, tabH;
  MsmTable<Fq2> tabB2;
  Affine<Fq> *fix1 = nullptr;      // alpha1, beta1
  Affine<Fq2> *fix2 = nullptr;     // beta2
  Affine<Fq> *d1tab = nullptr;
  Affine<Fq2> *d2tab = nullptr;
  // blinding override (tests): when set every proof uses these r, s
  bool fixed_rs = false;
  uint32_t fr[8], fs[8];
  // dense: the proof-independent-wire shortcut is off - every SMT level is hashed and the four witness
  // MSMs run over the full witness, as snarkjs / rapidsnark do (measurement aid: zkb_load_circuit_ex flag 1)
  bool dense = false;
  // generic witness path: a wasm that is not the census program (or ZKB_LOAD_GENERIC_WITNESS) runs as
  // a tape extracted from the wasm itself by symbolic execution; has_witness = the circuit can make witnesses at all
  bool has_witness = false, generic = false;
  TapeDev tape;
  std::vector<WitnessProgram::Input> gen_inputs;
  uint32_t gen_first_signal = 0;
  // H MSM bucket lists through the batched-affine pair tree (ZKB_AFFINE=1).  Off by default: measured on B200 it
  // executes 25 % fewer field products than the XYZZ kernel but moves ~320 B per addition through HBM and ends up
  // within 2 % of it, at 30 GB of extra workspace.
  bool affine = false;
  // batch workspace
  uint32_t cap = 0;                // proofs resident at once (witness group)
  uint32_t chunk = 0;              // proofs per NTT/MSM chunk
  Fr *inputs = nullptr, *wtns = nullptr, *rs = nullptr;
  Fr *gen_slots = nullptr;         // generic witness: value slots of a whole group (cap x n_slots), see run_witness_group
  uint32_t gen_slots_cap = 0;
  Lane lanes[MAX_LANES];
  int n_lanes = 0;
  XYZZ<Fq> *tconst1 = nullptr;     // sum_i tmpl_i * {A_i, B1_i, C_i}
  XYZZ<Fq2> *tconst2 = nullptr;    // sum_i tmpl_i * B2_i
  int *status = nullptr;
  // Large keys: every MSM is cut by point range into sub-MSMs of 2^17 points that run as extra batch items (a single
  // 4 M-point MSM would otherwise be 32,768 threads with 2,048 sequential adds each); 1 range for census-sized keys.
  uint32_t subW = 1, subH = 1, nsubW = 0, nsubH = 0, n_pad = 0;
  // sharded key: this rank owns ranges [loW, loW + cntW) of the witness MSMs and [loH, loH + cntH) of the H MSM
  uint32_t loW = 0, cntW = 1, loH = 0, cntH = 1;
  int shard_rank = 0, shard_n = 1;
  ShardSlot *xbuf = nullptr;       // rank 0: SHARD_MAX_RANKS slots (exported)
  ShardSlot *root_x = nullptr;     // where this rank publishes
  bool root_is_ipc = false;
  uint32_t epoch = 0;
  // witness slices of a sharded proof (see k_wt_gather): own slice buffer (+ flag word behind it), the peers' mappings
  uint8_t *wt_buf = nullptr;
  size_t wt_flag_off = 0;
  uint8_t *h_wt = nullptr;         // pinned staging of the own slice
  const uint8_t *wt_peer[SHARD_MAX_RANKS] = {};
  bool wt_peer_ipc[SHARD_MAX_RANKS] = {};
  int wt_attached = 0;
  uint32_t wt_epoch = 0;
  MsmCfg cfgW, cfgH;              // window sizes: witness MSMs (sparse after the template difference), H MSM (dense)
  uint8_t *out = nullptr;          // device results
  uint8_t *h_out = nullptr;        // pinned
  Fr *h_inputs = nullptr;          // pinned
  Fr *h_rs = nullptr;              // pinned
  int *h_status = nullptr;         // pinned
  std::mutex mu;
  uint32_t last_chunk_m = 0;
  bool last_affine = false;
  int last_lane = 0;
  OsRandom rng;
  // stage timing of the last device pass (ms)
  float t_witness = 0, t_abc = 0, t_ntt = 0, t_msm = 0, t_fin = 0;
  size_t out_stride() const { return 256 + 32 * (size_t)n_public; }
};

static bool random_fr(OsRandom &g, uint32_t out[8]) {
  static const uint32_t RMOD[8] = {0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u,
                                   0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
  for (;;) {
    if (!g.fill(out, 32)) return false;
    out[7] &= 0x3fffffffu;
    for (int i = 7; i >= 0; i--) {
      if (out[i] < RMOD[i]) return true;
      if (out[i] > RMOD[i]) break;
    }
  }
}

// the pair tree pays off when a launch has enough lists to fill the GPU; a single proof keeps the XYZZ kernel
static constexpr uint32_t AFFINE_MIN_ITEMS = 16, AFFINE_BATCH = 64;

static uint32_t env_u32(const char *name, uint32_t dflt) {
  const char *v = getenv(name);
  if (!v || !*v) return dflt;
  long x = atol(v);
  return x > 0 ? (uint32_t)x : dflt;
}

// proofs per NTT/MSM chunk: 128 for census-sized keys (2^17 rows), fewer for larger ones so that a chunk's working
// set (45 MB per proof at 2^17) stays bounded; a 2^22-row circuit runs one proof per chunk
static uint32_t default_chunk(const Circuit *c) {
  uint32_t fit = (uint32_t)(((size_t)1 << 24) / (c->domain ? c->domain : 1));
  if (fit < 1) fit = 1;
  if (fit > 128) fit = 128;
  return env_u32("ZKB_CHUNK", fit);
}

static void free_workspace(Circuit *c) {
  cudaFree(c->inputs); cudaFree(c->wtns); cudaFree(c->rs); cudaFree(c->status); cudaFree(c->out); cudaFree(c->gen_slots);
  c->inputs = c->wtns = c->rs = nullptr; c->status = nullptr; c->out = nullptr; c->gen_slots = nullptr; c->gen_slots_cap = 0;
  if (c->h_out) cudaFreeHost(c->h_out);
  if (c->h_inputs) cudaFreeHost(c->h_inputs);
  if (c->h_rs) cudaFreeHost(c->h_rs);
  if (c->h_status) cudaFreeHost(c->h_status);
  c->h_out = nullptr; c->h_inputs = nullptr; c->h_rs = nullptr; c->h_status = nullptr;
  for (int i = 0; i < MAX_LANES; i++) c->lanes[i].free_all();
  c->cap = 0;
  c->chunk = 0;
}

static int alloc_workspace(Circuit *c, uint32_t cap, uint32_t chunk);
static int ensure_workspace(Circuit *c, uint32_t cap, uint32_t chunk) {
  if (c->cap >= cap && c->chunk >= chunk) return ZKB_OK;
  free_workspace(c);
  int rc = alloc_workspace(c, cap, chunk);
  if (rc) { free_workspace(c); cudaGetLastError(); return rc; }   // (clears the allocation error: it is reported through rc)
  c->cap = cap;
  c->chunk = chunk;
  return ZKB_OK;
}

static int alloc_workspace(Circuit *c, uint32_t cap, uint32_t chunk) {
  // (re)allocate everything; sizes are small next to the 180 GB of HBM
  // more than one lane only pays when there is more than one chunk
  int want = (int)env_u32("ZKB_LANES", 4);   // measured on B200, 1,024-proof batch: 2 lanes 1312, 4 lanes 1356 proofs/s
  if (want > MAX_LANES) want = MAX_LANES;
  c->n_lanes = cap > chunk ? want : 1;
  CKR(cudaMalloc(&c->inputs, (size_t)cap * c->L.n_inputs * 32), "alloc inputs");
  CKR(cudaMalloc(&c->wtns, (size_t)cap * c->n_vars * 32), "alloc witness");
  CKR(cudaMalloc(&c->rs, (size_t)cap * 64), "alloc rs");
  CKR(cudaMalloc(&c->status, (size_t)cap * 4), "alloc status");
  CKR(cudaMalloc(&c->out, (size_t)cap * c->out_stride()), "alloc out");
  CKR(cudaMallocHost(&c->h_out, (size_t)cap * c->out_stride()), "alloc pinned out");
  CKR(cudaMallocHost(&c->h_inputs, (size_t)cap * c->L.n_inputs * 32), "alloc pinned inputs");
  CKR(cudaMallocHost(&c->h_rs, (size_t)cap * 64), "alloc pinned rs");
  CKR(cudaMallocHost(&c->h_status, (size_t)cap * 4), "alloc pinned status");
  for (int i = 0; i < c->n_lanes; i++) {
    Lane &ln = c->lanes[i];
    if (!ln.st) {
      CKR(cudaStreamCreateWithFlags(&ln.st, cudaStreamNonBlocking), "lane stream");
      CKR(cudaEventCreateWithFlags(&ln.done, cudaEventDisableTiming), "lane event");
      CKR(cudaStreamCreateWithFlags(&ln.sW, cudaStreamNonBlocking), "lane side stream");
      CKR(cudaStreamCreateWithFlags(&ln.s2, cudaStreamNonBlocking), "lane side stream");
      for (cudaEvent_t *e : {&ln.e_fork, &ln.e_sortW, &ln.e_W, &ln.e_2})
        CKR(cudaEventCreateWithFlags(e, cudaEventDisableTiming), "lane event");
    }
    CKR(cudaMalloc(&ln.abc, (size_t)chunk * 3 * c->domain * 32), "alloc abc");
    CKR(cudaMalloc(&ln.hs, (size_t)chunk * c->domain * 32), "alloc h");
    CKR(cudaMalloc(&ln.dw, (size_t)chunk * c->n_pad * 32), "alloc witness diff");
    CKR(cudaMemset(ln.dw, 0, (size_t)chunk * c->n_pad * 32), "clear witness diff");   // the padding stays zero
    if (c->has_witness)
      CKR(cudaMalloc(&ln.stage, (size_t)chunk * (c->generic ? c->tape.n_slots : c->L.n_signals) * 32), "alloc witness staging");
    CKR(cudaMalloc(&ln.g1out, (size_t)chunk * 4 * sizeof(XYZZ<Fq>)), "alloc g1out");
    CKR(cudaMalloc(&ln.g2out, (size_t)chunk * sizeof(XYZZ<Fq2>)), "alloc g2out");
    CKR(cudaMalloc(&ln.fin_scratch, (size_t)chunk * 30 * sizeof(XYZZ<Fq>)), "alloc finalize scratch");
    CKR(ln.sortW.alloc(c->nsubW, chunk * c->subW, c->cfgW), "alloc sortW");
    CKR(ln.sortH.alloc(c->nsubH, chunk * c->subH, c->cfgH), "alloc sortH");
    // (the pair tree's workspace is allocated on first use, see run_prove_chunk)
    CKR(ln.work1.alloc(chunk * c->subW * 3, c->cfgW), "alloc msm work g1");
    CKR(ln.workH.alloc(chunk * c->subH, c->cfgH), "alloc msm work g1 (H)");
    CKR(ln.work2.alloc(chunk * c->subW, c->cfgW), "alloc msm work g2");
    if (c->subW > 1 || c->subH > 1) {
      CKR(cudaMalloc(&ln.g1raw, (size_t)chunk * (3 * c->subW + c->subH) * sizeof(XYZZ<Fq>)), "alloc g1 range sums");
      CKR(cudaMalloc(&ln.g2raw, (size_t)chunk * c->subW * sizeof(XYZZ<Fq2>)), "alloc g2 range sums");
    }
  }
  return ZKB_OK;
}

// witness for proofs [first, first + n) of c->inputs, n <= chunk, staged in lane ln
static int run_witness(Circuit *c, Lane &ln, uint32_t first, uint32_t n, cudaStream_t st) {
  CKR(cudaMemsetAsync(c->status + first, 0, (size_t)n * 4, st), "memset status");
  if (c->generic) {
    CKR(tape_eval(c->tape, c->inputs + (size_t)first * c->L.n_inputs, ln.stage, c->wtns + (size_t)first * c->n_vars,
                  c->status + first, n, st), "generic witness (tape)");
    g_launches += 2;
    return ZKB_OK;
  }
  CKR(cudaMemsetAsync(ln.stage, 0, (size_t)n * c->L.n_signals * 32, st), "memset staging");
  k_witness<<<dim3((n * COOP_LANES + 31) / 32, WITNESS_TASKS), 32, 0, st>>>(c->L, c->consts, c->hc,
                                                               c->inputs + (size_t)first * c->L.n_inputs, ln.stage,
                                                               c->status + first, n, c->dense ? 0 : 1);
  k_witness_gather<<<dim3((c->n_vars + 255) / 256, n), 256, 0, st>>>(ln.stage, c->L.n_signals, c->wmap, c->tmpl,
                                                                     c->wtns + (size_t)first * c->n_vars, c->n_vars);
  g_launches += 2;
  const cudaError_t le = cudaGetLastError();
  return le == cudaSuccess ? ZKB_OK : cuda_fail(le, "witness launch");
}

// witness (optional) + Groth16 for proofs [first, first + m), m <= chunk, on lane ln
static int run_prove_chunk(Circuit *c, Lane &ln, uint32_t first, uint32_t m, bool with_witness, cudaEvent_t *ev) {
  cudaStream_t st = ln.st;
  if (with_witness) { int rc = run_witness(c, ln, first, m, st); if (rc) return rc; }
  if (ev) cudaEventRecord(ev[1], st);
  const Fr *w = c->wtns + (size_t)first * c->n_vars;
  // batch item p * sub + s = point range s of proof p (sub = 1: one item per proof)
  // (a sharded key proves one proof at a time and only sorts / sums its own ranges [lo, lo + cnt))
  const uint32_t sW = c->cntW, sH = c->cntH;
  const bool splitW = c->subW > 1, splitH = c->subH > 1;
  const size_t strW = splitW ? (size_t)c->cfgW.windows * c->nsubW : 0, strH = splitH ? (size_t)c->cfgH.windows * c->nsubH : 0;
  if (c->shard_n > 1 && m != 1) { set_error("a sharded key proves one witness per call"); return ZKB_ERROR; }
  MsmTable<Fq> tabs[3] = {c->tabA, c->tabB1, c->tabC};
  XYZZ<Fq> *outW = splitW ? ln.g1raw : ln.g1out;
  XYZZ<Fq> *outH = splitH ? ln.g1raw + (size_t)3 * c->chunk * c->subW : ln.g1out + (size_t)3 * c->chunk;
  XYZZ<Fq2> *out2 = splitW ? ln.g2raw : ln.g2out;
  dim3 g1((c->domain + 127) / 128, m);
  dim3 g2((c->domain + 255) / 256, m);
  const bool affineH = c->affine && m * sH >= AFFINE_MIN_ITEMS;
  if (affineH && !ln.affH.pa) {
    // first batch-shaped chunk on this lane: the pair tree's workspace, AFFINE_BATCH items at a time (7.6 GB at 2^17
    // points per item; a chunk of 128 proofs runs as two sub-batches)
    const uint32_t items = c->chunk * c->subH < AFFINE_BATCH ? c->chunk * c->subH : AFFINE_BATCH;
    cudaError_t e = ln.affH.alloc(c->nsubH * (uint32_t)c->cfgH.windows, items, c->cfgH);
    if (e != cudaSuccess) { ln.affH.free_all(); cudaGetLastError(); return cuda_fail(e, "alloc H pair tree"); }
  }
  c->last_affine = affineH;
  if (affineH) g_launches += msm_affine_launches(ln.affH, m * sH) - 2;   // instead of the two XYZZ accumulate kernels of the H MSM
  static const uint32_t fork_max = env_u32("ZKB_FORK_MAX", 32);
  if (!ev && m <= fork_max) {
    // Latency shape: after the witness, three independent pipelines.  st: A/B/C vectors -> coset transforms -> h ->
    // H MSM.  sW: witness difference -> digit sort -> A, B1, C sums.  s2: (after that sort) the G2 sum.
    cudaEventRecord(ln.e_fork, st);
    cudaStreamWaitEvent(ln.sW, ln.e_fork, 0);
    cudaStreamWaitEvent(ln.s2, ln.e_fork, 0);
    k_witness_diff<<<dim3((c->n_vars + 255) / 256, m), 256, 0, ln.sW>>>(w, c->n_vars, c->tconst1 ? c->tmpl : nullptr, ln.dw, c->n_pad, c->n_vars);
    CKR(ln.sortW.run(ln.dw + (size_t)c->loW * c->nsubW, c->nsubW, m * sW, ln.sW), "sort witness digits");
    cudaEventRecord(ln.e_sortW, ln.sW);
    cudaStreamWaitEvent(ln.s2, ln.e_sortW, 0);
    CKR(msm_accumulate<Fq2>(ln.sortW, &c->tabB2, 1, m * sW, ln.work2, 0, ln.s2, strW, sW), "msm accumulate g2 (B2)");
    CKR(msm_reduce<Fq2>(ln.work2, 0, m * sW, out2, ln.s2), "msm reduce g2");
    if (splitW) CKR(msm_fold_subs<Fq2>(out2, ln.g2out, m, sW, 1, ln.s2), "fold g2 ranges");
    cudaEventRecord(ln.e_2, ln.s2);
    CKR(msm_accumulate<Fq>(ln.sortW, tabs, 3, m * sW, ln.work1, 0, ln.sW, strW, sW), "msm accumulate g1 (A,B1,C)");
    CKR(msm_reduce<Fq>(ln.work1, 0, 3 * m * sW, outW, ln.sW), "msm reduce g1");
    if (splitW) CKR(msm_fold_subs<Fq>(outW, ln.g1out, m, sW, 3, ln.sW), "fold g1 ranges");
    cudaEventRecord(ln.e_W, ln.sW);
    if (c->n_long_rows)
      k_build_abc_long<<<dim3((c->n_long_rows * 32 + 127) / 128, m), 128, 0, st>>>(c->csrA, c->csrB, c->row_order, w, c->n_vars,
                                                                                  ln.abc, c->domain, c->n_long_rows);
    k_build_abc<<<g1, 128, 0, st>>>(c->csrA, c->csrB, c->row_order, w, c->n_vars, ln.abc, c->domain, c->n_long_rows);
    CKR(c->ntt.dif(ln.abc, 3 * m, c->domain, true, true, st), "ntt dif");
    CKR(c->ntt.dit(ln.abc, 3 * m, c->domain, false, st), "ntt dit");
    k_join<<<g2, 256, 0, st>>>(ln.abc, ln.hs, c->domain);
    CKR(ln.sortH.run(ln.hs + (size_t)c->loH * c->nsubH, c->nsubH, m * sH, st), "sort h digits");
    if (affineH) CKR(msm_accumulate_affine(ln.sortH, c->tabH, m * sH, ln.workH, 0, ln.affH, st, strH, sH), "msm accumulate g1 (H, pair tree)");
    else CKR(msm_accumulate<Fq>(ln.sortH, &c->tabH, 1, m * sH, ln.workH, 0, st, strH, sH), "msm accumulate g1 (H)");
    CKR(msm_reduce<Fq>(ln.workH, 0, m * sH, outH, st), "msm reduce g1 (H)");
    if (splitH) CKR(msm_fold_subs<Fq>(outH, ln.g1out + (size_t)3 * c->chunk, m, sH, 1, st), "fold g1 (H) ranges");
    cudaStreamWaitEvent(st, ln.e_W, 0);
    cudaStreamWaitEvent(st, ln.e_2, 0);
    g_launches += 1 + 4 + 1 + 1 + (splitW ? 2 : 0) + (splitH ? 1 : 0) + (c->n_long_rows ? 1 : 0);
  } else {
    if (c->n_long_rows)
      k_build_abc_long<<<dim3((c->n_long_rows * 32 + 127) / 128, m), 128, 0, st>>>(c->csrA, c->csrB, c->row_order, w, c->n_vars,
                                                                                  ln.abc, c->domain, c->n_long_rows);
    k_build_abc<<<g1, 128, 0, st>>>(c->csrA, c->csrB, c->row_order, w, c->n_vars, ln.abc, c->domain, c->n_long_rows);
    g_launches += 1 + 4 + 1 + (c->n_long_rows ? 1 : 0);   // build_abc (+ its long-row kernel), 2 DIF + 2 DIT passes, join
    if (ev) cudaEventRecord(ev[2], st);
    CKR(c->ntt.dif(ln.abc, 3 * m, c->domain, true, true, st), "ntt dif");
    CKR(c->ntt.dit(ln.abc, 3 * m, c->domain, false, st), "ntt dit");
    k_join<<<g2, 256, 0, st>>>(ln.abc, ln.hs, c->domain);
    if (ev) cudaEventRecord(ev[3], st);
    k_witness_diff<<<dim3((c->n_vars + 255) / 256, m), 256, 0, st>>>(w, c->n_vars, c->tconst1 ? c->tmpl : nullptr, ln.dw, c->n_pad, c->n_vars);
    g_launches += 1;
    CKR(ln.sortW.run(ln.dw + (size_t)c->loW * c->nsubW, c->nsubW, m * sW, st), "sort witness digits");
    CKR(ln.sortH.run(ln.hs + (size_t)c->loH * c->nsubH, c->nsubH, m * sH, st), "sort h digits");
    if (ev) cudaEventRecord(ev[4], st);
    // bucket sums: G1 over the witness difference (A, B1, C share one sort), H over h, G2 (B2) over the witness difference
    CKR(msm_accumulate<Fq>(ln.sortW, tabs, 3, m * sW, ln.work1, 0, st, strW, sW), "msm accumulate g1 (A,B1,C)");
    if (affineH) CKR(msm_accumulate_affine(ln.sortH, c->tabH, m * sH, ln.workH, 0, ln.affH, st, strH, sH), "msm accumulate g1 (H, pair tree)");
    else CKR(msm_accumulate<Fq>(ln.sortH, &c->tabH, 1, m * sH, ln.workH, 0, st, strH, sH), "msm accumulate g1 (H)");
    if (ev) cudaEventRecord(ev[5], st);
    CKR(msm_accumulate<Fq2>(ln.sortW, &c->tabB2, 1, m * sW, ln.work2, 0, st, strW, sW), "msm accumulate g2 (B2)");
    if (ev) cudaEventRecord(ev[6], st);
    CKR(msm_reduce<Fq>(ln.work1, 0, 3 * m * sW, outW, st), "msm reduce g1");
    CKR(msm_reduce<Fq>(ln.workH, 0, m * sH, outH, st), "msm reduce g1 (H)");
    CKR(msm_reduce<Fq2>(ln.work2, 0, m * sW, out2, st), "msm reduce g2");
    if (splitW) {
      CKR(msm_fold_subs<Fq>(outW, ln.g1out, m, sW, 3, st), "fold g1 ranges");
      CKR(msm_fold_subs<Fq2>(out2, ln.g2out, m, sW, 1, st), "fold g2 ranges");
      g_launches += 2;
    }
    if (splitH) {
      CKR(msm_fold_subs<Fq>(outH, ln.g1out + (size_t)3 * c->chunk, m, sH, 1, st), "fold g1 (H) ranges");
      g_launches += 1;
    }
  }
  if (c->shard_n > 1) {
    c->epoch++;
    g_launches += 1;
    if (c->shard_rank != 0) {
      k_shard_publish<<<1, 256, 0, st>>>(c->root_x + c->shard_rank, ln.g1out, ln.g1out + (size_t)3 * c->chunk, ln.g2out, c->epoch);
      if (ev) { cudaEventRecord(ev[7], st); cudaEventRecord(ev[8], st); }
      CKR(cudaGetLastError(), "shard publish");
      return ZKB_OK;                                     // rank 0 assembles the proof
    }
    k_shard_combine<<<1, 32, 0, st>>>(c->xbuf, c->shard_n, c->epoch, ln.g1out, ln.g1out + (size_t)3 * c->chunk, ln.g2out,
                                      c->status + first);
  }
  if (ev) cudaEventRecord(ev[7], st);
  g_launches += 2 * msm_sort_launches() + 4 + 2 + 6;   // 2 sorts, (2 + 1) x 2 accumulate, 3 x 2 reduce
  FinalizeParams P;
  P.g1 = ln.g1out;
  P.g1h = ln.g1out + (size_t)3 * c->chunk;
  P.g2 = ln.g2out;
  P.rs = c->rs + (size_t)first * 2;
  P.wtns = w;
  P.wtns_stride = c->n_vars;
  P.alpha1 = c->fix1;
  P.beta1 = c->fix1 + 1;
  P.d1tab = c->d1tab;
  P.beta2 = c->fix2;
  P.d2tab = c->d2tab;
  P.tconst1 = c->tconst1;
  P.tconst2 = c->tconst2;
  P.scratch = ln.fin_scratch;
  P.out = c->out + (size_t)first * c->out_stride();
  P.n_public = c->n_public;
  P.n = m;
  CKR(launch_finalize(P, st), "finalize");
  g_launches += 1;
  if (ev) cudaEventRecord(ev[8], st);
  CKR(cudaGetLastError(), "prove chunk launch");
  return ZKB_OK;
}

// full device pass over the n resident proofs; stage_ms (optional, 8 floats): witness, build_abc, ntt+join,
// msm sort, msm accumulate G1, msm accumulate G2, msm reduce, finalize - summed over chunks.  The instrumented pass is
// serial (one lane); without stage_ms chunks alternate over the lanes and overlap.
// Generic witness for proofs [0, n) in ONE launch on stream st.  A witness program is a chain of ~10^5 dependent levels:
// its run time hardly depends on how many proofs (warps) run it, so a whole group is evaluated at once instead of
// chunk by chunk (8 x 85 ms -> ~90 ms for 1,024 proofs of the census program).
static int run_witness_group(Circuit *c, uint32_t n, cudaStream_t st) {
  if (c->gen_slots_cap < n) {
    cudaFree(c->gen_slots);
    c->gen_slots = nullptr;
    c->gen_slots_cap = 0;
    CKR(cudaMalloc(&c->gen_slots, (size_t)n * c->tape.n_slots * 32), "alloc witness slots");
    c->gen_slots_cap = n;
  }
  CKR(cudaMemsetAsync(c->status, 0, (size_t)n * 4, st), "memset status");
  CKR(tape_eval(c->tape, c->inputs, c->gen_slots, c->wtns, c->status, n, st), "generic witness (tape)");
  g_launches += 2;
  return ZKB_OK;
}

static int prove_resident(Circuit *c, uint32_t n, bool with_witness, float *stage_ms) {
  cudaStream_t st = c->ctx->stream;
  struct EvBag {                         // destroyed on every exit path
    std::vector<cudaEvent_t> v;
    ~EvBag() { for (auto e : v) cudaEventDestroy(e); }
  } bag;
  std::vector<cudaEvent_t> &evs = bag.v;
  auto newev = [&]() { cudaEvent_t e; cudaEventCreate(&e); evs.push_back(e); return e; };
  cudaEvent_t start;
  CKR(cudaEventCreateWithFlags(&start, cudaEventDisableTiming), "event");
  evs.push_back(start);
  if (with_witness && c->generic && n > c->chunk && !stage_ms) {
    int rc = run_witness_group(c, n, st);
    if (rc) return rc;
    with_witness = false;
  }
  cudaEventRecord(start, st);            // inputs / blinding were queued on the context stream
  // an instrumented pass runs on one lane: event-bracketed stage times are then not shared with another stream
  const int n_lanes = stage_ms ? 1 : c->n_lanes;
  for (int i = 0; i < n_lanes; i++) cudaStreamWaitEvent(c->lanes[i].st, start, 0);
  std::vector<cudaEvent_t> chunk_ev;
  uint32_t k = 0;
  for (uint32_t first = 0; first < n; first += c->chunk, k++) {
    uint32_t m = n - first < c->chunk ? n - first : c->chunk;
    Lane &ln = c->lanes[k % n_lanes];
    cudaEvent_t ev[9];
    if (stage_ms) {
      for (int i = 0; i < 9; i++) { ev[i] = newev(); chunk_ev.push_back(ev[i]); }
      cudaEventRecord(ev[0], ln.st);
    }
    int rc = run_prove_chunk(c, ln, first, m, with_witness, stage_ms ? ev : nullptr);
    if (rc) { cudaDeviceSynchronize(); return rc; }    // side streams of a forked chunk are joined before the error returns
    c->last_chunk_m = m;
    c->last_lane = (int)(k % n_lanes);
  }
  for (int i = 0; i < n_lanes; i++) {
    cudaEventRecord(c->lanes[i].done, c->lanes[i].st);
    cudaStreamWaitEvent(st, c->lanes[i].done, 0);
  }
  CKR(cudaStreamSynchronize(st), "prove");
  if (stage_ms) {
    for (int i = 0; i < 8; i++) stage_ms[i] = 0;
    for (size_t q = 0; q + 8 < chunk_ev.size(); q += 9)
      for (int i = 0; i < 8; i++) { float t; cudaEventElapsedTime(&t, chunk_ev[q + i], chunk_ev[q + i + 1]); stage_ms[i] += t; }
  }
  return ZKB_OK;
}

static bool fill_blinding(Circuit *c, uint32_t n) {
  for (uint32_t i = 0; i < n; i++) {
    uint32_t *r = c->h_rs[2 * i].v, *s = c->h_rs[2 * i + 1].v;
    if (c->fixed_rs) { memcpy(r, c->fr, 32); memcpy(s, c->fs, 32); }
    else if (!random_fr(c->rng, r) || !random_fr(c->rng, s)) { set_error("getrandom failed: no blinding scalars"); return false; }
  }
  return true;
}

// inputs.json -> canonical values in main-signal order (n_inputs x 8 u32)
static int pack_inputs(const Circuit *c, const char *json, size_t len, uint32_t *dst, std::string &err) {
  std::map<std::string, std::vector<uint32_t>> m;
  if (!parse_inputs_json(json, len, m, err)) return ZKB_ERROR;
  size_t total = 0;
  if (c->generic) {                 // names resolved through the wasm's own input hashmap (FNV-1a-64 of the name)
    for (auto &kv : m) {
      const uint64_t h = fnv1a64_name(kv.first);
      const WitnessProgram::Input *in = nullptr;
      for (auto &g : c->gen_inputs) if (g.hash == h) in = &g;
      if (!in) { err = "inputs: unexpected signal (not an input of the circuit): " + kv.first; return ZKB_ERROR; }
      if (kv.second.size() != (size_t)in->size * 8) { err = "inputs: wrong number of values for " + kv.first; return ZKB_ERROR; }
      memcpy(dst + (size_t)(in->pos - c->gen_first_signal) * 8, kv.second.data(), kv.second.size() * 4);
      total += in->size;
    }
    if (m.size() != c->gen_inputs.size() || total != c->L.n_inputs) { err = "inputs: signal not found (an input of the circuit is missing)"; return ZKB_ERROR; }
    return ZKB_OK;
  }
  for (int k = 0; k < 12; k++) {
    auto it = m.find(INPUT_NAMES[k]);
    if (it == m.end()) { err = std::string("inputs: signal not found: ") + INPUT_NAMES[k]; return ZKB_ERROR; }
    if (it->second.size() != (size_t)c->in_size[k] * 8) { err = std::string("inputs: wrong number of values for ") + INPUT_NAMES[k]; return ZKB_ERROR; }
    memcpy(dst + (size_t)(c->in_pos[k] - 1) * 8, it->second.data(), it->second.size() * 4);
    total += c->in_size[k];
  }
  if (m.size() != 12) { err = "inputs: unexpected signal (not an input of the census circuit)"; return ZKB_ERROR; }
  if (total != c->L.n_inputs) { err = "inputs: size mismatch"; return ZKB_ERROR; }
  return ZKB_OK;
}

template <class T>
static cudaError_t upload(T **dptr, const void *src, size_t bytes) {
  cudaError_t e = cudaMalloc(dptr, bytes ? bytes : 16);
  if (e != cudaSuccess) return e;
  return cudaMemcpy(*dptr, src, bytes, cudaMemcpyHostToDevice);
}

static int load_circuit(Ctx *ctx, const uint8_t *zkey, size_t zkey_len, const uint8_t *wasm, size_t wasm_len,
                        Circuit **out, int shard_rank = 0, int shard_n = 1, uint32_t flags = 0) {
  CKR(cudaSetDevice(ctx->device), "set device");
  std::string err;
  ZkeyView z;
  if (!parse_zkey(zkey, zkey_len, z, err)) { set_error(err); return ZKB_ERROR; }
  std::unique_ptr<Circuit> c(new Circuit());
  c->ctx = ctx;
  c->n_vars = z.n_vars; c->n_public = z.n_public; c->domain = z.domain; c->power = z.power;
  c->dense = (flags & 1u) != 0 || env_u32("ZKB_DENSE", 0) != 0;
  { const char *a = getenv("ZKB_AFFINE"); c->affine = a && a[0] == '1'; }
  cudaStream_t st = ctx->stream;
  memset(&c->L, 0, sizeof c->L);

  // Which witness generator: the hand-written census kernel when the wasm is recognised as census.circom (layout
  // and input map checked below), the generic tape otherwise (or when flag 2 / ZKB_GENERIC_WITNESS asks for it).
  bool want_generic = wasm && ((flags & 2u) != 0 || env_u32("ZKB_GENERIC_WITNESS", 0) != 0);
  std::string census_why;
  if (wasm && !want_generic) {
    WasmCircuit w;
    auto is_census = [&]() -> bool {
      if (!parse_circom_wasm(wasm, wasm_len, w, census_why)) return false;
      uint32_t pos = 0, size = 0;
      if (!wasm_input_lookup(w, "censusSiblings", pos, size) || size < 4) { census_why = "censusSiblings missing"; return false; }
      if (!census_layout_build(c->L, size)) { census_why = "unsupported number of levels"; return false; }
      c->L.n_wires = w.n_wires;
      const uint32_t expect_pos[12] = {c->L.electionId, c->L.nullifier, c->L.voteHash, c->L.sikRoot, c->L.censusRoot,
                                       c->L.voteWeight, c->L.availableWeight, c->L.address, c->L.password,
                                       c->L.signature, c->L.censusSiblings, c->L.sikSiblings};
      const uint32_t expect_size[12] = {2, 1, 2, 1, 1, 1, 1, 1, 1, 1, size, size};
      for (int k = 0; k < 12; k++)
        if (!wasm_input_lookup(w, INPUT_NAMES[k], c->in_pos[k], c->in_size[k]) || c->in_pos[k] != expect_pos[k] ||
            c->in_size[k] != expect_size[k]) { census_why = std::string("input ") + INPUT_NAMES[k]; return false; }
      if (w.n_inputs != c->L.n_inputs || w.witness_map.back() >= c->L.n_signals || w.witness_map[0] != 0) { census_why = "signal layout mismatch"; return false; }
      if (z.n_public != 8) { census_why = "census circuit has 8 public signals; zkey says otherwise"; return false; }
      return true;
    };
    if (!is_census()) { want_generic = true; memset(&c->L, 0, sizeof c->L); }
    else {
    if (w.n_wires != z.n_vars) { set_error("zkey and wasm disagree on the number of wires"); return ZKB_INVALID_WITNESS_LENGTH; }
    c->has_witness = true;
    // constants
    std::vector<uint32_t> cbuf;
    std::vector<int> forms;
    uint32_t off = 0;
    for (int ti = 0; ti < 3; ti++) {
      uint32_t *offs[4] = {&c->L.pex[ti].c_off, &c->L.pex[ti].s_off, &c->L.pex[ti].m_off, &c->L.pex[ti].p_off};
      for (int k = 0; k < 4; k++) {
        *offs[k] = off;
        for (auto &pc : w.poseidon[ti].tab[k]) { cbuf.insert(cbuf.end(), pc.v, pc.v + 8); forms.push_back(pc.form); }
        off += (uint32_t)w.poseidon[ti].tab[k].size();
      }
    }
    int *dforms = nullptr;
    CKR(upload(&c->consts, cbuf.data(), cbuf.size() * 4), "upload poseidon constants");
    CKR(upload(&dforms, forms.data(), forms.size() * 4), "upload forms");
    k_consts_to_mont<<<(off + 127) / 128, 128, 0, st>>>(c->consts, dforms, off);
    CKR(upload(&c->wmap, w.witness_map.data(), w.witness_map.size() * 4), "upload witness map");
    CKR(cudaMalloc(&c->hc, 64), "alloc hc");
    k_hash_consts<<<1, 1, 0, st>>>(c->L, c->consts, c->hc);
    // template witness: the program itself on all-zero inputs, constant blocks included
    CKR(cudaMalloc(&c->tmpl, (size_t)c->n_vars * 32), "alloc template");
    Fr *zin = nullptr, *zstage = nullptr;
    int *zst = nullptr;
    CKR(cudaMalloc(&zin, (size_t)c->L.n_inputs * 32), "alloc");
    CKR(cudaMalloc(&zstage, (size_t)c->L.n_signals * 32), "alloc");
    CKR(cudaMalloc(&zst, 4), "alloc");
    CKR(cudaMemsetAsync(zin, 0, (size_t)c->L.n_inputs * 32, st), "memset");
    CKR(cudaMemsetAsync(zstage, 0, (size_t)c->L.n_signals * 32, st), "memset");
    CKR(cudaMemsetAsync(zst, 0, 4, st), "memset");
    k_witness<<<dim3(1, WITNESS_TASKS), 32, 0, st>>>(c->L, c->consts, c->hc, zin, zstage, zst, 1, 0);
    k_witness_gather<<<dim3((c->n_vars + 255) / 256, 1), 256, 0, st>>>(zstage, c->L.n_signals, c->wmap, nullptr, c->tmpl,
                                                                       c->n_vars);
    CKR(cudaStreamSynchronize(st), "witness template");
    cudaFree(zstage);
    cudaFree(zin); cudaFree(zst); cudaFree(dforms);
    }
  }
  if (want_generic) {
    WitnessProgram prog;
    if (!build_witness_program(wasm, wasm_len, prog, err)) {
      set_error(err + (census_why.empty() ? "" : " (not recognised as census.circom either: " + census_why + ")"));
      return ZKB_UNSUPPORTED_CIRCUIT;
    }
    if (prog.n_wires != z.n_vars) { set_error("zkey and wasm disagree on the number of wires"); return ZKB_INVALID_WITNESS_LENGTH; }
    CKR(c->tape.upload(prog, st), "upload witness program");
    c->generic = true;
    c->has_witness = true;
    c->gen_inputs = prog.inputs;
    c->gen_first_signal = prog.first_input_signal;
    c->L.n_inputs = prog.n_inputs;
    c->L.n_wires = prog.n_wires;
    // The template trick of SURVEY 8a W7 needs nothing census-specific: sum_i w_i P_i = sum_i t_i P_i + sum_i (w_i - t_i) P_i
    // holds for ANY vector t.  t = the program's witness on all-zero inputs (its asserts may fail; the values are still
    // defined); wires a proof shares with it drop out of the four witness MSMs.  Single-range keys only: a sparse
    // difference cannot be assumed for an arbitrary circuit, and a large dense one needs the point-range split.
    if (!c->dense && z.n_vars <= (1u << 18)) {
      Fr *zin = nullptr, *zslots = nullptr;
      int *zst = nullptr;
      CKR(cudaMalloc(&c->tmpl, (size_t)c->n_vars * 32), "alloc template");
      CKR(cudaMalloc(&zin, (size_t)prog.n_inputs * 32), "alloc");
      CKR(cudaMalloc(&zslots, (size_t)prog.n_slots * 32), "alloc");
      CKR(cudaMalloc(&zst, 4), "alloc");
      CKR(cudaMemsetAsync(zin, 0, (size_t)prog.n_inputs * 32, st), "memset");
      CKR(cudaMemsetAsync(zst, 0, 4, st), "memset");
      CKR(tape_eval(c->tape, zin, zslots, c->tmpl, zst, 1, st), "template witness (tape)");
      CKR(cudaStreamSynchronize(st), "template witness");
      cudaFree(zin); cudaFree(zslots); cudaFree(zst);
    }
  }

  // coefficient matrices (CSR by row)
  {
    CoefCsr A, B;
    build_csr(z, 0, A);
    build_csr(z, 1, B);
    size_t nA = A.wire.size(), nB = B.wire.size(), rp = (size_t)z.domain + 1;
    std::vector<uint32_t> ibuf;
    ibuf.insert(ibuf.end(), A.row_ptr.begin(), A.row_ptr.end());
    ibuf.insert(ibuf.end(), B.row_ptr.begin(), B.row_ptr.end());
    ibuf.insert(ibuf.end(), A.wire.begin(), A.wire.end());
    ibuf.insert(ibuf.end(), B.wire.begin(), B.wire.end());
    CKR(upload(&c->csr_buf, ibuf.data(), ibuf.size() * 4), "upload csr");
    std::vector<uint8_t> vbuf(A.value);
    vbuf.insert(vbuf.end(), B.value.begin(), B.value.end());
    CKR(upload(&c->csr_val, vbuf.data(), vbuf.size()), "upload csr values");
    c->csrA = {c->csr_buf, c->csr_buf + 2 * rp, c->csr_val};
    c->csrB = {c->csr_buf + rp, c->csr_buf + 2 * rp + nA, c->csr_val + nA};
    (void)nB;
    // rows by descending term count (counting sort, stable)
    std::vector<uint32_t> len(z.domain), order(z.domain);
    uint32_t maxlen = 0;
    for (uint32_t r = 0; r < z.domain; r++) {
      len[r] = (A.row_ptr[r + 1] - A.row_ptr[r]) + (B.row_ptr[r + 1] - B.row_ptr[r]);
      maxlen = std::max(maxlen, len[r]);
    }
    std::vector<uint32_t> start(maxlen + 2, 0);
    for (uint32_t r = 0; r < z.domain; r++) start[maxlen - len[r] + 1]++;
    for (uint32_t i = 0; i <= maxlen; i++) start[i + 1] += start[i];
    for (uint32_t r = 0; r < z.domain; r++) order[start[maxlen - len[r]]++] = r;
    c->n_long_rows = 0;
    while (c->n_long_rows < z.domain && len[order[c->n_long_rows]] > ABC_LONG_ROW) c->n_long_rows++;
    CKR(upload(&c->row_order, order.data(), order.size() * 4), "upload row order");
  }
  CKR(c->ntt.init(z.power, st), "ntt plan");
  // fixed-base MSM tables from sections 5-9
  {
    Affine<Fq> *d1 = nullptr;
    Affine<Fq2> *d2 = nullptr;
    const bool use_tmpl = c->tmpl && !c->dense;
    // census: a few thousand differing wires -> 13-bit windows (4,096 buckets); a generic program's difference may be
    // dense -> 16-bit windows whatever the template saves
    c->cfgW = msm_cfg((int)env_u32("ZKB_C_WITNESS", use_tmpl && !c->generic ? 13 : 16));
    c->cfgH = msm_cfg((int)env_u32("ZKB_C_H", 16));
    // point ranges (see Circuit::subW): 2^17 points each once a key has more than 2^18; the last range is padded
    // with points at infinity
    // (a key sharded over more ranks than it has 2^17-point ranges is cut finer, so that every rank owns a range)
    uint32_t SUB = 1u << 17;
    if (shard_n > 1)
      while (SUB > 4096 && ((z.n_vars + SUB - 1) / SUB < (uint32_t)shard_n || z.domain / SUB < (uint32_t)shard_n)) SUB >>= 1;
    c->nsubW = ((z.n_vars > 2 * SUB || shard_n > 1) && !use_tmpl) ? SUB : z.n_vars;
    c->subW = (z.n_vars + c->nsubW - 1) / c->nsubW;
    c->n_pad = c->subW * c->nsubW;
    c->nsubH = (z.domain > 2 * SUB || (shard_n > 1 && z.domain > SUB)) ? SUB : z.domain;
    c->subH = z.domain / c->nsubH;
    c->loW = c->loH = 0;
    c->cntW = c->subW;
    c->cntH = c->subH;
    c->shard_rank = shard_rank;
    c->shard_n = shard_n;
    if (shard_n > 1) {
      if (c->subW < (uint32_t)shard_n || c->subH < (uint32_t)shard_n) {
        set_error("key too small to shard: every rank needs at least one 4096-point range of each MSM");
        return ZKB_ERROR;
      }
      c->loW = (uint32_t)((uint64_t)c->subW * shard_rank / shard_n);
      c->cntW = (uint32_t)((uint64_t)c->subW * (shard_rank + 1) / shard_n) - c->loW;
      c->loH = (uint32_t)((uint64_t)c->subH * shard_rank / shard_n);
      c->cntH = (uint32_t)((uint64_t)c->subH * (shard_rank + 1) / shard_n) - c->loH;
      {
        const size_t max_slice = ((size_t)z.n_vars + shard_n - 1) / shard_n + 1;
        c->wt_flag_off = (max_slice * 32 + 255) / 256 * 256;
        CKR(cudaMalloc(&c->wt_buf, c->wt_flag_off + 256), "alloc witness slice");
        CKR(cudaMemsetAsync(c->wt_buf + c->wt_flag_off, 0, 256, st), "memset");
        CKR(cudaMallocHost(&c->h_wt, max_slice * 32), "alloc pinned witness slice");
        c->wt_peer[shard_rank] = c->wt_buf;
      }
      CKR(cudaMalloc(&c->xbuf, SHARD_MAX_RANKS * sizeof(ShardSlot)), "alloc shard exchange");
      CKR(cudaMemsetAsync(c->xbuf, 0, SHARD_MAX_RANKS * sizeof(ShardSlot), st), "memset");
      c->root_x = c->xbuf;
    }
    // n points after pad_front leading infinities, padded at the back to subs * n_sub
    // only ranges [lo, lo + cnt) are uploaded and expanded (all of them unless the key is sharded)
    auto build1 = [&](MsmTable<Fq> &t, const uint8_t *src, uint32_t n, uint32_t pad_front, MsmCfg cfg, uint32_t n_sub,
                      uint32_t subs, uint32_t lo, uint32_t cnt) -> int {
      std::vector<uint8_t> tmp;
      const uint8_t *p = src;
      const size_t total = (size_t)n_sub * subs;
      if (pad_front || total != n) { tmp.assign(total * 64, 0); memcpy(tmp.data() + (size_t)pad_front * 64, src, (size_t)n * 64); p = tmp.data(); }
      CKR(upload(&d1, p + (size_t)lo * n_sub * 64, (size_t)cnt * n_sub * 64), "upload bases");
      CKR(msm_build_table<Fq>(t, d1, n_sub, cfg, st, cnt), "build table");
      CKR(cudaStreamSynchronize(st), "build table");
      cudaFree(d1);
      return ZKB_OK;
    };
    int rc;
    if ((rc = build1(c->tabA, z.a, z.n_vars, 0, c->cfgW, c->nsubW, c->subW, c->loW, c->cntW))) return rc;
    if ((rc = build1(c->tabB1, z.b1, z.n_vars, 0, c->cfgW, c->nsubW, c->subW, c->loW, c->cntW))) return rc;
    if ((rc = build1(c->tabC, z.c, z.n_vars - z.n_public - 1, z.n_public + 1, c->cfgW, c->nsubW, c->subW, c->loW, c->cntW))) return rc;
    if ((rc = build1(c->tabH, z.h, z.domain, 0, c->cfgH, c->nsubH, c->subH, c->loH, c->cntH))) return rc;
    {
      std::vector<uint8_t> tmp;
      const uint8_t *p = z.b2;
      if (c->n_pad != z.n_vars) { tmp.assign((size_t)c->n_pad * 128, 0); memcpy(tmp.data(), z.b2, (size_t)z.n_vars * 128); p = tmp.data(); }
      CKR(upload(&d2, p + (size_t)c->loW * c->nsubW * 128, (size_t)c->cntW * c->nsubW * 128), "upload bases");
    }
    CKR(msm_build_table<Fq2>(c->tabB2, d2, c->nsubW, c->cfgW, st, c->cntW), "build table");
    CKR(cudaStreamSynchronize(st), "build table");
    cudaFree(d2);
  }
  // constant part of the four witness MSMs: sum_i tmpl_i * base_i (once per key)
  if (c->tmpl && !c->dense) {
    MsmSort s1;
    MsmWork<Fq> w1;
    MsmWork<Fq2> w2;
    CKR(s1.alloc(z.n_vars, 1, c->cfgW), "alloc");
    CKR(w1.alloc(3, c->cfgW), "alloc");
    CKR(w2.alloc(1, c->cfgW), "alloc");
    CKR(cudaMalloc(&c->tconst1, 3 * sizeof(XYZZ<Fq>)), "alloc");
    CKR(cudaMalloc(&c->tconst2, sizeof(XYZZ<Fq2>)), "alloc");
    CKR(s1.run(c->tmpl, z.n_vars, 1, st), "sort template");
    MsmTable<Fq> tabs[3] = {c->tabA, c->tabB1, c->tabC};
    CKR(msm_run<Fq>(s1, tabs, 3, 1, w1, c->tconst1, st), "template msm g1");
    CKR(msm_run<Fq2>(s1, &c->tabB2, 1, 1, w2, c->tconst2, st), "template msm g2");
    CKR(cudaStreamSynchronize(st), "template msm");
    s1.free_all(); w1.free_all(); w2.free_all();
  }
  // alpha1, beta1, beta2, delta tables
  {
    uint8_t f1[128];
    memcpy(f1, z.alpha1, 64);
    memcpy(f1 + 64, z.beta1, 64);
    CKR(upload(&c->fix1, f1, 128), "upload");
    CKR(upload(&c->fix2, z.beta2, 128), "upload");
    Affine<Fq> *dl1 = nullptr;
    Affine<Fq2> *dl2 = nullptr;
    CKR(upload(&dl1, z.delta1, 64), "upload");
    CKR(upload(&dl2, z.delta2, 128), "upload");
    CKR(cudaMalloc(&c->d1tab, 64 * 15 * sizeof(Affine<Fq>)), "alloc");
    CKR(cudaMalloc(&c->d2tab, 64 * 15 * sizeof(Affine<Fq2>)), "alloc");
    CKR(launch_fixed_tables(c->d1tab, dl1, c->d2tab, dl2, st), "delta tables");
    CKR(cudaStreamSynchronize(st), "delta tables");
    cudaFree(dl1); cudaFree(dl2);
  }
  CKR(cudaGetLastError(), "load circuit");
  *out = c.release();
  return ZKB_OK;
}

static void destroy_circuit(Circuit *c) {
  if (!c) return;
  cudaSetDevice(c->ctx->device);
  cudaFree(c->consts); cudaFree(c->hc); cudaFree(c->tmpl); cudaFree(c->wmap); cudaFree(c->csr_buf); cudaFree(c->row_order);
  cudaFree(c->csr_val); cudaFree(c->fix1); cudaFree(c->fix2); cudaFree(c->d1tab); cudaFree(c->d2tab);
  cudaFree(c->tconst1); cudaFree(c->tconst2);
  c->tape.free_all();
  if (c->root_is_ipc) cudaIpcCloseMemHandle(c->root_x);
  for (int r = 0; r < SHARD_MAX_RANKS; r++)
    if (c->wt_peer_ipc[r]) cudaIpcCloseMemHandle(const_cast<uint8_t *>(c->wt_peer[r]));
  cudaFree(c->wt_buf);
  if (c->h_wt) cudaFreeHost(c->h_wt);
  cudaFree(c->xbuf);
  cudaFree(c->tabA.tab); cudaFree(c->tabB1.tab); cudaFree(c->tabC.tab); cudaFree(c->tabH.tab); cudaFree(c->tabB2.tab);
  c->ntt.destroy();
  cudaFree(c->inputs); cudaFree(c->wtns); cudaFree(c->rs); cudaFree(c->status); cudaFree(c->out);
  if (c->h_out) cudaFreeHost(c->h_out);
  if (c->h_inputs) cudaFreeHost(c->h_inputs);
  if (c->h_rs) cudaFreeHost(c->h_rs);
  if (c->h_status) cudaFreeHost(c->h_status);
  for (int i = 0; i < MAX_LANES; i++) {
    c->lanes[i].free_all();
    if (c->lanes[i].st) {
      Lane &ln = c->lanes[i];
      cudaStreamDestroy(ln.st); cudaStreamDestroy(ln.sW); cudaStreamDestroy(ln.s2);
      for (cudaEvent_t e : {ln.done, ln.e_fork, ln.e_sortW, ln.e_W, ln.e_2}) cudaEventDestroy(e);
    }
  }
  delete c;
}

static int copy_out(const std::string &s, char *buf, size_t *size) {
  size_t need = s.size() + 1;
  if (!buf || *size < need) { *size = need; return ZKB_SHORT_BUFFER; }
  memcpy(buf, s.c_str(), need);
  *size = s.size();
  return ZKB_OK;
}

// upload h_inputs[0..n) + blinding, run, download results/status
static int prove_group(Circuit *c, uint32_t n, bool with_witness, float *stage_ms) {
  cudaStream_t st = c->ctx->stream;
  if (with_witness)
    CKR(cudaMemcpyAsync(c->inputs, c->h_inputs, (size_t)n * c->L.n_inputs * 32, cudaMemcpyHostToDevice, st), "h2d inputs");
  if (!fill_blinding(c, n)) return ZKB_ERROR;
  CKR(cudaMemcpyAsync(c->rs, c->h_rs, (size_t)n * 64, cudaMemcpyHostToDevice, st), "h2d rs");
  if (!with_witness) CKR(cudaMemsetAsync(c->status, 0, (size_t)n * 4, st), "memset status");
  int rc = prove_resident(c, n, with_witness, stage_ms);
  if (rc) return rc;
  CKR(cudaMemcpyAsync(c->h_out, c->out, (size_t)n * c->out_stride(), cudaMemcpyDeviceToHost, st), "d2h");
  if (with_witness) CKR(cudaMemcpyAsync(c->h_status, c->status, (size_t)n * 4, cudaMemcpyDeviceToHost, st), "d2h status");
  else memset(c->h_status, 0, (size_t)n * 4);
  CKR(cudaStreamSynchronize(st), "sync");
  return ZKB_OK;
}

}  // namespace zkb

using namespace zkb;

struct zkb_ctx { Ctx c; };
struct zkb_circuit { Circuit *c; };

extern "C" {

int zkb_ctx_create(int device, zkb_ctx **out) {
  if (require_device()) return ZKB_ERROR;
  int n = 0;
  cudaGetDeviceCount(&n);
  if (device < 0 || device >= n) { set_error("invalid device index"); return ZKB_ERROR; }
  CKR(cudaSetDevice(device), "set device");
  zkb_ctx *x = new zkb_ctx();
  x->c.device = device;
  cudaError_t e = cudaStreamCreateWithFlags(&x->c.stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) { delete x; return cuda_fail(e, "stream create"); }
  *out = x;
  return ZKB_OK;
}

void zkb_ctx_destroy(zkb_ctx *x) {
  if (!x) return;
  cudaSetDevice(x->c.device);
  cudaStreamDestroy(x->c.stream);
  delete x;
}

void *zkb_ctx_stream(zkb_ctx *x) { return (void *)x->c.stream; }

int zkb_load_circuit_ex(zkb_ctx *ctx, const void *zkey, size_t zkey_len, const void *wasm, size_t wasm_len,
                        uint32_t flags, zkb_circuit **out) {
  if (!ctx || !zkey || !out) { set_error("null argument"); return ZKB_ERROR; }
  Circuit *c = nullptr;
  int rc = load_circuit(&ctx->c, (const uint8_t *)zkey, zkey_len, (const uint8_t *)wasm, wasm_len, &c, 0, 1, flags);
  if (rc) return rc;
  *out = new zkb_circuit{c};
  return ZKB_OK;
}
int zkb_load_circuit(zkb_ctx *ctx, const void *zkey, size_t zkey_len, const void *wasm, size_t wasm_len,
                     zkb_circuit **out) {
  return zkb_load_circuit_ex(ctx, zkey, zkey_len, wasm, wasm_len, 0, out);
}

// A proving key sharded over `nranks` GPUs (one context / process per GPU): rank `rank` keeps the fixed-base tables of
// its point ranges only.  No witness generator (prove from a .wtns); every rank is given the same witness.
int zkb_load_circuit_shard(zkb_ctx *ctx, const void *zkey, size_t zkey_len, int rank, int nranks, zkb_circuit **out) {
  if (!ctx || !zkey || !out) { set_error("null argument"); return ZKB_ERROR; }
  if (nranks < 1 || nranks > SHARD_MAX_RANKS || rank < 0 || rank >= nranks) { set_error("bad shard rank"); return ZKB_ERROR; }
  Circuit *c = nullptr;
  int rc = load_circuit(&ctx->c, (const uint8_t *)zkey, zkey_len, nullptr, 0, &c, rank, nranks);
  if (rc) return rc;
  *out = new zkb_circuit{c};
  return ZKB_OK;
}
// rank 0: the 64-byte CUDA IPC handle of its exchange buffer
int zkb_shard_export(zkb_circuit *h, void *handle64) {
  if (!h || !handle64) { set_error("null argument"); return ZKB_ERROR; }
  Circuit *c = h->c;
  CKR(cudaSetDevice(c->ctx->device), "set device");
  if (!c->xbuf) { set_error("not a sharded key"); return ZKB_ERROR; }
  cudaIpcMemHandle_t mh;
  CKR(cudaIpcGetMemHandle(&mh, c->xbuf), "ipc export");
  memcpy(handle64, &mh, 64);
  return ZKB_OK;
}
// ranks > 0: map rank 0's exchange buffer (their partial sums are written there over NVLink)
int zkb_shard_attach(zkb_circuit *h, const void *handle64) {
  if (!h || !handle64) { set_error("null argument"); return ZKB_ERROR; }
  Circuit *c = h->c;
  CKR(cudaSetDevice(c->ctx->device), "set device");
  if (!c->xbuf) { set_error("not a sharded key"); return ZKB_ERROR; }
  cudaIpcMemHandle_t mh;
  memcpy(&mh, handle64, 64);
  void *p = nullptr;
  CKR(cudaIpcOpenMemHandle(&p, mh, cudaIpcMemLazyEnablePeerAccess), "ipc open (peer access to rank 0)");
  c->root_x = reinterpret_cast<ShardSlot *>(p);
  c->root_is_ipc = true;
  return ZKB_OK;
}
// same process (several contexts): publish straight into the root circuit's buffer
int zkb_shard_attach_local(zkb_circuit *h, zkb_circuit *root) {
  if (!h || !root) { set_error("null argument"); return ZKB_ERROR; }
  Circuit *c = h->c, *r = root->c;
  CKR(cudaSetDevice(c->ctx->device), "set device");
  if (!c->xbuf || !r->xbuf) { set_error("not a sharded key"); return ZKB_ERROR; }
  if (c->ctx->device != r->ctx->device) {
    int can = 0;
    cudaDeviceCanAccessPeer(&can, c->ctx->device, r->ctx->device);
    if (!can) { set_error("no peer access between the two GPUs"); return ZKB_ERROR; }
    cudaError_t e = cudaDeviceEnablePeerAccess(r->ctx->device, 0);
    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return cuda_fail(e, "enable peer access");
    cudaGetLastError();
  }
  c->root_x = r->xbuf;
  return ZKB_OK;
}

// Witness-slice exchange of a sharded key: every rank exports its slice buffer and maps the others'.  Once all
// nranks - 1 peers are attached, zkb_prove_wtns uploads only this rank's slice of the witness and gathers the rest
// over NVLink; until then every rank uploads the whole witness.
int zkb_shard_export_witness(zkb_circuit *h, void *handle64) {
  if (!h || !handle64) { set_error("null argument"); return ZKB_ERROR; }
  Circuit *c = h->c;
  CKR(cudaSetDevice(c->ctx->device), "set device");
  if (!c->wt_buf) { set_error("not a sharded key"); return ZKB_ERROR; }
  {
    // allocate the proving workspace now: cudaMalloc synchronises the device, which must not happen while a peer's
    // kernel is already spinning on this rank's "slice uploaded" flag
    std::lock_guard<std::mutex> g(c->mu);
    int rc = ensure_workspace(c, c->cap ? c->cap : 1, c->chunk ? c->chunk : default_chunk(c));
    if (rc) return rc;
  }
  cudaIpcMemHandle_t mh;
  CKR(cudaIpcGetMemHandle(&mh, c->wt_buf), "ipc export");
  memcpy(handle64, &mh, 64);
  return ZKB_OK;
}
int zkb_shard_attach_witness(zkb_circuit *h, int peer_rank, const void *handle64) {
  if (!h || !handle64) { set_error("null argument"); return ZKB_ERROR; }
  Circuit *c = h->c;
  CKR(cudaSetDevice(c->ctx->device), "set device");
  if (!c->wt_buf || peer_rank < 0 || peer_rank >= c->shard_n || peer_rank == c->shard_rank || c->wt_peer[peer_rank]) {
    set_error("bad peer rank for the witness exchange");
    return ZKB_ERROR;
  }
  cudaIpcMemHandle_t mh;
  memcpy(&mh, handle64, 64);
  void *p = nullptr;
  CKR(cudaIpcOpenMemHandle(&p, mh, cudaIpcMemLazyEnablePeerAccess), "ipc open (peer witness slice)");
  c->wt_peer[peer_rank] = (const uint8_t *)p;
  c->wt_peer_ipc[peer_rank] = true;
  c->wt_attached++;
  return ZKB_OK;
}
int zkb_shard_attach_witness_local(zkb_circuit *h, int peer_rank, zkb_circuit *peer) {
  if (!h || !peer) { set_error("null argument"); return ZKB_ERROR; }
  Circuit *c = h->c, *q = peer->c;
  CKR(cudaSetDevice(c->ctx->device), "set device");
  if (!c->wt_buf || !q->wt_buf || peer_rank != q->shard_rank || peer_rank == c->shard_rank || c->wt_peer[peer_rank]) {
    set_error("bad peer for the witness exchange");
    return ZKB_ERROR;
  }
  if (c->ctx->device != q->ctx->device) {
    int can = 0;
    cudaDeviceCanAccessPeer(&can, c->ctx->device, q->ctx->device);
    if (!can) { set_error("no peer access between the two GPUs"); return ZKB_ERROR; }
    cudaError_t e = cudaDeviceEnablePeerAccess(q->ctx->device, 0);
    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return cuda_fail(e, "enable peer access");
    cudaGetLastError();
  }
  {
    std::lock_guard<std::mutex> g(c->mu);
    int rc = ensure_workspace(c, c->cap ? c->cap : 1, c->chunk ? c->chunk : default_chunk(c));
    if (rc) return rc;
  }
  c->wt_peer[peer_rank] = q->wt_buf;
  c->wt_attached++;
  return ZKB_OK;
}

void zkb_circuit_destroy(zkb_circuit *h) {
  if (!h) return;
  destroy_circuit(h->c);
  delete h;
}

// info: n_vars, n_public, domain, n_inputs, n_levels_plus1, n_signals, chunk, cap
int zkb_circuit_info(zkb_circuit *h, uint32_t *info) {
  if (!h || !info) { set_error("null argument"); return ZKB_ERROR; }
  Circuit *c = h->c;
  info[0] = c->n_vars; info[1] = c->n_public; info[2] = c->domain; info[3] = c->L.n_inputs;
  info[4] = c->L.n; info[5] = c->L.n_signals; info[6] = c->chunk; info[7] = c->cap;
  return ZKB_OK;
}

// r32 / s32: canonical little-endian scalars < r; both NULL restores random blinding
int zkb_set_blinding(zkb_circuit *h, const uint8_t *r32, const uint8_t *s32) {
  if (!h) { set_error("null handle"); return ZKB_ERROR; }
  Circuit *c = h->c;
  std::lock_guard<std::mutex> g(c->mu);
  if (!r32 || !s32) { c->fixed_rs = false; return ZKB_OK; }
  memcpy(c->fr, r32, 32);
  memcpy(c->fs, s32, 32);
  c->fixed_rs = true;
  return ZKB_OK;
}

// ---- resident-input path (device-only timing of the hot path) ---------------------------------
// inputs: n x n_inputs canonical 32-byte values in main-signal order (electionId[2], nullifier, voteHash[2],
// sikRoot, censusRoot, voteWeight, availableWeight, address, password, signature, censusSiblings, sikSiblings)
int zkb_batch_set_inputs(zkb_circuit *h, int n, const void *inputs) {
  if (!h || !inputs || n <= 0) { set_error("batch_set_inputs: null handle / inputs or n <= 0"); return ZKB_ERROR; }
  Circuit *c = h->c;
  if (!c->has_witness) { set_error("circuit was loaded without a wasm: no witness generator"); return ZKB_ERROR; }
  std::lock_guard<std::mutex> g(c->mu);
  CKR(cudaSetDevice(c->ctx->device), "set device");
  uint32_t chunk = default_chunk(c);
  int rc = ensure_workspace(c, (uint32_t)n, chunk);
  if (rc) return rc;
  CKR(cudaMemcpyAsync(c->inputs, inputs, (size_t)n * c->L.n_inputs * 32, cudaMemcpyHostToDevice, c->ctx->stream), "h2d inputs");
  if (!fill_blinding(c, n)) return ZKB_ERROR;
  CKR(cudaMemcpyAsync(c->rs, c->h_rs, (size_t)n * 64, cudaMemcpyHostToDevice, c->ctx->stream), "h2d rs");
  CKR(cudaStreamSynchronize(c->ctx->stream), "sync");
  return ZKB_OK;
}

// runs witness + Groth16 on the resident inputs, results stay on the device.  stage_ms: 5 floats or NULL.
int zkb_batch_prove_resident(zkb_circuit *h, int n, float *stage_ms) {
  if (!h || n <= 0) { set_error("batch_prove_resident: null handle or n <= 0"); return ZKB_ERROR; }
  Circuit *c = h->c;
  std::lock_guard<std::mutex> g(c->mu);
  CKR(cudaSetDevice(c->ctx->device), "set device");
  if ((uint32_t)n > c->cap) { set_error("more proofs than resident inputs"); return ZKB_ERROR; }
  return prove_resident(c, (uint32_t)n, true, stage_ms);
}

// proofs256: n x 256 B; publics: n x n_public x 32 B; status: n ints (0 ok, 4 assert failed)
int zkb_batch_get_results(zkb_circuit *h, int n, void *proofs256, void *publics, int *status) {
  if (!h) { set_error("null handle"); return ZKB_ERROR; }
  Circuit *c = h->c;
  std::lock_guard<std::mutex> g(c->mu);
  if (n <= 0 || (uint32_t)n > c->cap) { set_error("batch_get_results: n outside the resident batch"); return ZKB_ERROR; }
  CKR(cudaSetDevice(c->ctx->device), "set device");
  cudaStream_t st = c->ctx->stream;
  CKR(cudaMemcpyAsync(c->h_out, c->out, (size_t)n * c->out_stride(), cudaMemcpyDeviceToHost, st), "d2h");
  CKR(cudaMemcpyAsync(c->h_status, c->status, (size_t)n * 4, cudaMemcpyDeviceToHost, st), "d2h status");
  CKR(cudaStreamSynchronize(st), "sync");
  for (int i = 0; i < n; i++) {
    const uint8_t *o = c->h_out + (size_t)i * c->out_stride();
    if (proofs256) memcpy((uint8_t *)proofs256 + (size_t)i * 256, o, 256);
    if (publics) memcpy((uint8_t *)publics + (size_t)i * 32 * c->n_public, o + 256, 32 * (size_t)c->n_public);
    if (status) status[i] = c->h_status[i];
  }
  return ZKB_OK;
}

// device copy of the n resident witnesses (n x n_vars x 32 B, canonical) - parity tests
int zkb_batch_get_witness(zkb_circuit *h, int first, int n, void *wtns) {
  if (!h || !wtns) { set_error("null argument"); return ZKB_ERROR; }
  Circuit *c = h->c;
  std::lock_guard<std::mutex> g(c->mu);
  if (first < 0 || n <= 0 || (uint64_t)first + (uint64_t)n > c->cap) { set_error("batch_get_witness: range outside the resident batch"); return ZKB_ERROR; }
  CKR(cudaSetDevice(c->ctx->device), "set device");
  CKR(cudaMemcpy(wtns, c->wtns + (size_t)first * c->n_vars, (size_t)n * c->n_vars * 32, cudaMemcpyDeviceToHost), "d2h witness");
  return ZKB_OK;
}

uint64_t zkb_launch_count(void) { return g_launches.load(); }

// Work actually executed by the MSM accumulate kernels for the last processed chunk (exact, counted on the device
// outside any timed region): out[0] = G1 mixed adds (A + B1 + C + H), out[1] = G2 mixed adds (B2), out[2] = sorted
// digit entries of the witness, out[3] = of the H scalars, out[4] = proofs in that chunk, out[5] = chunk capacity.
int zkb_work_counters(zkb_circuit *h, uint64_t *out) {
  Circuit *c = h->c;
  std::lock_guard<std::mutex> g(c->mu);
  CKR(cudaSetDevice(c->ctx->device), "set device");
  cudaStream_t st = c->ctx->stream;
  uint32_t m = c->last_chunk_m;
  if (!m) { set_error("no chunk processed yet"); return ZKB_ERROR; }
  if (c->subW > 1 || c->subH > 1) { set_error("work counters are implemented for single-range keys only"); return ZKB_ERROR; }
  unsigned long long t[5] = {0, 0, 0, 0, 0};
  Lane &ln = c->lanes[c->last_lane];
  CKR(msm_count_madds<Fq>(ln.sortW, c->tabA, m, &t[0], st), "count");
  CKR(msm_count_madds<Fq>(ln.sortW, c->tabB1, m, &t[1], st), "count");
  CKR(msm_count_madds<Fq>(ln.sortW, c->tabC, m, &t[2], st), "count");
  unsigned long long aff[3] = {0, 0, 0}, inversions = 0;
  if (c->last_affine) {
    CKR(msm_affine_counts(ln.sortH, ln.affH, m, aff, st), "count");
    t[3] = aff[1];
    inversions = aff[2];
  } else {
    CKR(msm_count_madds<Fq>(ln.sortH, c->tabH, m, &t[3], st), "count");
  }
  CKR(msm_count_madds<Fq2>(ln.sortW, c->tabB2, m, &t[4], st), "count");
  out[0] = t[0] + t[1] + t[2] + t[3];
  out[1] = t[4];
  uint64_t ew = 0, eh = 0;
  for (uint32_t b = 0; b < m; b++) {
    uint32_t v;
    CKR(cudaMemcpy(&v, ln.sortW.offsets + (size_t)b * (c->cfgW.buckets + 1) + c->cfgW.buckets, 4, cudaMemcpyDeviceToHost), "d2h");
    ew += v;
    CKR(cudaMemcpy(&v, ln.sortH.offsets + (size_t)b * (c->cfgH.buckets + 1) + c->cfgH.buckets, 4, cudaMemcpyDeviceToHost), "d2h");
    eh += v;
  }
  out[2] = ew; out[3] = eh; out[4] = m; out[5] = c->chunk;
  out[6] = aff[0]; out[7] = inversions;
  return ZKB_OK;
}

// MSM partial sums of the first proof of the last chunk, affine canonical: pi_a'(64) pi_b1'(64) pi_b'(128)
// pi_c'(64) pi_h(64) = 384 B; and that proof's H scalars (domain x 32 B canonical) when h_out != NULL.
int zkb_debug_partials(zkb_circuit *h, void *out384, void *h_out) {
  Circuit *c = h->c;
  std::lock_guard<std::mutex> g(c->mu);
  CKR(cudaSetDevice(c->ctx->device), "set device");
  XYZZ<Fq> g1[4];
  XYZZ<Fq2> g2;
  Lane &ln = c->lanes[c->last_lane];
  CKR(cudaMemcpy(g1, ln.g1out, 3 * sizeof(XYZZ<Fq>), cudaMemcpyDeviceToHost), "d2h");
  CKR(cudaMemcpy(g1 + 3, ln.g1out + (size_t)3 * c->chunk, sizeof(XYZZ<Fq>), cudaMemcpyDeviceToHost), "d2h");
  CKR(cudaMemcpy(&g2, ln.g2out, sizeof(XYZZ<Fq2>), cudaMemcpyDeviceToHost), "d2h");
  if (c->tconst1) {   // the device sums cover w - tmpl; add the per-key constant part back (host arithmetic, debug only)
    XYZZ<Fq> t1[3];
    XYZZ<Fq2> t2;
    CKR(cudaMemcpy(t1, c->tconst1, sizeof t1, cudaMemcpyDeviceToHost), "d2h");
    CKR(cudaMemcpy(&t2, c->tconst2, sizeof t2, cudaMemcpyDeviceToHost), "d2h");
    for (int i = 0; i < 3; i++) g1[i].add(t1[i]);
    g2.add(t2);
  }
  uint8_t *o = (uint8_t *)out384;
  auto put1 = [&](const XYZZ<Fq> &p, uint8_t *dst) {   // host (portable) arithmetic, debug only
    Affine<Fq> a = p.to_affine();
    Fq x = a.x.from_mont(), y = a.y.from_mont();
    memcpy(dst, x.v, 32); memcpy(dst + 32, y.v, 32);
  };
  put1(g1[0], o); put1(g1[1], o + 64); put1(g1[2], o + 256); put1(g1[3], o + 320);
  Affine<Fq2> b = g2.to_affine();
  Fq cc[4] = {b.x.a.from_mont(), b.x.b.from_mont(), b.y.a.from_mont(), b.y.b.from_mont()};
  memcpy(o + 128, cc, 128);
  if (h_out) CKR(cudaMemcpy(h_out, ln.hs, (size_t)c->domain * 32, cudaMemcpyDeviceToHost), "d2h h");
  return ZKB_OK;
}

// Batched Poseidon over the circuit's own constants (arity 2, 3 or 4): the hash of the census / SIK trees
// (arbo HashFunctionPoseidon, internal/helpers.go:45-49) - used by the synthetic census generator.
// in: n x arity canonical 32-byte values; out: n x 32 bytes.
int zkb_poseidon_hash(zkb_circuit *h, int arity, int n, const void *in, void *out) {
  Circuit *c = h->c;
  if (!c->consts) { set_error("circuit was loaded without a wasm: no Poseidon constants"); return ZKB_ERROR; }
  if (arity < 2 || arity > 4 || n < 0) { set_error("poseidon: arity must be 2, 3 or 4"); return ZKB_ERROR; }
  if (n == 0) return ZKB_OK;
  std::lock_guard<std::mutex> g(c->mu);
  CKR(cudaSetDevice(c->ctx->device), "set device");
  cudaStream_t st = c->ctx->stream;
  Fr *din = nullptr, *dout = nullptr;
  CKR(cudaMalloc(&din, (size_t)n * arity * 32), "alloc");
  CKR(cudaMalloc(&dout, (size_t)n * 32), "alloc");
  CKR(cudaMemcpyAsync(din, in, (size_t)n * arity * 32, cudaMemcpyHostToDevice, st), "h2d");
  unsigned grid = (unsigned)((n + 63) / 64);
  if (arity == 2) k_poseidon_batch<3><<<grid, 64, 0, st>>>(c->L, c->consts, din, dout, (uint32_t)n);
  else if (arity == 3) k_poseidon_batch<4><<<grid, 64, 0, st>>>(c->L, c->consts, din, dout, (uint32_t)n);
  else k_poseidon_batch<5><<<grid, 64, 0, st>>>(c->L, c->consts, din, dout, (uint32_t)n);
  CKR(cudaGetLastError(), "poseidon launch");
  CKR(cudaMemcpyAsync(out, dout, (size_t)n * 32, cudaMemcpyDeviceToHost, st), "d2h");
  CKR(cudaStreamSynchronize(st), "sync");
  cudaFree(din); cudaFree(dout);
  return ZKB_OK;
}

}  // extern "C"

// ---- census / SIK tree builder ----------------------------------------------------------------
// arbo semantics (internal/helpers.go:36-85, GenTree + GenProof): the path of a key is its bits LSB first; an empty
// subtree hashes to 0, a subtree with one key is the leaf H(key, value, 1), any other node is H(left, right); the
// siblings of a key are the other branch at every depth until the key is alone.  The structure (a trie over the
// bit-reversed keys) is laid out on the host in one sort + one pass; every hash runs on the GPU, one batched Poseidon
// launch for all leaves and one per tree level, deepest first.
static int poseidon_device(Circuit *c, int arity, uint32_t n, const Fr *din, Fr *dout, cudaStream_t st) {
  const unsigned grid = (n + 63) / 64;
  if (arity == 2) k_poseidon_batch<3><<<grid, 64, 0, st>>>(c->L, c->consts, din, dout, n);
  else if (arity == 3) k_poseidon_batch<4><<<grid, 64, 0, st>>>(c->L, c->consts, din, dout, n);
  else k_poseidon_batch<5><<<grid, 64, 0, st>>>(c->L, c->consts, din, dout, n);
  const cudaError_t le = cudaGetLastError();
  return le == cudaSuccess ? ZKB_OK : cuda_fail(le, "poseidon launch");
}

extern "C" int zkb_census_tree(zkb_circuit *h, int n_keys, const void *keys32, const void *values32, int n_levels,
                               void *root32, void *siblings) {
  if (!h || n_keys <= 0 || !keys32 || !values32 || !root32 || n_levels < 1 || n_levels > 255) { set_error("census_tree: bad argument"); return ZKB_ERROR; }
  Circuit *c = h->c;
  if (!c->consts) { set_error("circuit was loaded without the census wasm: no Poseidon constants"); return ZKB_ERROR; }
  const uint32_t n = (uint32_t)n_keys;
  const uint8_t *K = (const uint8_t *)keys32, *V = (const uint8_t *)values32;
  auto bit = [&](uint32_t i, int d) -> int { return (K[(size_t)i * 32 + (d >> 3)] >> (d & 7)) & 1; };
  // order the keys by their LSB-first bit string
  std::vector<uint32_t> ord(n);
  for (uint32_t i = 0; i < n; i++) ord[i] = i;
  std::sort(ord.begin(), ord.end(), [&](uint32_t a, uint32_t b) {
    for (int byte = 0; byte < 32; byte++) {
      uint8_t x = K[(size_t)a * 32 + byte], y = K[(size_t)b * 32 + byte];
      if (x != y) {
        const uint8_t diff = x ^ y, low = diff & (uint8_t)(-(int8_t)diff);     // lowest differing bit decides
        return (x & low) == 0;
      }
    }
    return false;
  });
  for (uint32_t i = 1; i < n; i++)
    if (memcmp(K + (size_t)ord[i] * 32, K + (size_t)ord[i - 1] * 32, 32) == 0) { set_error("census_tree: duplicate key"); return ZKB_ERROR; }
  struct Node { uint32_t lo, hi; int depth; int left, right; };          // children: node index, -1 = empty
  std::vector<Node> nodes;
  nodes.push_back({0, n, 0, -1, -1});
  int max_depth = 0;
  for (size_t q = 0; q < nodes.size(); q++) {
    Node nd = nodes[q];
    if (nd.hi - nd.lo <= 1) continue;
    if (nd.depth >= n_levels) { set_error("census_tree: two keys share their first nLevels bits"); return ZKB_ERROR; }
    uint32_t mid = nd.lo;
    while (mid < nd.hi && !bit(ord[mid], nd.depth)) mid++;                // keys with bit `depth` = 0 come first
    if (mid > nd.lo) { nodes[q].left = (int)nodes.size(); nodes.push_back({nd.lo, mid, nd.depth + 1, -1, -1}); }
    if (mid < nd.hi) { nodes[q].right = (int)nodes.size(); nodes.push_back({mid, nd.hi, nd.depth + 1, -1, -1}); }
    max_depth = std::max(max_depth, nd.depth + 1);
  }
  std::lock_guard<std::mutex> g(c->mu);
  CKR(cudaSetDevice(c->ctx->device), "set device");
  cudaStream_t st = c->ctx->stream;
  const size_t nn = nodes.size();
  std::vector<Fr> hash(nn, Fr::zero());
  Fr *din = nullptr, *dout = nullptr;
  CKR(cudaMalloc(&din, (size_t)std::max<size_t>(nn, n) * 3 * 32), "alloc");
  CKR(cudaMalloc(&dout, (size_t)std::max<size_t>(nn, n) * 32), "alloc");
  std::vector<Fr> in, out;
  auto run = [&](int arity, uint32_t cnt) -> int {
    out.resize(cnt);
    CKR(cudaMemcpyAsync(din, in.data(), (size_t)cnt * arity * 32, cudaMemcpyHostToDevice, st), "h2d");
    int rc = poseidon_device(c, arity, cnt, din, dout, st);
    if (rc) return rc;
    CKR(cudaMemcpyAsync(out.data(), dout, (size_t)cnt * 32, cudaMemcpyDeviceToHost, st), "d2h");
    CKR(cudaStreamSynchronize(st), "sync");
    return ZKB_OK;
  };
  // leaves: H(key, value, 1)
  {
    std::vector<size_t> idx;
    in.clear();
    for (size_t q = 0; q < nn; q++)
      if (nodes[q].hi - nodes[q].lo == 1) {
        idx.push_back(q);
        Fr k, v, one = Fr::zero();
        memcpy(k.v, K + (size_t)ord[nodes[q].lo] * 32, 32);
        memcpy(v.v, V + (size_t)ord[nodes[q].lo] * 32, 32);
        one.v[0] = 1;
        in.push_back(k); in.push_back(v); in.push_back(one);
      }
    int rc = run(3, (uint32_t)idx.size());
    if (rc) { cudaFree(din); cudaFree(dout); return rc; }
    for (size_t t = 0; t < idx.size(); t++) hash[idx[t]] = out[t];
  }
  // internal nodes, deepest level first
  std::vector<std::vector<size_t>> by_depth(max_depth + 1);
  for (size_t q = 0; q < nn; q++)
    if (nodes[q].hi - nodes[q].lo > 1) by_depth[nodes[q].depth].push_back(q);
  for (int d = max_depth; d >= 0; d--) {
    if (by_depth[d].empty()) continue;
    in.clear();
    for (size_t q : by_depth[d]) {
      in.push_back(nodes[q].left >= 0 ? hash[nodes[q].left] : Fr::zero());
      in.push_back(nodes[q].right >= 0 ? hash[nodes[q].right] : Fr::zero());
    }
    int rc = run(2, (uint32_t)by_depth[d].size());
    if (rc) { cudaFree(din); cudaFree(dout); return rc; }
    for (size_t t = 0; t < by_depth[d].size(); t++) hash[by_depth[d][t]] = out[t];
  }
  cudaFree(din); cudaFree(dout);
  memcpy(root32, hash[0].v, 32);
  if (siblings) {
    uint8_t *S = (uint8_t *)siblings;
    const size_t stride = (size_t)(n_levels + 1) * 32;
    memset(S, 0, (size_t)n * stride);
    // walk every node once: the sibling of all keys under a child is the other child's hash
    for (size_t q = 0; q < nn; q++) {
      const Node &nd = nodes[q];
      if (nd.hi - nd.lo <= 1) continue;
      for (int side = 0; side < 2; side++) {
        const int ch = side ? nd.right : nd.left, other = side ? nd.left : nd.right;
        if (ch < 0) continue;
        const Fr sib = other >= 0 ? hash[other] : Fr::zero();
        for (uint32_t t = nodes[ch].lo; t < nodes[ch].hi; t++) memcpy(S + (size_t)ord[t] * stride + (size_t)nd.depth * 32, sib.v, 32);
      }
    }
  }
  return ZKB_OK;
}

// ---- reference-shaped entry points ---------------------------------------------------------------
// Host threads of the JSON stages.  Not OpenMP: a caller's OMP_NUM_THREADS (torchrun exports 1) must not serialise
// the parsing of a thousand documents.  ZKB_HOST_THREADS overrides; default min(hardware threads, 16).
static unsigned host_threads() {
  static const unsigned n = []() {
    unsigned e = env_u32("ZKB_HOST_THREADS", 0);
    if (e) return e;
    unsigned h = std::thread::hardware_concurrency();
    if (!h) h = 4;
    return h > 16 ? 16u : h;
  }();
  return n;
}
template <class F>
static void parallel_for(uint32_t n, unsigned nthreads, F fn) {
  if (!n) return;
  const unsigned t = nthreads < n ? nthreads : n;
  if (t <= 1) { for (uint32_t i = 0; i < n; i++) fn(i); return; }
  std::atomic<uint32_t> next{0};
  auto body = [&]() { for (;;) { uint32_t i = next.fetch_add(1); if (i >= n) break; fn(i); } };
  std::vector<std::thread> th;
  for (unsigned k = 1; k < t; k++) th.emplace_back(body);
  body();
  for (auto &x : th) x.join();
}

// One group of m <= cap documents, host and device overlapped chunk by chunk: worker threads parse the documents in
// order while this thread, as soon as a chunk's documents are parsed, queues that chunk's H2D copy, proving pipeline
// and D2H copy on its lane; the proofs of a chunk are formatted as soon as its lane reports it done, while later
// chunks are still on the GPU.  (Round 1 ran parse -> H2D -> prove -> D2H -> format strictly in sequence.)
static int fullprove_group(Circuit *c, uint32_t m, const char *const *docs, const size_t *lens, char *proofs,
                           size_t proof_stride, char *publics, size_t public_stride, int *status) {
  cudaStream_t st = c->ctx->stream;
  const uint32_t chunk = c->chunk, nchunks = (m + chunk - 1) / chunk, n_in = c->L.n_inputs;
  std::vector<int> bad(m, 0);
  std::vector<std::string> errs(m);
  if (!fill_blinding(c, m)) return ZKB_ERROR;
  CKR(cudaMemcpyAsync(c->rs, c->h_rs, (size_t)m * 64, cudaMemcpyHostToDevice, st), "h2d rs");
  struct EvBag {
    std::vector<cudaEvent_t> v;
    ~EvBag() { for (auto e : v) cudaEventDestroy(e); }
  } bag;
  bag.v.resize(nchunks + 1);
  for (auto &e : bag.v) CKR(cudaEventCreateWithFlags(&e, cudaEventDisableTiming), "event");
  cudaEvent_t ev_rs = bag.v[nchunks];
  CKR(cudaEventRecord(ev_rs, st), "event record");
  std::unique_ptr<std::atomic<uint32_t>[]> parsed(new std::atomic<uint32_t>[nchunks]);
  for (uint32_t k = 0; k < nchunks; k++) parsed[k].store(0);
  std::atomic<uint32_t> next{0};
  auto parse_body = [&]() {
    for (;;) {
      const uint32_t i = next.fetch_add(1);
      if (i >= m) break;
      uint32_t *dst = c->h_inputs[(size_t)i * n_in].v;
      memset(dst, 0, (size_t)n_in * 32);
      if (pack_inputs(c, docs[i], lens[i], dst, errs[i])) bad[i] = 1;
      parsed[i / chunk].fetch_add(1, std::memory_order_release);
    }
  };
  const unsigned nt = m >= 16 ? host_threads() : 1;
  std::vector<std::thread> workers;
  if (nt <= 1) parse_body();
  else for (unsigned k = 0; k + 1 < nt; k++) workers.emplace_back(parse_body);
  int rc = ZKB_OK;
  const bool group_witness = c->generic && nchunks > 1;      // see run_witness_group
  if (group_witness) {
    for (uint32_t k = 0; k < nchunks; k++) {
      const uint32_t mk = m - k * chunk < chunk ? m - k * chunk : chunk;
      while (parsed[k].load(std::memory_order_acquire) < mk) std::this_thread::yield();
    }
    // (no early return here: the parser threads are joined below)
    cudaError_t e = cudaMemcpyAsync(c->inputs, c->h_inputs, (size_t)m * n_in * 32, cudaMemcpyHostToDevice, st);
    rc = e == cudaSuccess ? run_witness_group(c, m, st) : cuda_fail(e, "h2d inputs");
    cudaEventRecord(ev_rs, st);
  }
  for (uint32_t k = 0; k < nchunks && rc == ZKB_OK; k++) {
    const uint32_t first = k * chunk, mk = m - first < chunk ? m - first : chunk;
    while (parsed[k].load(std::memory_order_acquire) < mk) std::this_thread::yield();
    Lane &ln = c->lanes[k % (uint32_t)c->n_lanes];
    cudaStreamWaitEvent(ln.st, ev_rs, 0);
    cudaError_t e = group_witness ? cudaSuccess
                                  : cudaMemcpyAsync(c->inputs + (size_t)first * n_in, c->h_inputs + (size_t)first * n_in,
                                                    (size_t)mk * n_in * 32, cudaMemcpyHostToDevice, ln.st);
    if (e != cudaSuccess) { rc = cuda_fail(e, "h2d inputs"); break; }
    rc = run_prove_chunk(c, ln, first, mk, !group_witness, nullptr);
    if (rc) break;
    c->last_chunk_m = mk;
    c->last_lane = (int)(k % (uint32_t)c->n_lanes);
    cudaMemcpyAsync(c->h_out + (size_t)first * c->out_stride(), c->out + (size_t)first * c->out_stride(),
                    (size_t)mk * c->out_stride(), cudaMemcpyDeviceToHost, ln.st);
    cudaMemcpyAsync(c->h_status + first, c->status + first, (size_t)mk * 4, cudaMemcpyDeviceToHost, ln.st);
    cudaEventRecord(bag.v[k], ln.st);
  }
  for (auto &w : workers) w.join();
  if (rc) { cudaDeviceSynchronize(); return rc; }
  for (uint32_t k = 0; k < nchunks; k++) {
    const uint32_t first = k * chunk, mk = m - first < chunk ? m - first : chunk;
    CKR(cudaEventSynchronize(bag.v[k]), "prove");
    parallel_for(mk, nt, [&](uint32_t q) {
      const uint32_t i = first + q;
      const uint8_t *o = c->h_out + (size_t)i * c->out_stride();
      int stt = bad[i] ? ZKB_ERROR : c->h_status[i];
      status[i] = stt;
      char *pb = proofs + (size_t)i * proof_stride, *qb = publics + (size_t)i * public_stride;
      pb[0] = 0;
      qb[0] = 0;
      if (stt == 0) {
        std::string pj = proof_to_json(o, false), sj = publics_to_json(o + 256, c->n_public);
        if (pj.size() + 1 <= proof_stride && sj.size() + 1 <= public_stride) {
          memcpy(pb, pj.c_str(), pj.size() + 1);
          memcpy(qb, sj.c_str(), sj.size() + 1);
        } else {
          status[i] = ZKB_SHORT_BUFFER;
        }
      }
    });
  }
  for (uint32_t i = 0; i < m; i++)
    if (bad[i]) set_error(errs[i]);
  return ZKB_OK;
}

extern "C" {

// n inputs.json documents -> n proof.json / public.json strings (NUL terminated) at the given strides.
// status[i]: 0 ok, 1 malformed inputs, 4 circuit assert failed.  Returns 0 unless the batch as a whole failed.
int zkb_fullprove_batch(zkb_circuit *h, int n, const char *const *inputs_json, const size_t *inputs_len, char *proofs,
                        size_t proof_stride, char *publics, size_t public_stride, int *status) {
  if (!h || (n > 0 && (!inputs_json || !inputs_len || !proofs || !publics || !status))) { set_error("null argument"); return ZKB_ERROR; }
  Circuit *c = h->c;
  if (!c->has_witness) { set_error("circuit was loaded without a wasm: no witness generator"); return ZKB_ERROR; }
  if (n <= 0) return ZKB_OK;
  if (proof_stride < 1 || public_stride < 1) { set_error("zero output stride"); return ZKB_ERROR; }
  std::lock_guard<std::mutex> g(c->mu);
  CKR(cudaSetDevice(c->ctx->device), "set device");
  uint32_t chunk = default_chunk(c), group = env_u32("ZKB_GROUP", 1024);
  uint32_t cap = (uint32_t)n < group ? (uint32_t)n : group;
  int rc = ensure_workspace(c, cap > c->cap ? cap : c->cap, chunk);
  if (rc) return rc;
  for (uint32_t first = 0; first < (uint32_t)n; first += c->cap) {
    uint32_t m = (uint32_t)n - first < c->cap ? (uint32_t)n - first : c->cap;
    rc = fullprove_group(c, m, inputs_json + first, inputs_len + first, proofs + (size_t)first * proof_stride, proof_stride,
                         publics + (size_t)first * public_stride, public_stride, status + first);
    if (rc) return rc;
  }
  return ZKB_OK;
}

// One proof: the drop-in for prover.Prove(zkey, wasm, inputs) once the circuit is loaded.
// proof_size / public_size: in = buffer capacity, out = bytes written (or needed on ZKB_SHORT_BUFFER).
int zkb_fullprove(zkb_circuit *h, const char *inputs_json, size_t inputs_len, char *proof_buf, size_t *proof_size,
                  char *public_buf, size_t *public_size, char *err, size_t errmax) {
  char pb[1024], qb[2048];
  int status = 0;
  const char *ins[1] = {inputs_json};
  size_t lens[1] = {inputs_len};
  int rc = zkb_fullprove_batch(h, 1, ins, lens, pb, sizeof pb, qb, sizeof qb, &status);
  if (rc == 0 && status != 0) rc = status;
  if (rc) {
    if (rc == ZKB_ASSERT_FAILED) set_error("circuit assert failed (witness generation, exception code 4)");
    if (err && errmax) snprintf(err, errmax, "%s", zkb_last_error());
    return rc;
  }
  int r1 = copy_out(pb, proof_buf, proof_size), r2 = copy_out(qb, public_buf, public_size);
  if (r1 || r2) {
    if (err && errmax) snprintf(err, errmax, "short buffer");
    return ZKB_SHORT_BUFFER;
  }
  return ZKB_OK;
}

// .wtns (snarkjs/circom binary witness) for one inputs.json; *wtns_len in = capacity, out = size
int zkb_witness(zkb_circuit *h, const char *inputs_json, size_t inputs_len, void *wtns_out, size_t *wtns_len) {
  Circuit *c = h->c;
  if (!c->has_witness) { set_error("circuit was loaded without a wasm: no witness generator"); return ZKB_ERROR; }
  std::lock_guard<std::mutex> g(c->mu);
  CKR(cudaSetDevice(c->ctx->device), "set device");
  size_t need = 4 + 4 + 4 + 12 + (4 + 32 + 4) + 12 + (size_t)c->n_vars * 32;
  if (!wtns_out || *wtns_len < need) { *wtns_len = need; return ZKB_SHORT_BUFFER; }
  int rc = ensure_workspace(c, c->cap ? c->cap : 1, c->chunk ? c->chunk : default_chunk(c));
  if (rc) return rc;
  std::string err;
  memset(c->h_inputs, 0, (size_t)c->L.n_inputs * 32);
  if (pack_inputs(c, inputs_json, inputs_len, c->h_inputs[0].v, err)) { set_error(err); return ZKB_ERROR; }
  cudaStream_t st = c->ctx->stream;
  CKR(cudaMemcpyAsync(c->inputs, c->h_inputs, (size_t)c->L.n_inputs * 32, cudaMemcpyHostToDevice, st), "h2d");
  if ((rc = run_witness(c, c->lanes[0], 0, 1, st))) return rc;
  uint8_t *o = (uint8_t *)wtns_out;
  static const uint32_t RMOD[8] = {0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u,
                                   0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
  auto p32 = [&](uint32_t v) { memcpy(o, &v, 4); o += 4; };
  auto p64 = [&](uint64_t v) { memcpy(o, &v, 8); o += 8; };
  memcpy(o, "wtns", 4); o += 4;
  p32(2); p32(2);
  p32(1); p64(40);
  p32(32); memcpy(o, RMOD, 32); o += 32; p32(c->n_vars);
  p32(2); p64((uint64_t)c->n_vars * 32);
  CKR(cudaMemcpyAsync(o, c->wtns, (size_t)c->n_vars * 32, cudaMemcpyDeviceToHost, st), "d2h witness");
  CKR(cudaMemcpyAsync(c->h_status, c->status, 4, cudaMemcpyDeviceToHost, st), "d2h status");
  CKR(cudaStreamSynchronize(st), "sync");
  *wtns_len = need;
  if (c->h_status[0]) { set_error("circuit assert failed (exception code 4)"); return ZKB_ASSERT_FAILED; }
  return ZKB_OK;
}

// Groth16 from a caller-supplied witness (.wtns bytes): what go-rapidsnark's Groth16ProverRaw does.
int zkb_prove_wtns_stages(zkb_circuit *h, const void *wtns, size_t wtns_size, char *proof_buf, size_t *proof_size,
                          char *public_buf, size_t *public_size, float *stage_ms);
int zkb_prove_wtns(zkb_circuit *h, const void *wtns, size_t wtns_size, char *proof_buf, size_t *proof_size,
                   char *public_buf, size_t *public_size) {
  return zkb_prove_wtns_stages(h, wtns, wtns_size, proof_buf, proof_size, public_buf, public_size, nullptr);
}
// same, with the device stage times of the pass (8 floats as zkb_batch_prove_resident; the witness slot is 0)
int zkb_prove_wtns_stages(zkb_circuit *h, const void *wtns, size_t wtns_size, char *proof_buf, size_t *proof_size,
                          char *public_buf, size_t *public_size, float *stage_ms) {
  Circuit *c = h->c;
  const uint8_t *b = (const uint8_t *)wtns;
  const bool resident = wtns == nullptr;   // NULL: prove again from the witness already on the device (measurement aid)
  if (resident) {
    if (!c->wtns) { set_error("no witness resident on the device yet"); return ZKB_ERROR; }
    std::lock_guard<std::mutex> g(c->mu);
    CKR(cudaSetDevice(c->ctx->device), "set device");
    if (c->shard_n > 1) CKR(cudaMemsetAsync(c->status, 0, 4, c->ctx->stream), "memset status");
    int rc = prove_group(c, 1, false, stage_ms);
    if (rc) return rc;
    if (c->shard_n > 1) {
      int stt = 0;
      CKR(cudaMemcpy(&stt, c->status, 4, cudaMemcpyDeviceToHost), "d2h status");
      if (stt == 7) { set_error("sharded prove: timed out waiting for a peer's partial sums"); return ZKB_ERROR; }
      if (c->shard_rank != 0) { if (proof_size) *proof_size = 0; if (public_size) *public_size = 0; return ZKB_OK; }
    }
    std::string pj = proof_to_json(c->h_out, false), sj = publics_to_json(c->h_out + 256, c->n_public);
    int r1 = copy_out(pj, proof_buf, proof_size), r2 = copy_out(sj, public_buf, public_size);
    return (r1 || r2) ? ZKB_SHORT_BUFFER : ZKB_OK;
  }
  if (wtns_size < 12 || memcmp(b, "wtns", 4) != 0) { set_error("wtns: bad magic"); return ZKB_ERROR; }
  static const uint32_t RMOD[8] = {0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u,
                                   0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
  uint32_t nsec;
  memcpy(&nsec, b + 8, 4);
  size_t p = 12;
  const uint8_t *data = nullptr;
  uint64_t data_sz = 0;
  uint32_t nw = 0;
  bool have_hdr = false;
  for (uint32_t i = 0; i < nsec && p + 12 <= wtns_size; i++) {
    uint32_t id;
    uint64_t sz;
    memcpy(&id, b + p, 4);
    memcpy(&sz, b + p + 4, 8);
    p += 12;
    if (sz > wtns_size - p) { set_error("wtns: truncated"); return ZKB_ERROR; }
    if (id == 1) {
      uint32_t n8 = 0;
      if (sz < 40) { set_error("wtns: header section too short"); return ZKB_ERROR; }
      memcpy(&n8, b + p, 4);
      if (n8 != 32 || memcmp(b + p + 4, RMOD, 32) != 0) { set_error("wtns: not a BN254 scalar-field witness (n8 / prime)"); return ZKB_ERROR; }
      memcpy(&nw, b + p + 36, 4);
      have_hdr = true;
    }
    if (id == 2) { data = b + p; data_sz = sz; }
    p += sz;
  }
  if (!have_hdr) { set_error("wtns: no header section"); return ZKB_ERROR; }
  if (!data) { set_error("wtns: no data section"); return ZKB_ERROR; }
  if (nw != c->n_vars) { set_error("wtns: witness length does not match the zkey"); return ZKB_INVALID_WITNESS_LENGTH; }
  if (data_sz < (uint64_t)nw * 32) { set_error("wtns: data section shorter than nVars x 32 bytes"); return ZKB_INVALID_WITNESS_LENGTH; }
  // canonical values only (they feed Montgomery products and MSM digits); a sharded key with its peers attached
  // checks and uploads only its own slice, the rest arrives over NVLink
  const bool slices = c->shard_n > 1 && c->wt_attached == c->shard_n - 1;
  const uint32_t lo = slices ? (uint32_t)(((uint64_t)nw * c->shard_rank) / c->shard_n) : 0;
  const uint32_t hi = slices ? (uint32_t)(((uint64_t)nw * (c->shard_rank + 1)) / c->shard_n) : nw;
  std::atomic<int> bad_value{0};
  const uint32_t piece = 1u << 15;
  parallel_for((hi - lo + piece - 1) / piece, nw >= (1u << 18) ? host_threads() : 1, [&](uint32_t q) {
    const uint32_t a = lo + q * piece, b = std::min(hi, a + piece);
    for (uint32_t i = a; i < b; i++) {
      uint32_t top;
      memcpy(&top, data + (size_t)i * 32 + 28, 4);
      if (top < RMOD[7]) continue;
      uint32_t v[8];
      memcpy(v, data + (size_t)i * 32, 32);
      bool lt = false;
      for (int k = 7; k >= 0; k--) { if (v[k] != RMOD[k]) { lt = v[k] < RMOD[k]; break; } }
      if (!lt) bad_value.store(1);
    }
    if (slices) memcpy(c->h_wt + (size_t)(a - lo) * 32, data + (size_t)a * 32, (size_t)(b - a) * 32);   // pinned staging
  });
  if (bad_value.load()) { set_error("wtns: value not in [0, r)"); return ZKB_ERROR; }
  std::lock_guard<std::mutex> g(c->mu);
  CKR(cudaSetDevice(c->ctx->device), "set device");
  int rc = ensure_workspace(c, c->cap ? c->cap : 1, c->chunk ? c->chunk : default_chunk(c));
  if (rc) return rc;
  int *wt_status = reinterpret_cast<int *>(c->wt_buf + c->wt_flag_off + 64);   // set to 7 when a peer's slice never arrives
  if (slices) {
    cudaStream_t st = c->ctx->stream;
    c->wt_epoch++;
    CKR(cudaMemsetAsync(wt_status, 0, 4, st), "memset");
    CKR(cudaMemcpyAsync(c->wt_buf, c->h_wt, (size_t)(hi - lo) * 32, cudaMemcpyHostToDevice, st), "h2d witness slice");
    k_wt_set_flag<<<1, 1, 0, st>>>(reinterpret_cast<uint32_t *>(c->wt_buf + c->wt_flag_off), c->wt_epoch);
    WitnessPeers P;
    for (int r = 0; r < SHARD_MAX_RANKS; r++) {
      P.slice[r] = reinterpret_cast<const Fr *>(c->wt_peer[r]);
      P.flag[r] = c->wt_peer[r] ? reinterpret_cast<const uint32_t *>(c->wt_peer[r] + c->wt_flag_off) : nullptr;
    }
    k_wt_wait<<<1, 32, 0, st>>>(P, c->shard_n, c->wt_epoch, wt_status);
    k_wt_gather<<<(nw + 255) / 256, 256, 0, st>>>(P, c->shard_n, nw, c->wtns);
    g_launches += 3;
    CKR(cudaGetLastError(), "witness slice exchange");
  } else {
    CKR(cudaMemcpyAsync(c->wtns, data, (size_t)nw * 32, cudaMemcpyHostToDevice, c->ctx->stream), "h2d witness");
  }
  if (c->shard_n > 1) CKR(cudaMemsetAsync(c->status, 0, 4, c->ctx->stream), "memset status");
  if ((rc = prove_group(c, 1, false, stage_ms))) return rc;
  if (c->shard_n > 1) {
    int stt = 0;
    CKR(cudaMemcpy(&stt, c->status, 4, cudaMemcpyDeviceToHost), "d2h status");
    if (stt == 7) { set_error("sharded prove: timed out waiting for a peer's partial sums"); return ZKB_ERROR; }
    if (slices) {
      CKR(cudaMemcpy(&stt, wt_status, 4, cudaMemcpyDeviceToHost), "d2h status");
      if (stt == 7) { set_error("sharded prove: timed out waiting for a peer's witness slice"); return ZKB_ERROR; }
    }
    if (c->shard_rank != 0) {                            // only rank 0 assembles and returns the proof
      if (proof_size) *proof_size = 0;
      if (public_size) *public_size = 0;
      if (proof_buf) proof_buf[0] = 0;
      if (public_buf) public_buf[0] = 0;
      return ZKB_OK;
    }
  }
  std::string pj = proof_to_json(c->h_out, false), sj = publics_to_json(c->h_out + 256, c->n_public);
  int r1 = copy_out(pj, proof_buf, proof_size), r2 = copy_out(sj, public_buf, public_size);
  return (r1 || r2) ? ZKB_SHORT_BUFFER : ZKB_OK;
}

// ---- rapidsnark prover.h ------------------------------------------------------------------------------------------
// Key cache of the rapidsnark-shaped entry points: one resident circuit per process, identified by the key's size
// and a 64-bit hash over ALL of its bytes.  The full hash (8 bytes per step, ~10 ms for the 55 MB census key) runs
// when a (pointer, size) pair is first seen or when its cheap fingerprint changed; the fingerprint checked on every
// call covers sections 1-3 completely (protocol, Groth16 header with alpha/beta/gamma/delta, IC: two keys of one
// circuit from different setups always differ there) plus every 4,099th byte of the rest.
static uint64_t mix64(uint64_t h, uint64_t v) {
  h ^= v;
  h *= 0x9E3779B97F4A7C15ull;
  return h ^ (h >> 29);
}
static uint64_t hash_all(const uint8_t *b, size_t n) {
  uint64_t h = 0x243F6A8885A308D3ull ^ n;
  size_t i = 0;
  for (; i + 8 <= n; i += 8) { uint64_t v; memcpy(&v, b + i, 8); h = mix64(h, v); }
  uint64_t tail = 0;
  if (i < n) memcpy(&tail, b + i, n - i);
  return mix64(h, tail);
}
static uint64_t key_fingerprint(const uint8_t *b, size_t n) {
  // sections 1..3 sit at the front of a snarkjs zkey: 12-byte file header, then {id, len, payload}
  size_t p = 12, head = n < 12 ? n : 12;
  for (int s = 0; s < 3 && p + 12 <= n; s++) {
    uint64_t sz;
    memcpy(&sz, b + p + 4, 8);
    if (sz > n - p - 12) break;
    p += 12 + (size_t)sz;
    head = p;
  }
  if (head > ((size_t)1 << 20)) head = (size_t)1 << 20;
  uint64_t h = hash_all(b, head);
  for (size_t i = head; i < n; i += 4099) h = mix64(h, b[i]);
  return h;
}
struct KeyCache {
  std::mutex mu;
  zkb_ctx *ctx = nullptr;
  zkb_circuit *circuit = nullptr;
  const void *ptr = nullptr;
  size_t size = 0;
  uint64_t fingerprint = 0, full = 0;
};
static KeyCache g_keys;

// the circuit for this key, loading it when it is not the resident one (g_keys.mu held)
static int cached_circuit(const void *zkey, size_t size, zkb_circuit **out) {
  KeyCache &k = g_keys;
  if (!zkey || size < 12) { set_error("zkey: empty buffer"); return ZKB_ERROR; }
  if (!k.ctx && zkb_ctx_create((int)env_u32("ZKB_DEVICE", 0) % (zkb_device_count() ? zkb_device_count() : 1), &k.ctx)) return ZKB_ERROR;
  const uint64_t fp = key_fingerprint((const uint8_t *)zkey, size);
  if (k.circuit && k.size == size && k.fingerprint == fp && k.ptr == zkey) { *out = k.circuit; return ZKB_OK; }
  const uint64_t full = hash_all((const uint8_t *)zkey, size);
  if (!(k.circuit && k.size == size && k.full == full)) {
    if (k.circuit) { zkb_circuit_destroy(k.circuit); k.circuit = nullptr; }
    int rc = zkb_load_circuit(k.ctx, zkey, size, nullptr, 0, &k.circuit);
    if (rc) return rc;
  }
  k.ptr = zkey; k.size = size; k.fingerprint = fp; k.full = full;
  *out = k.circuit;
  return ZKB_OK;
}

// rapidsnark's prover.h entry point, same symbol and signature (go-rapidsnark links against it):
// returns 0 PROVER_OK, 1 PROVER_ERROR, 2 PROVER_ERROR_SHORT_BUFFER.
int groth16_prover(const void *zkey_buffer, unsigned long zkey_size, const void *wtns_buffer, unsigned long wtns_size,
                   char *proof_buffer, unsigned long *proof_size, char *public_buffer, unsigned long *public_size,
                   char *error_msg, unsigned long error_msg_maxsize) {
  auto fail = [&](int rc) {
    if (error_msg && error_msg_maxsize) snprintf(error_msg, error_msg_maxsize, "%s", zkb_last_error());
    return rc == ZKB_SHORT_BUFFER ? 2 : 1;
  };
  if (!proof_size || !public_size) { set_error("null size pointer"); return fail(1); }
  std::lock_guard<std::mutex> g(g_keys.mu);
  zkb_circuit *c = nullptr;
  int rc = cached_circuit(zkey_buffer, zkey_size, &c);
  if (rc) return fail(rc);
  size_t ps = *proof_size, qs = *public_size;
  rc = zkb_prove_wtns(c, wtns_buffer, wtns_size, proof_buffer, &ps, public_buffer, &qs);
  *proof_size = ps;
  *public_size = qs;
  if (rc) return fail(rc);
  return 0;
}

// Later rapidsnark releases (prover.h of v0.0.2+, bound by go-rapidsnark/prover >= v0.0.11) added these; exported so
// that newer wrappers link too.  `unsigned long` and their `unsigned long long` are the same 64-bit type on LP64.
void groth16_proof_size(unsigned long *proof_size) {
  if (proof_size) *proof_size = 810;      // 8 coordinates of <= 77 digits + JSON punctuation, as rapidsnark reports
}
int groth16_public_size_for_zkey_buf(const void *zkey_buffer, unsigned long zkey_size, unsigned long *public_size,
                                     char *error_msg, unsigned long error_msg_maxsize) {
  ZkeyView z;
  std::string err;
  if (!zkey_buffer || !public_size || !parse_zkey((const uint8_t *)zkey_buffer, zkey_size, z, err)) {
    if (error_msg && error_msg_maxsize) snprintf(error_msg, error_msg_maxsize, "%s", err.empty() ? "null argument" : err.c_str());
    return 1;
  }
  *public_size = 4 + (unsigned long)z.n_public * 82;     // ["<=77 digits",...] + NUL
  return 0;
}
static bool read_file(const char *path, std::vector<uint8_t> &out, std::string &err) {
  FILE *f = path ? fopen(path, "rb") : nullptr;
  if (!f) { err = std::string("cannot open ") + (path ? path : "(null)"); return false; }
  fseek(f, 0, SEEK_END);
  long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  out.resize(n > 0 ? (size_t)n : 0);
  bool ok = n >= 0 && fread(out.data(), 1, out.size(), f) == out.size();
  fclose(f);
  if (!ok) err = std::string("cannot read ") + path;
  return ok;
}
int groth16_public_size_for_zkey_file(const char *zkey_fname, unsigned long *public_size, char *error_msg,
                                      unsigned long error_msg_maxsize) {
  std::vector<uint8_t> buf;
  std::string err;
  if (!read_file(zkey_fname, buf, err)) {
    if (error_msg && error_msg_maxsize) snprintf(error_msg, error_msg_maxsize, "%s", err.c_str());
    return 1;
  }
  return groth16_public_size_for_zkey_buf(buf.data(), buf.size(), public_size, error_msg, error_msg_maxsize);
}
// The file is re-read only when its path, size or mtime changed since the last call.
int groth16_prover_zkey_file(const char *zkey_file_path, const void *wtns_buffer, unsigned long wtns_size,
                             char *proof_buffer, unsigned long *proof_size, char *public_buffer,
                             unsigned long *public_size, char *error_msg, unsigned long error_msg_maxsize) {
  static std::mutex fmu;
  static std::string cached_path;
  static std::vector<uint8_t> cached_bytes;
  static long long cached_mtime = 0, cached_len = -1;
  std::lock_guard<std::mutex> g(fmu);
  struct stat sb;
  if (!zkey_file_path || stat(zkey_file_path, &sb) != 0) {
    if (error_msg && error_msg_maxsize) snprintf(error_msg, error_msg_maxsize, "cannot stat %s", zkey_file_path ? zkey_file_path : "(null)");
    return 1;
  }
  const long long mt = (long long)sb.st_mtim.tv_sec * 1000000000ll + sb.st_mtim.tv_nsec;
  if (cached_path != zkey_file_path || cached_mtime != mt || cached_len != (long long)sb.st_size) {
    std::string err;
    if (!read_file(zkey_file_path, cached_bytes, err)) {
      if (error_msg && error_msg_maxsize) snprintf(error_msg, error_msg_maxsize, "%s", err.c_str());
      cached_path.clear();
      return 1;
    }
    cached_path = zkey_file_path; cached_mtime = mt; cached_len = (long long)sb.st_size;
  }
  return groth16_prover(cached_bytes.data(), cached_bytes.size(), wtns_buffer, wtns_size, proof_buffer, proof_size,
                        public_buffer, public_size, error_msg, error_msg_maxsize);
}

}  // extern "C"
