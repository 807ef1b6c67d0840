// Device-resident raw MSM / NTT sessions (BASELINE.json config 5): synthetic bases and scalars are
// generated on the GPU, a large G1 MSM is split BY POINT RANGE into sub-MSMs of <= 2^17 points that run as batch items
// of the same kernels the prover uses, and on several GPUs every rank takes a contiguous range of sub-MSMs and PUSHES
// its partial sum straight into rank 0's memory over NVLink (CUDA IPC mapping, st.release.sys flag); rank 0's combine
// kernel waits on the flags and adds.  No NCCL and no host round trip on the data path ("gather 8 points
// via P2P, 7 adds").
//
// The 4-step multi-GPU NTT lives here too: N = N1 * N2, rank g owns N2/G columns; local length-N1 transforms, then each
// rank READS its slice of every peer's buffer through the IPC mappings (the all-to-all is fused into the transposing
// load, twiddle multiply included), then local length-N2 transforms.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <vector>
#include "common.h"
#include "ntt.cuh"
#include "msm.cuh"

namespace zkb {

#define CKR(x, what) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return cuda_fail(e_, what); } while (0)

static constexpr int MAX_RANKS = 16;
static constexpr long long SPIN_LIMIT_CYCLES = 8000000000LL;   // ~4 s at 1.9 GHz: a stuck peer ends in an error, not a hang

// ---- synthetic data ------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
template <class F>
__device__ __forceinline__ F hash_field(uint64_t seed, uint64_t idx, uint32_t ctr, uint32_t top_mask) {
  F r;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    uint64_t h = splitmix64(seed ^ splitmix64(idx * 4 + i) ^ ((uint64_t)ctr << 40));
    r.v[2 * i] = (uint32_t)h;
    r.v[2 * i + 1] = (uint32_t)(h >> 32);
  }
  r.v[7] &= top_mask;
  return r;
}

// base idx = try-and-increment hash to BN254 G1: x = H(seed, idx, ctr), y = sqrt(x^3 + 3) when it exists
// (q = 3 mod 4: sqrt(a) = a^((q+1)/4)).  Stored affine, Montgomery form (x is taken as the Montgomery residue).
__global__ void k_gen_points(Affine<Fq> *out, uint64_t first, uint64_t count, uint64_t seed) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  // (q + 1) / 4
  const uint32_t E[8] = {0xb61f3f52u, 0x4f082305u, 0x5a1c72a3u, 0x65e05aa4u, 0xa0605617u, 0x6e14116du, 0xb84c680au, 0x0c19139cu};
  Fq three = Fq::one() + Fq::one() + Fq::one();
  for (uint32_t ctr = 0;; ctr++) {
    Fq x = hash_field<Fq>(seed, first + i, ctr, 0x3fffffffu);
    Fq::final_sub(x.v);                                  // < 2^254 < 2q
    Fq rhs = x.sqrc().mulc(x) + three;
    Fq y = Fq::one();
    for (int b = 255; b >= 0; b--) {
      y = y.sqrc();
      if ((E[b >> 5] >> (b & 31)) & 1) y = y.mulc(rhs);
    }
    if (y.sqrc() == rhs) {
      if (splitmix64(seed + 77 + first + i) & 1) y = y.neg();
      out[i].x = x;
      out[i].y = y;
      return;
    }
  }
}
// scalar idx: uniform in [0, 2^253) (< r), canonical form
__global__ void k_gen_scalars(Fr *out, uint64_t first, uint64_t count, uint64_t seed) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  out[i] = hash_field<Fr>(seed ^ 0x5ca1ab1eull, first + i, 0, 0x1fffffffu);
}

// ---- partial-sum exchange ------------------------------------------------------------------------
struct alignas(256) ExSlot {
  XYZZ<Fq> p;
  uint32_t flag;
};

// fold `n` XYZZ points into one (one CTA of 256 threads)
__global__ void __launch_bounds__(256) k_fold(const XYZZ<Fq> *in, uint32_t n, XYZZ<Fq> *out) {
  __shared__ XYZZ<Fq> sm[256];
  const uint32_t t = threadIdx.x;
  XYZZ<Fq> acc = XYZZ<Fq>::infinity();
  for (uint32_t i = t; i < n; i += 256) xyzz_add_ni(&acc, in + i);
  sm[t] = acc;
  __syncthreads();
  for (uint32_t stride = 128; stride >= 1; stride >>= 1) {
    if (t < stride) {
      xyzz_add_ni(&acc, sm + t + stride);
      sm[t] = acc;
    }
    __syncthreads();
  }
  if (t == 0) *out = acc;
}

// write this rank's partial sum into the root's slot (peer memory when slot belongs to another GPU), then release
// the epoch flag system-wide
__global__ void k_publish(ExSlot *slot, const XYZZ<Fq> *partial, uint32_t epoch) {
  const uint32_t t = threadIdx.x;
  const uint32_t *src = reinterpret_cast<const uint32_t *>(partial);
  uint32_t *dst = reinterpret_cast<uint32_t *>(&slot->p);
  if (t < sizeof(XYZZ<Fq>) / 4) dst[t] = src[t];
  __threadfence_system();
  __syncthreads();
  if (t == 0) asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(&slot->flag), "r"(epoch) : "memory");
}

// root: wait until every slot carries `epoch`, add the partial sums in slot order, emit the affine canonical result
__global__ void k_combine(ExSlot *slots, int nslots, uint32_t epoch, Affine<Fq> *out, int *status) {
  if (threadIdx.x || blockIdx.x) return;
  const long long t0 = clock64();
  for (int s = 0; s < nslots; s++) {
    uint32_t f;
    do {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(f) : "l"(&slots[s].flag) : "memory");
      if (f != epoch && clock64() - t0 > SPIN_LIMIT_CYCLES) { *status = 1 + s; return; }
    } while (f != epoch);
  }
  XYZZ<Fq> acc = XYZZ<Fq>::infinity();
  for (int s = 0; s < nslots; s++) {
    XYZZ<Fq> p = slots[s].p;
    xyzz_add_ni(&acc, &p);
  }
  Affine<Fq> a;
  xyzz_to_affine_ni(&acc, &a);
  a.x = a.x.from_mont();
  a.y = a.y.from_mont();
  *out = a;
  *status = 0;
}

__global__ void k_affine_from_mont(const Affine<Fq> *in, Affine<Fq> *out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Affine<Fq> a = in[i];
  a.x = a.x.from_mont();
  a.y = a.y.from_mont();
  out[i] = a;
}

struct MsmSession {
  int device = 0;
  cudaStream_t st = nullptr;
  uint32_t n_sub = 0, subs = 0;       // this rank: subs sub-MSMs of n_sub points
  bool vb = false;                    // variable base: no window table, `windows` bucket sets per sub-MSM, Horner at the end
  int windows = 0;                    // windows of the digit decomposition (cfg.windows is 1 in the variable-base form)
  uint32_t items = 0;                 // batch items of one run: subs (fixed base) or subs * windows (variable base)
  uint64_t first = 0, total = 0;      // global index of this rank's first point; points in the whole MSM
  uint64_t seed = 0;
  MsmCfg cfg;
  Affine<Fq> *bases = nullptr;
  Fr *scalars = nullptr;
  MsmTable<Fq> tab;
  MsmSort sort;
  MsmWork<Fq> work;
  XYZZ<Fq> *sub_out = nullptr, *partial = nullptr;
  ExSlot *xbuf = nullptr;             // root: MAX_RANKS slots (exported); others: unused
  ExSlot *root_x = nullptr;           // where this rank publishes (own xbuf on the root, IPC mapping elsewhere)
  bool root_is_ipc = false;
  int slot = 0;
  uint32_t epoch = 0;
  Affine<Fq> *result = nullptr;
  int *status = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr, e2 = nullptr;
  float table_ms = 0, gen_ms = 0;
};


// ---- 4-step NTT across GPUs ----------------------------------------------------------------------------
// N = N1 * N2, x[n1 * N2 + n2].  Rank g owns the columns n2 in [g N2/G, (g+1) N2/G), each stored contiguously:
// B_g[c][n1].  (1) forward DIF of length N1 on every column: B_g[c][p] = Y[n2][k1 = bitrev(p)].  (2) release flag.
// (3) rank g' takes positions p in [g' N1/G, (g'+1) N1/G) of EVERY rank's B (32 x 32 tiles read through the IPC
// mappings, multiplied by omega_N^(n2 k1), written transposed): C_g'[p_local][n2].  (4) forward DIF of length N2 on
// every row: C_g'[p_local][q] = X[bitrev_N1(p) + N1 * bitrev_N2(q)].
struct NttCtl { uint32_t ready; uint32_t pad[63]; };   // head of the exported allocation (256 B), then B

__global__ void k_set_flag(uint32_t *flag, uint32_t epoch) {
  __threadfence_system();
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(epoch) : "memory");
}
struct PeerPtrs { const uint8_t *base[MAX_RANKS]; };
__global__ void k_wait_flags(PeerPtrs peers, int n, uint32_t epoch, int *status) {
  if (threadIdx.x || blockIdx.x) return;
  const long long t0 = clock64();
  for (int g = 0; g < n; g++) {
    const uint32_t *flag = &reinterpret_cast<const NttCtl *>(peers.base[g])->ready;
    uint32_t f;
    do {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(f) : "l"(flag) : "memory");
      if (f != epoch && clock64() - t0 > SPIN_LIMIT_CYCLES) { *status = 1 + g; return; }
    } while (f != epoch);
  }
}

__device__ __forceinline__ Fr ld_fr_volatile(const Fr *p) {   // peer data: bypass L1 (no stale lines across epochs)
  const uint4 *q = reinterpret_cast<const uint4 *>(p);
  uint4 a = __ldcg(q), b = __ldcg(q + 1);
  Fr r;
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}

// grid (N2 / 32, (N1/G) / 32); block 32 x 8.  tile: 32 columns n2 x 32 positions p.
__global__ void __launch_bounds__(256) k_gather_transpose(PeerPtrs peers, Fr *C, int l1, int l2, uint32_t cols_per_rank,
                                                          uint32_t p0, const Fr *__restrict__ tw_lo,
                                                          const Fr *__restrict__ tw_hi, int status_ok_only,
                                                          const int *status) {
  __shared__ uint32_t tile[8][32][33];   // limb-planar, padded
  if (status_ok_only && *status) return;
  const uint32_t N1 = 1u << l1, N2 = 1u << l2;
  const uint32_t n2_0 = blockIdx.x * 32, pl_0 = blockIdx.y * 32;
  const uint32_t tx = threadIdx.x, ty = threadIdx.y;
  const uint64_t nmask = ((uint64_t)N1 << l2) - 1;
  for (uint32_t r = ty; r < 32; r += 8) {              // row r of the tile = column n2_0 + r, lanes = 32 consecutive p
    const uint32_t n2 = n2_0 + r, g = n2 / cols_per_rank, c = n2 % cols_per_rank;
    const Fr *B = reinterpret_cast<const Fr *>(peers.base[g] + sizeof(NttCtl));
    const uint32_t p = p0 + pl_0 + tx;
    Fr v = ld_fr_volatile(B + (size_t)c * N1 + p);
    const uint32_t k1 = __brev(p) >> (32 - l1);
    const uint64_t e = ((uint64_t)n2 * k1) & nmask;
    Fr w = tw_lo[e & 8191] * tw_hi[e >> 13];
    v = v * w;
#pragma unroll
    for (int l = 0; l < 8; l++) tile[l][r][tx] = v.v[l];
  }
  __syncthreads();
  for (uint32_t r = ty; r < 32; r += 8) {              // row r = position pl_0 + r, lanes = 32 consecutive n2
    Fr v;
#pragma unroll
    for (int l = 0; l < 8; l++) v.v[l] = tile[l][tx][r];
    uint4 *q = reinterpret_cast<uint4 *>(C + (size_t)(pl_0 + r) * N2 + n2_0 + tx);
    q[0] = make_uint4(v.v[0], v.v[1], v.v[2], v.v[3]);
    q[1] = make_uint4(v.v[4], v.v[5], v.v[6], v.v[7]);
  }
  (void)N2;
}

// tw[e] = base^(e << shift), e < count (Montgomery form); pw[k] = omega_N^(2^k)
__global__ void k_pow_table(Fr *out, const Fr *pw, uint32_t count, int shift) {
  uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= count) return;
  Fr acc = Fr::one();
  uint64_t x = (uint64_t)e << shift;
  for (int k = 0; x; k++, x >>= 1)
    if (x & 1) acc = acc * pw[k];
  out[e] = acc;
}
__global__ void k_root_pw(Fr *pw, int logn) {        // pw[k] = omega_{2^logn}^(2^k), omega_{2^28} = 5^((r-1)/2^28)
  if (threadIdx.x || blockIdx.x) return;
  const uint32_t W28[8] = {0x80d13d9cu, 0x636e7355u, 0x2445ffd6u, 0xa22bf374u, 0x1eb203d8u, 0x56452ac0u, 0x2963f9e7u, 0x1860ef94u};
  Fr g;
  for (int i = 0; i < 8; i++) g.v[i] = W28[i];
  for (int i = 28; i > logn; i--) g = g.sqr();
  for (int k = 0; k < 32; k++) { pw[k] = g; g = g.sqr(); }
}
// B_g[c][n1] = to_mont(hash(n1 * N2 + first_col + c))
__global__ void k_fill_columns(Fr *B, int l1, int l2, uint32_t first_col, uint32_t cols, uint64_t seed) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ((size_t)cols << l1)) return;
  uint32_t c = (uint32_t)(i >> l1), n1 = (uint32_t)(i & ((1u << l1) - 1));
  uint64_t idx = ((uint64_t)n1 << l2) + first_col + c;
  B[i] = hash_field<Fr>(seed ^ 0x5ca1ab1eull, idx, 0, 0x1fffffffu).to_mont();
}

struct NttDist {
  int device = 0, rank = 0, nranks = 1, logn = 0, l1 = 0, l2 = 0;
  uint64_t seed = 0;
  cudaStream_t st = nullptr;
  uint8_t *alloc = nullptr;          // [NttCtl][B]
  Fr *B = nullptr, *C = nullptr, *tw_lo = nullptr, *tw_hi = nullptr;
  NttPlan plan1, plan2;
  PeerPtrs peers;
  bool ipc[MAX_RANKS];
  uint32_t epoch = 0;
  int *status = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr, e2 = nullptr, e3 = nullptr;
};
}  // namespace zkb

using namespace zkb;

extern "C" {

struct zkb_msm_session;

// One rank's share of a 2^logn-point synthetic G1 MSM: points [rank * N / nranks, (rank + 1) * N / nranks), cut into
// sub-MSMs of at most 2^17 points.  Generates bases + scalars on the device and builds the fixed-base window table.
void zkb_msm_session_destroy(zkb_msm_session *h);
int zkb_msm_session_create(int device, int logn, int rank, int nranks, uint64_t seed, int window_bits,
                           zkb_msm_session **out) {
  if (require_device()) return ZKB_ERROR;
  if (logn < 10 || logn > 27 || nranks < 1 || nranks > MAX_RANKS || rank < 0 || rank >= nranks ||
      (nranks & (nranks - 1)) || (window_bits & 0xff) < 12 || (window_bits & 0xff) > 16 || (window_bits & ~0x1ff)) {
    set_error("msm session: logn in [10,27], nranks a power of two <= 16, window in [12,16] (+ 0x100: variable base)");
    return ZKB_ERROR;
  }
  const bool vb = (window_bits & 0x100) != 0;
  window_bits &= 0xff;
  CKR(cudaSetDevice(device), "set device");
  MsmSession *s = new MsmSession();
  // any failure below releases what was allocated so far
#define CKS(x, what) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { zkb_msm_session_destroy(reinterpret_cast<zkb_msm_session *>(s)); return cuda_fail(e_, what); } } while (0)
  s->device = device;
  s->total = 1ull << logn;
  uint64_t mine = s->total / nranks;
  // fixed base: sub-MSMs of 2^17 points (16 digits per point land in one bucket set); variable base: every window has
  // its own bucket set, so a sub-MSM takes 2^20 points to fill 2^15 buckets with ~32 entries each
  const uint32_t sub_max = vb ? (1u << 20) : (1u << 17);
  s->n_sub = (uint32_t)(mine < sub_max ? mine : sub_max);
  s->subs = (uint32_t)(mine / s->n_sub);
  s->vb = vb;
  s->first = mine * rank;
  s->seed = seed;
  s->slot = rank;
  s->cfg = msm_cfg(window_bits);
  s->windows = s->cfg.windows;
  if (vb) s->cfg.windows = 1;
  s->items = vb ? s->subs * (uint32_t)s->windows : s->subs;
  CKS(cudaStreamCreateWithFlags(&s->st, cudaStreamNonBlocking), "stream");
  cudaEventCreate(&s->e0); cudaEventCreate(&s->e1); cudaEventCreate(&s->e2);
  CKS(cudaMalloc(&s->bases, mine * sizeof(Affine<Fq>)), "alloc bases");
  CKS(cudaMalloc(&s->scalars, mine * sizeof(Fr)), "alloc scalars");
  cudaEventRecord(s->e0, s->st);
  k_gen_points<<<(unsigned)((mine + 127) / 128), 128, 0, s->st>>>(s->bases, s->first, mine, seed);
  k_gen_scalars<<<(unsigned)((mine + 255) / 256), 256, 0, s->st>>>(s->scalars, s->first, mine, seed);
  cudaEventRecord(s->e1, s->st);
  if (vb) { s->tab.tab = s->bases; s->tab.n = s->n_sub; s->tab.cfg = s->cfg; }      // the bases ARE the table
  else CKS(msm_build_table<Fq>(s->tab, s->bases, s->n_sub, s->cfg, s->st, s->subs), "build table");
  cudaEventRecord(s->e2, s->st);
  CKS(s->sort.alloc(s->n_sub, s->items, s->cfg), "alloc sort");
  CKS(s->work.alloc(s->items, s->cfg), "alloc buckets");
  CKS(cudaMalloc(&s->sub_out, (size_t)s->items * sizeof(XYZZ<Fq>)), "alloc");
  CKS(cudaMalloc(&s->partial, sizeof(XYZZ<Fq>)), "alloc");
  CKS(cudaMalloc(&s->xbuf, MAX_RANKS * sizeof(ExSlot)), "alloc exchange");
  CKS(cudaMemsetAsync(s->xbuf, 0, MAX_RANKS * sizeof(ExSlot), s->st), "memset");
  CKS(cudaMalloc(&s->result, sizeof(Affine<Fq>)), "alloc");
  CKS(cudaMalloc(&s->status, 4), "alloc");
  s->root_x = s->xbuf;
  CKS(cudaStreamSynchronize(s->st), "msm session setup");
  cudaEventElapsedTime(&s->gen_ms, s->e0, s->e1);
  cudaEventElapsedTime(&s->table_ms, s->e1, s->e2);
#undef CKS
  *out = reinterpret_cast<zkb_msm_session *>(s);
  return ZKB_OK;
}

void zkb_msm_session_destroy(zkb_msm_session *h) {
  MsmSession *s = reinterpret_cast<MsmSession *>(h);
  if (!s) return;
  cudaSetDevice(s->device);
  if (s->st) cudaStreamSynchronize(s->st);
  if (s->root_is_ipc) cudaIpcCloseMemHandle(s->root_x);
  if (!s->vb) cudaFree(s->tab.tab);
  cudaFree(s->bases); cudaFree(s->scalars); cudaFree(s->sub_out); cudaFree(s->partial);
  cudaFree(s->xbuf); cudaFree(s->result); cudaFree(s->status);
  s->sort.free_all();
  s->work.free_all();
  if (s->e0) { cudaEventDestroy(s->e0); cudaEventDestroy(s->e1); cudaEventDestroy(s->e2); }
  if (s->st) cudaStreamDestroy(s->st);
  cudaGetLastError();
  delete s;
}

// info[6] = points of this rank, sub-MSM size, sub-MSMs, window bits, table build ms (x1000), data generation ms (x1000)
int zkb_msm_session_info(zkb_msm_session *h, uint64_t *info) {
  MsmSession *s = reinterpret_cast<MsmSession *>(h);
  info[0] = (uint64_t)s->n_sub * s->subs; info[1] = s->n_sub; info[2] = s->subs; info[3] = (uint64_t)s->cfg.c;
  info[4] = (uint64_t)(s->table_ms * 1000); info[5] = (uint64_t)(s->gen_ms * 1000);
  return ZKB_OK;
}

// 64-byte CUDA IPC handle of this rank's exchange buffer (call on the root, hand the bytes to the other ranks)
int zkb_msm_session_export(zkb_msm_session *h, void *handle64) {
  MsmSession *s = reinterpret_cast<MsmSession *>(h);
  CKR(cudaSetDevice(s->device), "set device");
  cudaIpcMemHandle_t mh;
  CKR(cudaIpcGetMemHandle(&mh, s->xbuf), "ipc export");
  static_assert(sizeof(mh) == 64, "CUDA IPC handle size");
  memcpy(handle64, &mh, 64);
  return ZKB_OK;
}
// non-root ranks: map the root's exchange buffer; this rank's partial sums are then written there over NVLink
int zkb_msm_session_attach(zkb_msm_session *h, const void *handle64) {
  MsmSession *s = reinterpret_cast<MsmSession *>(h);
  CKR(cudaSetDevice(s->device), "set device");
  cudaIpcMemHandle_t mh;
  memcpy(&mh, handle64, 64);
  void *p = nullptr;
  CKR(cudaIpcOpenMemHandle(&p, mh, cudaIpcMemLazyEnablePeerAccess), "ipc open (peer access to the root GPU)");
  s->root_x = reinterpret_cast<ExSlot *>(p);
  s->root_is_ipc = true;
  return ZKB_OK;
}
// same process, several contexts (tests / single-process drivers): publish into another session's buffer directly
int zkb_msm_session_attach_local(zkb_msm_session *h, zkb_msm_session *root) {
  MsmSession *s = reinterpret_cast<MsmSession *>(h), *r = reinterpret_cast<MsmSession *>(root);
  CKR(cudaSetDevice(s->device), "set device");
  if (s->device != r->device) {
    int can = 0;
    cudaDeviceCanAccessPeer(&can, s->device, r->device);
    if (!can) { set_error("no peer access between the two GPUs"); return ZKB_ERROR; }
    cudaError_t e = cudaDeviceEnablePeerAccess(r->device, 0);
    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return cuda_fail(e, "enable peer access");
    cudaGetLastError();
  }
  s->root_x = r->xbuf;
  return ZKB_OK;
}

// queue this rank's share: sort, bucket accumulation, bucket reduction, fold, publish into the root's slot.
// Asynchronous; *ms (optional) makes it wait and returns the device time of this rank's part.
int zkb_msm_session_run(zkb_msm_session *h, float *ms) {
  MsmSession *s = reinterpret_cast<MsmSession *>(h);
  CKR(cudaSetDevice(s->device), "set device");
  s->epoch++;
  cudaEventRecord(s->e0, s->st);
  if (s->vb) {
    // item j * subs + t = window j of sub-MSM t; its bases are block t (items % subs) of the rank's points
    CKR(s->sort.run_vb(s->scalars, s->n_sub, s->subs, s->windows, s->st), "sort (variable base)");
    CKR(msm_accumulate<Fq>(s->sort, &s->tab, 1, s->items, s->work, 0, s->st, (size_t)s->n_sub, s->subs), "accumulate");
    CKR(msm_reduce<Fq>(s->work, 0, s->items, s->sub_out, s->st), "reduce");
    CKR(msm_horner(s->sub_out, s->partial, 1, s->subs, s->windows, s->cfg.c, s->st), "horner");
  } else {
    CKR(s->sort.run(s->scalars, s->n_sub, s->subs, s->st), "sort");
    CKR(msm_accumulate<Fq>(s->sort, &s->tab, 1, s->subs, s->work, 0, s->st, (size_t)s->cfg.windows * s->n_sub), "accumulate");
    CKR(msm_reduce<Fq>(s->work, 0, s->subs, s->sub_out, s->st), "reduce");
    k_fold<<<1, 256, 0, s->st>>>(s->sub_out, s->subs, s->partial);
  }
  k_publish<<<1, 64, 0, s->st>>>(s->root_x + s->slot, s->partial, s->epoch);
  cudaEventRecord(s->e1, s->st);
  CKR(cudaGetLastError(), "msm session launch");
  if (ms) {
    CKR(cudaStreamSynchronize(s->st), "msm session run");
    cudaEventElapsedTime(ms, s->e0, s->e1);
  }
  return ZKB_OK;
}

// root: wait for the nslots partial sums of the current epoch, add them, return the affine canonical point (64 B).
// *ms (optional): device time from the start of this rank's run() to the end of the combine.
int zkb_msm_session_combine(zkb_msm_session *h, int nslots, void *out64, float *ms) {
  MsmSession *s = reinterpret_cast<MsmSession *>(h);
  CKR(cudaSetDevice(s->device), "set device");
  if (nslots < 1 || nslots > MAX_RANKS) { set_error("combine: bad slot count"); return ZKB_ERROR; }
  k_combine<<<1, 32, 0, s->st>>>(s->xbuf, nslots, s->epoch, s->result, s->status);
  cudaEventRecord(s->e2, s->st);
  int st = -1;
  CKR(cudaMemcpyAsync(&st, s->status, 4, cudaMemcpyDeviceToHost, s->st), "d2h");
  CKR(cudaMemcpyAsync(out64, s->result, 64, cudaMemcpyDeviceToHost, s->st), "d2h");
  CKR(cudaStreamSynchronize(s->st), "combine");
  if (st != 0) { set_error("combine: timed out waiting for the partial sum of rank " + std::to_string(st - 1)); return ZKB_ERROR; }
  if (ms) cudaEventElapsedTime(ms, s->e0, s->e2);
  return ZKB_OK;
}

// executed mixed adds of this rank's share (exact, counted on the device; measurement aid)
int zkb_msm_session_madds(zkb_msm_session *h, uint64_t *madds) {
  MsmSession *s = reinterpret_cast<MsmSession *>(h);
  CKR(cudaSetDevice(s->device), "set device");
  unsigned long long t = 0;
  // every sub-MSM has its own table block; none of the synthetic bases is infinity, so count entries - buckets used
  std::vector<uint32_t> off((size_t)s->items * (s->cfg.buckets + 1));
  CKR(cudaMemcpy(off.data(), s->sort.offsets, off.size() * 4, cudaMemcpyDeviceToHost), "d2h");
  for (uint32_t b = 0; b < s->items; b++) {
    const uint32_t *o = off.data() + (size_t)b * (s->cfg.buckets + 1);
    for (uint32_t i = 0; i < s->cfg.buckets; i++) {
      uint32_t len = o[i + 1] - o[i];
      if (len > 1) t += len - 1;
    }
  }
  *madds = t;
  return ZKB_OK;
}

// parity aid: this rank's bases (64 B affine canonical) and scalars (32 B canonical) back to the host
int zkb_msm_session_read(zkb_msm_session *h, void *bases_out, void *scalars_out) {
  MsmSession *s = reinterpret_cast<MsmSession *>(h);
  CKR(cudaSetDevice(s->device), "set device");
  size_t n = (size_t)s->n_sub * s->subs;
  Affine<Fq> *tmp = nullptr;
  CKR(cudaMalloc(&tmp, n * sizeof(Affine<Fq>)), "alloc");
  k_affine_from_mont<<<(unsigned)((n + 127) / 128), 128, 0, s->st>>>(s->bases, tmp, n);
  CKR(cudaMemcpyAsync(bases_out, tmp, n * 64, cudaMemcpyDeviceToHost, s->st), "d2h");
  CKR(cudaMemcpyAsync(scalars_out, s->scalars, n * 32, cudaMemcpyDeviceToHost, s->st), "d2h");
  CKR(cudaStreamSynchronize(s->st), "read");
  cudaFree(tmp);
  return ZKB_OK;
}

// ---- single-GPU NTT sweep point: resident data, `iters` inverse (DIF) + forward (DIT) transforms of nvec vectors ----
int zkb_ntt_bench(int device, int logn, int nvec, int iters, float *dif_ms, float *dit_ms) {
  if (require_device()) return ZKB_ERROR;
  CKR(cudaSetDevice(device), "set device");
  NttPlan plan;
  CKR(plan.init(logn, 0), "ntt plan (logn in [12, 27])");
  size_t n = (size_t)1 << logn, total = n * nvec;
  Fr *d = nullptr;
  CKR(cudaMalloc(&d, total * 32), "alloc");
  k_gen_scalars<<<(unsigned)((total + 255) / 256), 256>>>(d, 0, total, 1);
  CKR(fr_to_mont(d, total, 0), "to_mont");
  cudaEvent_t e0, e1, e2;
  cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2);
  CKR(plan.dif(d, nvec, n, true, true, 0), "dif");       // warm-up
  CKR(plan.dit(d, nvec, n, false, 0), "dit");
  cudaEventRecord(e0);
  for (int i = 0; i < iters; i++) CKR(plan.dif(d, nvec, n, true, true, 0), "dif");
  cudaEventRecord(e1);
  for (int i = 0; i < iters; i++) CKR(plan.dit(d, nvec, n, false, 0), "dit");
  cudaEventRecord(e2);
  CKR(cudaEventSynchronize(e2), "ntt bench");
  float a = 0, b = 0;
  cudaEventElapsedTime(&a, e0, e1);
  cudaEventElapsedTime(&b, e1, e2);
  *dif_ms = a / iters;
  *dit_ms = b / iters;
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2);
  cudaFree(d);
  plan.destroy();
  return ZKB_OK;
}


// ---- 4-step NTT over `nranks` GPUs (one session per rank; see NttDist above) -----------------------------------
struct zkb_ntt_dist;
int zkb_ntt_dist_create(int device, int logn, int rank, int nranks, uint64_t seed, zkb_ntt_dist **out) {
  if (require_device()) return ZKB_ERROR;
  if (logn < 22 || logn > 30 || nranks < 1 || nranks > MAX_RANKS || (nranks & (nranks - 1)) || rank < 0 || rank >= nranks) {
    set_error("ntt dist: logn in [22,30], nranks a power of two <= 16");
    return ZKB_ERROR;
  }
  CKR(cudaSetDevice(device), "set device");
  NttDist *s = new NttDist();
  s->device = device; s->rank = rank; s->nranks = nranks; s->logn = logn; s->seed = seed;
  s->l2 = logn / 2; s->l1 = logn - s->l2;
  const size_t per = ((size_t)1 << logn) / nranks;
  CKR(cudaStreamCreateWithFlags(&s->st, cudaStreamNonBlocking), "stream");
  cudaEventCreate(&s->e0); cudaEventCreate(&s->e1); cudaEventCreate(&s->e2); cudaEventCreate(&s->e3);
  CKR(cudaMalloc(&s->alloc, sizeof(NttCtl) + per * 32), "alloc B");
  CKR(cudaMemsetAsync(s->alloc, 0, sizeof(NttCtl), s->st), "memset");
  s->B = reinterpret_cast<Fr *>(s->alloc + sizeof(NttCtl));
  CKR(cudaMalloc(&s->C, per * 32), "alloc C");
  CKR(cudaMalloc(&s->status, 4), "alloc");
  CKR(cudaMemsetAsync(s->status, 0, 4, s->st), "memset");
  CKR(s->plan1.init(s->l1, s->st), "ntt plan 1");
  CKR(s->plan2.init(s->l2, s->st), "ntt plan 2");
  Fr *pw = nullptr;
  CKR(cudaMalloc(&pw, 32 * 32), "alloc");
  CKR(cudaMalloc(&s->tw_lo, 8192 * 32), "alloc");
  const uint32_t hi = (uint32_t)(((size_t)1 << logn) >> 13);
  CKR(cudaMalloc(&s->tw_hi, (size_t)hi * 32), "alloc");
  k_root_pw<<<1, 1, 0, s->st>>>(pw, logn);
  k_pow_table<<<8192 / 256, 256, 0, s->st>>>(s->tw_lo, pw, 8192, 0);
  k_pow_table<<<(hi + 255) / 256, 256, 0, s->st>>>(s->tw_hi, pw, hi, 13);
  CKR(cudaStreamSynchronize(s->st), "ntt dist setup");
  cudaFree(pw);
  for (int g = 0; g < MAX_RANKS; g++) { s->peers.base[g] = nullptr; s->ipc[g] = false; }
  s->peers.base[rank] = s->alloc;
  *out = reinterpret_cast<zkb_ntt_dist *>(s);
  return ZKB_OK;
}
void zkb_ntt_dist_destroy(zkb_ntt_dist *h) {
  NttDist *s = reinterpret_cast<NttDist *>(h);
  if (!s) return;
  cudaSetDevice(s->device);
  cudaStreamSynchronize(s->st);
  for (int g = 0; g < MAX_RANKS; g++)
    if (s->ipc[g]) cudaIpcCloseMemHandle(const_cast<uint8_t *>(s->peers.base[g]));
  cudaFree(s->alloc); cudaFree(s->C); cudaFree(s->status); cudaFree(s->tw_lo); cudaFree(s->tw_hi);
  s->plan1.destroy(); s->plan2.destroy();
  cudaEventDestroy(s->e0); cudaEventDestroy(s->e1); cudaEventDestroy(s->e2); cudaEventDestroy(s->e3);
  cudaStreamDestroy(s->st);
  delete s;
}
int zkb_ntt_dist_export(zkb_ntt_dist *h, void *handle64) {
  NttDist *s = reinterpret_cast<NttDist *>(h);
  CKR(cudaSetDevice(s->device), "set device");
  cudaIpcMemHandle_t mh;
  CKR(cudaIpcGetMemHandle(&mh, s->alloc), "ipc export");
  memcpy(handle64, &mh, 64);
  return ZKB_OK;
}
int zkb_ntt_dist_attach(zkb_ntt_dist *h, int peer_rank, const void *handle64) {
  NttDist *s = reinterpret_cast<NttDist *>(h);
  CKR(cudaSetDevice(s->device), "set device");
  if (peer_rank < 0 || peer_rank >= s->nranks || peer_rank == s->rank) { set_error("ntt dist: bad peer rank"); return ZKB_ERROR; }
  cudaIpcMemHandle_t mh;
  memcpy(&mh, handle64, 64);
  void *p = nullptr;
  CKR(cudaIpcOpenMemHandle(&p, mh, cudaIpcMemLazyEnablePeerAccess), "ipc open (peer access)");
  s->peers.base[peer_rank] = reinterpret_cast<const uint8_t *>(p);
  s->ipc[peer_rank] = true;
  return ZKB_OK;
}
int zkb_ntt_dist_attach_local(zkb_ntt_dist *h, int peer_rank, zkb_ntt_dist *peer) {
  NttDist *s = reinterpret_cast<NttDist *>(h), *o = reinterpret_cast<NttDist *>(peer);
  CKR(cudaSetDevice(s->device), "set device");
  if (s->device != o->device) {
    int can = 0;
    cudaDeviceCanAccessPeer(&can, s->device, o->device);
    if (!can) { set_error("no peer access between the two GPUs"); return ZKB_ERROR; }
    cudaError_t e = cudaDeviceEnablePeerAccess(o->device, 0);
    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return cuda_fail(e, "enable peer access");
    cudaGetLastError();
  }
  s->peers.base[peer_rank] = o->alloc;
  return ZKB_OK;
}
// (re)generate this rank's columns of the synthetic input
int zkb_ntt_dist_fill(zkb_ntt_dist *h) {
  NttDist *s = reinterpret_cast<NttDist *>(h);
  CKR(cudaSetDevice(s->device), "set device");
  const uint32_t cols = (1u << s->l2) / s->nranks;
  const size_t per = (size_t)cols << s->l1;
  k_fill_columns<<<(unsigned)((per + 255) / 256), 256, 0, s->st>>>(s->B, s->l1, s->l2, s->rank * cols, cols, s->seed);
  CKR(cudaStreamSynchronize(s->st), "fill");
  return ZKB_OK;
}
// queue one transform (asynchronous).  All ranks must have called fill() (and be past a host barrier) before.
int zkb_ntt_dist_run(zkb_ntt_dist *h) {
  NttDist *s = reinterpret_cast<NttDist *>(h);
  CKR(cudaSetDevice(s->device), "set device");
  for (int g = 0; g < s->nranks; g++)
    if (!s->peers.base[g]) { set_error("ntt dist: peer " + std::to_string(g) + " not attached"); return ZKB_ERROR; }
  const uint32_t N1 = 1u << s->l1, N2 = 1u << s->l2, cols = N2 / s->nranks, rows = N1 / s->nranks;
  s->epoch++;
  cudaEventRecord(s->e0, s->st);
  CKR(s->plan1.dif(s->B, (int)cols, N1, false, false, s->st), "column transforms");
  k_set_flag<<<1, 1, 0, s->st>>>(&reinterpret_cast<NttCtl *>(s->alloc)->ready, s->epoch);
  cudaEventRecord(s->e1, s->st);
  k_wait_flags<<<1, 32, 0, s->st>>>(s->peers, s->nranks, s->epoch, s->status);
  k_gather_transpose<<<dim3(N2 / 32, rows / 32), dim3(32, 8), 0, s->st>>>(s->peers, s->C, s->l1, s->l2, cols,
                                                                             s->rank * rows, s->tw_lo, s->tw_hi, 1, s->status);
  cudaEventRecord(s->e2, s->st);
  CKR(s->plan2.dif(s->C, (int)rows, N2, false, false, s->st), "row transforms");
  cudaEventRecord(s->e3, s->st);
  CKR(cudaGetLastError(), "ntt dist launch");
  return ZKB_OK;
}
// wait for the queued transform; ms[4] = total, column transforms, wait + exchange (P2P gather), row transforms
int zkb_ntt_dist_sync(zkb_ntt_dist *h, float *ms) {
  NttDist *s = reinterpret_cast<NttDist *>(h);
  CKR(cudaSetDevice(s->device), "set device");
  int st = -1;
  CKR(cudaMemcpyAsync(&st, s->status, 4, cudaMemcpyDeviceToHost, s->st), "d2h");
  CKR(cudaStreamSynchronize(s->st), "ntt dist run");
  if (st != 0) { set_error("ntt dist: timed out waiting for rank " + std::to_string(st - 1)); return ZKB_ERROR; }
  if (ms) {
    cudaEventElapsedTime(ms + 0, s->e0, s->e3);
    cudaEventElapsedTime(ms + 1, s->e0, s->e1);
    cudaEventElapsedTime(ms + 2, s->e1, s->e2);
    cudaEventElapsedTime(ms + 3, s->e2, s->e3);
  }
  return ZKB_OK;
}
// which = 0: this rank's input columns B_g[c][n1] (only valid right after fill()); 1: its output rows
// C_g[p_local][q] = X[bitrev_N1(rank * N1/G + p_local) + N1 * bitrev_N2(q)].  Canonical form, N/G x 32 bytes.
int zkb_ntt_dist_read(zkb_ntt_dist *h, int which, void *out) {
  NttDist *s = reinterpret_cast<NttDist *>(h);
  CKR(cudaSetDevice(s->device), "set device");
  const size_t per = ((size_t)1 << s->logn) / s->nranks;
  Fr *tmp = nullptr;
  CKR(cudaMalloc(&tmp, per * 32), "alloc");
  CKR(cudaMemcpyAsync(tmp, which ? s->C : s->B, per * 32, cudaMemcpyDeviceToDevice, s->st), "copy");
  CKR(fr_from_mont(tmp, per, s->st), "from_mont");
  CKR(cudaMemcpyAsync(out, tmp, per * 32, cudaMemcpyDeviceToHost, s->st), "d2h");
  CKR(cudaStreamSynchronize(s->st), "read");
  cudaFree(tmp);
  return ZKB_OK;
}
// parity aid: forward transform of one resident vector by the single-GPU plan (natural in, BIT-REVERSED out):
// data = N x 32 B canonical on the host, in place
int zkb_raw_ntt_dif_forward(void *data, int logn) {
  if (require_device()) return ZKB_ERROR;
  NttPlan plan;
  CKR(plan.init(logn, 0), "ntt plan");
  size_t n = (size_t)1 << logn;
  Fr *d = nullptr;
  CKR(cudaMalloc(&d, n * 32), "alloc");
  CKR(cudaMemcpy(d, data, n * 32, cudaMemcpyHostToDevice), "h2d");
  CKR(fr_to_mont(d, n, 0), "to_mont");
  CKR(plan.dif(d, 1, n, false, false, 0), "dif");
  CKR(fr_from_mont(d, n, 0), "from_mont");
  CKR(cudaMemcpy(data, d, n * 32, cudaMemcpyDeviceToHost), "d2h");
  cudaFree(d);
  plan.destroy();
  return ZKB_OK;
}

}  // extern "C"
