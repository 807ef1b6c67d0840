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

}  // namespace zkb

using namespace zkb;

extern "C" {

struct zkb_msm_session;

// One rank's share of a 2^logn-point synthetic G1 MSM: points [rank * N / nranks, (rank + 1) * N / nranks), cut into
// sub-MSMs of at most 2^17 points.  Generates bases + scalars on the device and builds the fixed-base window table.
int zkb_msm_session_create(int device, int logn, int rank, int nranks, uint64_t seed, int window_bits,
                           zkb_msm_session **out) {
  if (require_device()) return ZKB_ERROR;
  if (logn < 10 || logn > 27 || nranks < 1 || nranks > MAX_RANKS || rank < 0 || rank >= nranks ||
      (nranks & (nranks - 1)) || window_bits < 12 || window_bits > 16) {
    set_error("msm session: logn in [10,27], nranks a power of two <= 16, window in [12,16]");
    return ZKB_ERROR;
  }
  CKR(cudaSetDevice(device), "set device");
  MsmSession *s = new MsmSession();
  s->device = device;
  s->total = 1ull << logn;
  uint64_t mine = s->total / nranks;
  s->n_sub = (uint32_t)(mine < (1u << 17) ? mine : (1u << 17));
  s->subs = (uint32_t)(mine / s->n_sub);
  s->first = mine * rank;
  s->seed = seed;
  s->slot = rank;
  s->cfg = msm_cfg(window_bits);
  CKR(cudaStreamCreateWithFlags(&s->st, cudaStreamNonBlocking), "stream");
  cudaEventCreate(&s->e0); cudaEventCreate(&s->e1); cudaEventCreate(&s->e2);
  CKR(cudaMalloc(&s->bases, mine * sizeof(Affine<Fq>)), "alloc bases");
  CKR(cudaMalloc(&s->scalars, mine * sizeof(Fr)), "alloc scalars");
  cudaEventRecord(s->e0, s->st);
  k_gen_points<<<(unsigned)((mine + 127) / 128), 128, 0, s->st>>>(s->bases, s->first, mine, seed);
  k_gen_scalars<<<(unsigned)((mine + 255) / 256), 256, 0, s->st>>>(s->scalars, s->first, mine, seed);
  cudaEventRecord(s->e1, s->st);
  CKR(msm_build_table<Fq>(s->tab, s->bases, s->n_sub, s->cfg, s->st, s->subs), "build table");
  cudaEventRecord(s->e2, s->st);
  CKR(s->sort.alloc(s->n_sub, s->subs, s->cfg), "alloc sort");
  CKR(s->work.alloc(s->subs, s->cfg), "alloc buckets");
  CKR(cudaMalloc(&s->sub_out, (size_t)s->subs * sizeof(XYZZ<Fq>)), "alloc");
  CKR(cudaMalloc(&s->partial, sizeof(XYZZ<Fq>)), "alloc");
  CKR(cudaMalloc(&s->xbuf, MAX_RANKS * sizeof(ExSlot)), "alloc exchange");
  CKR(cudaMemsetAsync(s->xbuf, 0, MAX_RANKS * sizeof(ExSlot), s->st), "memset");
  CKR(cudaMalloc(&s->result, sizeof(Affine<Fq>)), "alloc");
  CKR(cudaMalloc(&s->status, 4), "alloc");
  s->root_x = s->xbuf;
  CKR(cudaStreamSynchronize(s->st), "msm session setup");
  cudaEventElapsedTime(&s->gen_ms, s->e0, s->e1);
  cudaEventElapsedTime(&s->table_ms, s->e1, s->e2);
  *out = reinterpret_cast<zkb_msm_session *>(s);
  return ZKB_OK;
}

void zkb_msm_session_destroy(zkb_msm_session *h) {
  MsmSession *s = reinterpret_cast<MsmSession *>(h);
  if (!s) return;
  cudaSetDevice(s->device);
  cudaStreamSynchronize(s->st);
  if (s->root_is_ipc) cudaIpcCloseMemHandle(s->root_x);
  cudaFree(s->bases); cudaFree(s->scalars); cudaFree(s->tab.tab); cudaFree(s->sub_out); cudaFree(s->partial);
  cudaFree(s->xbuf); cudaFree(s->result); cudaFree(s->status);
  s->sort.free_all();
  s->work.free_all();
  cudaEventDestroy(s->e0); cudaEventDestroy(s->e1); cudaEventDestroy(s->e2);
  cudaStreamDestroy(s->st);
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
  CKR(s->sort.run(s->scalars, s->n_sub, s->subs, s->st), "sort");
  CKR(msm_accumulate<Fq>(s->sort, &s->tab, 1, s->subs, s->work, 0, s->st, (size_t)s->cfg.windows * s->n_sub), "accumulate");
  CKR(msm_reduce<Fq>(s->work, 0, s->subs, s->sub_out, s->st), "reduce");
  k_fold<<<1, 256, 0, s->st>>>(s->sub_out, s->subs, s->partial);
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
  std::vector<uint32_t> off((size_t)s->subs * (s->cfg.buckets + 1));
  CKR(cudaMemcpy(off.data(), s->sort.offsets, off.size() * 4, cudaMemcpyDeviceToHost), "d2h");
  for (uint32_t b = 0; b < s->subs; b++) {
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

}  // extern "C"
