// Raw kernel entry points of the C ABI (include/zkcensus_b200.h, "raw" section): field ops, Fr NTT
// and G1/G2 MSM on caller-provided host buffers, plus the integer-pipe micro-benchmark.  These are
// BASELINE.json config 5 ("raw BN254 G1 MSM and Fr NTT sweep") and what the kernel parity tests call.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <string>
#include "common.h"
#include "ntt.cuh"
#include "msm.cuh"

namespace zkb {

thread_local std::string g_last_error;
void set_error(const std::string &s) { g_last_error = s; }

int cuda_fail(cudaError_t e, const char *what) {
  set_error(std::string(what) + ": " + cudaGetErrorString(e));
  return ZKB_ERROR;
}

int require_device() {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    set_error("no CUDA device available: this library has no CPU fallback");
    return ZKB_ERROR;
  }
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  if (p.major < 10) {
    set_error("device is not sm_100 class (built for sm_100a only)");
    return ZKB_ERROR;
  }
  return ZKB_OK;
}

// ---- field ops ------------------------------------------------------------------------------
template <class F>
__global__ void k_field_op(int op, const F *a, const F *b, F *out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  F x = a[i], y = b[i], r;
  switch (op) {
    case 0: r = x * y; break;
    case 1: r = x + y; break;
    case 2: r = x - y; break;
    case 3: r = x.inv(); break;
    case 4: r = x.sqr(); break;
    case 5: r = x.to_mont(); break;
    case 6: r = x.from_mont(); break;
    default: r = x.neg(); break;
  }
  out[i] = r;
}

// dependent modmul chains: each thread runs `iters` multiplications on 4 independent chains
template <class F>
__global__ void __launch_bounds__(256) k_bench_modmul(F *out, int iters) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  F a = F::one(), b = F::r2(), c = F::r2(), d = F::one();
  a.v[0] ^= t; b.v[1] ^= t; c.v[2] ^= t; d.v[3] ^= t;
  a.v[7] &= 0x0fffffffu; b.v[7] &= 0x0fffffffu; c.v[7] &= 0x0fffffffu; d.v[7] &= 0x0fffffffu;
  for (int i = 0; i < iters; i++) {
    a = a * b;
    b = b * c;
    c = c * d;
    d = d * a;
  }
  if (a.v[0] == 0x12345678u && d.v[5] == 7u) out[t] = a + b + c + d;   // practically never; keeps the work live
}

template <class F>
__global__ void k_xyzz_to_affine_raw(const XYZZ<F> *in, Affine<F> *out, uint32_t n) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Affine<F> a;
  xyzz_to_affine_ni(in + i, &a);
  out[i] = a;
}

__global__ void k_fq_from_mont(Fq *x, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] = x[i].from_mont();
}
__global__ void k_fq_to_mont(Fq *x, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] = x[i].to_mont();
}
__global__ void k_fr_scale(Fr *x, const Fr *k, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] = x[i] * *k;
}

struct DevBuf {
  void *p = nullptr;
  ~DevBuf() { if (p) cudaFree(p); }
  cudaError_t alloc(size_t n) { return cudaMalloc(&p, n ? n : 16); }
  template <class T> T *as() { return reinterpret_cast<T *>(p); }
};

}  // namespace zkb

using namespace zkb;

#define CKR(x, what) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return cuda_fail(e_, what); } while (0)

template <class F>
static int raw_msm(const void *bases, size_t n, const void *scalars, int nbatch, void *out, float *kernel_ms,
                   float *table_ms, uint32_t flags = 0) {
  if (require_device()) return ZKB_ERROR;
  const char *ce = getenv("ZKB_RAW_MSM_C");
  int cw = ce ? atoi(ce) : 16;
  if (cw < 12 || cw > 16) { set_error("ZKB_RAW_MSM_C must be in [12, 16]"); return ZKB_ERROR; }
  const MsmCfg cfg = msm_cfg(cw);
  const size_t psz = sizeof(Affine<F>);
  DevBuf db, ds, dout, daff;
  CKR(db.alloc(n * psz), "alloc");
  CKR(ds.alloc(n * 32 * nbatch), "alloc");
  CKR(dout.alloc(sizeof(XYZZ<F>) * nbatch), "alloc");
  CKR(daff.alloc(psz * nbatch), "alloc");
  CKR(cudaMemcpy(db.p, bases, n * psz, cudaMemcpyHostToDevice), "h2d");
  CKR(cudaMemcpy(ds.p, scalars, n * 32 * nbatch, cudaMemcpyHostToDevice), "h2d");
  size_t nfq = n * psz / 32;
  k_fq_to_mont<<<(unsigned)((nfq + 255) / 256), 256>>>(db.as<Fq>(), nfq);
  cudaEvent_t e0, e1, e2;
  cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2);
  MsmTable<F> tab;
  cudaEventRecord(e0);
  CKR(msm_build_table<F>(tab, db.as<Affine<F>>(), (uint32_t)n, cfg, 0), "table");
  cudaEventRecord(e1);
  MsmSort sort;
  MsmWork<F> work;
  CKR(sort.alloc((uint32_t)n, nbatch, cfg), "sort alloc");
  CKR(work.alloc(nbatch, cfg), "work alloc");
  CKR(cudaEventSynchronize(e1), "table build");
  int reps = kernel_ms ? 2 : 1;
  for (int r = 0; r < reps; r++) {
    cudaEventRecord(e1);
    CKR(sort.run(ds.as<Fr>(), n, nbatch, 0), "sort");
    if constexpr (sizeof(F) == 32) {
      if (flags & 2u) {                      // variable-base: no window table, one bucket set per window, Horner at the end
        MsmCfg c1 = cfg;
        c1.windows = 1;
        MsmSort vs;
        MsmWork<F> vw;
        MsmTable<F> bt;
        bt.tab = db.as<Affine<F>>(); bt.n = (uint32_t)n; bt.cfg = c1;
        DevBuf wsum;
        const uint32_t items = (uint32_t)nbatch * (uint32_t)cfg.windows;
        CKR(vs.alloc((uint32_t)n, items, c1), "sort alloc");
        CKR(vw.alloc(items, c1), "work alloc");
        CKR(wsum.alloc(sizeof(XYZZ<F>) * items), "alloc");
        CKR(vs.run_vb(ds.as<Fr>(), n, nbatch, cfg.windows, 0), "sort (variable base)");
        CKR(msm_accumulate<F>(vs, &bt, 1, items, vw, 0, 0, 0, items), "accumulate (variable base)");
        CKR(msm_reduce<F>(vw, 0, items, wsum.as<XYZZ<F>>(), 0), "reduce (variable base)");
        CKR(msm_horner(wsum.as<XYZZ<F>>(), dout.as<XYZZ<F>>(), nbatch, 1, cfg.windows, cfg.c, 0), "horner");
        cudaEventRecord(e2);
        CKR(cudaEventSynchronize(e2), "msm run");
        vs.free_all();
        vw.free_all();
        continue;
      }
      if (flags & 1u) {                      // bucket lists through the batched-affine pair tree
        MsmAffineWs ws;
        CKR(ws.alloc((uint32_t)n * (uint32_t)cfg.windows, nbatch > 2 ? 2 : nbatch, cfg), "pair tree alloc");   // 3 items: two sub-batches
        if ((flags >> 8) & 15u) ws.levels = (int)((flags >> 8) & 15u) > MsmAffineWs::MAX_LEVELS ? MsmAffineWs::MAX_LEVELS : (int)((flags >> 8) & 15u);
        if (flags >> 16) for (int l = 0; l < MsmAffineWs::MAX_LEVELS; l++) ws.group[l] = flags >> 16;
        CKR(msm_accumulate_affine(sort, tab, nbatch, work, 0, ws, 0), "msm (pair tree)");
        CKR(msm_reduce<F>(work, 0, nbatch, dout.as<XYZZ<F>>(), 0), "msm reduce");
        CKR(cudaDeviceSynchronize(), "msm (pair tree)");
        ws.free_all();
        cudaEventRecord(e2);
        CKR(cudaEventSynchronize(e2), "msm run");
        continue;
      }
    }
    CKR(msm_run<F>(sort, &tab, 1, nbatch, work, dout.as<XYZZ<F>>(), 0), "msm");
    cudaEventRecord(e2);
    CKR(cudaEventSynchronize(e2), "msm run");
  }
  k_xyzz_to_affine_raw<F><<<(nbatch + 31) / 32, 32>>>(dout.as<XYZZ<F>>(), daff.as<Affine<F>>(), nbatch);
  size_t ofq = (size_t)nbatch * psz / 32;
  k_fq_from_mont<<<(unsigned)((ofq + 255) / 256), 256>>>(daff.as<Fq>(), ofq);
  CKR(cudaGetLastError(), "launch");
  CKR(cudaMemcpy(out, daff.p, psz * nbatch, cudaMemcpyDeviceToHost), "d2h");
  if (table_ms) cudaEventElapsedTime(table_ms, e0, e1);
  if (kernel_ms) cudaEventElapsedTime(kernel_ms, e1, e2);
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2);
  sort.free_all();
  work.free_all();
  cudaFree(tab.tab);
  return ZKB_OK;
}

extern "C" {

const char *zkb_last_error(void) { return g_last_error.c_str(); }

int zkb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

// field: 0 = Fq, 1 = Fr.  op: 0 mul 1 add 2 sub 3 inv 4 sqr 5 to_mont 6 from_mont 7 neg.
// a, b, out: n elements of 8 x u32 (operands are taken as Montgomery residues as-is).
int zkb_raw_field_op(int field, int op, const void *a, const void *b, void *out, size_t n) {
  if (require_device()) return ZKB_ERROR;
  DevBuf da, db, dout;
  CKR(da.alloc(n * 32), "alloc"); CKR(db.alloc(n * 32), "alloc"); CKR(dout.alloc(n * 32), "alloc");
  CKR(cudaMemcpy(da.p, a, n * 32, cudaMemcpyHostToDevice), "h2d");
  CKR(cudaMemcpy(db.p, b, n * 32, cudaMemcpyHostToDevice), "h2d");
  unsigned g = (unsigned)((n + 127) / 128);
  if (field == 0) k_field_op<Fq><<<g, 128>>>(op, da.as<Fq>(), db.as<Fq>(), dout.as<Fq>(), n);
  else k_field_op<Fr><<<g, 128>>>(op, da.as<Fr>(), db.as<Fr>(), dout.as<Fr>(), n);
  CKR(cudaGetLastError(), "launch");
  CKR(cudaMemcpy(out, dout.p, n * 32, cudaMemcpyDeviceToHost), "d2h");
  return ZKB_OK;
}

// Integer-pipe micro-benchmark: returns modmul/s through *modmul_per_s (Fq when field = 0).
int zkb_bench_modmul(int field, int iters, int blocks_per_sm, double *modmul_per_s, double *ms) {
  if (require_device()) return ZKB_ERROR;
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int blocks = p.multiProcessorCount * blocks_per_sm, threads = 256;
  DevBuf out;
  CKR(out.alloc((size_t)blocks * threads * 32), "alloc");
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int rep = 0; rep < 2; rep++) {
    cudaEventRecord(e0);
    if (field == 0) k_bench_modmul<Fq><<<blocks, threads>>>(out.as<Fq>(), iters);
    else k_bench_modmul<Fr><<<blocks, threads>>>(out.as<Fr>(), iters);
    cudaEventRecord(e1);
    CKR(cudaEventSynchronize(e1), "bench");
  }
  float t = 0;
  cudaEventElapsedTime(&t, e0, e1);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  *ms = t;
  *modmul_per_s = 4.0 * iters * (double)blocks * threads / (t * 1e-3);
  return ZKB_OK;
}

// In-place NTT of `nvec` vectors of 2^logn canonical (non-Montgomery) Fr values, natural order in
// and out.  inverse != 0: inverse transform including the 1/N factor.  *kernel_ms (optional)
// receives the device time of the transform kernels only (data already resident).
int zkb_raw_ntt(void *data, int logn, int nvec, int inverse, float *kernel_ms) {
  if (require_device()) return ZKB_ERROR;
  size_t n = (size_t)1 << logn, total = n * nvec;
  NttPlan plan;
  CKR(plan.init(logn, 0), "ntt plan");
  DevBuf d, tmp, ninv;
  CKR(d.alloc(total * 32), "alloc"); CKR(tmp.alloc(total * 32), "alloc"); CKR(ninv.alloc(32), "alloc");
  CKR(cudaMemcpy(d.p, data, total * 32, cudaMemcpyHostToDevice), "h2d");
  CKR(fr_to_mont(d.as<Fr>(), total, 0), "to_mont");
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  Fr *res;
  cudaEventRecord(e0);
  if (!inverse) {
    CKR(ntt_bitrev(tmp.as<Fr>(), d.as<Fr>(), logn, nvec, n, 0), "bitrev");
    CKR(plan.dit(tmp.as<Fr>(), nvec, n, false, 0), "dit");
    res = tmp.as<Fr>();
  } else {
    CKR(plan.dif(d.as<Fr>(), nvec, n, true, false, 0), "dif");
    CKR(ntt_bitrev(tmp.as<Fr>(), d.as<Fr>(), logn, nvec, n, 0), "bitrev");
    // 1/N = coset_scale[0] (inc^0 * N^-1)
    k_fr_scale<<<(unsigned)((total + 255) / 256), 256>>>(tmp.as<Fr>(), plan.coset_scale, total);
    res = tmp.as<Fr>();
  }
  cudaEventRecord(e1);
  CKR(fr_from_mont(res, total, 0), "from_mont");
  CKR(cudaMemcpy(data, res, total * 32, cudaMemcpyDeviceToHost), "d2h");
  if (kernel_ms) cudaEventElapsedTime(kernel_ms, e0, e1);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  plan.destroy();
  return ZKB_OK;
}

// The prover's transform chain on raw vectors (for parity tests of the fused path):
// out = NTT( coset_shift( iNTT(in) ) ), natural order in and out, canonical values.
int zkb_raw_coset_ntt(void *data, int logn, int nvec) {
  if (require_device()) return ZKB_ERROR;
  size_t n = (size_t)1 << logn, total = n * nvec;
  NttPlan plan;
  CKR(plan.init(logn, 0), "ntt plan");
  DevBuf d;
  CKR(d.alloc(total * 32), "alloc");
  CKR(cudaMemcpy(d.p, data, total * 32, cudaMemcpyHostToDevice), "h2d");
  CKR(fr_to_mont(d.as<Fr>(), total, 0), "to_mont");
  CKR(plan.dif(d.as<Fr>(), nvec, n, true, true, 0), "dif");
  CKR(plan.dit(d.as<Fr>(), nvec, n, false, 0), "dit");
  CKR(fr_from_mont(d.as<Fr>(), total, 0), "from_mont");
  CKR(cudaMemcpy(data, d.p, total * 32, cudaMemcpyDeviceToHost), "d2h");
  plan.destroy();
  return ZKB_OK;
}

// bases: n affine points, canonical coordinates (x, y), 64 B each, all-zero = infinity.
// scalars: nbatch vectors of n canonical 32-byte scalars (same bases, independent scalar sets).
// out: nbatch affine canonical points.
int zkb_raw_msm_g1(const void *bases, size_t n, const void *scalars, int nbatch, void *out, float *kernel_ms,
                   float *table_ms) {
  return raw_msm<Fq>(bases, n, scalars, nbatch, out, kernel_ms, table_ms);
}
// flags bit 1: variable-base MSM (no window table: one bucket set per window, window sums combined by Horner's rule)
// flags bit 0: sum the bucket lists with the batched-affine pair tree (the H MSM's batch path); bits 8-11: tree levels
// (0 = default 3); bits 16-31: additions per field inversion (0 = default 512)
int zkb_raw_msm_g1_ex(const void *bases, size_t n, const void *scalars, int nbatch, void *out, float *kernel_ms,
                      float *table_ms, uint32_t flags) {
  return raw_msm<Fq>(bases, n, scalars, nbatch, out, kernel_ms, table_ms, flags);
}
// G2: points are (x.c0, x.c1, y.c0, y.c1), 128 B each.
int zkb_raw_msm_g2(const void *bases, size_t n, const void *scalars, int nbatch, void *out, float *kernel_ms,
                   float *table_ms) {
  return raw_msm<Fq2>(bases, n, scalars, nbatch, out, kernel_ms, table_ms);
}

}  // extern "C"
