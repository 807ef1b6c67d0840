// Fr NTT kernels for the QAP quotient: radix-2 stages grouped into passes of up to
// 9 stages; each pass stages a 2048-element tile (64 KiB, limb-planar) in shared memory, runs its
// butterflies there and makes exactly one coalesced HBM read and one HBM write of the vector.
//
//   inverse transform = decimation-in-frequency passes (natural in -> bit-reversed out)
//   forward transform = decimation-in-time passes     (bit-reversed in -> natural out)
// so the prover's iNTT -> coset shift -> NTT chain needs no permutation pass: the shift table is
// indexed by the bit-reversed position and the H-MSM consumes the natural-order result.
//
// Roots are snarkjs' / ptau's: omega_{2^28} = 5^((r-1)/2^28), omega_{2^k} by repeated squaring.
#include "ntt.cuh"
#include <cstdio>
#include <cstdlib>

namespace zkb {

static constexpr int TILE_LOG = 11;
static constexpr int TILE = 1 << TILE_LOG;      // elements per CTA tile

// omega_{2^28} in Montgomery form
__device__ __constant__ uint32_t OMEGA28[8] = {0x80d13d9cu, 0x636e7355u, 0x2445ffd6u, 0xa22bf374u,
                                              0x1eb203d8u, 0x56452ac0u, 0x2963f9e7u, 0x1860ef94u};

// pw[k] = base^(2^k), k < 32 ; out[e] = base^e for e < count (square-and-multiply over pw)
__global__ void k_powers(Fr *out, const Fr *pw, size_t count, Fr lead) {
  size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= count) return;
  Fr acc = lead;
  size_t x = e;
  for (int k = 0; x; k++, x >>= 1)
    if (x & 1) acc = acc * pw[k];
  out[e] = acc;
}

// pw[k] = g^(2^k) where g = omega_{2^logn} (inverse ? g^-1 : g)
__global__ void k_root_powers(Fr *pw, int logn, int inverse, Fr *ninv_out) {
  if (threadIdx.x || blockIdx.x) return;
  Fr g;
  for (int i = 0; i < 8; i++) g.v[i] = OMEGA28[i];
  for (int i = 28; i > logn; i--) g = g.sqr();
  if (inverse) g = g.inv();
  for (int k = 0; k < 32; k++) {
    pw[k] = g;
    g = g.sqr();
  }
  if (ninv_out) {
    // 2^-logn
    Fr two = Fr::one() + Fr::one(), t = Fr::one();
    for (int i = 0; i < logn; i++) t = t * two;
    *ninv_out = t.inv();
  }
}

// scale[p] = ninv * inc^(bitrev_logn(p))   (coset shift applied to a bit-reversed coefficient vector)
__global__ void k_coset_scale(Fr *out, const Fr *pw_inc, const Fr *ninv, int logn) {
  size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >> logn) return;
  uint32_t e = __brev((uint32_t)p) >> (32 - logn);
  Fr acc = *ninv;
  for (int k = 0; e; k++, e >>= 1)
    if (e & 1) acc = acc * pw_inc[k];
  out[p] = acc;
}

__device__ __forceinline__ Fr lds_fr(const uint32_t *sm, int idx) {
  Fr r;
#pragma unroll
  for (int l = 0; l < 8; l++) r.v[l] = sm[l * TILE + idx];
  return r;
}
__device__ __forceinline__ void sts_fr(uint32_t *sm, int idx, const Fr &x) {
#pragma unroll
  for (int l = 0; l < 8; l++) sm[l * TILE + idx] = x.v[l];
}
__device__ __forceinline__ Fr ldg_fr(const Fr *p) {
  const uint4 *q = reinterpret_cast<const uint4 *>(p);
  uint4 a = q[0], b = q[1];
  Fr r;
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
__device__ __forceinline__ void stg_fr(Fr *p, const Fr &x) {
  uint4 *q = reinterpret_cast<uint4 *>(p);
  q[0] = make_uint4(x.v[0], x.v[1], x.v[2], x.v[3]);
  q[1] = make_uint4(x.v[4], x.v[5], x.v[6], x.v[7]);
}

// One pass over stage bits [lo, hi) of a 2^logn transform on `nvec` vectors laid `vec_stride`
// elements apart.  DIF=true: stages hi-1 .. lo with (u,v) -> (u+v, (u-v)w).  DIF=false (DIT):
// stages lo .. hi-1 with (u,v) -> (u+vw, u-vw).  tw[e] = omega_N^e (or its inverse), e < N/2.
// `scale` (optional) multiplies the element stored at position p by scale[p].
template <bool DIF, int UNR, int THREADS>
__global__ void __launch_bounds__(THREADS, UNR ? 3 : 1) k_ntt_pass(Fr *data, size_t vec_stride, int logn, int lo, int hi,
                                                           const Fr *__restrict__ tw, const Fr *__restrict__ scale) {
  extern __shared__ uint32_t sm[];
  const int rows_log = hi - lo, rows = 1 << rows_log;
  const int g_log = TILE_LOG - rows_log, G = 1 << g_log;   // groups per tile (columns or chunks); all index math below
  const int rows_mask = rows - 1, g_mask = G - 1;            // is shifts and masks (powers of two)
  const bool col = lo > 0;
  Fr *vec = data + (size_t)blockIdx.y * vec_stride;
  const int tile = blockIdx.x;
  // tile -> (hi_idx, low_grp)
  size_t base;
  int low0 = 0;                                       // low bits (below lo) of group 0's index
  if (col) {
    int lg_log = lo - g_log;                         // log2 of the tiles per high index
    int hi_idx = tile >> lg_log, low_grp = tile & ((1 << lg_log) - 1);
    low0 = low_grp * G;
    base = ((size_t)hi_idx << hi) + low0;
  } else {
    base = (size_t)tile * TILE;
  }
  // ---- load -----------------------------------------------------------------------------
  for (int i = threadIdx.x; i < TILE; i += THREADS) {
    int k, g;
    size_t addr;
    int sidx;
    if (col) { g = i & g_mask; k = i >> g_log; addr = base + ((size_t)k << lo) + g; sidx = i; }
    else { k = i & rows_mask; g = i >> rows_log; addr = base + i; sidx = i; }
    sts_fr(sm, sidx, ldg_fr(vec + addr));
  }
  __syncthreads();
  const int stride = col ? G : 1;
  // ---- butterflies ----------------------------------------------------------------------
  for (int st = 0; st < rows_log; st++) {
    const int sb = DIF ? (rows_log - 1 - st) : st;    // bit (within the tile rows) of this stage
    const int s = lo + sb;                            // global stage: half distance 2^s
    const int d = 1 << sb;
#pragma unroll (UNR ? UNR : 1)
    for (int b = threadIdx.x; b < TILE / 2; b += THREADS) {
      int kk, g;
      if (col) { g = b & g_mask; kk = b >> g_log; } else { kk = b & (rows_mask >> 1); g = b >> (rows_log - 1); }
      int k = ((kk >> sb) << (sb + 1)) | (kk & (d - 1));
      int i0 = col ? ((k << g_log) + g) : ((g << rows_log) + k);
      int i1 = i0 + d * stride;
      Fr u = lds_fr(sm, i0), v = lds_fr(sm, i1);
      if (s == 0) {                                   // half distance 1: every twiddle is omega^0 = 1 (block-uniform)
        sts_fr(sm, i0, u + v);
        sts_fr(sm, i1, u - v);
        continue;
      }
      size_t j = ((size_t)(k & (d - 1)) << lo) + (col ? (low0 + g) : 0);
      size_t e = j << (logn - 1 - s);
      Fr w = ldg_fr(tw + e);
      if (DIF) {
        sts_fr(sm, i0, u + v);
        sts_fr(sm, i1, (u - v) * w);
      } else {
        Fr t = v * w;
        sts_fr(sm, i0, u + t);
        sts_fr(sm, i1, u - t);
      }
    }
    __syncthreads();
  }
  // ---- store ----------------------------------------------------------------------------
  for (int i = threadIdx.x; i < TILE; i += THREADS) {
    int k, g;
    size_t addr;
    int sidx;
    if (col) { g = i & g_mask; k = i >> g_log; addr = base + ((size_t)k << lo) + g; sidx = i; }
    else { k = i & rows_mask; g = i >> rows_log; addr = base + i; sidx = i; }
    Fr x = lds_fr(sm, sidx);
    if (scale) x = x * ldg_fr(scale + addr);
    stg_fr(vec + addr, x);
  }
}

// The same pass with TWO stages per barrier (radix 4): a thread takes the four elements that differ in two adjacent
// stage bits from shared memory, runs both butterfly stages on them in registers and stores them back - half the
// shared-memory traffic, barriers and index arithmetic of k_ntt_pass; an odd stage count ends with one radix-2 stage.
// Twiddles are the ones the two radix-2 stages would load: two for the stage of the upper bit (its two butterflies sit
// at rows k and k + d0), one for the stage of the lower bit (both of its butterflies have the same row bits below it).
template <bool DIF>
__global__ void __launch_bounds__(256, 3) k_ntt_pass4(Fr *data, size_t vec_stride, int logn, int lo, int hi,
                                                      const Fr *__restrict__ tw, const Fr *__restrict__ scale) {
  constexpr int THREADS = 256;
  extern __shared__ uint32_t sm[];
  const int rows_log = hi - lo, rows = 1 << rows_log;
  const int g_log = TILE_LOG - rows_log, G = 1 << g_log;
  const int rows_mask = rows - 1, g_mask = G - 1;
  const bool col = lo > 0;
  Fr *vec = data + (size_t)blockIdx.y * vec_stride;
  const int tile = blockIdx.x;
  size_t base;
  int low0 = 0;
  if (col) {
    int lg_log = lo - g_log;
    int hi_idx = tile >> lg_log, low_grp = tile & ((1 << lg_log) - 1);
    low0 = low_grp * G;
    base = ((size_t)hi_idx << hi) + low0;
  } else {
    base = (size_t)tile * TILE;
  }
  for (int i = threadIdx.x; i < TILE; i += THREADS) {
    size_t addr;
    if (col) { int g = i & g_mask, k = i >> g_log; addr = base + ((size_t)k << lo) + g; }
    else addr = base + i;
    sts_fr(sm, i, ldg_fr(vec + addr));
  }
  __syncthreads();
  const int stride = col ? G : 1;
  // twiddle of the butterfly whose lower row is `krow` in the stage of row bit `sb` (column g)
  auto twd = [&](int krow, int sb, int g) -> Fr {
    const size_t j = ((size_t)(krow & ((1 << sb) - 1)) << lo) + (col ? (low0 + g) : 0);
    return ldg_fr(tw + (j << (logn - 1 - (lo + sb))));
  };
  int st = 0;
  for (; st + 1 < rows_log; st += 2) {
    const int sb0 = DIF ? rows_log - 2 - st : st, sb1 = sb0 + 1;     // the two row bits of this double stage
    const int d0 = 1 << sb0;
    const bool unit0 = lo + sb0 == 0;                                 // global stage 0: every twiddle is 1
#pragma unroll 2
    for (int q = threadIdx.x; q < TILE / 4; q += THREADS) {
      int kk, g;
      if (col) { g = q & g_mask; kk = q >> g_log; } else { kk = q & (rows_mask >> 2); g = q >> (rows_log - 2); }
      const int k = ((kk >> sb0) << (sb0 + 2)) | (kk & (d0 - 1));    // row with bits sb0, sb1 clear
      const int i00 = col ? ((k << g_log) + g) : ((g << rows_log) + k);
      const int i01 = i00 + d0 * stride, i10 = i00 + 2 * d0 * stride, i11 = i10 + d0 * stride;
      const Fr x00 = lds_fr(sm, i00), x01 = lds_fr(sm, i01), x10 = lds_fr(sm, i10), x11 = lds_fr(sm, i11);
      if (DIF) {
        const Fr a0 = x00 + x10, a2 = (x00 - x10) * twd(k, sb1, g);
        const Fr a1 = x01 + x11, a3 = (x01 - x11) * twd(k + d0, sb1, g);
        sts_fr(sm, i00, a0 + a1);
        sts_fr(sm, i10, a2 + a3);
        if (unit0) {
          sts_fr(sm, i01, a0 - a1);
          sts_fr(sm, i11, a2 - a3);
        } else {
          const Fr w0 = twd(k, sb0, g);
          sts_fr(sm, i01, (a0 - a1) * w0);
          sts_fr(sm, i11, (a2 - a3) * w0);
        }
      } else {
        Fr t0 = x01, t1 = x11;
        if (!unit0) {
          const Fr w0 = twd(k, sb0, g);
          t0 = t0 * w0;
          t1 = t1 * w0;
        }
        const Fr b0 = x00 + t0, b1 = x00 - t0, b2 = x10 + t1, b3 = x10 - t1;
        const Fr u0 = b2 * twd(k, sb1, g), u1 = b3 * twd(k + d0, sb1, g);
        sts_fr(sm, i00, b0 + u0);
        sts_fr(sm, i10, b0 - u0);
        sts_fr(sm, i01, b1 + u1);
        sts_fr(sm, i11, b1 - u1);
      }
    }
    __syncthreads();
  }
  if (st < rows_log) {                                                // odd count: the last stage alone
    const int sb = DIF ? (rows_log - 1 - st) : st;
    const int s = lo + sb, d = 1 << sb;
#pragma unroll 4
    for (int b = threadIdx.x; b < TILE / 2; b += THREADS) {
      int kk, g;
      if (col) { g = b & g_mask; kk = b >> g_log; } else { kk = b & (rows_mask >> 1); g = b >> (rows_log - 1); }
      const int k = ((kk >> sb) << (sb + 1)) | (kk & (d - 1));
      const int i0 = col ? ((k << g_log) + g) : ((g << rows_log) + k);
      const int i1 = i0 + d * stride;
      const Fr u = lds_fr(sm, i0), v = lds_fr(sm, i1);
      if (s == 0) {
        sts_fr(sm, i0, u + v);
        sts_fr(sm, i1, u - v);
        continue;
      }
      const Fr w = twd(k, sb, g);
      if (DIF) {
        sts_fr(sm, i0, u + v);
        sts_fr(sm, i1, (u - v) * w);
      } else {
        const Fr t = v * w;
        sts_fr(sm, i0, u + t);
        sts_fr(sm, i1, u - t);
      }
    }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < TILE; i += THREADS) {
    size_t addr;
    if (col) { int g = i & g_mask, k = i >> g_log; addr = base + ((size_t)k << lo) + g; }
    else addr = base + i;
    Fr x = lds_fr(sm, i);
    if (scale) x = x * ldg_fr(scale + addr);
    stg_fr(vec + addr, x);
  }
}

// out[bitrev(i)] = in[i]
__global__ void k_bitrev(Fr *out, const Fr *in, int logn, size_t vec_stride) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >> logn) return;
  size_t j = __brev((uint32_t)i) >> (32 - logn);
  stg_fr(out + blockIdx.y * vec_stride + j, ldg_fr(in + blockIdx.y * vec_stride + i));
}

__global__ void k_to_mont(Fr *x, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) stg_fr(x + i, ldg_fr(x + i).to_mont());
}
__global__ void k_from_mont(Fr *x, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) stg_fr(x + i, ldg_fr(x + i).from_mont());
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return e_; } while (0)

cudaError_t NttPlan::init(int logn_, cudaStream_t st) {
  logn = logn_;
  if (logn < TILE_LOG || logn > 27) return cudaErrorInvalidValue;    // 2^11 = one tile, one pass over stages [0, 11)
  size_t n = (size_t)1 << logn;
  CK(cudaMalloc(&tw_fwd, (n / 2) * sizeof(Fr)));
  CK(cudaMalloc(&tw_inv, (n / 2) * sizeof(Fr)));
  CK(cudaMalloc(&coset_scale, n * sizeof(Fr)));
  Fr *pw, *ninv;
  CK(cudaMalloc(&pw, 32 * sizeof(Fr)));
  CK(cudaMalloc(&ninv, sizeof(Fr)));
  Fr one = Fr::one();
  int th = 256;
  k_root_powers<<<1, 1, 0, st>>>(pw, logn, 0, ninv);
  k_powers<<<(unsigned)((n / 2 + th - 1) / th), th, 0, st>>>(tw_fwd, pw, n / 2, one);
  k_root_powers<<<1, 1, 0, st>>>(pw, logn, 1, nullptr);
  k_powers<<<(unsigned)((n / 2 + th - 1) / th), th, 0, st>>>(tw_inv, pw, n / 2, one);
  // inc = omega_{2^(logn+1)}  (snarkjs: Fr.w[power+1])
  k_root_powers<<<1, 1, 0, st>>>(pw, logn + 1, 0, nullptr);
  k_coset_scale<<<(unsigned)((n + th - 1) / th), th, 0, st>>>(coset_scale, pw, ninv, logn);
  CK(cudaStreamSynchronize(st));
  CK(cudaFree(pw));
  CK(cudaFree(ninv));
#define ZKB_NTT_ATTR(U, T)                                                                                          \
  CK((cudaFuncSetAttribute(k_ntt_pass<true, U, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE * 32)));      \
  CK((cudaFuncSetAttribute(k_ntt_pass<false, U, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE * 32)));
  ZKB_NTT_ATTR(0, 256) ZKB_NTT_ATTR(4, 256) ZKB_NTT_ATTR(8, 128) ZKB_NTT_ATTR(4, 128) ZKB_NTT_ATTR(2, 512)
#undef ZKB_NTT_ATTR
  CK((cudaFuncSetAttribute(k_ntt_pass4<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE * 32)));
  CK((cudaFuncSetAttribute(k_ntt_pass4<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE * 32)));
  return cudaGetLastError();
}

void NttPlan::destroy() {
  if (tw_fwd) cudaFree(tw_fwd);
  if (tw_inv) cudaFree(tw_inv);
  if (coset_scale) cudaFree(coset_scale);
  tw_fwd = tw_inv = coset_scale = nullptr;
}

// stage ranges, top-down: column passes of 9 stages while more than 11 remain (hi >= 11 keeps a tile
// inside one high index), then one contiguous pass over the remaining low stages [0, <=11).
static int split_passes(int logn, int lo[4], int hi[4]) {
  int np = 0, top = logn;
  while (top > TILE_LOG) {
    lo[np] = top - 9;
    hi[np] = top;
    top -= 9;
    np++;
  }
  lo[np] = 0;
  hi[np] = top;
  return np + 1;
}

// butterflies of one thread and stage the compiler may interleave (independent products hide each other's carry-chain
// latency; three 64 KB tiles per SM leave 85 registers per thread).  ZKB_NTT_UNROLL overrides the default.
// Threads per 2048-element tile and butterflies of one thread the compiler may interleave (independent products hide
// each other's carry-chain latency; three 64 KB tiles per SM leave 85 registers per thread).  Measured on B200, ntt_join
// per 512 proofs (profiles/r02_ntt_variant_sweep.log, r02_ntt_radix4_sweep.log): 256 threads, loop not unrolled 68.7 ms;
// 256 x 4 interleaved 65.9 ms; 128 x 8 69.6; 128 x 4 71.2; 512 x 2 70.9; radix 4 (k_ntt_pass4, two stages per barrier)
// 63.7 - 65.4 ms (default).  ZKB_NTT_VARIANT selects another one.
static int ntt_variant() {
  static const int v = getenv("ZKB_NTT_VARIANT") ? atoi(getenv("ZKB_NTT_VARIANT")) : 0;
  return v;
}
template <bool DIF>
static void launch_pass(dim3 grid, cudaStream_t st, Fr *data, size_t vec_stride, int logn, int lo, int hi, const Fr *tw,
                        const Fr *scale) {
#define ZKB_NTT_GO(U, T) k_ntt_pass<DIF, U, T><<<grid, T, TILE * 32, st>>>(data, vec_stride, logn, lo, hi, tw, scale)
  switch (ntt_variant()) {
    case 1: ZKB_NTT_GO(0, 256); break;
    case 2: ZKB_NTT_GO(8, 128); break;
    case 3: ZKB_NTT_GO(4, 128); break;
    case 4: ZKB_NTT_GO(2, 512); break;
    case 5: ZKB_NTT_GO(4, 256); break;     // radix 2, four butterflies interleaved
    default: k_ntt_pass4<DIF><<<grid, 256, TILE * 32, st>>>(data, vec_stride, logn, lo, hi, tw, scale); break;   // radix 4
  }
#undef ZKB_NTT_GO
}

// DIF passes: natural in -> bit-reversed out.  inverse selects the twiddle table; with_coset_scale
// multiplies by n^-1 * inc^bitrev(p) in the last pass.
cudaError_t NttPlan::dif(Fr *data, int nvec, size_t vec_stride, bool inverse, bool with_coset_scale,
                         cudaStream_t st) const {
  int lo[4], hi[4];
  int np = split_passes(logn, lo, hi);
  size_t n = (size_t)1 << logn;
  dim3 grid((unsigned)(n / TILE), nvec);
  for (int p = 0; p < np; p++) {
    const Fr *sc = (with_coset_scale && p == np - 1) ? coset_scale : nullptr;
    launch_pass<true>(grid, st, data, vec_stride, logn, lo[p], hi[p], inverse ? tw_inv : tw_fwd, sc);
  }
  return cudaGetLastError();
}

// DIT passes: bit-reversed in -> natural out
cudaError_t NttPlan::dit(Fr *data, int nvec, size_t vec_stride, bool inverse, cudaStream_t st) const {
  int lo[4], hi[4];
  int np = split_passes(logn, lo, hi);
  size_t n = (size_t)1 << logn;
  dim3 grid((unsigned)(n / TILE), nvec);
  for (int p = np - 1; p >= 0; p--)
    launch_pass<false>(grid, st, data, vec_stride, logn, lo[p], hi[p], inverse ? tw_inv : tw_fwd, nullptr);
  return cudaGetLastError();
}

cudaError_t ntt_bitrev(Fr *out, const Fr *in, int logn, int nvec, size_t vec_stride, cudaStream_t st) {
  size_t n = (size_t)1 << logn;
  dim3 grid((unsigned)((n + 255) / 256), nvec);
  k_bitrev<<<grid, 256, 0, st>>>(out, in, logn, vec_stride);
  return cudaGetLastError();
}
cudaError_t fr_to_mont(Fr *x, size_t n, cudaStream_t st) {
  k_to_mont<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(x, n);
  return cudaGetLastError();
}
cudaError_t fr_from_mont(Fr *x, size_t n, cudaStream_t st) {
  k_from_mont<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(x, n);
  return cudaGetLastError();
}

}  // namespace zkb
