// Batched fixed-base Pippenger MSM over BN254 G1/G2 - see msm.cu
#pragma once
#include <cuda_runtime.h>
#include "ec.cuh"

namespace zkb {

// Window configuration: c-bit signed digits, `windows` of them cover 254 bits, 2^(c-1) buckets (|d| - 1).
struct MsmCfg {
  int c = 16, windows = 16;
  uint32_t buckets = 1u << 15;
  uint32_t parts() const { return buckets / 8; }    // capacity of the level-1 partial arrays (chunks of >= 8 buckets)
};
// c in [12, 16] (the reduction needs >= 2048 buckets); the top window must have room for the signed-digit carry
inline MsmCfg msm_cfg(int c) {
  MsmCfg k;
  k.c = c;
  k.windows = (254 + c - 1) / c;
  if (254 - (k.windows - 1) * c >= c - 1) k.windows++;
  k.buckets = 1u << (c - 1);
  return k;
}

// Fixed-base table: tab[j * n + k] = 2^(c j) * base_k (affine, Montgomery), j < windows
template <class F>
struct MsmTable {
  Affine<F> *tab = nullptr;
  uint32_t n = 0;
  MsmCfg cfg;
};

// One sorted digit set per batch item: for every scalar k and window j with digit d != 0 an entry
// (j * n + k) | (d < 0 ? 1u<<31 : 0) filed under bucket |d| - 1.
struct MsmSort {
  uint32_t n = 0;          // scalars per batch item
  uint32_t batch = 0;      // capacity (batch items)
  MsmCfg cfg;
  uint32_t *counts = nullptr;   // [batch][buckets]      bucket sizes
  uint32_t *offsets = nullptr;  // [batch][buckets + 1]  exclusive scan
  uint32_t *cursor = nullptr;   // [batch][buckets]      scatter cursors
  uint32_t *entries = nullptr;  // [batch][n * windows]
  uint32_t *order = nullptr;    // [batch][buckets]      bucket ids by descending size
  uint32_t *n_long = nullptr;   // [batch]               how many leading buckets of `order` a whole warp sums
  uint32_t max_long = 256;      // upper bound of n_long for the last run (grid of the warp-per-bucket kernel)
  cudaError_t alloc(uint32_t n, uint32_t batch, MsmCfg cfg);
  void free_all();
  // scalars: [batch] vectors of n canonical 256-bit values, `scalar_stride` elements apart
  cudaError_t run(const Fr *scalars, size_t scalar_stride, uint32_t nbatch, cudaStream_t st);
  // variable-base form: see k_digits_vb in msm.cu
  cudaError_t run_vb(const Fr *scalars, size_t scalar_stride, uint32_t nsets, int windows, cudaStream_t st);
};

int msm_sort_launches();   // kernels one MsmSort::run launches
// window sums of a variable-base MSM -> the MSM: out[s] = sum_j 2^(c j) sum_t in[j * nsets * fold + s * fold + t]
cudaError_t msm_horner(const XYZZ<Fq> *in, XYZZ<Fq> *out, uint32_t nsets, uint32_t fold, int windows, int c, cudaStream_t st);

// workspace for bucket sums + reduction of `slots` simultaneous (batch item, table) pairs
template <class F>
struct MsmWork {
  uint32_t slots = 0;
  MsmCfg cfg;
  XYZZ<F> *buckets = nullptr;   // [slots][buckets]
  XYZZ<F> *part_r = nullptr;    // [slots][parts]
  XYZZ<F> *part_s = nullptr;    // [slots][parts]
  cudaError_t alloc(uint32_t slots, MsmCfg cfg);
  void free_all();
};

// Batched-affine pair tree for dense G1 bucket lists (the H MSM: ~64 entries per bucket).  Level l halves every
// bucket's list by adding neighbours pairwise IN AFFINE coordinates; the divisions of a thread's `group` additions
// share one field inversion (Montgomery's trick, prefix products parked in HBM), so an addition costs 6 field products
// + inversion / group instead of the 10 of an XYZZ mixed addition.  After `levels` halvings the short remainders are
// summed by the XYZZ kernel.  pa / pb ping-pong the level outputs.
struct MsmAffineWs {
  static constexpr int MAX_LEVELS = 4;
  uint32_t batch = 0, buckets = 0;   // batch: items the workspace holds at once (larger launches run in sub-batches)
  size_t cap_a = 0, cap_b = 0;       // points per batch item in pa (levels 1, 3) and pb (levels 2, 4)
  int levels = 3;
  uint32_t group[MAX_LEVELS] = {512, 512, 512, 512};
  Affine<Fq> *pa = nullptr, *pb = nullptr;
  Fq *park = nullptr;                // [batch][cap_b] parked prefix products of levels >= 1 (level 0 parks in pb)
  uint32_t *lvl_off = nullptr;       // [MAX_LEVELS][batch][buckets + 1] exclusive scans of the per-level list lengths
  cudaError_t alloc(uint32_t n_entries, uint32_t batch, MsmCfg cfg);
  void free_all();
};
// bucket sums of one G1 table through the pair tree -> work.buckets[slot0 + b] (same contract as msm_accumulate, ntab = 1)
cudaError_t msm_accumulate_affine(const MsmSort &sort, const MsmTable<Fq> &table, uint32_t nbatch, MsmWork<Fq> &work,
                                  uint32_t slot0, MsmAffineWs &ws, cudaStream_t st, size_t tab_batch_stride = 0,
                                  uint32_t tab_mod = 0);
int msm_affine_launches(const MsmAffineWs &ws, uint32_t nbatch);
// executed work of the pair tree over `sort`: out[0] = affine additions, out[1] = XYZZ mixed additions of the tail,
// out[2] = field inversions
cudaError_t msm_affine_counts(const MsmSort &sort, const MsmAffineWs &ws, uint32_t nbatch, unsigned long long *out3,
                              cudaStream_t st);

// build the window table from n affine bases (Montgomery form, zkey layout; (0,0) = infinity).  subs > 1: `bases`
// holds subs consecutive ranges of n points and the table is subs consecutive [windows][n] blocks (a large MSM split
// by point range into sub-MSMs that run as batch items, see tab_batch_stride below).
template <class F>
cudaError_t msm_build_table(MsmTable<F> &t, const Affine<F> *bases, uint32_t n, MsmCfg cfg, cudaStream_t st,
                            uint32_t subs = 1);

// bucket sums of (batch item b, table t) -> work.buckets[slot0 + b*ntab + t].  Tables, sort and work must share cfg.
template <class F>
// tab_batch_stride: batch item b reads its bases at table + (b % tab_mod) * tab_batch_stride (0: every item shares
// the table; tab_mod = 0 means nbatch).  With a key split into `subs` point ranges, batch item p * subs + s is range
// s of proof p and tab_mod = subs.
cudaError_t msm_accumulate(const MsmSort &sort, const MsmTable<F> *tables, int ntab, uint32_t nbatch, MsmWork<F> &work,
                           uint32_t slot0, cudaStream_t st, size_t tab_batch_stride = 0, uint32_t tab_mod = 0);
// out[p * ntab + t] = sum_s in[(p * subs + s) * ntab + t]
template <class F>
cudaError_t msm_fold_subs(const XYZZ<F> *in, XYZZ<F> *out, uint32_t nproofs, uint32_t subs, uint32_t ntab, cudaStream_t st);
// out[i] = sum_b (b+1) * buckets[slot0 + i][b], i < nslots
template <class F>
cudaError_t msm_reduce(MsmWork<F> &work, uint32_t slot0, uint32_t nslots, XYZZ<F> *out, cudaStream_t st);
// accumulate + reduce: out[b * ntab + t]
template <class F>
cudaError_t msm_run(const MsmSort &sort, const MsmTable<F> *tables, int ntab, uint32_t nbatch, MsmWork<F> &work,
                    XYZZ<F> *out, cudaStream_t st);
// exact number of mixed adds an accumulate launch over (sort, table) executes (measurement aid, not timed)
template <class F>
cudaError_t msm_count_madds(const MsmSort &sort, const MsmTable<F> &table, uint32_t nbatch, unsigned long long *host_total,
                            cudaStream_t st);

}  // namespace zkb
