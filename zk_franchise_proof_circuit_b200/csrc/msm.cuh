// Batched fixed-base Pippenger MSM over BN254 G1/G2 - see msm.cu
#pragma once
#include <cuda_runtime.h>
#include "ec.cuh"

namespace zkb {

static constexpr int MSM_C = 16;                       // window bits
static constexpr int MSM_WINDOWS = 16;                 // ceil(254 / 16); top window holds 14 bits
static constexpr int MSM_BUCKETS = 1 << (MSM_C - 1);   // signed digits: |d| in 1..2^15 -> bucket |d|-1

// Fixed-base table: tab[j * n + k] = 2^(16 j) * base_k (affine, Montgomery), j < 16
template <class F>
struct MsmTable {
  Affine<F> *tab = nullptr;
  uint32_t n = 0;
};

// One sorted digit set: for every scalar k and window j with digit d != 0 an entry
// (j * n + k) | (d < 0 ? 1u<<31 : 0) filed under bucket |d| - 1.
struct MsmSort {
  uint32_t n = 0;          // scalars per batch item
  uint32_t batch = 0;      // capacity (batch items)
  uint32_t *counts = nullptr;   // [batch][BUCKETS]      bucket sizes
  uint32_t *offsets = nullptr;  // [batch][BUCKETS + 1]  exclusive scan
  uint32_t *cursor = nullptr;   // [batch][BUCKETS]      scatter cursors
  uint32_t *entries = nullptr;  // [batch][n * WINDOWS]
  cudaError_t alloc(uint32_t n, uint32_t batch);
  void free_all();
  // scalars: [batch] vectors of n canonical 256-bit values, `scalar_stride` elements apart
  cudaError_t run(const Fr *scalars, size_t scalar_stride, uint32_t nbatch, cudaStream_t st);
};

// workspace for bucket sums + reduction of `slots` simultaneous (batch item, table) pairs
template <class F>
struct MsmWork {
  uint32_t slots = 0;
  XYZZ<F> *buckets = nullptr;   // [slots][BUCKETS]
  XYZZ<F> *part_r = nullptr;    // [slots][1024]
  XYZZ<F> *part_s = nullptr;    // [slots][1024]
  cudaError_t alloc(uint32_t slots);
  void free_all();
};

// build the 16-window table from n affine bases (Montgomery form, zkey layout; (0,0) = infinity)
template <class F>
cudaError_t msm_build_table(MsmTable<F> &t, const Affine<F> *bases, uint32_t n, cudaStream_t st);

// For each batch item b < nbatch and each table t < ntab: out[(b * ntab + t)] = sum over the sorted
// entries of `sort` item b of +-tab_t[entry].  All tables must have n == sort.n.
template <class F>
cudaError_t msm_run(const MsmSort &sort, const MsmTable<F> *tables, int ntab, uint32_t nbatch, MsmWork<F> &work,
                    XYZZ<F> *out, cudaStream_t st);

// split form: bucket sums of (batch item b, table t) go to work.buckets[slot0 + b*ntab + t]; reduce separately
template <class F>
cudaError_t msm_accumulate(const MsmSort &sort, const MsmTable<F> *tables, int ntab, uint32_t nbatch, MsmWork<F> &work,
                           uint32_t slot0, cudaStream_t st);
template <class F>
cudaError_t msm_reduce(MsmWork<F> &work, uint32_t slot0, uint32_t nslots, XYZZ<F> *out, cudaStream_t st);
// exact number of mixed adds an accumulate launch over (sort, table) executes (measurement aid, not timed)
template <class F>
cudaError_t msm_count_madds(const MsmSort &sort, const MsmTable<F> &table, uint32_t nbatch, unsigned long long *host_total,
                            cudaStream_t st);

}  // namespace zkb
