// BN254 prime-field arithmetic for sm_100a: 8 x 32-bit limbs, Montgomery form with R = 2^256.
//
// The same R as snarkjs / rapidsnark / the circom wasm use, so zkey points, zkey coefficients and
// the wasm's Poseidon constants are consumed without conversion.
//
// Device multiply: interleaved (CIOS) Montgomery product built from mad.lo.cc / madc.hi.cc carry
// chains.  The products of even-indexed and odd-indexed limbs are kept in two separate 8-limb
// accumulators (`ev`, `od`; od is offset by one limb) so that every 64-bit product lands on an
// aligned limb pair and each row is ONE carry chain; ptxas turns each lo/hi pair into a single
// IMAD.WIDE.U32 with carry-in/out.  The one-limb right shift of the reduction is absorbed by
// swapping the roles of the two accumulators every row.  Integer pipe only - there is no dense
// contraction here for tensor cores to do.  SASS per product: 121 IMAD.WIDE + 16 IMAD / IMAD.HI = 137
// multiplier instructions ("modmul" in the roofline bookkeeping).
//
// Device square (sqr_dev): the 512-bit square from 36 instead of 64 limb products (fp_sqr.inc, generated and
// checked by tools/gen_sqr.py), then the reduction rows of the product alone: 117 multiplier instructions.
//
// Inline-asm rule used throughout: an asm block holds a whole carry chain (the CC flag never crosses
// statements) and every pure output that is written before the last input is read is early-clobber
// ("=&r"), otherwise ptxas/nvcc may give it the register of a not-yet-consumed input.
//
// The host versions (plain 64-bit C) exist for two reasons: host-side table preparation, and so
// that the per-thread algorithms built on top (curve formulas, Poseidon, the census witness
// program) can be unit-tested on a machine without a GPU.  They are not a product fallback: every
// product entry point runs the device path and fails if no sm_100 device is present.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define ZKB_HD __host__ __device__ __forceinline__
#define ZKB_D __device__ __forceinline__
#else
#define ZKB_HD inline
#define ZKB_D inline
#endif

namespace zkb {

struct FqParams {
  ZKB_HD static constexpr uint32_t mod(int i) {
    constexpr uint32_t m[8] = {0xd87cfd47u, 0x3c208c16u, 0x6871ca8du, 0x97816a91u,
                               0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
    return m[i];
  }
  ZKB_HD static constexpr uint32_t one(int i) {  // R mod q
    constexpr uint32_t m[8] = {0xc58f0d9du, 0xd35d438du, 0xf5c70b3du, 0x0a78eb28u,
                               0x7879462cu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
    return m[i];
  }
  ZKB_HD static constexpr uint32_t r2(int i) {  // R^2 mod q
    constexpr uint32_t m[8] = {0x538afa89u, 0xf32cfc5bu, 0xd44501fbu, 0xb5e71911u,
                               0x0a417ff6u, 0x47ab1effu, 0xcab8351fu, 0x06d89f71u};
    return m[i];
  }
  static constexpr uint32_t INV = 0xe4866389u;  // -q^-1 mod 2^32
};

struct FrParams {
  ZKB_HD static constexpr uint32_t mod(int i) {
    constexpr uint32_t m[8] = {0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u,
                               0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
    return m[i];
  }
  ZKB_HD static constexpr uint32_t one(int i) {  // R mod r
    constexpr uint32_t m[8] = {0x4ffffffbu, 0xac96341cu, 0x9f60cd29u, 0x36fc7695u,
                               0x7879462eu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
    return m[i];
  }
  ZKB_HD static constexpr uint32_t r2(int i) {  // R^2 mod r
    constexpr uint32_t m[8] = {0xae216da7u, 0x1bb8e645u, 0xe35c59e3u, 0x53fe3ab1u,
                               0x53bb8085u, 0x8c49833du, 0x7f4e44a5u, 0x0216d0b1u};
    return m[i];
  }
  static constexpr uint32_t INV = 0xefffffffu;  // -r^-1 mod 2^32
};

template <class P> struct Fp;
#if defined(__CUDACC__)
// Out-of-line Montgomery product.  The curve formulas (10-30 products per point operation) call this instead of
// inlining ~190 SASS instructions per product: a bucket-accumulation loop body then fits the instruction cache
// (inlined it was ~57 KB and the kernel ran instruction-fetch bound, 8x below the IMAD rate).
template <class P> __device__ __noinline__ Fp<P> fp_mul_call(Fp<P> a, Fp<P> b);
template <class P> __device__ __noinline__ Fp<P> fp_sqr_call(Fp<P> a);
#endif

template <class P>
struct alignas(16) Fp {
  uint32_t v[8];

  ZKB_HD static Fp zero() {
    Fp r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = 0;
    return r;
  }
  ZKB_HD static Fp one() {
    Fp r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = P::one(i);
    return r;
  }
  ZKB_HD static Fp r2() {
    Fp r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = P::r2(i);
    return r;
  }
  ZKB_HD bool is_zero() const {
    uint32_t t = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) t |= v[i];
    return t == 0;
  }
  ZKB_HD bool operator==(const Fp &o) const {
    uint32_t t = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) t |= v[i] ^ o.v[i];
    return t == 0;
  }
  ZKB_HD bool operator!=(const Fp &o) const { return !(*this == o); }

  // ---- r = a - mod if a >= mod (a < 2*mod) ------------------------------------------------
  ZKB_HD static void final_sub(uint32_t a[8]) {
    uint32_t t[8];
#if defined(__CUDA_ARCH__)
    uint32_t borrow;
    asm("sub.cc.u32 %0, %9, %17;\n\t"
        "subc.cc.u32 %1, %10, %18;\n\t"
        "subc.cc.u32 %2, %11, %19;\n\t"
        "subc.cc.u32 %3, %12, %20;\n\t"
        "subc.cc.u32 %4, %13, %21;\n\t"
        "subc.cc.u32 %5, %14, %22;\n\t"
        "subc.cc.u32 %6, %15, %23;\n\t"
        "subc.cc.u32 %7, %16, %24;\n\t"
        "subc.u32 %8, 0, 0;\n\t"
        : "=&r"(t[0]), "=&r"(t[1]), "=&r"(t[2]), "=&r"(t[3]), "=&r"(t[4]), "=&r"(t[5]), "=&r"(t[6]), "=&r"(t[7]),
          "=&r"(borrow)
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
          "r"(P::mod(0)), "r"(P::mod(1)), "r"(P::mod(2)), "r"(P::mod(3)), "r"(P::mod(4)), "r"(P::mod(5)),
          "r"(P::mod(6)), "r"(P::mod(7)));
    if (borrow == 0) {
#pragma unroll
      for (int i = 0; i < 8; i++) a[i] = t[i];
    }
#else
    uint64_t br = 0;
    for (int i = 0; i < 8; i++) {
      uint64_t d = (uint64_t)a[i] - P::mod(i) - br;
      t[i] = (uint32_t)d;
      br = (d >> 32) & 1;
    }
    if (!br)
      for (int i = 0; i < 8; i++) a[i] = t[i];
#endif
  }

  ZKB_HD Fp operator+(const Fp &o) const {
    Fp r;
#if defined(__CUDA_ARCH__)
    asm("add.cc.u32 %0, %8, %16;\n\t"
        "addc.cc.u32 %1, %9, %17;\n\t"
        "addc.cc.u32 %2, %10, %18;\n\t"
        "addc.cc.u32 %3, %11, %19;\n\t"
        "addc.cc.u32 %4, %12, %20;\n\t"
        "addc.cc.u32 %5, %13, %21;\n\t"
        "addc.cc.u32 %6, %14, %22;\n\t"
        "addc.u32 %7, %15, %23;\n\t"
        : "=&r"(r.v[0]), "=&r"(r.v[1]), "=&r"(r.v[2]), "=&r"(r.v[3]), "=&r"(r.v[4]), "=&r"(r.v[5]), "=&r"(r.v[6]),
          "=&r"(r.v[7])
        : "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(o.v[0]),
          "r"(o.v[1]), "r"(o.v[2]), "r"(o.v[3]), "r"(o.v[4]), "r"(o.v[5]), "r"(o.v[6]), "r"(o.v[7]));
#else
    uint64_t c = 0;
    for (int i = 0; i < 8; i++) {
      c += (uint64_t)v[i] + o.v[i];
      r.v[i] = (uint32_t)c;
      c >>= 32;
    }
#endif
    final_sub(r.v);  // both operands < mod < 2^254, so no carry out of limb 7
    return r;
  }

  ZKB_HD Fp operator-(const Fp &o) const {
    Fp r;
#if defined(__CUDA_ARCH__)
    uint32_t borrow;
    asm("sub.cc.u32 %0, %9, %17;\n\t"
        "subc.cc.u32 %1, %10, %18;\n\t"
        "subc.cc.u32 %2, %11, %19;\n\t"
        "subc.cc.u32 %3, %12, %20;\n\t"
        "subc.cc.u32 %4, %13, %21;\n\t"
        "subc.cc.u32 %5, %14, %22;\n\t"
        "subc.cc.u32 %6, %15, %23;\n\t"
        "subc.cc.u32 %7, %16, %24;\n\t"
        "subc.u32 %8, 0, 0;\n\t"
        : "=&r"(r.v[0]), "=&r"(r.v[1]), "=&r"(r.v[2]), "=&r"(r.v[3]), "=&r"(r.v[4]), "=&r"(r.v[5]), "=&r"(r.v[6]),
          "=&r"(r.v[7]), "=&r"(borrow)
        : "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(o.v[0]),
          "r"(o.v[1]), "r"(o.v[2]), "r"(o.v[3]), "r"(o.v[4]), "r"(o.v[5]), "r"(o.v[6]), "r"(o.v[7]));
    if (borrow) {
      asm("add.cc.u32 %0, %0, %8;\n\t"
          "addc.cc.u32 %1, %1, %9;\n\t"
          "addc.cc.u32 %2, %2, %10;\n\t"
          "addc.cc.u32 %3, %3, %11;\n\t"
          "addc.cc.u32 %4, %4, %12;\n\t"
          "addc.cc.u32 %5, %5, %13;\n\t"
          "addc.cc.u32 %6, %6, %14;\n\t"
          "addc.u32 %7, %7, %15;\n\t"
          : "+r"(r.v[0]), "+r"(r.v[1]), "+r"(r.v[2]), "+r"(r.v[3]), "+r"(r.v[4]), "+r"(r.v[5]), "+r"(r.v[6]),
            "+r"(r.v[7])
          : "r"(P::mod(0)), "r"(P::mod(1)), "r"(P::mod(2)), "r"(P::mod(3)), "r"(P::mod(4)), "r"(P::mod(5)),
            "r"(P::mod(6)), "r"(P::mod(7)));
    }
#else
    uint64_t br = 0;
    for (int i = 0; i < 8; i++) {
      uint64_t d = (uint64_t)v[i] - o.v[i] - br;
      r.v[i] = (uint32_t)d;
      br = (d >> 32) & 1;
    }
    if (br) {
      uint64_t c = 0;
      for (int i = 0; i < 8; i++) {
        c += (uint64_t)r.v[i] + P::mod(i);
        r.v[i] = (uint32_t)c;
        c >>= 32;
      }
    }
#endif
    return r;
  }

  ZKB_HD Fp neg() const { return zero() - *this; }
  // branch-free c ? a : b
  ZKB_HD static Fp select(bool c, const Fp &a, const Fp &b) {
    Fp r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = c ? a.v[i] : b.v[i];
    return r;
  }
  ZKB_HD Fp dbl() const { return *this + *this; }

#if defined(__CUDA_ARCH__)
  // acc[0..7] += {a0,a2,a4,a6} * b on aligned limb pairs (one carry chain); carry-out added to `top`
  ZKB_D static void cmad_row(uint32_t acc[8], uint32_t a0, uint32_t a2, uint32_t a4, uint32_t a6, uint32_t b,
                             uint32_t &top) {
    asm("mad.lo.cc.u32 %0, %9, %13, %0;\n\t"
        "madc.hi.cc.u32 %1, %9, %13, %1;\n\t"
        "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"
        "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
        "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"
        "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
        "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"
        "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
        "addc.u32 %8, %8, 0;\n\t"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]),
          "+r"(acc[7]), "+r"(top)
        : "r"(a0), "r"(a2), "r"(a4), "r"(a6), "r"(b));
  }
  // same, carry-out discarded (provably zero for the odd accumulator)
  ZKB_D static void cmad_row_nc(uint32_t acc[8], uint32_t a0, uint32_t a2, uint32_t a4, uint32_t a6, uint32_t b) {
    asm("mad.lo.cc.u32 %0, %8, %12, %0;\n\t"
        "madc.hi.cc.u32 %1, %8, %12, %1;\n\t"
        "madc.lo.cc.u32 %2, %9, %12, %2;\n\t"
        "madc.hi.cc.u32 %3, %9, %12, %3;\n\t"
        "madc.lo.cc.u32 %4, %10, %12, %4;\n\t"
        "madc.hi.cc.u32 %5, %10, %12, %5;\n\t"
        "madc.lo.cc.u32 %6, %11, %12, %6;\n\t"
        "madc.hi.u32 %7, %11, %12, %7;\n\t"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]),
          "+r"(acc[7])
        : "r"(a0), "r"(a2), "r"(a4), "r"(a6), "r"(b));
  }
  // x0 += y[1] (carry feeds the chain);  y <- (y >> 64 bits) + {a1,a3,a5,a7} * b
  ZKB_D static void shift_mad_row(uint32_t &x0, uint32_t y[8], uint32_t a1, uint32_t a3, uint32_t a5, uint32_t a7,
                                  uint32_t b) {
    asm("add.cc.u32 %0, %0, %10;\n\t"
        "madc.lo.cc.u32 %1, %16, %20, %11;\n\t"
        "madc.hi.cc.u32 %2, %16, %20, %12;\n\t"
        "madc.lo.cc.u32 %3, %17, %20, %13;\n\t"
        "madc.hi.cc.u32 %4, %17, %20, %14;\n\t"
        "madc.lo.cc.u32 %5, %18, %20, %15;\n\t"
        "madc.hi.cc.u32 %6, %18, %20, %9;\n\t"
        "madc.lo.cc.u32 %7, %19, %20, 0;\n\t"
        "madc.hi.u32 %8, %19, %20, 0;\n\t"
        : "+r"(x0), "=&r"(y[0]), "=&r"(y[1]), "=&r"(y[2]), "=&r"(y[3]), "=&r"(y[4]), "=&r"(y[5]), "=&r"(y[6]),
          "=&r"(y[7])
        : "r"(y[7]), "r"(y[1]), "r"(y[2]), "r"(y[3]), "r"(y[4]), "r"(y[5]), "r"(y[6]), "r"(a1), "r"(a3), "r"(a5),
          "r"(a7), "r"(b));
  }
#endif

  // Montgomery product: this * o / R mod p
  ZKB_HD Fp operator*(const Fp &o) const {
    Fp res;
#if defined(__CUDA_ARCH__)
    const uint32_t *a = v, *b = o.v;
    uint32_t ev[8], od[8];
    // row 0: ev = even-limb products, od = odd-limb products (od is one limb up)
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
      ev[j] = a[j] * b[0];
      ev[j + 1] = __umulhi(a[j], b[0]);
      od[j] = a[j + 1] * b[0];
      od[j + 1] = __umulhi(a[j + 1], b[0]);
    }
    {
      uint32_t m = ev[0] * P::INV;
      cmad_row_nc(od, P::mod(1), P::mod(3), P::mod(5), P::mod(7), m);
      cmad_row(ev, P::mod(0), P::mod(2), P::mod(4), P::mod(6), m, od[7]);
    }
#pragma unroll
    for (int i = 1; i < 8; i++) {
      // roles swap every row: X is the limb-aligned accumulator, Y the one being shifted down
      uint32_t *X = (i & 1) ? od : ev;
      uint32_t *Y = (i & 1) ? ev : od;
      shift_mad_row(X[0], Y, a[1], a[3], a[5], a[7], b[i]);
      cmad_row(X, a[0], a[2], a[4], a[6], b[i], Y[7]);
      uint32_t m = X[0] * P::INV;
      cmad_row_nc(Y, P::mod(1), P::mod(3), P::mod(5), P::mod(7), m);
      cmad_row(X, P::mod(0), P::mod(2), P::mod(4), P::mod(6), m, Y[7]);
    }
    // after 8 rows X = od (i = 7): result = od>>32 ... roles: last X is od, last Y is ev
    // T = (X >> 32) + Y  with X = od, Y = ev
    asm("add.cc.u32 %0, %0, %8;\n\t"
        "addc.cc.u32 %1, %1, %9;\n\t"
        "addc.cc.u32 %2, %2, %10;\n\t"
        "addc.cc.u32 %3, %3, %11;\n\t"
        "addc.cc.u32 %4, %4, %12;\n\t"
        "addc.cc.u32 %5, %5, %13;\n\t"
        "addc.cc.u32 %6, %6, %14;\n\t"
        "addc.u32 %7, %7, 0;\n\t"
        : "+r"(ev[0]), "+r"(ev[1]), "+r"(ev[2]), "+r"(ev[3]), "+r"(ev[4]), "+r"(ev[5]), "+r"(ev[6]), "+r"(ev[7])
        : "r"(od[1]), "r"(od[2]), "r"(od[3]), "r"(od[4]), "r"(od[5]), "r"(od[6]), "r"(od[7]));
#pragma unroll
    for (int i = 0; i < 8; i++) res.v[i] = ev[i];
    final_sub(res.v);
#else
    uint32_t t[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 8; i++) {
      uint64_t c = 0;
      for (int j = 0; j < 8; j++) {
        c += (uint64_t)v[j] * o.v[i] + t[j];
        t[j] = (uint32_t)c;
        c >>= 32;
      }
      c += t[8];
      t[8] = (uint32_t)c;
      t[9] = (uint32_t)(c >> 32);
      uint32_t m = t[0] * P::INV;
      c = ((uint64_t)m * P::mod(0) + t[0]) >> 32;
      for (int j = 1; j < 8; j++) {
        c += (uint64_t)m * P::mod(j) + t[j];
        t[j - 1] = (uint32_t)c;
        c >>= 32;
      }
      c += t[8];
      t[7] = (uint32_t)c;
      t[8] = t[9] + (uint32_t)(c >> 32);
    }
    for (int i = 0; i < 8; i++) res.v[i] = t[i];
    final_sub(res.v);
#endif
    return res;
  }
#if defined(__CUDA_ARCH__)
  // Montgomery square: the 512-bit square from 36 limb products (fp_sqr.inc, generated and checked by
  // tools/gen_sqr.py), then the reduction rows of operator* without their a * b_i products - 100 IMAD.WIDE instead of
  // 128.  (T_lo + m p) / R + T_hi < p + 1 + p^2 / R < 2 p for both BN254 fields: one final subtraction.
  ZKB_D Fp sqr_dev() const {
    const uint32_t *a = v;
    uint32_t E[16], O[16];
#pragma unroll
    for (int i = 0; i < 16; i++) E[i] = O[i] = 0;
#include "fp_sqr.inc"
    uint32_t ev[8], od[8];
#pragma unroll
    for (int i = 0; i < 8; i++) ev[i] = O[i];
    {
      const uint32_t m = ev[0] * P::INV;
#pragma unroll
      for (int j = 0; j < 8; j += 2) {
        od[j] = P::mod(j + 1) * m;
        od[j + 1] = __umulhi(P::mod(j + 1), m);
      }
      cmad_row(ev, P::mod(0), P::mod(2), P::mod(4), P::mod(6), m, od[7]);
    }
#pragma unroll
    for (int i = 1; i < 8; i++) {
      uint32_t *X = (i & 1) ? od : ev;
      uint32_t *Y = (i & 1) ? ev : od;
      const uint32_t m = (X[0] + Y[1]) * P::INV;
      shift_mad_row(X[0], Y, P::mod(1), P::mod(3), P::mod(5), P::mod(7), m);
      cmad_row(X, P::mod(0), P::mod(2), P::mod(4), P::mod(6), m, Y[7]);
    }
    // (od >> 32) + ev + T_hi
    asm("add.cc.u32 %0, %0, %8;\n\t"
        "addc.cc.u32 %1, %1, %9;\n\t"
        "addc.cc.u32 %2, %2, %10;\n\t"
        "addc.cc.u32 %3, %3, %11;\n\t"
        "addc.cc.u32 %4, %4, %12;\n\t"
        "addc.cc.u32 %5, %5, %13;\n\t"
        "addc.cc.u32 %6, %6, %14;\n\t"
        "addc.u32 %7, %7, 0;\n\t"
        : "+r"(ev[0]), "+r"(ev[1]), "+r"(ev[2]), "+r"(ev[3]), "+r"(ev[4]), "+r"(ev[5]), "+r"(ev[6]), "+r"(ev[7])
        : "r"(od[1]), "r"(od[2]), "r"(od[3]), "r"(od[4]), "r"(od[5]), "r"(od[6]), "r"(od[7]));
    asm("add.cc.u32 %0, %0, %8;\n\t"
        "addc.cc.u32 %1, %1, %9;\n\t"
        "addc.cc.u32 %2, %2, %10;\n\t"
        "addc.cc.u32 %3, %3, %11;\n\t"
        "addc.cc.u32 %4, %4, %12;\n\t"
        "addc.cc.u32 %5, %5, %13;\n\t"
        "addc.cc.u32 %6, %6, %14;\n\t"
        "addc.u32 %7, %7, %15;\n\t"
        : "+r"(ev[0]), "+r"(ev[1]), "+r"(ev[2]), "+r"(ev[3]), "+r"(ev[4]), "+r"(ev[5]), "+r"(ev[6]), "+r"(ev[7])
        : "r"(O[8]), "r"(O[9]), "r"(O[10]), "r"(O[11]), "r"(O[12]), "r"(O[13]), "r"(O[14]), "r"(O[15]));
    Fp res;
#pragma unroll
    for (int i = 0; i < 8; i++) res.v[i] = ev[i];
    final_sub(res.v);
    return res;
  }
#endif
  ZKB_HD Fp sqr() const {
#if defined(__CUDA_ARCH__) && !defined(ZKB_NO_FAST_SQR)
    return sqr_dev();
#else
    return *this * *this;
#endif
  }
  // call-based variants (see fp_mul_call)
  ZKB_HD Fp mulc(const Fp &o) const {
#if defined(__CUDA_ARCH__)
    return fp_mul_call<P>(*this, o);
#else
    return *this * o;
#endif
  }
  ZKB_HD Fp sqrc() const {
#if defined(__CUDA_ARCH__)
    return fp_sqr_call<P>(*this);
#else
    return *this * *this;
#endif
  }

  // normal form <-> Montgomery form
  ZKB_HD Fp to_mont() const { return *this * r2(); }
  ZKB_HD Fp from_mont() const {
    Fp o = zero();
    o.v[0] = 1;
    return *this * o;
  }

  // this^e (e: 8 limbs little endian, plain integer), square-and-multiply from the top bit
  ZKB_HD Fp pow(const uint32_t e[8]) const {
    Fp r = one();
    for (int i = 255; i >= 0; i--) {
      r = r.sqr();
      if ((e[i >> 5] >> (i & 31)) & 1) r = r * *this;
    }
    return r;
  }
  // Fermat inverse (0 -> 0)
  ZKB_HD Fp inv() const {
    uint32_t e[8];
    uint64_t br = 2;
    for (int i = 0; i < 8; i++) {
      uint64_t d = (uint64_t)P::mod(i) - br;
      e[i] = (uint32_t)d;
      br = (d >> 32) & 1;
    }
    return pow(e);
  }
};

#if defined(__CUDACC__)
template <class P> __device__ __noinline__ Fp<P> fp_mul_call(Fp<P> a, Fp<P> b) { return a * b; }
template <class P> __device__ __noinline__ Fp<P> fp_sqr_call(Fp<P> a) { return a.sqr(); }
#endif

typedef Fp<FqParams> Fq;
typedef Fp<FrParams> Fr;

}  // namespace zkb
