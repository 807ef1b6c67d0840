// Census witness program: the signal-level evaluation of `circuit/census.circom:49-115`
// (ZkFranchiseProofCircuit) and the circomlib 2.0.5 templates it instantiates, written as per-thread
// host/device code.  It reproduces what the reference's circuit.wasm computes:
// every circom signal is produced at circom's own signal number and stored through the wasm's
// witness->signal table, so the output is the 82,754-wire witness in wasm order, canonical form.
//
// One proof is split into five independent tasks (one thread each in the batched kernel); the two "chain" tasks
// carry the sequential Poseidon work, everything that does not depend on it (bit decompositions, the IsZero
// inversions, the state machine) runs beside them:
//   task 0: censusVerifier chain: hash1New, levels[160..0] (switcher + Poseidon2 per level), checkRoot
//   task 1: sik = Poseidon3(address,password,signature), then the sikVerifier chain
//   task 2: main inputs, checkWeight = LessEqThan(252), computedNullifier = Poseidon4, checkNullifier
//   task 3 / 4: censusVerifier / sikVerifier side: own inputs, Num2Bits_strict + AliasCheck, SMTLevIns (IsZero
//               inverses), SMTVerifierSM, areKeyEquals, keysOk (and the oldKey = 0 sub-circuits in dense mode)
// Signals are written to a per-proof staging array indexed by circom signal number (one 32-byte store, tagged
// "written" / "Montgomery" in the two spare top bits); witness_gather_one then builds each wire from its signal
// (or from the template when the signal was not produced), converting out of Montgomery form in parallel.
// Below a leaf's insertion level every SMT level hashes Poseidon2(0,0): those 770-signal blocks (and
// the oldKey=0 / oldValue=0 sub-circuits) are identical in every proof.  With skip_const they are not
// recomputed: the output buffer is pre-filled from a template produced once per circuit by running
// this same program with skip_const = false.  The decision is made on the DATA
// (both hash inputs zero), never on an assumed tree depth, so the witness stays bit-exact.
//
// Assert semantics follow the wasm: a failed `===` sets status 4 (exceptionHandler(4)).
#pragma once
#include "fp.cuh"

#if defined(__CUDACC__)
#define ZKB_HDN __host__ __device__ __noinline__
#else
#define ZKB_HDN inline
#endif

namespace zkb {

struct PexLayout {           // PoseidonEx(t-1, 1) block, offsets relative to its first signal
  int t, rp;
  uint32_t size;
  uint32_t ark, mix, mixlast, mixs, sigf, sigp;
  uint32_t c_off, s_off, m_off, p_off;   // offsets (in Fr elements) into the constants buffer
};

struct VerifierLayout {      // SMTVerifier(n), absolute signal numbers
  uint32_t base;             // enabled, root, siblings[n], oldKey, oldValue, isOld0, key, value, fnc
  uint32_t areKeyEquals, checkRoot, hash1New, hash1Old, keysOk, levels, n2bNew, n2bOld, sm, smtLevIns;
};

struct CensusLayout {
  uint32_t n;                // levels = nLevels + 1
  uint32_t n_signals, n_wires, n_inputs;
  PexLayout pex[3];          // t = 3, 4, 5
  uint32_t level_size;       // 789
  VerifierLayout census, sik;
  uint32_t checkNullifier, checkWeight, computedNullifier, sikHash;
  // main input signal numbers
  uint32_t electionId, nullifier, voteHash, sikRoot, censusRoot, voteWeight, availableWeight, address, password,
      signature, censusSiblings, sikSiblings;
};

static inline uint32_t pex_size(int t, int rp) { return (t + 1) + 16 * t + 14 * t + (t + 1) + 2 * t * rp + 32 * t + 4 * rp; }

// Fills everything except the constants offsets.  Returns false when n is unsupported.
static inline bool census_layout_build(CensusLayout &L, uint32_t n_levels_plus1) {
  const uint32_t n = n_levels_plus1;
  if (n < 4 || n > 254) return false;
  L.n = n;
  const int rps[3] = {57, 56, 60};
  for (int i = 0; i < 3; i++) {
    PexLayout &p = L.pex[i];
    int t = 3 + i;
    p.t = t;
    p.rp = rps[i];
    p.ark = t + 1;
    p.mix = p.ark + 16 * t;
    p.mixlast = p.mix + 14 * t;
    p.mixs = p.mixlast + (t + 1);
    p.sigf = p.mixs + 2 * t * p.rp;
    p.sigp = p.sigf + 32 * t;
    p.size = p.sigp + 4 * p.rp;
  }
  const uint32_t poseidon2 = 3 + L.pex[0].size;            // Poseidon(2): out, inputs[2], pEx
  const uint32_t poseidon3 = 4 + L.pex[1].size;
  const uint32_t poseidon4 = 5 + L.pex[2].size;
  const uint32_t hash1 = 3 + poseidon3, hash2 = 3 + poseidon2;
  L.level_size = 13 + hash2 + 6;
  const uint32_t n2b_strict = 255 + (254 + (383 + 136)) + 255;
  const uint32_t levins = 3 * n + 3 * n;
  const uint32_t verifier = (n + 8) + 6 + 6 + 2 * hash1 + 20 + n * L.level_size + 2 * n2b_strict + 15 * n + levins;
  uint32_t s = 1;
  L.electionId = s; s += 2;
  L.nullifier = s; s += 1;
  L.voteHash = s; s += 2;
  L.sikRoot = s++; L.censusRoot = s++; L.voteWeight = s++; L.availableWeight = s++;
  L.address = s++; L.password = s++; L.signature = s++;
  L.censusSiblings = s; s += n;
  L.sikSiblings = s; s += n;
  L.n_inputs = s - 1;
  auto place = [&](VerifierLayout &v) {
    v.base = s;
    uint32_t q = s + n + 8;
    v.areKeyEquals = q; q += 6;
    v.checkRoot = q; q += 6;
    v.hash1New = q; q += hash1;
    v.hash1Old = q; q += hash1;
    v.keysOk = q; q += 20;
    v.levels = q; q += n * L.level_size;
    v.n2bNew = q; q += n2b_strict;
    v.n2bOld = q; q += n2b_strict;
    v.sm = q; q += 15 * n;
    v.smtLevIns = q; q += levins;
    s += verifier;
  };
  place(L.census);
  L.checkNullifier = s; s += 6;
  L.checkWeight = s; s += 3 + 3 + 254;
  L.computedNullifier = s; s += poseidon4;
  L.sikHash = s; s += poseidon3;
  place(L.sik);
  L.n_signals = s;
  return true;
}

// ---------------------------------------------------------------------------------------------
// Env: where signals go.  consts = Poseidon tables (Montgomery form), sig2wire = inverse of the
// wasm's witness->signal table (-1 = signal eliminated), out = this proof's witness (canonical).
// ---------------------------------------------------------------------------------------------
static constexpr uint32_t TAG_WRITTEN = 0x40000000u, TAG_MONT = 0x80000000u;   // top bits of limb 7 (values < 2^254)

struct WitnessEnv {
  const CensusLayout *L;
  const Fr *consts;
  Fr *stage;          // [n_signals] this proof's staging array (zeroed before the tasks run)
  int status;
  // Lane group of a hash-chain task on the device (COOP_LANES lanes run the task's scalar code redundantly and share
  // each Poseidon permutation, see poseidon_ex_coop): li = lane within the group, glane0 = warp lane of the group's
  // lane 0, gmask = the group's shuffle mask.  coop = false (host, side tasks): one thread, everything below is
  // the plain scalar program.
  uint32_t li = 0, glane0 = 0, gmask = 0;
  bool coop = false;

  // scalar-section stores: only the group's lane 0 writes
  ZKB_HD void put_norm(uint32_t sig, const Fr &v) { if (!li) putf_norm(sig, v); }      // v already canonical
  ZKB_HD void put(uint32_t sig, const Fr &v_mont) { if (!li) putf(sig, v_mont); }      // v in Montgomery form
  ZKB_HD void put_u32(uint32_t sig, uint32_t x) { if (!li) putf_u32(sig, x); }
  // unconditional stores (the calling lane owns the value)
  ZKB_HD void putf_norm(uint32_t sig, const Fr &v) {
    Fr t = v;
    t.v[7] |= TAG_WRITTEN;
    stage[sig] = t;
  }
  ZKB_HD void putf(uint32_t sig, const Fr &v_mont) {
    Fr t = v_mont;
    t.v[7] |= TAG_WRITTEN | TAG_MONT;
    stage[sig] = t;
  }
  ZKB_HD void putf_u32(uint32_t sig, uint32_t x) {
    Fr v = Fr::zero();
    v.v[0] = x;
    v.v[7] = TAG_WRITTEN;
    stage[sig] = v;
  }
  ZKB_HD void fail() { status = 4; }
};
static constexpr uint32_t COOP_LANES = 8;

// wire value from the staged signal (canonical form), or the template's when the signal was not produced
ZKB_HD Fr witness_gather_one(const Fr *stage, uint32_t sig, const Fr *tmpl, uint32_t wire) {
  Fr x = stage[sig];
  const uint32_t tag = x.v[7];
  if (!(tag & TAG_WRITTEN)) return tmpl ? tmpl[wire] : Fr::zero();
  x.v[7] &= ~(TAG_WRITTEN | TAG_MONT);
  return (tag & TAG_MONT) ? x.from_mont() : x;
}

ZKB_HDN Fr fr_sbox(WitnessEnv &e, uint32_t sig, const Fr &x) {   // Sigma {out,in,in2,in4}
  Fr x2 = x.sqr(), x4 = x2.sqr(), y = x4 * x;
  e.put(sig + 1, x);
  e.put(sig + 2, x2);
  e.put(sig + 3, x4);
  e.put(sig, y);
  return y;
}

// PoseidonEx(T-1 inputs, 1 output), initialState = 0.  `base` = first signal of the PoseidonEx block.
// emit = false computes the hash only.
template <int T>
ZKB_HDN Fr poseidon_ex(WitnessEnv &e, uint32_t base, const Fr *inputs, bool emit) {
  const PexLayout &pl = e.L->pex[T - 3];
  const Fr *C = e.consts + pl.c_off, *S = e.consts + pl.s_off, *M = e.consts + pl.m_off, *Pm = e.consts + pl.p_off;
  const int RP = pl.rp;
  Fr st[T], tmp[T];
  st[0] = Fr::zero();
  for (int j = 1; j < T; j++) st[j] = inputs[j - 1];
  if (emit) {
    for (int j = 1; j < T; j++) e.put(base + j, inputs[j - 1]);
    e.put_u32(base + T, 0);
  }
  auto ark = [&](int k, int coff) {       // Ark k: {out[T], in[T]}
    uint32_t s = base + pl.ark + k * 2 * T;
    for (int j = 0; j < T; j++) {
      if (emit) e.put(s + T + j, st[j]);
      st[j] = st[j] + C[coff + j];
      if (emit) e.put(s + j, st[j]);
    }
  };
  auto mix = [&](int k, const Fr *Mx) {   // Mix k: out[i] = sum_j M[j][i] in[j]
    uint32_t s = base + pl.mix + k * 2 * T;
    for (int i = 0; i < T; i++) {
      Fr acc = Mx[i] * st[0];
      for (int j = 1; j < T; j++) acc = acc + Mx[j * T + i] * st[j];
      tmp[i] = acc;
    }
    for (int j = 0; j < T; j++) {
      if (emit) { e.put(s + T + j, st[j]); e.put(s + j, tmp[j]); }
      st[j] = tmp[j];
    }
  };
  auto sigma_full = [&](int r) {
    for (int j = 0; j < T; j++) {
      uint32_t s = base + pl.sigf + (r * T + j) * 4;
      if (emit) st[j] = fr_sbox(e, s, st[j]);
      else { Fr x2 = st[j].sqr(); st[j] = x2.sqr() * st[j]; }
    }
  };
  ark(0, 0);
  for (int r = 0; r < 3; r++) {
    sigma_full(r);
    ark(r + 1, (r + 1) * T);
    mix(r, M);
  }
  sigma_full(3);
  ark(4, 4 * T);
  mix(3, Pm);
  for (int r = 0; r < RP; r++) {
    uint32_t ss = base + pl.sigp + r * 4, ms = base + pl.mixs + r * 2 * T;
    Fr y;
    if (emit) y = fr_sbox(e, ss, st[0]);
    else { Fr x2 = st[0].sqr(); y = x2.sqr() * st[0]; }
    st[0] = y + C[5 * T + r];
    const Fr *Sr = S + (2 * T - 1) * r;
    Fr o0 = Sr[0] * st[0];
    for (int i = 1; i < T; i++) o0 = o0 + Sr[i] * st[i];
    if (emit) {
      for (int j = 0; j < T; j++) e.put(ms + T + j, st[j]);
      e.put(ms, o0);
    }
    for (int i = 1; i < T; i++) {
      st[i] = st[i] + st[0] * Sr[T + i - 1];
      if (emit) e.put(ms + i, st[i]);
    }
    st[0] = o0;
  }
  for (int r = 0; r < 3; r++) {
    sigma_full(4 + r);
    ark(5 + r, 5 * T + RP + r * T);
    mix(4 + r, M);
  }
  sigma_full(7);
  Fr o = M[0] * st[0];
  for (int j = 1; j < T; j++) o = o + M[j * T] * st[j];
  if (emit) {
    uint32_t s = base + pl.mixlast;
    for (int j = 0; j < T; j++) e.put(s + 1 + j, st[j]);
    e.put(s, o);
    e.put(base, o);
  }
  return o;
}

#if defined(__CUDA_ARCH__)
// The same permutation shared by the COOP_LANES = 8 lanes of a group (device only; T = 3 or 4, 2T - 1 <= 8).  A
// single thread runs a permutation as one dependent chain of ~600 / 770 products; here lane j < T owns state
// element j, the S-boxes of a full round run side by side, and in a partial round the 2T - 1 products of MixS
// (T for out[0], T - 1 for the other outputs) are one product per lane, so a partial round costs 3 + 1 product
// times instead of 3 + (2T - 1): 276 / 280 instead of 594 / 772 product times per permutation.  Every lane executes
// the same instruction stream (no lane-dependent branch around a product or a shuffle); signals are stored by the
// lane that owns them.  Returns the hash on every lane.
__device__ __forceinline__ Fr shfl_fr(uint32_t mask, const Fr &x, uint32_t src) {
  Fr r;
#pragma unroll
  for (int l = 0; l < 8; l++) r.v[l] = __shfl_sync(mask, x.v[l], src);
  return r;
}
__device__ __noinline__ Fr fr_sbox_coop(WitnessEnv &e, uint32_t sig, const Fr &x, bool store) {
  Fr x2 = x.sqr(), x4 = x2.sqr(), y = x4 * x;
  if (store) {
    e.putf(sig + 1, x);
    e.putf(sig + 2, x2);
    e.putf(sig + 3, x4);
    e.putf(sig, y);
  }
  return y;
}
template <int T>
__device__ __noinline__ Fr poseidon_ex_coop(WitnessEnv &e, uint32_t base, const Fr *inputs, bool emit) {
  static_assert(2 * T - 1 <= (int)COOP_LANES, "group too narrow");
  const PexLayout &pl = e.L->pex[T - 3];
  const Fr *C = e.consts + pl.c_off, *S = e.consts + pl.s_off, *M = e.consts + pl.m_off, *Pm = e.consts + pl.p_off;
  const int RP = pl.rp;
  const uint32_t li = e.li, m = e.gmask, l0 = e.glane0;
  const bool own = li < (uint32_t)T, wr = own && emit;
  const uint32_t lj = own ? li : 0;                       // clamped element index for table reads
  Fr s = Fr::zero();
  if (li >= 1 && own) s = inputs[li - 1];
  if (emit && li == 0) {
    for (int j = 1; j < T; j++) e.putf(base + j, inputs[j - 1]);
    e.putf_u32(base + T, 0);
  }
  auto ark = [&](int k, int coff) {       // Ark k: {out[T], in[T]}
    uint32_t sg = base + pl.ark + k * 2 * T;
    if (wr) e.putf(sg + T + li, s);
    s = s + C[coff + lj];
    if (wr) e.putf(sg + li, s);
  };
  auto mix = [&](int k, const Fr *Mx) {   // Mix k: out[i] = sum_j M[j][i] in[j]; lane i forms out[i]
    uint32_t sg = base + pl.mix + k * 2 * T;
    Fr acc = Fr::zero();
#pragma unroll
    for (int j = 0; j < T; j++) acc = acc + Mx[j * T + lj] * shfl_fr(m, s, l0 + j);
    if (wr) { e.putf(sg + T + li, s); e.putf(sg + li, acc); }
    s = acc;
  };
  auto sigma_full = [&](int r) { s = fr_sbox_coop(e, base + pl.sigf + (r * T + lj) * 4, s, wr); };
  ark(0, 0);
  for (int r = 0; r < 3; r++) {
    sigma_full(r);
    ark(r + 1, (r + 1) * T);
    mix(r, M);
  }
  sigma_full(3);
  ark(4, 4 * T);
  mix(3, Pm);
  // Partial rounds in THREE product times.  With x = st[0], c = C[5T + r] and the sparse row S:
  //   out[0] = S[0] (x^5 + c) + sum_{j>=1} S[j] st[j],      out[i] = st[i] + S[T+i-1] (x^5 + c).
  // Lane 0 walks the S-box chain x^2, x^4, x^5 (its signals are needed); beside it, in the same three product slots,
  // lanes 1..T-1 form S[j] st[j], and T helper lanes (lane T for out[0], lane T+i for out[i]) form S[.] x, S[.] c and
  // finally x^4 (S[.] x) = S[.] x^5, so that everything is ready when the S-box is.  All exact field identities.
  static_assert(2 * T <= (int)COOP_LANES, "helper lanes");
  for (int r = 0; r < RP; r++) {
    const uint32_t ss = base + pl.sigp + r * 4, ms = base + pl.mixs + r * 2 * T;
    const Fr *Sr = S + (2 * T - 1) * r;
    const Fr c = C[5 * T + r];
    const bool helper = li >= (uint32_t)T && li < (uint32_t)(2 * T);
    // the S coefficient this lane works with: lanes j < T: S[j]; helper lane T: S[0]; helper lane T+i: S[T+i-1]
    const uint32_t ci = li < (uint32_t)T ? li : (li == (uint32_t)T ? 0u : (helper ? li - 1 : 0u));
    const Fr coef = Sr[ci];
    const Fr x = shfl_fr(m, s, l0);                       // st[0] before the S-box
    // slot 1: lane 0: x^2; lanes 1..T-1: S[j] st[j]; helpers: S[.] x
    const Fr p1 = (li == 0 ? x : coef) * (li == 0 || helper ? x : s);
    // slot 2: lane 0: x^4; helpers: S[.] c
    const Fr p2 = (li == 0 ? p1 : coef) * (li == 0 ? p1 : c);
    const Fr x4 = shfl_fr(m, p2, l0);
    // slot 3: lane 0: x^5; helpers: x^4 (S[.] x) = S[.] x^5
    const Fr p3 = x4 * (li == 0 ? x : p1);
    if (emit && li == 0) {                                // Sigma {out, in, in2, in4}
      e.putf(ss + 1, x);
      e.putf(ss + 2, p1);
      e.putf(ss + 3, p2);
      e.putf(ss, p3);
    }
    const Fr hsum = p3 + p2;                              // helpers: S[.] (x^5 + c)
    // out[0] on lane 0: helper lane T's term + the S[j] st[j] of lanes 1..T-1
    Fr o0 = shfl_fr(m, hsum, l0 + T);
#pragma unroll
    for (int j = 1; j < T; j++) o0 = o0 + shfl_fr(m, p1, l0 + j);
    const Fr q = shfl_fr(m, hsum, l0 + ((li >= 1 && own) ? (uint32_t)T + li : (uint32_t)T));
    if (li == 0) {
      const Fr in0 = p3 + c;                              // st[0] after the S-box and the round constant
      if (emit) { e.putf(ms + T, in0); e.putf(ms, o0); }
      s = o0;
    } else if (own) {
      if (emit) e.putf(ms + T + li, s);
      s = s + q;
      if (emit) e.putf(ms + li, s);
    }
  }
  for (int r = 0; r < 3; r++) {
    sigma_full(4 + r);
    ark(5 + r, 5 * T + RP + r * T);
    mix(4 + r, M);
  }
  sigma_full(7);
  const Fr prod = M[lj * T] * s;
  Fr o = shfl_fr(m, prod, l0);
#pragma unroll
  for (int j = 1; j < T; j++) o = o + shfl_fr(m, prod, l0 + j);
  if (emit) {
    uint32_t sg = base + pl.mixlast;
    if (own) e.putf(sg + 1 + li, s);
    if (li == 0) { e.putf(sg, o); e.putf(base, o); }
  }
  return o;
}
#endif

// Poseidon(T-1) = {out, inputs[T-1]} + pEx
template <int T>
ZKB_HD Fr poseidon_comp(WitnessEnv &e, uint32_t base, const Fr *inputs, bool emit) {
#if defined(__CUDA_ARCH__)
  if constexpr (2 * T - 1 <= (int)COOP_LANES) {
    if (e.coop) {
      Fr oc = poseidon_ex_coop<T>(e, base + T, inputs, emit);
      if (emit) {
        e.put(base, oc);
        for (int j = 0; j < T - 1; j++) e.put(base + 1 + j, inputs[j]);
      }
      return oc;
    }
  }
#endif
  Fr o = poseidon_ex<T>(e, base + T, inputs, emit);
  if (emit) {
    e.put(base, o);
    for (int j = 0; j < T - 1; j++) e.put(base + 1 + j, inputs[j]);
  }
  return o;
}

// canonical 256-bit helpers ---------------------------------------------------------------------
ZKB_HD uint32_t bit_of(const Fr &x, int i) { return (x.v[i >> 5] >> (i & 31)) & 1u; }

// IsZero {out,in,inv}; x in Montgomery form.  Returns out (0/1).
ZKB_HDN uint32_t is_zero_comp(WitnessEnv &e, uint32_t sig, const Fr &x, bool emit = true) {
  uint32_t z = x.is_zero() ? 1u : 0u;
  if (!emit) return z;
  e.put_u32(sig, z);
  e.put(sig + 1, x);
  if (z) e.put_u32(sig + 2, 0);
  else e.put(sig + 2, x.inv());
  return z;
}

// Num2Bits(nbits) {out[nbits], in}: x canonical.  Returns false when x does not fit (assert).
ZKB_HD bool num2bits_comp(WitnessEnv &e, uint32_t sig, const Fr &x, int nbits) {
  for (int i = 0; i < nbits; i++) e.put_u32(sig + i, bit_of(x, i));
  e.put_norm(sig + nbits, x);
  for (int i = nbits; i < 256; i++)
    if (bit_of(x, i)) return false;
  return true;
}

// Num2Bits_strict {out[254], in} + aliasCheck {in[254]} + compConstant(-1) + n2b(254).  key canonical.
ZKB_HDN void num2bits_strict_comp(WitnessEnv &e, uint32_t sig, const Fr &key) {
  const uint32_t alias = sig + 255, cc = alias + 254, n2b = sig + 255 + 773;
  for (int i = 0; i < 254; i++) {
    uint32_t b = bit_of(key, i);
    e.put_u32(sig + i, b);
    e.put_u32(alias + i, b);
    e.put_u32(cc + 1 + i, b);
    e.put_u32(n2b + i, b);
  }
  e.put_norm(sig + 254, key);
  e.put_norm(n2b + 254, key);
  // CompConstant(ct = r - 1): parts[i] from bit pairs; a = 2^i, b = 2^128 - 2^i; values as plain integers
  uint32_t sout[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 127; i++) {
    uint32_t clsb = (FrParams::mod((2 * i) >> 5) >> ((2 * i) & 31)) & 1u;       // bits of r - 1: r is odd, so
    uint32_t cmsb = (FrParams::mod((2 * i + 1) >> 5) >> ((2 * i + 1) & 31)) & 1u;
    if (i == 0) clsb = 0;                                                        // (r-1) clears bit 0 only
    uint32_t sl = bit_of(key, 2 * i), sm = bit_of(key, 2 * i + 1);
    int kind;  // 0: zero, 1: a, 2: b
    if (!cmsb && !clsb) kind = (sl | sm) ? 2 : 0;
    else if (!cmsb && clsb) kind = sm ? 2 : (sl ? 0 : 1);
    else if (cmsb && !clsb) kind = sm ? (sl ? 2 : 0) : 1;
    else kind = (sm & sl) ? 0 : 1;
    uint32_t part[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (kind == 1) part[i >> 5] = 1u << (i & 31);
    else if (kind == 2) {
      // 2^128 - 2^i : bits i..127 set
      for (int w = 0; w < 4; w++) {
        uint32_t m = 0xffffffffu;
        int lo = w * 32;
        if (i >= lo + 32) m = 0;
        else if (i > lo) m = 0xffffffffu << (i - lo);
        part[w] = m;
      }
    }
    Fr pv;
    for (int w = 0; w < 8; w++) pv.v[w] = part[w];
    e.put_norm(cc + 255 + i, pv);
    uint64_t c = 0;
    for (int w = 0; w < 8; w++) {
      c += (uint64_t)sout[w] + part[w];
      sout[w] = (uint32_t)c;
      c >>= 32;
    }
  }
  Fr sv;
  for (int w = 0; w < 8; w++) sv.v[w] = sout[w];
  e.put_norm(cc + 382, sv);
  if (!num2bits_comp(e, cc + 383, sv, 135)) e.fail();
  uint32_t outbit = bit_of(sv, 127);
  e.put_u32(cc, outbit);
  if (outbit) e.fail();           // AliasCheck: compConstant.out === 0
}

// SMTHash1 {out,key,value} + Poseidon(3); inputs Montgomery
ZKB_HD Fr smt_hash1(WitnessEnv &e, uint32_t sig, const Fr &key, const Fr &value, bool emit) {
  Fr in[3] = {key, value, Fr::one()};
  Fr o = poseidon_comp<4>(e, sig + 3, in, emit);
  if (emit) { e.put(sig, o); e.put(sig + 1, key); e.put(sig + 2, value); }
  return o;
}

// SMTVerifier(n) with enabled = 1, fnc = 0, oldKey = oldValue = isOld0 = 0 (census.circom:79-103).
// key_n/value_n/root_n/siblings canonical.  h00 = Poseidon2(0,0), h001 = Poseidon3(0,0,1) (Montgomery).
// chain = the part that carries the sequential hashing (hash1New, levels, checkRoot, the `value` input);
// side = everything else.  Both recompute the cheap 0/1 bookkeeping (zero flags, levIns, state machine) privately.
ZKB_HDN void smt_verifier(WitnessEnv &e, const VerifierLayout &V, const Fr &key_n, const Fr &value_n, const Fr &root_n,
                         const Fr *siblings_n, const Fr &h00, const Fr &h001, bool skip_const, bool chain, bool side) {
  const uint32_t n = e.L->n;
  const uint32_t b = V.base;
  const Fr key = key_n.to_mont(), root = root_n.to_mont();
  Fr old1leaf = h001, new1leaf = Fr::zero();
  if (side) {
    e.put_u32(b, 1);
    e.put_norm(b + 1, root_n);
    for (uint32_t i = 0; i < n; i++) e.put_norm(b + 2 + i, siblings_n[i]);
    e.put_u32(b + n + 2, 0);   // oldKey
    e.put_u32(b + n + 3, 0);   // oldValue
    e.put_u32(b + n + 4, 0);   // isOld0
    e.put_norm(b + n + 5, key_n);
    e.put_u32(b + n + 7, 0);   // fnc
    // hash1Old = SMTHash1(0, 0), n2bOld = Num2Bits_strict(0): constant
    if (!skip_const) {
      smt_hash1(e, V.hash1Old, Fr::zero(), Fr::zero(), true);
      num2bits_strict_comp(e, V.n2bOld, Fr::zero());
    }
    num2bits_strict_comp(e, V.n2bNew, key_n);
  }
  if (chain) {
    e.put_norm(b + n + 6, value_n);
    new1leaf = smt_hash1(e, V.hash1New, key, value_n.to_mont(), true);
  }
  // smtLevIns {levIns[n], enabled, siblings[n], done[n-1]} + isZero[n]
  const uint32_t li = V.smtLevIns;
  if (side) e.put_u32(li + n, 1);
  // bitmask of zero siblings (n <= 254)
  uint32_t zmask[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (uint32_t i = 0; i < n; i++) zmask[i >> 5] |= (siblings_n[i].is_zero() ? 1u : 0u) << (i & 31);
  if (side) {
    // isZero[i] = IsZero {out, in, inv} for every sibling.  The inverses of all non-zero siblings come from ONE field
    // inversion (Montgomery's trick): the running product before sibling i is parked in its `inv` slot on the way
    // up and replaced by the inverse on the way down.  (One Fermat inversion per sibling was 0.11 ms each: 1.4 ms
    // of this task for a 13-level path, 17 ms for a 160-level one.)
    Fr run = Fr::one();
    for (uint32_t i = 0; i < n; i++) {
      const uint32_t sg = li + 3 * n + 3 * i, z = (zmask[i >> 5] >> (i & 31)) & 1u;
      e.put_norm(li + n + 1 + i, siblings_n[i]);
      e.put_u32(sg, z);
      if (z) { e.put_u32(sg + 1, 0); e.put_u32(sg + 2, 0); continue; }
      const Fr sm = siblings_n[i].to_mont();
      e.put(sg + 1, sm);
      e.put(sg + 2, run);
      run = run * sm;
    }
    Fr inv = run.inv();
    for (int i = (int)n - 1; i >= 0; i--) {
      if ((zmask[i >> 5] >> (i & 31)) & 1u) continue;
      const uint32_t sg = li + 3 * n + 3 * i;
      Fr pre = e.stage[sg + 2];
      pre.v[7] &= ~(TAG_WRITTEN | TAG_MONT);
      e.put(sg + 2, inv * pre);
      inv = inv * siblings_n[i].to_mont();
    }
  }
  auto isz = [&](uint32_t i) { return (zmask[i >> 5] >> (i & 31)) & 1u; };
  if (side && !isz(n - 1)) e.fail();                          // (isZero[n-1].out - 1) * enabled === 0
  // levIns / done (all 0/1); levmask bit i = levIns[i]
  uint32_t levmask[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  {
    uint32_t lev = 1u - isz(n - 2);
    if (side) e.put_u32(li + n - 1, lev);
    levmask[(n - 1) >> 5] |= lev << ((n - 1) & 31);
    uint32_t done = lev;
    if (side) e.put_u32(li + 2 * n + 1 + (n - 2), done);
    for (uint32_t i = n - 2; i > 0; i--) {
      lev = (1u - done) * (1u - isz(i - 1));
      if (side) e.put_u32(li + i, lev);
      levmask[i >> 5] |= lev << (i & 31);
      done = lev + done;
      if (side) e.put_u32(li + 2 * n + 1 + (i - 1), done);
    }
    lev = 1u - done;
    if (side) e.put_u32(li, lev);
    levmask[0] |= lev;
  }
  // state machines sm[i] (is0 = 0, fnc = 0)
  uint32_t topmask[8] = {0, 0, 0, 0, 0, 0, 0, 0}, inewmask[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  uint32_t namask[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  {
    uint32_t p_top = 1, p_i0 = 0, p_iold = 0, p_inew = 0, p_na = 0;
    for (uint32_t i = 0; i < n; i++) {
      uint32_t s = V.sm + 15 * i;
      uint32_t lev = (levmask[i >> 5] >> (i & 31)) & 1u;
      uint32_t ptl = p_top * lev;
      uint32_t st_top = p_top - ptl, st_inew = ptl, st_iold = 0, st_i0 = 0;
      uint32_t st_na = p_na + p_inew + p_iold + p_i0;
      if (side) {
        e.put_u32(s + 0, st_top); e.put_u32(s + 1, st_i0); e.put_u32(s + 2, st_iold); e.put_u32(s + 3, st_inew);
        e.put_u32(s + 4, st_na); e.put_u32(s + 5, 0); e.put_u32(s + 6, lev); e.put_u32(s + 7, 0);
        e.put_u32(s + 8, p_top); e.put_u32(s + 9, p_i0); e.put_u32(s + 10, p_iold); e.put_u32(s + 11, p_inew);
        e.put_u32(s + 12, p_na); e.put_u32(s + 13, ptl); e.put_u32(s + 14, 0);
        if (i == n - 1 && st_na + st_iold + st_inew + st_i0 != 1) e.fail();
      }
      topmask[i >> 5] |= st_top << (i & 31);
      inewmask[i >> 5] |= st_inew << (i & 31);
      namask[i >> 5] |= (st_na & 1u) << (i & 31);
      p_top = st_top; p_i0 = st_i0; p_iold = st_iold; p_inew = st_inew; p_na = st_na;
    }
  }
  // levels n-1 .. 0
  Fr child = Fr::zero();
  for (int i = (int)n - 1; chain && i >= 0; i--) {
    const uint32_t s = V.levels + (uint32_t)i * e.L->level_size;
    const uint32_t st_top = (topmask[i >> 5] >> (i & 31)) & 1u, st_inew = (inewmask[i >> 5] >> (i & 31)) & 1u;
    const uint32_t st_na = (namask[i >> 5] >> (i & 31)) & 1u;
    const uint32_t lrbit = bit_of(key_n, i);
    const Fr sib = siblings_n[i].to_mont();
    e.put_u32(s + 1, st_top); e.put_u32(s + 2, 0); e.put_u32(s + 3, 0); e.put_u32(s + 4, st_inew);
    e.put_u32(s + 5, st_na);
    e.put_norm(s + 6, siblings_n[i]);
    e.put(s + 7, old1leaf);
    e.put(s + 8, new1leaf);
    e.put_u32(s + 9, lrbit);
    e.put(s + 10, child);
    // switcher {outL,outR,sel,L,R,aux}
    const uint32_t sw = s + 13 + (3 + 3 + e.L->pex[0].size);
    Fr aux = lrbit ? (sib - child) : Fr::zero();
    Fr outL = aux + child, outR = sib - aux;
    e.put(sw, outL); e.put(sw + 1, outR); e.put_u32(sw + 2, lrbit); e.put(sw + 3, child); e.put(sw + 4, sib);
    e.put(sw + 5, aux);
    // proofHash = SMTHash2 {out,L,R} + Poseidon(2)
    Fr h;
    const uint32_t ph = s + 13;
    if (skip_const && outL.is_zero() && outR.is_zero()) {
      h = h00;                      // block left as pre-filled from the template
    } else {
      Fr in[2] = {outL, outR};
      h = poseidon_comp<3>(e, ph + 3, in, true);
      e.put(ph, h); e.put(ph + 1, outL); e.put(ph + 2, outR);
    }
    Fr aux0 = st_top ? h : Fr::zero();
    e.put(s + 11, aux0);
    e.put_u32(s + 12, 0);           // aux[1] = old1leaf * st_iold, st_iold = 0
    Fr rt = st_inew ? aux0 + new1leaf : aux0;
    e.put(s, rt);
    child = rt;
  }
  // areKeyEquals = IsEqual(oldKey = 0, key) {out,in[2]} + isz
  if (side) {
    uint32_t s = V.areKeyEquals;
    e.put_u32(s + 1, 0);
    e.put_norm(s + 2, key_n);
    uint32_t z = is_zero_comp(e, s + 3, key);
    e.put_u32(s, z);
    // keysOk = MultiAND(4)(fnc=0, 1-isOld0=1, areKeyEquals.out, enabled=1)
    uint32_t k = V.keysOk;
    e.put_u32(k, 0);
    e.put_u32(k + 1, 0); e.put_u32(k + 2, 1); e.put_u32(k + 3, z); e.put_u32(k + 4, 1);
    e.put_u32(k + 5, 0); e.put_u32(k + 6, 0); e.put_u32(k + 7, z);              // and2 {out,a,b}
    e.put_u32(k + 8, 0); e.put_u32(k + 9, 0); e.put_u32(k + 10, 1);             // ands[0] {out,in[2]}
    e.put_u32(k + 11, 0); e.put_u32(k + 12, 0); e.put_u32(k + 13, 1);           //   and1 {out,a,b}
    e.put_u32(k + 14, z); e.put_u32(k + 15, z); e.put_u32(k + 16, 1);           // ands[1]
    e.put_u32(k + 17, z); e.put_u32(k + 18, z); e.put_u32(k + 19, 1);           //   and1
  }
  // checkRoot = ForceEqualIfEnabled(1, levels[0].root, root)
  if (chain) {
    uint32_t s = V.checkRoot;
    e.put_u32(s, 1);
    e.put(s + 1, child);
    e.put_norm(s + 2, root_n);
    Fr d = root - child;
    uint32_t z = is_zero_comp(e, s + 3, d);
    if (!z) e.fail();
  }
}

// Task 2: main inputs, checkWeight, computedNullifier, checkNullifier.  in = the 2n+12 canonical inputs in
// main-signal order (signal 1 + k).
ZKB_HDN void census_main_task(WitnessEnv &e, const Fr *in) {
  const CensusLayout &L = *e.L;
  e.put_u32(0, 1);
  for (uint32_t k = 0; k < L.n_inputs; k++) e.put_norm(1 + k, in[k]);
  const Fr &voteWeight = in[L.voteWeight - 1], &availableWeight = in[L.availableWeight - 1];
  // checkWeight = LessEqThan(252) {out,in[2]} + lt = LessThan(252) {out,in[2]} + n2b = Num2Bits(253)
  {
    uint32_t s = L.checkWeight, lt = s + 3, n2b = lt + 3;
    Fr one = Fr::zero();
    one.v[0] = 1;
    Fr in1p = availableWeight + one;                 // canonical arithmetic: + and - are form-agnostic
    Fr two252 = Fr::zero();
    two252.v[7] = 1u << 28;
    Fr x = voteWeight + two252 - in1p;
    e.put_norm(s + 1, voteWeight); e.put_norm(s + 2, availableWeight);
    e.put_norm(lt + 1, voteWeight); e.put_norm(lt + 2, in1p);
    if (!num2bits_comp(e, n2b, x, 253)) e.fail();
    uint32_t o = 1u - bit_of(x, 252);
    e.put_u32(lt, o);
    e.put_u32(s, o);
    if (o != 1) e.fail();                            // checkWeight.out === 1
  }
  Fr hin[4] = {in[L.signature - 1].to_mont(), in[L.password - 1].to_mont(), in[L.electionId - 1].to_mont(),
               in[L.electionId].to_mont()};
  Fr nul = poseidon_comp<5>(e, L.computedNullifier, hin, true);
  {
    uint32_t s = L.checkNullifier;
    Fr given = in[L.nullifier - 1].to_mont();
    e.put_u32(s, 1);
    e.put(s + 1, nul);
    e.put_norm(s + 2, in[L.nullifier - 1]);
    uint32_t z = is_zero_comp(e, s + 3, given - nul);
    if (!z) e.fail();
  }
}

// which: 0 = census tree, 1 = SIK tree; chain / side: see smt_verifier
ZKB_HD void census_tree_task(WitnessEnv &e, int which, bool chain, bool side, const Fr *in, const Fr &h00,
                             const Fr &h001, bool skip_const) {
  const CensusLayout &L = *e.L;
  const Fr &address = in[L.address - 1];
  if (which == 0) {
    smt_verifier(e, L.census, address, in[L.availableWeight - 1], in[L.censusRoot - 1], in + (L.censusSiblings - 1),
                 h00, h001, skip_const, chain, side);
  } else {
    Fr sik = Fr::zero();
    if (chain) {
      Fr hin[3] = {address.to_mont(), in[L.password - 1].to_mont(), in[L.signature - 1].to_mont()};
      sik = poseidon_comp<4>(e, L.sikHash, hin, true).from_mont();
    }
    smt_verifier(e, L.sik, address, sik, in[L.sikRoot - 1], in + (L.sikSiblings - 1), h00, h001, skip_const, chain,
                 side);
  }
}

// task ids of the batched kernel
static constexpr int WITNESS_TASKS = 5;
ZKB_HD void census_witness_task(WitnessEnv &e, int task, const Fr *in, const Fr &h00, const Fr &h001, bool skip_const) {
  if (task == 2) census_main_task(e, in);
  else if (task < 2) census_tree_task(e, task, true, false, in, h00, h001, skip_const);
  else census_tree_task(e, task - 3, false, true, in, h00, h001, skip_const);
}

}  // namespace zkb
