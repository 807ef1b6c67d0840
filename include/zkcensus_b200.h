/* libzkcensus_b200 - C ABI of the B200-native census Groth16 prover.
 *
 * Drop-in boundary for the proving path of vocdoni/zk-franchise-proof-circuit: the native library a
 * cgo / N-API / ctypes binding loads in place of go-rapidsnark's static libs (Go) or snarkjs' wasm
 * engine (JS).  Plain C: pointers and sizes only, caller-owned buffers, no C++ or torch types.
 * Citations are paths in the reference repository.
 *
 * Ownership: every input pointer is borrowed for the duration of the call; whatever must persist
 * (key tables, witness constants) is copied to the device inside zkb_load_circuit.  Outputs go to
 * caller-allocated buffers with in/out sizes, as in rapidsnark's prover.h.
 * Threading: calls may arrive on arbitrary OS threads (cgo); each call binds its device itself and
 * calls on one circuit handle are serialised internally.
 * There is NO CPU fallback: without an sm_100 GPU every compute entry point returns ZKB_ERROR.
 */
#ifndef ZKCENSUS_B200_H
#define ZKCENSUS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Status codes (0/1/2 coincide with rapidsnark's PROVER_OK / PROVER_ERROR / PROVER_ERROR_SHORT_BUFFER;
 * 4 is the circom wasm's exceptionHandler(4) "constraint assert failed"). */
#define ZKB_OK 0
#define ZKB_ERROR 1
#define ZKB_SHORT_BUFFER 2
#define ZKB_INVALID_WITNESS_LENGTH 3
#define ZKB_ASSERT_FAILED 4
#define ZKB_UNSUPPORTED_CIRCUIT 5
#define ZKB_INVALID_PROOF 6

typedef struct zkb_ctx zkb_ctx;         /* one GPU + its stream */
typedef struct zkb_circuit zkb_circuit; /* a proving key (+ witness calculator) resident on that GPU */

/* NUL-terminated description of the last error on the calling thread. */
const char *zkb_last_error(void);
int zkb_device_count(void);

/* One context per GPU; multi-GPU = one process (or one context) per device, proofs sharded by the caller. */
int zkb_ctx_create(int device, zkb_ctx **out);
void zkb_ctx_destroy(zkb_ctx *ctx);
void *zkb_ctx_stream(zkb_ctx *ctx); /* cudaStream_t the pipeline runs on (for external CUDA-event timing) */

/* Parses the snarkjs .zkey and the circom .wasm exactly as the reference reads them from
 * artifacts/<name>/<env>/<nLevels>/{proving_key.zkey,circuit.wasm} (zk_census_test.go:81-84) and makes the
 * key device-resident (fixed-base MSM tables, CSR coefficients, NTT twiddles, Poseidon constants,
 * witness template).  wasm may be NULL: then only zkb_prove_wtns / groth16_prover are available.
 * Witness generation: a wasm recognised as census.circom runs on the hand-written census kernel; any other circom-2
 * witness calculator is executed once, symbolically, at load time and its field operations become a straight-line
 * program the GPU evaluates per proof.  Returns ZKB_UNSUPPORTED_CIRCUIT only for programs outside
 * that extractor's subset (control flow or array indexing that depends on a signal value); zkb_last_error says which. */
int zkb_load_circuit(zkb_ctx *ctx, const void *zkey, size_t zkey_len, const void *wasm, size_t wasm_len,
                     zkb_circuit **out);
/* flags bit 0 (ZKB_LOAD_DENSE): switch the proof-independent-wire shortcut off - every SMT level is
 * hashed and the four witness MSMs run over all nVars wires, i.e. the work snarkjs / rapidsnark do.  Measurement aid:
 * bench.py reports this point next to the default so the GPU speed-up can be separated from the algorithmic one. */
#define ZKB_LOAD_DENSE 1u
/* flags bit 1 (ZKB_LOAD_GENERIC_WITNESS): use the generic (wasm-extracted) witness program even for the census wasm -
 * the parity test of the generic path against the reference's own witness calculator. */
#define ZKB_LOAD_GENERIC_WITNESS 2u
int zkb_load_circuit_ex(zkb_ctx *ctx, const void *zkey, size_t zkey_len, const void *wasm, size_t wasm_len,
                        uint32_t flags, zkb_circuit **out);
void zkb_circuit_destroy(zkb_circuit *c);
/* info[8] = nVars, nPublic, domainSize, nInputs, nLevels+1, nSignals, chunk, resident capacity */
int zkb_circuit_info(zkb_circuit *c, uint32_t *info);

/* Test hook: pin the Groth16 blinding scalars r, s (canonical little-endian, < r) so that pi_a/pi_b/pi_c
 * can be compared bit-for-bit with snarkjs groth16.prove on the same inputs.  NULL, NULL = random again. */
int zkb_set_blinding(zkb_circuit *c, const uint8_t *r32, const uint8_t *s32);

/* --- reference-shaped entry points ---------------------------------------------------------------- */

/* prover.Prove(zkey, wasm, inputs) (zk_census_test.go:89) / groth16.fullProve (ts_inputs/src/example.ts:358):
 * inputs.json in, proof.json + public.json out (compact JSON, the bytes (*Proof).Bytes() returns,
 * zk_census_test.go:93-100).  *proof_size / *public_size: capacity in, bytes written (or needed) out.
 * A failed circuit assert returns ZKB_ASSERT_FAILED (the wasm's exception code 4). */
int zkb_fullprove(zkb_circuit *c, const char *inputs_json, size_t inputs_len, char *proof_buf, size_t *proof_size,
                  char *public_buf, size_t *public_size, char *err, size_t errmax);

/* n independent proofs in one call.  Outputs are NUL-terminated strings at proofs + i*proof_stride and
 * publics + i*public_stride (1024 bytes hold a proof.json; public.json needs up to 96 bytes per public signal + 64:
 * a short stride gives status ZKB_SHORT_BUFFER for that proof); status[i] per proof, the batch continues past failures. */
int zkb_fullprove_batch(zkb_circuit *c, int n, const char *const *inputs_json, const size_t *inputs_len, char *proofs,
                        size_t proof_stride, char *publics, size_t public_stride, int *status);

/* The witness step alone: .wtns bytes for one inputs.json (what go-rapidsnark/witness CalculateWTNSBin and
 * circom_runtime calculateWTNSBin return when they run circuit.wasm). */
int zkb_witness(zkb_circuit *c, const char *inputs_json, size_t inputs_len, void *wtns_out, size_t *wtns_len);

/* The Groth16 step alone from a caller-supplied .wtns (go-rapidsnark prover.Groth16ProverRaw). */
int zkb_prove_wtns(zkb_circuit *c, const void *wtns, size_t wtns_size, char *proof_buf, size_t *proof_size,
                   char *public_buf, size_t *public_size);

/* same, plus the device stage times of the pass (8 floats as zkb_batch_prove_resident; measurement aid).
 * wtns == NULL: prove again from the witness the previous call left on the device (device-resident timing). */
int zkb_prove_wtns_stages(zkb_circuit *c, const void *wtns, size_t wtns_size, char *proof_buf, size_t *proof_size,
                          char *public_buf, size_t *public_size, float *stage_ms);

/* A proving key sharded over nranks GPUs (BASELINE.json configs[3]: one large proof, MSMs split by point range):
 * rank `rank` keeps the window tables of its 2^17-point ranges only, every rank is given the same .wtns and computes
 * the H scalars itself (the three coset transforms are ~6 ms at 2^22), ranks > 0 write their five partial sums into
 * rank 0's exchange buffer over NVLink (zkb_shard_export on rank 0 -> zkb_shard_attach on the others), rank 0 adds
 * them and assembles the proof.  zkb_prove_wtns(_stages) is the call on every rank; ranks > 0 return empty strings.
 * The caller places a barrier between two proofs (a slot holds one epoch's partial sums). */
int zkb_load_circuit_shard(zkb_ctx *ctx, const void *zkey, size_t zkey_len, int rank, int nranks, zkb_circuit **out);
int zkb_shard_export(zkb_circuit *c, void *handle64);
int zkb_shard_attach(zkb_circuit *c, const void *handle64);
int zkb_shard_attach_local(zkb_circuit *c, zkb_circuit *root);
/* Witness-slice exchange of a sharded key (optional; without it every rank uploads the whole .wtns): every rank exports
 * the buffer that holds ITS slice of the witness (values [nVars r / G, nVars (r+1) / G)) and attaches the others'.
 * Once all G - 1 peers are attached, zkb_prove_wtns on rank r checks and uploads only that slice and gathers the rest
 * from the peers over NVLink (P2P loads through the IPC mappings, flag-synchronised). */
int zkb_shard_export_witness(zkb_circuit *c, void *handle64);
int zkb_shard_attach_witness(zkb_circuit *c, int peer_rank, const void *handle64);
int zkb_shard_attach_witness_local(zkb_circuit *c, int peer_rank, zkb_circuit *peer);

/* rapidsnark prover.h, same symbol and signature, so go-rapidsnark's cgo wrapper links unchanged
 * (go.mod:30 github.com/iden3/go-rapidsnark/prover v0.0.9).  Returns 0 / 1 / 2. */
int groth16_prover(const void *zkey_buffer, unsigned long zkey_size, const void *wtns_buffer, unsigned long wtns_size,
                   char *proof_buffer, unsigned long *proof_size, char *public_buffer, unsigned long *public_size,
                   char *error_msg, unsigned long error_msg_maxsize);

/* Entry points later rapidsnark releases added to prover.h (bound by newer go-rapidsnark/prover versions); same
 * names and argument order, `unsigned long` == their `unsigned long long` on LP64.  Return 0 / 1 / 2 as above. */
void groth16_proof_size(unsigned long *proof_size);
int groth16_public_size_for_zkey_buf(const void *zkey_buffer, unsigned long zkey_size, unsigned long *public_size,
                                     char *error_msg, unsigned long error_msg_maxsize);
int groth16_public_size_for_zkey_file(const char *zkey_fname, unsigned long *public_size, char *error_msg,
                                      unsigned long error_msg_maxsize);
int groth16_prover_zkey_file(const char *zkey_file_path, const void *wtns_buffer, unsigned long wtns_size,
                             char *proof_buffer, unsigned long *proof_size, char *public_buffer,
                             unsigned long *public_size, char *error_msg, unsigned long error_msg_maxsize);

/* Groth16 verification on the GPU (one proof per thread): (*Proof).Verify(vkey) of zk_census_test.go:122 /
 * `snarkjs groth16 verify`.  Documents are the reference's JSON files (verification_key.json, signals.json,
 * proof.json).  zkb_verify returns ZKB_OK when the proof is valid and ZKB_INVALID_PROOF when it is not;
 * the batch form writes ok[i] = 1 / 0.  Encodings outside [0,q) / [0,r) and points off the curve or outside the
 * G2 subgroup are rejected. */
int zkb_verify(const char *vkey_json, size_t vkey_len, const char *public_json, size_t public_len,
               const char *proof_json, size_t proof_len);
int zkb_verify_batch(const char *vkey_json, size_t vkey_len, int n, const char *const *publics_json,
                     const size_t *publics_len, const char *const *proofs_json, const size_t *proofs_len, int *ok);

/* The same check on binary results (the layout of zkb_batch_get_results): proofs256 = n x 256 B, publics = n x nPublic x 32 B. */
int zkb_verify_batch_bin(const char *vkey_json, size_t vkey_len, int n, const void *publics, const void *proofs256, int *ok);

/* `snarkjs zkey export verificationkey` (circuit/circuit-compiler.sh:128-134): the verification_key.json of a proving
 * key, the exact text snarkjs writes (incl. vk_alphabeta_12 in ffjavascript's convention).  *out_len: capacity in,
 * bytes written (or needed, with ZKB_SHORT_BUFFER) out.  Host arithmetic only (one pairing); needs no GPU. */
int zkb_export_vkey(const void *zkey, size_t zkey_len, char *out, size_t *out_len);

/* Batched Poseidon with the circuit's constants (arity 2..4): the hash function of the census and SIK trees
 * (arbo.HashFunctionPoseidon, internal/helpers.go:45-49; circomlibjs in ts_inputs/src/inputs.ts:16,33).
 * in: n x arity canonical 32-byte values, out: n x 32 bytes. */
int zkb_poseidon_hash(zkb_circuit *c, int arity, int n, const void *in, void *out);

/* Poseidon sparse Merkle tree of a census (arbo semantics of internal/helpers.go:36-85 GenTree /
 * GenProof): n distinct 32-byte little-endian keys with their values -> root and, per key, its nLevels + 1 siblings
 * (zero padded: what MockInputs puts into censusSiblings / sikSiblings, internal/inputs.go:44-72).  The trie is laid
 * out on the host (C++), all hashing runs on the GPU (one batched launch per tree level).  siblings may be NULL. */
int zkb_census_tree(zkb_circuit *c, int n_keys, const void *keys32, const void *values32, int n_levels, void *root32,
                    void *siblings);

/* --- resident-input batch path (inputs already in HBM; what bench.py times as `value`) -------------- */

/* inputs: n x nInputs canonical 32-byte little-endian values in the circuit's main-signal order
 * (electionId[2], nullifier, voteHash[2], sikRoot, censusRoot, voteWeight, availableWeight, address, password,
 * signature, censusSiblings[nLevels+1], sikSiblings[nLevels+1]; circuit/census.circom:51-67 with the public
 * list of circuit/circuit-compiler.sh:85-88 first). */
int zkb_batch_set_inputs(zkb_circuit *c, int n, const void *inputs);
/* witness + Groth16 for the n resident inputs, results stay on the device.  stage_ms: NULL or 8 floats
 * (witness, buildABC, NTT+join, MSM sort, MSM accumulate G1, MSM accumulate G2, MSM reduce, finalize) measured
 * with CUDA events on the pipeline stream. */
int zkb_batch_prove_resident(zkb_circuit *c, int n, float *stage_ms);
/* proofs256: n x 256 B (A.x A.y B.x.c0 B.x.c1 B.y.c0 B.y.c1 C.x C.y canonical LE); publics: n x nPublic x 32 B */
int zkb_batch_get_results(zkb_circuit *c, int n, void *proofs256, void *publics, int *status);
int zkb_batch_get_witness(zkb_circuit *c, int first, int n, void *wtns);
/* measurement aids: kernels launched so far by the proving pipeline; executed MSM work of the last chunk
 * (out[8]: G1 XYZZ mixed adds, G2 mixed adds, witness digit entries, H digit entries, proofs in chunk, chunk capacity,
 * affine additions of the H MSM's pair tree (6 field products each), field inversions of the pair tree) */
uint64_t zkb_launch_count(void);
int zkb_work_counters(zkb_circuit *c, uint64_t *out);
/* debugging aid for parity tests: the five MSM partial sums (pi_a' pi_b1' pi_b' pi_c' pi_h, 384 B affine canonical)
 * and the H scalars (domainSize x 32 B) of the first proof of the last processed chunk */
int zkb_debug_partials(zkb_circuit *c, void *out384, void *h_out);

/* --- raw kernels on caller data (BASELINE.json config 5) --------------------------------------------- */

/* field: 0 = Fq, 1 = Fr; op: 0 mul 1 add 2 sub 3 inv 4 sqr 5 to_mont 6 from_mont 7 neg; 8 x u32 Montgomery residues */
int zkb_raw_field_op(int field, int op, const void *a, const void *b, void *out, size_t n);
int zkb_bench_modmul(int field, int iters, int blocks_per_sm, double *modmul_per_s, double *ms);
/* nvec vectors of 2^logn canonical Fr values, natural order in and out */
int zkb_raw_ntt(void *data, int logn, int nvec, int inverse, float *kernel_ms);
int zkb_raw_coset_ntt(void *data, int logn, int nvec);
/* bases: n canonical affine points (64 B G1 / 128 B G2, zeros = infinity); scalars: nbatch x n x 32 B */
int zkb_raw_msm_g1(const void *bases, size_t n, const void *scalars, int nbatch, void *out, float *kernel_ms,
                   float *table_ms);
int zkb_raw_msm_g2(const void *bases, size_t n, const void *scalars, int nbatch, void *out, float *kernel_ms,
                   float *table_ms);
/* flags bit 1: variable-base MSM (no window table; one bucket set per window, window sums combined by Horner's rule);
 * flags bit 0: bucket lists summed by the batched-affine pair tree (opt-in path of the prover's H MSM in batch shape);
 * bits 8-11: tree levels (0 = 3), bits 16-31: additions sharing one field inversion (0 = 512) */
int zkb_raw_msm_g1_ex(const void *bases, size_t n, const void *scalars, int nbatch, void *out, float *kernel_ms,
                      float *table_ms, uint32_t flags);


/* --- device-resident raw sessions: large MSM split by point range, across GPUs (BASELINE.json configs 4-5) -------
 * The reference has no multi-GPU code; this is the "single large-circuit MSM split by point range with per-GPU
 * partial sums combined via P2P copy over NVLink" of BASELINE.json's north_star.  One process per GPU: every rank
 * creates a session for its point range (synthetic bases by try-and-increment hash-to-curve and scalars are generated
 * on the device, the window table is built once), rank 0 exports the CUDA IPC handle of its exchange buffer, the
 * other ranks attach to it, then each step every rank calls run() (asynchronous; its last kernel writes the rank's
 * partial sum into rank 0's memory and releases a flag) and rank 0 calls combine(). */
typedef struct zkb_msm_session zkb_msm_session;
/* window_bits: 12..16, + 0x100 for the variable-base form (no window table: the bases are used as they are) */
int zkb_msm_session_create(int device, int logn, int rank, int nranks, uint64_t seed, int window_bits,
                           zkb_msm_session **out);
void zkb_msm_session_destroy(zkb_msm_session *s);
/* info[6] = points of this rank, sub-MSM size, sub-MSMs, window bits, table build us, data generation us */
int zkb_msm_session_info(zkb_msm_session *s, uint64_t *info);
int zkb_msm_session_export(zkb_msm_session *s, void *handle64);
int zkb_msm_session_attach(zkb_msm_session *s, const void *handle64);
int zkb_msm_session_attach_local(zkb_msm_session *s, zkb_msm_session *root); /* same process, another GPU or the same */
int zkb_msm_session_run(zkb_msm_session *s, float *ms);
int zkb_msm_session_combine(zkb_msm_session *s, int nslots, void *out64, float *ms);
int zkb_msm_session_madds(zkb_msm_session *s, uint64_t *madds);
int zkb_msm_session_read(zkb_msm_session *s, void *bases_out, void *scalars_out);
/* 4-step NTT of 2^logn (>= 2^22) values over nranks GPUs, one session per rank: rank g owns N2/G columns of the
 * N1 x N2 input, runs the length-N1 column transforms, then READS its N1/G positions of every rank's columns through
 * CUDA IPC mappings (the all-to-all transpose and the omega_N^(n2 k1) twiddles are fused into that load), then runs
 * the length-N2 row transforms.  Output rows, concatenated over ranks, are the transform in bit-reversed order
 * (exactly what the single-GPU decimation-in-frequency plan produces).  Between two transforms the caller places a
 * host barrier (the column buffers are reused). */
typedef struct zkb_ntt_dist zkb_ntt_dist;
int zkb_ntt_dist_create(int device, int logn, int rank, int nranks, uint64_t seed, zkb_ntt_dist **out);
void zkb_ntt_dist_destroy(zkb_ntt_dist *s);
int zkb_ntt_dist_export(zkb_ntt_dist *s, void *handle64);
int zkb_ntt_dist_attach(zkb_ntt_dist *s, int peer_rank, const void *handle64);
int zkb_ntt_dist_attach_local(zkb_ntt_dist *s, int peer_rank, zkb_ntt_dist *peer);
int zkb_ntt_dist_fill(zkb_ntt_dist *s);
int zkb_ntt_dist_run(zkb_ntt_dist *s);
int zkb_ntt_dist_sync(zkb_ntt_dist *s, float *ms4); /* total, columns, wait + exchange, rows */
int zkb_ntt_dist_read(zkb_ntt_dist *s, int which, void *out);
int zkb_raw_ntt_dif_forward(void *data, int logn); /* single GPU, natural in, bit-reversed out (parity aid) */
/* resident NTT timing: ms per inverse (DIF + coset scale) and per forward (DIT) transform of nvec x 2^logn values */
int zkb_ntt_bench(int device, int logn, int nvec, int iters, float *dif_ms, float *dit_ms);

#ifdef __cplusplus
}
#endif
#endif /* ZKCENSUS_B200_H */
