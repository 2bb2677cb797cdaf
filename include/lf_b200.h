/* lf_b200.h -- C ABI of liblf_b200.so, the B200 (sm_100a) implementation of the LatticeFold prover hot path.
 *
 * The reference (NethermindEth/latticefold, Rust) has no FFI or plugin registry: the hot path sits behind generic
 * Rust functions.  Each entry point below replaces one of those call sites; the Rust-side binding a
 * maintainer would add (a `lf-b200-sys` crate + wrapper types) is shown in INTEGRATION.md.  File:line citations are
 * relative to the reference tree.
 *
 * Conventions
 *   - Ring element on the host side: D consecutive little-endian u64 limbs, canonical (value in [0,p)).
 *       NTT form:          limb index = slot * TAU + l          (what `RqNTT` holds; transcript/poseidon.rs:40-47)
 *       coefficient form:  limb index = power of X              (what `RqPoly` holds)
 *     Host vectors are contiguous arrays of such elements (the memory image of `Vec<R>` after `into_bigint()`).
 *   - Device vectors are opaque (`lf_vec`): limb-plane ("SoA") layout, see DESIGN.md.
 *   - Every function returns an `lf_status`; 0 is success, negative values mirror the reference's error enums.
 *     Nothing aborts; `lf_last_error(ctx)` gives the message of the last failure on that context.
 *   - One context = one (GPU, ring).  Calls on one context must be serialised by the caller; different contexts
 *     are independent (the reference calls `commit` from rayon workers: give each worker its own stream context
 *     or batch with lf_commit_batch).
 *   - There is no CPU fallback: without a CUDA device lf_ctx_create fails with LF_ERR_CUDA.
 */
#ifndef LF_B200_H
#define LF_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int32_t lf_status;
enum {
    LF_OK = 0,
    LF_ERR_WRONG_WITNESS_LEN = -1, /* CommitmentError::WrongWitnessLength      commitment.rs:14-26            */
    LF_ERR_LENGTHS_NOT_EQUAL = -2, /* CSError::LengthsNotEqual                 arith/error.rs:8-29            */
    LF_ERR_MLE_LEN = -3,           /* MleEvaluationError::IncorrectLength      utils/mle_helpers.rs:21-63     */
    LF_ERR_INVALID_SIZE_BOUNDS = -4, /* LatticefoldError / sanity_check        nifs.rs:165-173                */
    LF_ERR_INCORRECT_LENGTH = -5,  /* *Error::IncorrectLength                  nifs/error.rs:13-66            */
    LF_ERR_SUMCHECK_MISUSE = -6,   /* the panics of sumcheck/prover.rs:41,63,74,80 reported as a status       */
    LF_ERR_UNSUPPORTED = -8,       /* ring / parameter outside what this build implements                     */
    LF_ERR_DOES_NOT_FIT = -9,      /* a coefficient needs more digits than requested                          */
    LF_ERR_SUMCHECK_FAILED = -10,  /* SumCheckError::SumCheckFailed / evaluation-claim mismatch (verifier)    nifs/error.rs:13-66 */
    LF_ERR_RECOMPOSED = -11,       /* DecompositionError::RecomposedError (verifier)                          nifs/error.rs:35-48 */
    LF_ERR_CUDA = -20,
    LF_ERR_INVALID_ARG = -21
};

enum { LF_RING_GOLDILOCKS = 0, LF_RING_BABYBEAR = 1, LF_RING_FROG = 2 };  /* cyclotomic-rings/src/rings/{goldilocks,babybear,frog}.rs:9-20 */
enum { LF_FORM_NTT = 0, LF_FORM_COEFF = 1 };

typedef struct lf_ctx lf_ctx;
typedef struct lf_vec lf_vec;           /* device vector of ring elements                                         */
typedef struct lf_ajtai lf_ajtai;       /* AjtaiCommitmentScheme<R>{matrix}         commitment_scheme.rs:17-35    */
typedef struct lf_sparse lf_sparse;     /* SparseMatrix<R> as CSR                   arith/r1cs.rs:188-223         */
typedef struct lf_sumcheck lf_sumcheck; /* IPForMLSumcheck ProverState              sumcheck/prover.rs:19-31      */
typedef struct lf_transcript lf_transcript; /* PoseidonTranscript<R, CS> (host)     transcript/poseidon.rs:18-27  */

typedef struct { uint64_t p; int32_t d, n_slots, tau; uint64_t nu; } lf_ring_info;

/* ---- context ------------------------------------------------------------------------------------------------ */
lf_status lf_ring_describe(int32_t ring_id, lf_ring_info* out);
lf_status lf_ctx_create(int32_t ring_id, int32_t device, lf_ctx** out);
void lf_ctx_destroy(lf_ctx* ctx);
const char* lf_last_error(const lf_ctx* ctx);            /* ctx may be NULL: last creation error               */
lf_status lf_ctx_sync(lf_ctx* ctx);
void* lf_ctx_stream(lf_ctx* ctx);                        /* the cudaStream_t every call of this context uses   */
/* number of kernels this context has launched since creation (bench.py's gpu_launches) */
uint64_t lf_ctx_launches(const lf_ctx* ctx);
/* measurement aid: with profiling on, every kernel launch is bracketed by CUDA events on the context's stream;
 * the report is one line "kernel_name launches total_ms" per kernel.  Off by default (it adds two event records per launch). */
/* Representation of the WITNESS-SIZED host vectors (lf_vec_upload / lf_vec_download, lf_ajtai_create, lf_sparse_create values,
 * lf_prover_upload_witness / lf_witness_download_f, the witness arguments of lf_nifs_prove / lf_linearize):
 *   LF_REPR_CANONICAL  (default) canonical little-endian limbs, the image of a Vec<R> after into_bigint()
 *   LF_REPR_MONTGOMERY           ark-ff 0.4 Montgomery limbs (a * 2^64 mod p per base-field limb): the memory of a Vec<R> as it is
 * The conversion rides on the layout kernels of the copy.  Everything small (commitments, proofs, challenges, LCCCS fields,
 * transcript data) is canonical in either mode.                                                                              */
enum { LF_REPR_CANONICAL = 0, LF_REPR_MONTGOMERY = 1 };
lf_status lf_ctx_set_bulk_repr(lf_ctx* ctx, int32_t repr);
lf_status lf_ctx_profile(lf_ctx* ctx, int32_t enable);
lf_status lf_ctx_profile_report(lf_ctx* ctx, char* buf, size_t buf_len);

/* ---- multi-GPU: one context per rank; the witness-column axis (Ajtai commit) and the hypercube's high bits (sumcheck,
 * MLE evaluation) are sharded across the ranks, each followed by one small all-reduce.  The collective is
 * supplied by the host side (torch.distributed over NCCL in bench.py): op 0 = in-place sum of `words` u64 lanes,
 * op 1 = in-place all-gather (buffer of world * words, this rank's part at rank * words).  Returns 0 on success.
 * After this call lf_commit / lf_commit_batch / lf_mle_eval_batch treat their vectors as this rank's slice and return the
 * all-reduced result, and an lf_prover created on the context shards the whole step.                                   */
typedef int32_t (*lf_collective_fn)(void* user, int32_t op, void* device_ptr, size_t words);
lf_status lf_ctx_set_shard(lf_ctx* ctx, int32_t rank, int32_t world, lf_collective_fn fn, void* user);
/* Preferred on real multi-GPU boxes: the library opens its own NCCL communicator (rank 0 makes the id with
 * lf_nccl_unique_id and the host side broadcasts the 128 bytes), and every collective is enqueued on the context's stream
 * right behind the kernel that produced its input -- no host synchronisation, no callback.                              */
lf_status lf_nccl_unique_id(uint8_t* out128);
lf_status lf_ctx_set_shard_nccl(lf_ctx* ctx, int32_t rank, int32_t world, const uint8_t* id128);
uint64_t lf_ctx_collectives(const lf_ctx* ctx);
/* NVLink peer memory for the small all-reduces (one per commit batch / sumcheck round / evaluation batch): every rank exports a
 * mailbox region as a 64-byte CUDA IPC handle, the host side all-gathers the handles, and after the import the final reduction
 * of those ops and their all-reduce run as ONE kernel that stores its sums into every peer's mailbox (k_reduce_allreduce_p2p)
 * instead of split-limbs + ncclAllReduce + combine.  Call after lf_ctx_set_shard_nccl, on every rank, in the same order.     */
lf_status lf_ctx_p2p_export(lf_ctx* ctx, uint8_t* out_handle64);
lf_status lf_ctx_p2p_import(lf_ctx* ctx, int32_t rank, int32_t world, const uint8_t* handles /* world x 64 bytes; NULL = switch the peer path off again */);

/* ---- vectors: Vec<R> <-> device ------------------------------------------------------------------------------- */
lf_status lf_vec_upload(lf_ctx* ctx, const uint64_t* host, size_t n, int32_t form, lf_vec** out);
lf_status lf_vec_download(lf_ctx* ctx, const lf_vec* v, uint64_t* host);
size_t lf_vec_len(const lf_vec* v);
int32_t lf_vec_form(const lf_vec* v);
void lf_vec_free(lf_ctx* ctx, lf_vec* v);

/* ---- a2/a3: CRT::elementwise_crt / ICRT::elementwise_icrt (arith.rs:232,238,300,327; commitment_scheme.rs:85,110) */
lf_status lf_crt(lf_ctx* ctx, const lf_vec* in_coeff, lf_vec** out_ntt);
lf_status lf_icrt(lf_ctx* ctx, const lf_vec* in_ntt, lf_vec** out_coeff);

/* ---- a4: balanced decompositions (arith.rs:235,305,330; decomposition/utils.rs:23-49)                           */
/* gadget_decompose(B, L): element i -> elements [i*L, (i+1)*L), digit l has weight B^l                            */
lf_status lf_gadget_decompose(lf_ctx* ctx, const lf_vec* in_coeff, uint64_t B, int32_t L, lf_vec** out_coeff);
/* gadget_recompose(B, L): out[i] = sum_l in[i*L+l] * B^l (either form)                                            */
lf_status lf_gadget_recompose(lf_ctx* ctx, const lf_vec* in, uint64_t B, int32_t L, lf_vec** out);
/* decompose_to_vec(b, K).transpose(): K vectors, piece k has weight b^k (decomposition.rs:162-167,288-292)        */
lf_status lf_decompose_to_vec(lf_ctx* ctx, const lf_vec* in_coeff, uint64_t b, int32_t K, lf_vec** out_coeff_k);

/* ---- a5: Witness::get_fhat (arith.rs:273-297): tau NTT-form MLE tables from one coefficient-form vector         */
lf_status lf_fhat(lf_ctx* ctx, const lf_vec* in_coeff, lf_vec** out_tau);

/* ---- a1/a12: AjtaiCommitmentScheme (commitment_scheme.rs:17-114)                                                */
lf_status lf_ajtai_create(lf_ctx* ctx, size_t kappa, size_t n, const uint64_t* host_matrix_ntt, lf_ajtai** out);
void lf_ajtai_free(lf_ctx* ctx, lf_ajtai* a);
size_t lf_ajtai_kappa(const lf_ajtai* a);
size_t lf_ajtai_width(const lf_ajtai* a);
/* commit / commit_ntt: out_host = kappa ring elements.  Length mismatch -> LF_ERR_WRONG_WITNESS_LEN               */
lf_status lf_commit(lf_ctx* ctx, const lf_ajtai* a, const lf_vec* f_ntt, uint64_t* out_host);
/* the K-1 commits of one decomposition in a single pass over the matrix (decomposition.rs:178-201)                */
lf_status lf_commit_batch(lf_ctx* ctx, const lf_ajtai* a, const lf_vec* const* f_ntt, int32_t count, uint64_t* out_host);
/* commit_coeff (commitment_scheme.rs:80-87): CRT of every element, then commit                                      */
lf_status lf_commit_coeff(lf_ctx* ctx, const lf_ajtai* a, const lf_vec* f_coeff, uint64_t* out_host);
/* decompose_and_commit_coeff / _ntt (commitment_scheme.rs:89-114): G_B^{-1} of a coefficient-form (or, after ICRT, an NTT-form)
 * vector of n / L elements -- every element becomes L consecutive digit elements -- then commit_coeff.  A coefficient that does
 * not fit L digits of base B -> LF_ERR_DOES_NOT_FIT; (n / L) mismatch -> LF_ERR_WRONG_WITNESS_LEN                               */
lf_status lf_decompose_and_commit_coeff(lf_ctx* ctx, const lf_ajtai* a, const lf_vec* f_coeff, uint64_t B, int32_t L, uint64_t* out_host);
lf_status lf_decompose_and_commit_ntt(lf_ctx* ctx, const lf_ajtai* a, const lf_vec* w_ntt, uint64_t B, int32_t L, uint64_t* out_host);
/* the commitments of all K pieces of decompose_to_vec(b, K) of a coefficient-form witness (decompose_witness + commit_witnesses,
 * decomposition.rs:162-201, without the y_0 shortcut): out_host = K x kappa ring elements.  On the Goldilocks ring the pieces are
 * committed straight from their int8 digits as an integer GEMM on the tensor cores (csrc/commit_mma.cuh)                         */
lf_status lf_commit_pieces(lf_ctx* ctx, const lf_ajtai* a, const lf_vec* f_coeff, uint64_t b, int32_t K, uint64_t* out_host);

/* ---- a6: mat_vec_mul / calculate_Mz_mles (arith/utils.rs:52-65; mle_helpers.rs:137-146)                          */
lf_status lf_sparse_create(lf_ctx* ctx, size_t nrows, size_t ncols, const uint64_t* row_ptr, const uint64_t* col,
                           const uint64_t* val_host_ntt, lf_sparse** out);
void lf_sparse_free(lf_ctx* ctx, lf_sparse* m);
lf_status lf_spmv(lf_ctx* ctx, const lf_sparse* m, const lf_vec* z_ntt, lf_vec** out_ntt);

/* ---- a7: build_eq_x_r (sumcheck/utils.rs:100-170): r = s ring elements on the host, r[0] on bit 0               */
lf_status lf_eq_table(lf_ctx* ctx, const uint64_t* r_host, int32_t s, lf_vec** out_ntt);

/* ---- a9: evaluate_mles (mle_helpers.rs:65-88).  MLEs shorter than 2^s have an implicit zero tail; a vector longer
 *      than 2^s or a point of the wrong length gives LF_ERR_MLE_LEN                                               */
lf_status lf_mle_eval_batch(lf_ctx* ctx, const lf_vec* const* mles, int32_t count, int32_t num_vars,
                            const uint64_t* point_host, int32_t point_len, uint64_t* out_host);

/* ---- a11: compute_f_0 and the Horner MLE combinations (folding.rs:208-226,258-268) as one primitive              */
lf_status lf_lincomb(lf_ctx* ctx, const uint64_t* coeffs_host_ntt, const lf_vec* const* vecs, int32_t count, lf_vec** out);

/* ---- a8: MLSumcheck (utils/sumcheck.rs:53-80, sumcheck/prover.rs:56-162)
 * A Rust closure cannot cross the ABI, so the combination functions the reference uses are enumerated:            */
enum {
    LF_COMB_PRODUCTS = 0, /* sum_i coef_i * prod_{j in idx_i} v_j                    sumcheck/utils.rs:60-73       */
    LF_COMB_LIN = 1,      /* (sum_i c_i prod_{j in S_i} v_j) * v_last                linearization/utils.rs:90-107 */
    LF_COMB_FOLD = 2      /* v0 v1 + v2 v3 + v4 * sum_k sum_d mu_k^{d+1} f(f^2-1)..(f^2-(b-1)^2)  folding/utils.rs:273-325 */
};
typedef struct {
    int32_t kind;
    int32_t n_terms;            /* PRODUCTS / LIN                                                                 */
    const uint64_t* coef_host;  /* n_terms ring elements (NTT form)                                               */
    const int32_t* idx;         /* concatenated MLE indices                                                       */
    const int32_t* idx_len;     /* per-term count                                                                 */
    int32_t n_mu, b;            /* FOLD: 2K challenges mu (ring elements, slot-constant), small base b             */
    const uint64_t* mu_host;
} lf_comb;
/* takes ownership of the MLE vectors (they are folded in place and freed by lf_sumcheck_free), like the reference's
 * `Vec<DenseMultilinearExtension<R>>` argument.  All MLEs have num_vars variables; shorter vectors are zero-padded. */
lf_status lf_sumcheck_begin(lf_ctx* ctx, lf_vec** mles, int32_t n_mles, int32_t num_vars, int32_t degree,
                            const lf_comb* comb, lf_sumcheck** out);
/* one prove_round: prev_challenge = TAU limbs of the verifier message (NULL in round 1);
 * out_evals_host = (degree+1) ring elements                                                                      */
lf_status lf_sumcheck_round(lf_sumcheck* sc, const uint64_t* prev_challenge_sf, uint64_t* out_evals_host);
/* applies the last challenge and returns mle_k(r) for every MLE (what the caller would otherwise recompute with
 * evaluate_mles at the sumcheck point)                                                                           */
lf_status lf_sumcheck_finish(lf_sumcheck* sc, const uint64_t* last_challenge_sf, uint64_t* out_final_host);
void lf_sumcheck_free(lf_sumcheck* sc);

/* ---- a14: host transcript (the reference keeps Fiat-Shamir on the CPU; transcript/poseidon.rs:29-75)             */
lf_status lf_transcript_create(int32_t ring_id, lf_transcript** out);
lf_status lf_transcript_clone(const lf_transcript* t, lf_transcript** out);
void lf_transcript_free(lf_transcript* t);
void lf_transcript_absorb(lf_transcript* t, const uint64_t* ring_elems, size_t count);
void lf_transcript_absorb_base(lf_transcript* t, const uint64_t* limbs, size_t count);
void lf_transcript_absorb_tag(lf_transcript* t, const char* tag);
void lf_transcript_get_challenge(lf_transcript* t, uint64_t* out_sf);               /* TAU limbs                  */
void lf_transcript_get_short_challenge(lf_transcript* t, uint64_t* out_coeffs);     /* D coefficients             */
uint64_t lf_transcript_permutations(const lf_transcript* t);
/* which dense-layer implementation the host Poseidon of the Goldilocks ring runs on this machine: "avx512-ifma" or "scalar"
 * (same results; LF_POSEIDON_SCALAR=1 in the environment forces the scalar one)                                           */
const char* lf_host_poseidon_backend(void);

/* ---- a13: rot_lin_combination (cyclotomic-rings/src/rotation.rs:45-104), host                                    */
lf_status lf_rot_lin_combination(int32_t ring_id, const uint64_t* rho_coeff, const uint64_t* theta_ntt, int32_t count, uint64_t* out);

/* ---- the prover step: NIFSProver::prove (nifs.rs:48-103) = linearization + 2 x decomposition + folding            */
typedef struct { uint64_t nrows, ncols; const uint64_t* row_ptr; const uint64_t* col; const uint64_t* val; } lf_csr;
/* flat, host-side description of one step.  Pointers may be NULL where noted.                                      */
typedef struct {
    int32_t ring; int32_t L, K; uint64_t B_lo, B_hi, b;       /* DecompositionParams  decomposition_parameters.rs:11-20 */
    uint64_t kappa, n; const uint64_t* A;                     /* Ajtai matrix (host, NTT form) or NULL when a prepared lf_ajtai is passed */
    uint64_t m, n_ccs, l, t, q, d, s;                         /* CCS shape            arith.rs:51-74                */
    const lf_csr* M; const int32_t* S_flat; const int32_t* S_len; const uint64_t* c;
    const uint64_t *acc_r, *acc_v, *acc_cm, *acc_u, *acc_x_w, *acc_h;   /* LCCCS accumulator  arith.rs:193-206      */
    const uint64_t* w_acc_f;                                  /* accumulator witness f (NTT form, n elements)       */
    const uint64_t *cm_i_cm, *cm_i_x_ccs;                     /* incoming CCCS        arith.rs:180-185              */
    const uint64_t* w_i_f;                                    /* incoming witness f (NTT form, n elements)          */
} lf_problem;

typedef struct lf_prover lf_prover;   /* device-resident static state: Ajtai matrix + CCS matrices + workspaces     */
/* uploads the static inputs (matrix, CCS).  The reference builds those outside the timed closure as well
 * (benches/utils.rs:640-660).                                                                                      */
lf_status lf_prover_create(lf_ctx* ctx, const lf_problem* shape, lf_prover** out);
void lf_prover_free(lf_prover* p);
uint64_t lf_proof_words(const lf_problem* shape);     /* u64 words of the serialised LFProof                        */
uint64_t lf_lcccs_words(const lf_problem* shape);     /* u64 words of a serialised LCCCS: r, v, cm, u, x_w, h        */
/* Witness::from_w_ccs on the device (arith.rs:230-248): returns f (NTT form) = CRT(gadget_decompose(ICRT(w_ccs)))   */
lf_status lf_witness_f_from_w_ccs(lf_ctx* ctx, const uint64_t* w_ccs_host, size_t W, uint64_t B, int32_t L, uint64_t* f_host);
/* LFLinearizationProver::prove (linearization.rs:145-189) on (cm_i, w_i): LCCCS + its proof part                    */
lf_status lf_linearize(lf_prover* p, const lf_problem* in, lf_transcript* t, uint64_t* out_lcccs, uint64_t* out_lin_proof);
/* one full step from HOST inputs: uploads w_acc_f / w_i_f, proves, downloads proof, folded LCCCS and (if out_f is
 * not NULL) the folded witness f_0.  Proof layout: lin{msgs,v,u} | dec_acc{x,y,u,v per piece} | dec_new | fold{msgs,theta,eta} */
lf_status lf_nifs_prove(lf_prover* p, const lf_problem* in, lf_transcript* t, uint64_t* out_proof, uint64_t* out_lcccs, uint64_t* out_f);
/* NIFSVerifier::verify (nifs.rs:117-162): host code like the reference's (no witness-sized data, no GPU needed).  `in` carries
 * the CCS shape (M may be NULL), the accumulator and cm_i; A and the witnesses are not read.  LF_OK = accepted, out_lcccs (may be
 * NULL) receives the folded instance; a rejected proof returns LF_ERR_SUMCHECK_FAILED / LF_ERR_RECOMPOSED / LF_ERR_INCORRECT_LENGTH
 * with the reason in lf_last_error(NULL).                                                                                       */
lf_status lf_nifs_verify(const lf_problem* in, lf_transcript* t, const uint64_t* proof, uint64_t* out_lcccs);
/* LFLinearizationVerifier::verify (nifs/linearization.rs:192-285) on the linearization part of a proof (msgs, v, u), host code.   */
lf_status lf_linearization_verify(const lf_problem* in, lf_transcript* t, const uint64_t* lin_proof, uint64_t* out_lcccs);
/* Proof wire format: the bytes of `LFProof::serialize_with_mode(.., Compress::Yes)` (nifs.rs:28-34,
 * examples/e2e.rs:126-146; ark-serialize 0.4: u64-LE length prefixes, canonical little-endian field elements) <-> the flat u64 proof
 * of lf_nifs_prove.  Host code.  Deserialisation validates every length prefix against the problem's shape and rejects non-canonical
 * field elements (LF_ERR_INCORRECT_LENGTH / LF_ERR_INVALID_ARG), as arkworks does.                                                 */
uint64_t lf_proof_wire_bytes(const lf_problem* shape);
lf_status lf_proof_serialize(const lf_problem* shape, const uint64_t* proof_words, uint8_t* out_bytes);
lf_status lf_proof_deserialize(const lf_problem* shape, const uint8_t* bytes, uint64_t n_bytes, uint64_t* out_proof_words);
/* the same step with both witnesses already resident in HBM (bench.py's `value`): witnesses are handles made by
 * lf_prover_upload_witness; the folded witness stays on the device and is returned as a new handle                 */
typedef struct lf_witness lf_witness;
lf_status lf_prover_upload_witness(lf_prover* p, const uint64_t* f_host_ntt, lf_witness** out);
void lf_witness_free(lf_prover* p, lf_witness* w);
/* LFLinearizationProver::prove on a resident witness, and Witness::commit (arith.rs:357-362) = A f of a resident witness
 * (BASELINE configs[2]: commit + linearization sumcheck)                                                                     */
lf_status lf_linearize_resident(lf_prover* p, const lf_problem* in, const lf_witness* w_i, lf_transcript* t, uint64_t* out_lcccs, uint64_t* out_lin_proof);
lf_status lf_witness_commit(lf_prover* p, const lf_witness* w, uint64_t* out_host);
lf_status lf_witness_download_f(lf_prover* p, const lf_witness* w, uint64_t* f_host);
lf_status lf_nifs_prove_resident(lf_prover* p, const lf_problem* in, const lf_witness* w_acc, const lf_witness* w_i,
                                 lf_transcript* t, uint64_t* out_proof, uint64_t* out_lcccs, lf_witness** out_w);
/* per-phase device/host timings of the last step in milliseconds: [0] linearization [1] decomposition x2 [2] folding
 * [3] host transcript [4] total wall                                                                               */
lf_status lf_prover_last_timings(const lf_prover* p, double* out5);
/* diagnostic: with LF_TIMING_DETAIL=1 in the environment the step synchronises at phase marks; one line "mark ms" each */
lf_status lf_prover_timing_detail(const lf_prover* p, char* buf, size_t buf_len);

/* ---- K12: batched negacyclic NTT over Z_p[X]/(X^N + 1), N = 2^8 .. 2^16 (BASELINE.json configs[4]).
 * The reference has no transform of this shape -- its "NTT form" is the CRT of the degree-24/72/16 rings (lf_crt / lf_icrt,
 * a2/a3 above, stark-rings CRT::elementwise_crt) -- so the definition is the textbook one:
 *     forward  A[k] = sum_j a[j] psi^(j(2k+1)),   inverse  a[j] = N^-1 sum_k A[k] psi^(-j(2k+1)),   natural order both sides,
 * psi = the primitive 2N-th root of unity returned by lf_ntt_root (rule: latticefold_b200/csrc/ntt.cuh).  Elements are canonical
 * little-endian words: uint64_t for LF_FIELD_GOLDILOCKS, uint32_t for LF_FIELD_BABYBEAR; a batch is `batch` polynomials of N
 * words back to back.  The *_device entry points take plain device pointers (16-byte aligned) and run on the context's stream,
 * in place if d_in == d_out; the *_host ones copy in, transform and copy out.  The ring of `ctx` is irrelevant here.            */
enum { LF_FIELD_GOLDILOCKS = 0, LF_FIELD_BABYBEAR = 1 };
typedef struct lf_ntt_plan lf_ntt_plan;
lf_status lf_ntt_root(int32_t field, int32_t log_n, uint64_t* psi_out);
lf_status lf_ntt_plan_create(lf_ctx* ctx, int32_t field, int32_t log_n, lf_ntt_plan** out);
void lf_ntt_plan_free(lf_ctx* ctx, lf_ntt_plan* plan);
lf_status lf_ntt_forward_device(lf_ctx* ctx, const lf_ntt_plan* plan, const void* d_in, void* d_out, size_t batch);
lf_status lf_ntt_inverse_device(lf_ctx* ctx, const lf_ntt_plan* plan, const void* d_in, void* d_out, size_t batch);
lf_status lf_ntt_forward_host(lf_ctx* ctx, const lf_ntt_plan* plan, const void* h_in, void* h_out, size_t batch);
lf_status lf_ntt_inverse_host(lf_ctx* ctx, const lf_ntt_plan* plan, const void* h_in, void* h_out, size_t batch);
/* slot-wise product of two transformed batches, and the full negacyclic ring product INTT(NTT(a) . NTT(b))                     */
lf_status lf_ntt_pointwise_mul_device(lf_ctx* ctx, const lf_ntt_plan* plan, const void* d_a, const void* d_b, void* d_out, size_t batch);
lf_status lf_ntt_negacyclic_mul_host(lf_ctx* ctx, const lf_ntt_plan* plan, const void* h_a, const void* h_b, void* h_out, size_t batch);

/* ---- LatticeFold+ consumers of the same kernels (crates/latticefold-plus) on the coefficient-form ring
 * R = Z_q[X]/(X^16 + 1) (LF_RING_FROG, `frog_ring::RqPoly`, BaseRing = Fq).  Ring elements cross as 16 canonical coefficients.
 * The transcript is the same Poseidon sponge (latticefold-plus/src/transcript.rs:16-56: absorb = the coefficients, a challenge is
 * ONE field element squeezed and absorbed back): pass an lf_transcript created for LF_RING_FROG.
 *
 * Results are flat u64 images:
 *   set check   (setchk.rs:28-36 `Out`):  [nvars, n_mat, ncols, n_vec, n_M]  r[nvars]  sumcheck messages [nvars][4][16]
 *                                         e[(1 + n_M)][n_mat][ncols][16]  b[n_vec][16]
 *   range check (rgchk.rs:50-64 `Dcom`):  [L, k, l, kappa, b]  the set-check image, then per instance
 *                                         v[16] a[1 + n_M] b[1 + n_M][16] c[1 + n_M][16] cm_f[kappa][16] C_Mf[kappa][16] cm_mtau[kappa][16]
 * Conventions of the un-vendored stark-rings pieces (exp(0) = 1, element order of `split`): DESIGN.md "Conventions".           */
typedef struct { int32_t kind; int32_t pad; lf_csr m; const uint64_t* v; uint64_t n; } lf_plus_set; /* MonomialSet  setchk.rs:17-21: kind 0 = Matrix(m), 1 = Vector(v, n elements) */
typedef struct lf_plus_mat lf_plus_mat;   /* Matrix<R> kappa x n, coefficient form, device resident (the matrix A of from_f)  */
typedef struct lf_plus_rg lf_plus_rg;     /* RgInstance<R>                            rgchk.rs:41-48                          */
typedef struct lf_plus_vec lf_plus_vec;   /* Vec<R>, n x 16 coefficients, device resident (the witness of a LinB, lin.rs:38-42) */
/* Transcript::get_challenge of latticefold-plus/src/transcript.rs:46-55 (extension degree 1)                                  */
void lf_transcript_get_challenge_base(lf_transcript* t, uint64_t* out1);
/* In::set_check                              setchk.rs:59-262.  LF_ERR_INVALID_ARG when out_cap is too small (*out_len = needed) */
lf_status lf_plus_set_check(lf_ctx* ctx, lf_transcript* t, int32_t nvars, const lf_plus_set* sets, int32_t n_sets,
                            const lf_csr* M, int32_t n_M, uint64_t* out, uint64_t out_cap, uint64_t* out_len);
/* Out::verify                                setchk.rs:264-344 (host).  LF_OK = accepted, LF_ERR_SUMCHECK_FAILED = rejected   */
lf_status lf_plus_set_check_verify(lf_transcript* t, const uint64_t* words, uint64_t len);
lf_status lf_plus_mat_create(lf_ctx* ctx, uint64_t kappa, uint64_t n, const uint64_t* host_coeff, lf_plus_mat** out);
void lf_plus_mat_free(lf_ctx* ctx, lf_plus_mat* a);
/* RgInstance::from_f                         rgchk.rs:259-336: digits of cf(f), M_f = exp(D_f), A * M_f, split, cm_f, C_Mf, cm_mtau.
 * LF_ERR_DOES_NOT_FIT when a coefficient of f needs more than k digits in base b                                              */
lf_status lf_plus_rg_from_f(lf_ctx* ctx, const lf_plus_mat* A, const uint64_t* f_coeff, uint64_t n, uint64_t b, int32_t k, int32_t l, lf_plus_rg** out);
/* tau[n], fcoms[3][kappa][16] = cm_f, C_Mf, cm_mtau, comM[k][kappa][16][16] = A * M_f[kk]; any pointer may be NULL             */
lf_status lf_plus_rg_read(const lf_plus_rg* inst, uint64_t* tau, uint64_t* fcoms, uint64_t* comM);
void lf_plus_rg_free(lf_ctx* ctx, lf_plus_rg* inst);
/* Rg::range_check                            rgchk.rs:75-187                                                                   */
lf_status lf_plus_range_check(lf_ctx* ctx, lf_transcript* t, int32_t nvars, lf_plus_rg* const* inst, int32_t L,
                              const lf_csr* M, int32_t n_M, uint64_t* out, uint64_t out_cap, uint64_t* out_len);
/* Dcom::verify                               rgchk.rs:190-246 (host).  LF_OK / LF_ERR_SUMCHECK_FAILED / LF_ERR_RECOMPOSED (a psi check) */
lf_status lf_plus_range_check_verify(lf_transcript* t, const uint64_t* words, uint64_t len);
/* Cm::prove                                  cm.rs:57-203: range check, h = M_f s', comh, the two degree-2 sumchecks over
 * [eq | tau, m_tau, f, h | M_i-images | t(0), t(1)] and g = s0 tau + s1 m_tau + s2 f + h.  Images:
 *   CmProof (cm.rs:31-37):  the Dcom image | comh[L][kappa][16] | sumcheck messages [2][nvars][3][16] | evaluations [2][L][1 + n_M][4][16]
 *   ComX    (cm.rs:45-50):  cm_g[L][kappa][16] | ro[nvars][2] | vo[L][1 + n_M][2][16]          (lf_plus_comx_words words)
 * g_host (L x n x 16 words, Com::g) may be NULL.                                                                                 */
uint64_t lf_plus_comx_words(int32_t nvars, int32_t L, uint64_t kappa, int32_t n_M);
lf_status lf_plus_cm_prove(lf_ctx* ctx, lf_transcript* t, int32_t nvars, lf_plus_rg* const* inst, int32_t L, const lf_csr* M, int32_t n_M,
                           uint64_t* proof, uint64_t proof_cap, uint64_t* proof_len, uint64_t* comx, uint64_t* g_host);
/* CmProof::verify                            cm.rs:349-535 (host; only the number of matrices is read from M there).  comx_out may be NULL */
lf_status lf_plus_cm_verify(lf_transcript* t, const uint64_t* proof, uint64_t len, int32_t n_M, uint64_t* comx_out);
/* ComR1CS::linearize                         r1cs.rs:72-134: g* = A|B|C f, the degree-3 sumcheck of eq(r, x) (ga gb - gc) over ring-valued tables
 * (real negacyclic products) and the evaluations at its point.  image = [nvars] ro[nvars] messages[nvars][4][16] v[16] va vb vc;
 * the LinB the reference returns is (f, cm_f, r = (ro, ro), v = the four evaluations twice)                                             */
lf_status lf_plus_r1cs_linearize(lf_ctx* ctx, lf_transcript* t, const lf_csr* abc /* A, B, C */, const uint64_t* f, uint64_t n, uint64_t* out, uint64_t out_cap, uint64_t* out_len);
/* ComR1CSProof::verify                       r1cs.rs:136-162 (host)                                                                    */
lf_status lf_plus_r1cs_linearize_verify(lf_transcript* t, const uint64_t* words, uint64_t len);
/* Mlin::mlin                                 mlin.rs:41-106: from_f on each of the L witnesses (fs: L x n x 16), Cm::prove, and the sums
 * over the instances.  linb2x (LinB2X) = cm_g[kappa][16] | ro[nvars][2] | vo[1 + n_M][2][16]; g_host (n x 16, LinB2::g) may be NULL     */
lf_status lf_plus_mlin(lf_ctx* ctx, lf_transcript* t, const lf_plus_mat* A, const uint64_t* fs, int32_t L, uint64_t n, uint64_t b, int32_t k, int32_t l,
                       const lf_csr* M, int32_t n_M, uint64_t* proof, uint64_t proof_cap, uint64_t* proof_len, uint64_t* linb2x, uint64_t* g_host);
/* Decomp::decompose                          decomp.rs:32-99: F = decompose_to_vec(f, B, 2), C_i = A F_i, v_i = evaluations of F_i and of
 * every M_j F_i at the two points r_pairs (nvars x 2 field elements).  proof (DecompProof) = C0[kappa][16] | C1 | v0[1 + n_M][2][16] | v1;
 * F_host (2 x n x 16) may be NULL.  LF_ERR_DOES_NOT_FIT when a coefficient needs more than two digits                                   */
lf_status lf_plus_decompose(lf_ctx* ctx, const lf_plus_mat* A, const uint64_t* f, uint64_t n, const uint64_t* r_pairs, const lf_csr* M, int32_t n_M,
                            uint64_t B, uint64_t* proof, uint64_t* F_host);
/* DecompProof::verify                        decomp.rs:102-126 (host): LF_OK / LF_ERR_RECOMPOSED                                         */
lf_status lf_plus_decompose_verify(const uint64_t* proof, uint64_t kappa, int32_t n_M, const uint64_t* cm_f, const uint64_t* v, uint64_t B);
/* Device-resident witnesses for the flow of PlusProver::prove (plus.rs:80-117): a LinB's f is uploaded once, mlin leaves g on the device,
 * decompose leaves the two digit vectors (the next accumulator) there; the *_v entry points are the ones above on such vectors.          */
lf_status lf_plus_vec_upload(lf_ctx* ctx, const uint64_t* host, uint64_t n, lf_plus_vec** out);
lf_status lf_plus_vec_download(lf_ctx* ctx, const lf_plus_vec* v, uint64_t* host);
uint64_t lf_plus_vec_len(const lf_plus_vec* v);
void lf_plus_vec_free(lf_ctx* ctx, lf_plus_vec* v);
lf_status lf_plus_r1cs_linearize_v(lf_ctx* ctx, lf_transcript* t, const lf_csr* abc, const lf_plus_vec* f, uint64_t* out, uint64_t out_cap, uint64_t* out_len);
lf_status lf_plus_mlin_v(lf_ctx* ctx, lf_transcript* t, const lf_plus_mat* A, const lf_plus_vec* const* fs, int32_t L, uint64_t b, int32_t k, int32_t l,
                         const lf_csr* M, int32_t n_M, uint64_t* proof, uint64_t proof_cap, uint64_t* proof_len, uint64_t* linb2x, lf_plus_vec** g_out);
lf_status lf_plus_decompose_v(lf_ctx* ctx, const lf_plus_mat* A, const lf_plus_vec* f, const uint64_t* r_pairs, const lf_csr* M, int32_t n_M, uint64_t B,
                              uint64_t* proof, lf_plus_vec** F0, lf_plus_vec** F1);
/* Static matrices (the M of PlusProver::init, the R1CS matrices): lf_plus_csr_pin keeps a validated, device-resident copy; every lf_plus_*
 * entry point that is later handed the same host arrays uses it instead of uploading again.  The arrays must stay alive and unchanged until
 * lf_plus_csr_unpin (or the context is destroyed).                                                                                      */
lf_status lf_plus_csr_pin(lf_ctx* ctx, const lf_csr* m);
lf_status lf_plus_csr_unpin(lf_ctx* ctx, const lf_csr* m);
/* utils.rs:74-86 tensor(r) (host): out has 2^n entries                                                                         */
lf_status lf_plus_tensor(const uint64_t* r, int32_t n, uint64_t* out);

#ifdef __cplusplus
}
#endif
#endif /* LF_B200_H */
