/*
 * libboojum_b200 -- C-ABI of the B200-native backend for Boojum's polynomial-commitment hot path.
 *
 * The reference (matter-labs/era-boojum, Rust) has no FFI for this path: the work sits behind generic
 * traits.  Each entry point below is what a Rust `extern "C"` shim binds in place of the
 * cited reference function.  Conventions:
 *   - every call returns an int32 status (BJ_OK == 0, < 0 error); nothing aborts or throws across the ABI;
 *     bj_last_error(ctx) holds a message for the last failure on that context;
 *   - a bj_ctx is bound to one CUDA device and one stream; calls on one context are issued in order on that
 *     stream and are asynchronous unless stated (use bj_ctx_synchronize); one host thread per context;
 *   - field elements are little-endian u64; inputs may be non-canonical (any u64 congruent mod
 *     p = 2^64 - 2^32 + 1, as the reference tolerates, src/field/goldilocks/mod.rs:147-171); outputs are always
 *     CANONICAL (< p), i.e. exactly the values the reference serialises (mod.rs:99-107);
 *   - Fp2 elements are (c0, c1) pairs; Fp2 vectors are two separate u64 columns (SoA), as in the reference;
 *   - Poseidon2 digests are 4 x u64;
 *   - pointers named d_* are DEVICE pointers (cudaMalloc / bj_alloc / torch tensor data_ptr); h_* are host.
 *   - there is no CPU fallback: without a CUDA device bj_ctx_create fails with BJ_ERR_NO_DEVICE;
 *   - objects created through a context (bj_fri_oracles, bj_setup) hold device memory from that context's private
 *     stream-ordered pool: free them before bj_ctx_destroy.  A bj_proof is host memory only.
 */
#ifndef BOOJUM_B200_H
#define BOOJUM_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define BJ_API __attribute__((visibility("default")))
#else
#define BJ_API
#endif

#define BJ_OK 0
#define BJ_ERR_INVALID_ARG (-1)
#define BJ_ERR_CUDA (-2)
#define BJ_ERR_NO_DEVICE (-3)
#define BJ_ERR_OOM (-4)
#define BJ_ERR_UNSUPPORTED (-5)

#define BJ_GOLDILOCKS_P 0xFFFFFFFF00000001ull

typedef struct bj_ctx bj_ctx;

/* ---- context (replaces Worker, src/worker/mod.rs:5-87, as the executor handle) ---- */
BJ_API const char* bj_version(void);
BJ_API const char* bj_status_string(int32_t status);
/* stream: a cudaStream_t owned by the caller (e.g. torch.cuda.current_stream().cuda_stream); NULL is the CUDA
 * legacy default stream.  Kernels only ever run on this stream; the host-buffer entry points additionally use two
 * private copy streams so that upload, transform and download of successive column chunks overlap. */
BJ_API int32_t bj_ctx_create(int32_t device, void* stream, bj_ctx** out_ctx);
BJ_API int32_t bj_ctx_destroy(bj_ctx* ctx);
BJ_API int32_t bj_ctx_set_stream(bj_ctx* ctx, void* stream);
BJ_API int32_t bj_ctx_synchronize(bj_ctx* ctx);
/* Multi-GPU proving (one process and one context per GPU): declare that this context is rank `rank` of `world` (a power
 * of two, <= the LDE factor 2^log_lde of the proof).  Every buffer on the LDE domain then holds only the cosets j = rank (mod world), stored
 * [local coset][row]: bj_lde produces just those cosets, and bj_quotient_copy_permutation,
 * bj_quotient_divide_by_vanishing, bj_deep_quotient_group and bj_fri_fold take local buffers (sizes in their signatures
 * stay the GLOBAL domain sizes) and use the domain points of the owned cosets.  Merkle trees are built per rank over the
 * local leaves (a coset is a contiguous subtree: leaf index = coset * n + row, src/cs/implementations/proof.rs:89-91),
 * so caps and query paths are gathered by the caller.  bj_barycentric_evaluate reads local slot 0 (the global coset `rank`) and
 * uses that coset's shift, so any rank can open any column - the same value comes out.
 * The reference has no counterpart (its Worker is one machine's thread pool).  Default: rank 0 of 1. */
BJ_API int32_t bj_ctx_set_coset_shard(bj_ctx* ctx, uint32_t rank, uint32_t world, uint32_t log_lde);
BJ_API const char* bj_last_error(const bj_ctx* ctx);
/* number of kernels this library launched through ctx so far (for launch accounting) */
BJ_API uint64_t bj_launch_count(const bj_ctx* ctx);

/* ---- multi-GPU: communicator of the coset-sharded prover (one process - or one thread - per GPU) ----
 * Creating a communicator on a context declares its coset shard (bj_ctx_set_coset_shard(ctx, rank, world, log_lde)) and makes
 * bj_setup_create / bj_prove / bj_do_fri on that context run SHARDED: every rank passes the same full witness, keeps the LDE
 * cosets j = rank (mod world) of every committed polynomial, builds the Merkle subtrees of its cosets, and all ranks return
 * the same proof (identical to the single-GPU proof).  What crosses GPUs: cap digests of every oracle, the quotient cosets
 * (one all-gather of 2 * Q * n u64 - they are interpolated together, prover.rs:1399-1467), the openings (computed by the owner
 * of coset 0), the last FRI codeword and the query answers.  Requirements: world a power of two <= the LDE factor,
 * merkle_tree_cap_size >= the LDE factor.
 *   NCCL transport: rank 0 calls bj_comm_unique_id and hands the 128 bytes to the other ranks by any side channel (MPI, TCP,
 *   torch.distributed ...); every rank then calls bj_comm_create_nccl.  libnccl.so.2 is loaded at run time (the copy already
 *   in the process is reused); BJ_ERR_UNSUPPORTED if it is absent.
 *   Local transport: bj_comm_group_create(world) once, bj_comm_create_local per rank (ranks = host threads whose contexts sit
 *   on one device): lets one GPU run the sharded driver end to end.
 * Destroy the communicator before its context.  The raw collectives are exported for host code that shards other stages. */
typedef struct bj_comm bj_comm;
typedef struct bj_comm_group bj_comm_group;
#define BJ_COMM_UNIQUE_ID_BYTES 128
BJ_API int32_t bj_comm_unique_id(uint8_t out[BJ_COMM_UNIQUE_ID_BYTES]);
BJ_API int32_t bj_comm_create_nccl(bj_ctx* ctx, const uint8_t unique_id[BJ_COMM_UNIQUE_ID_BYTES], uint32_t rank, uint32_t world,
                            uint32_t log_lde, bj_comm** out);
BJ_API int32_t bj_comm_group_create(uint32_t world, bj_comm_group** out);
BJ_API void bj_comm_group_destroy(bj_comm_group* group);
BJ_API int32_t bj_comm_create_local(bj_ctx* ctx, bj_comm_group* group, uint32_t rank, uint32_t log_lde, bj_comm** out);
BJ_API int32_t bj_comm_destroy(bj_comm* comm);
BJ_API uint32_t bj_comm_rank(const bj_comm* comm);
BJ_API uint32_t bj_comm_world(const bj_comm* comm);
/* d_recv[r * n .. (r + 1) * n) = rank r's d_send[0 .. n): ncclAllGather on the context's stream (asynchronous); the local
 * transport completes before returning.  d_send may be its own slot of d_recv. */
BJ_API int32_t bj_comm_all_gather(bj_comm* comm, const uint64_t* d_send, uint64_t* d_recv, uint64_t n_u64_per_rank);
/* the same for small host buffers (staged through the device); synchronises */
BJ_API int32_t bj_comm_all_gather_host(bj_comm* comm, const uint64_t* h_send, uint64_t* h_recv, uint64_t n_u64_per_rank);
BJ_API int32_t bj_comm_broadcast_host(bj_comm* comm, uint64_t* h_buf, uint64_t n_u64, uint32_t root);

/* ---- device memory (GoodAllocator hook, src/cs/traits/mod.rs:13-15) ---- */
BJ_API int32_t bj_alloc(bj_ctx* ctx, size_t bytes, void** d_ptr);
BJ_API int32_t bj_free(bj_ctx* ctx, void* d_ptr);
BJ_API int32_t bj_upload(bj_ctx* ctx, void* d_dst, const void* h_src, size_t bytes);   /* async on ctx stream */
BJ_API int32_t bj_download(bj_ctx* ctx, void* h_dst, const void* d_src, size_t bytes); /* async on ctx stream */
BJ_API int32_t bj_alloc_host_pinned(size_t bytes, void** h_ptr);
BJ_API int32_t bj_free_host_pinned(void* h_ptr);

/* ---- twiddles: precompute_twiddles_for_fft::<_,_,_,INVERSED> (src/cs/implementations/utils.rs:88-125) ----
 * Copies tab[i] = w^bitrev_{n/2}(i), i < n/2 (w = omega_n or omega_n^-1) into d_out (n/2 u64).  The library
 * caches its own tables; this export exists for parity tests and for callers that want the reference table. */
BJ_API int32_t bj_twiddles(bj_ctx* ctx, uint32_t log_n, int32_t inverse, uint64_t* d_out);

/* ---- NTT: PrimeFieldLikeVectorized::fft_natural_to_bitreversed / ifft_natural_to_natural
 *      (src/field/traits/field_like.rs:139-161 -> src/fft/mod.rs:398-411, 464-491) ----
 * In place on n_cols columns of 2^log_n elements; column c starts at d_data + c*col_stride (elements).
 * forward: out[bitrev(k)] = sum_i a_i (coset w^k)^i.  inverse: natural-order values on coset<w> -> natural-order
 * monomial coefficients (network with w^-1, bit reversal, scaling by coset^-i n^-1).  coset == 1 means none. */
BJ_API int32_t bj_ntt_natural_to_bitreversed(bj_ctx* ctx, uint64_t* d_data, uint32_t log_n, uint32_t n_cols,
                                      uint64_t col_stride, uint64_t coset);
BJ_API int32_t bj_intt_natural_to_natural(bj_ctx* ctx, uint64_t* d_data, uint32_t log_n, uint32_t n_cols,
                                   uint64_t col_stride, uint64_t coset);
/* bitreverse_enumeration_inplace (src/fft/mod.rs:41-155), batched */
BJ_API int32_t bj_bitreverse(bj_ctx* ctx, uint64_t* d_data, uint32_t log_n, uint32_t n_cols, uint64_t col_stride);

/* ---- LDE: transform_raw_storages_to_lde / transform_monomials_to_lde (src/cs/implementations/utils.rs:270-403)
 * d_in : n_cols columns (column c at d_in + c*in_col_stride) of 2^log_n Lagrange values in natural row order
 *        (or monomial coefficients if from_monomials != 0).  Not modified.
 * d_out: [col][coset j][row], 2^log_lde cosets of 2^log_n values each; coset j is evaluated on
 *        7 * w_{nL}^{bitrev_L(j)} * <w_n>, values bit-reversed within the coset (ArcGenericLdeStorage layout,
 *        src/cs/implementations/polynomial/lde.rs:161-170).  Column c starts at d_out + c * (n << log_lde). */
BJ_API int32_t bj_lde(bj_ctx* ctx, const uint64_t* d_in, uint64_t in_col_stride, uint64_t* d_out, uint32_t log_n,
               uint32_t log_lde, uint32_t n_cols, int32_t from_monomials);

/* ---- Poseidon2 Merkle tree: MerkleTreeWithCap::construct / construct_by_chunking /
 *      construct_by_chunking_from_flat_sources / continue_from_leaf_hashes (src/cs/oracle/merkle_tree.rs:78-449)
 *      with H = GoldilocksPoseidon2Sponge<AbsorptionModeOverwrite> (src/cs/oracle/mod.rs:114-175) ----
 * h_sources: HOST array of n_sources DEVICE pointers; source s is a flat array of n_leaves*elems_per_leaf u64
 *            (a column's cosets flattened coset-major).  Leaf m absorbs, for s = 0..n_sources-1 in order,
 *            source_s[m*elems_per_leaf .. (m+1)*elems_per_leaf).
 * d_leaf_hashes: n_leaves digests.  d_nodes: concatenated levels n_leaves/2, n_leaves/4, ..., cap_size digests
 *            (node_hashes_enumerated_from_leafs); total n_leaves - cap_size digests.  The cap is the last level
 *            (or the leaf hashes when n_leaves == cap_size). */
BJ_API int32_t bj_merkle_build_poseidon2(bj_ctx* ctx, const uint64_t* const* h_sources, uint32_t n_sources,
                                  uint64_t n_leaves, uint32_t elems_per_leaf, uint32_t cap_size,
                                  uint64_t* d_leaf_hashes, uint64_t* d_nodes);
/* Same tree with H = blake2::Blake2s256 (src/cs/oracle/mod.rs:179-245): leaf = Blake2s-256 over the LE bytes of the reduced
 * elements, node = Blake2s-256(left || right); 32-byte digests stored as 4 LE u64 (same [n][4] layout). */
BJ_API int32_t bj_merkle_build_blake2s(bj_ctx* ctx, const uint64_t* const* h_sources, uint32_t n_sources, uint64_t n_leaves,
                                uint32_t elems_per_leaf, uint32_t cap_size, uint64_t* d_leaf_hashes, uint64_t* d_nodes);
/* impl TreeHasher for sha3::Keccak256 (src/cs/oracle/mod.rs:247-313): same layout and arguments, Keccak-256 digests */
BJ_API int32_t bj_merkle_build_keccak256(bj_ctx* ctx, const uint64_t* const* h_sources, uint32_t n_sources, uint64_t n_leaves,
                                uint32_t elems_per_leaf, uint32_t cap_size, uint64_t* d_leaf_hashes, uint64_t* d_nodes);
/* TreeHasher::hash_into_leaf on rows given contiguously: n_rows rows of row_len u64 (row-major) -> digests */
BJ_API int32_t bj_poseidon2_hash_rows(bj_ctx* ctx, const uint64_t* d_rows, uint64_t n_rows, uint32_t row_len,
                               uint64_t* d_digests);
/* raw permutation on n_states states of 12 u64 (src/implementations/poseidon2/state_generic_impl.rs:219-233) */
BJ_API int32_t bj_poseidon2_permute(bj_ctx* ctx, uint64_t* d_states, uint64_t n_states);

/* ---- FRI fold: fold_multiple / interpolate_flattened_cosets (src/cs/implementations/fri/mod.rs:362-474, 587-678)
 * One oracle step = `log_fold` (1..3) successive fold-by-2 of a flat Fp2 vector of 2^log_m values:
 *   out[i] = (f[2i] + f[2i+1]) + alpha * (f[2i] - f[2i+1]) * R[i] * kappa,
 *   R = inverse twiddle table of the full LDE domain (prefix), kappa = *coset_inv, squared after every fold;
 *   fold j uses challenge alpha^(2^j).  h_alpha = (c0, c1) of the first challenge.  On return *h_coset_inv_io holds
 *   the updated kappa (as the reference's `coset_inverse.square()` leaves it).  Output length 2^(log_m-log_fold). */
BJ_API int32_t bj_fri_fold(bj_ctx* ctx, const uint64_t* d_c0, const uint64_t* d_c1, uint32_t log_m, uint32_t log_fold,
                    const uint64_t h_alpha[2], uint64_t* h_coset_inv_io, uint64_t* d_out_c0, uint64_t* d_out_c1);

/* ---- batch inverse: batch_inverse_inplace / batch_inverse_inplace_in_extension (src/cs/implementations/utils.rs:405-600)
 * In place.  The reference panics on a zero element ("must be called on sets without zeroes", utils.rs:425-427); here a
 * zero is mapped to zero and does not disturb the other elements. */
BJ_API int32_t bj_batch_inverse(bj_ctx* ctx, uint64_t* d_data, uint64_t n);
BJ_API int32_t bj_batch_inverse_ext(bj_ctx* ctx, uint64_t* d_c0, uint64_t* d_c1, uint64_t n);

/* ---- DEEP: quotening_operation_in_extension (src/cs/implementations/prover.rs:2523-2706), one call per opening point
 * For every point t of the LDE domain (2^log_rows = n*L points, coset-major, bit-reversed inside a coset; x(t) =
 * 7 * w_{nL}^{bitrev(t)}):   acc[t] += ( sum_i ch_i * (f_i(t) - v_i) ) / (x(t) - at)      in Fp2.
 * h_src_c0 / h_src_c1: HOST arrays of n_src DEVICE pointers to the c0 / c1 columns (each n*L u64, LDE layout);
 *   h_src_c1[i] == NULL marks a base-field polynomial (prover.rs:2655-2677).
 * h_values_at / h_challenges: n_src (c0, c1) pairs: f_i(at) and the challenge coefficient of term i.
 * d_acc_c0 / d_acc_c1: the Fp2 codeword accumulated across calls (zero-initialised by the caller). */
BJ_API int32_t bj_deep_quotient_group(bj_ctx* ctx, const uint64_t* const* h_src_c0, const uint64_t* const* h_src_c1,
                               uint32_t n_src, const uint64_t* h_values_at, const uint64_t* h_challenges,
                               const uint64_t h_at[2], uint32_t log_rows, uint64_t* d_acc_c0, uint64_t* d_acc_c1);

/* ---- stage 2, copy-permutation argument: compute_partial_products_in_extension
 *      (src/cs/implementations/copy_permutation.rs:649-766; rational :114-248, grand product :425-510) ----
 * Inputs are Lagrange columns on the trace domain in natural row order (n = 2^log_n values each): the copy-permutation
 * (variable) columns and their sigma columns; non-residues k_j from bj_non_residues_for_copy_permutation.
 * chunk_size = quotient degree (columns multiplied per partial product).
 * Outputs: z (c0, c1) = exclusive prefix product of prod_j (w_j + beta k_j x + gamma)/(w_j + beta sigma_j + gamma), and the
 * ceil(n_cols/chunk_size) - 1 partial products, laid out [partial][c0|c1][n] in d_partials.
 * Returns BJ_ERR_INVALID_ARG if the grand product over the domain is not 1 (the reference asserts, :479). Synchronises. */
BJ_API int32_t bj_non_residues_for_copy_permutation(uint64_t domain_size, uint32_t num_columns, uint64_t* h_out);
BJ_API int32_t bj_copy_permutation_stage2(bj_ctx* ctx, const uint64_t* const* h_variable_cols, const uint64_t* const* h_sigma_cols,
                                   uint32_t n_cols, const uint64_t* h_non_residues, const uint64_t h_beta[2],
                                   const uint64_t h_gamma[2], uint32_t log_n, uint32_t chunk_size, uint64_t* d_z_c0,
                                   uint64_t* d_z_c1, uint64_t* d_partials);

/* ---- lookup argument over specialised columns, table id in a constant column
 *      (LookupParameters::UseSpecializedColumnsWithTableIdAsConstant; src/cs/implementations/lookup_argument_in_ext.rs) ----
 * stage 2 (compute_lookup_poly_pairs_specialized, :320-947), trace domain, natural order:
 *   A_i[r] = 1/(beta + sum_j gamma^j col_{i,j}[r] + gamma^w table_id[r]),  B[r] = m[r]/(beta + sum_j gamma^j t_j[r]).
 * h_lookup_cols: n_subarguments*width device pointers (the sub-arguments' variable columns, in order); d_table_id_col: the
 * constant column with the table id (NULL if none); h_table_cols: the n_table_cols = width (+1) lookup-table setup columns.
 * d_out: [n_subarguments + 1][c0|c1][n]  (A_0, ..., A_{k-1}, B). */
BJ_API int32_t bj_lookup_polys_specialized(bj_ctx* ctx, const uint64_t* const* h_lookup_cols, uint32_t n_subarguments, uint32_t width,
                                    const uint64_t* d_table_id_col, const uint64_t* const* h_table_cols, uint32_t n_table_cols,
                                    const uint64_t* d_multiplicity, const uint64_t h_beta[2], const uint64_t h_gamma[2],
                                    uint32_t log_n, uint64_t* d_out);
/* quotient terms (compute_quotient_terms_for_lookup_specialized, :949-1319) on the first n_points = Q*n points of the LDE:
 *   q += alpha_i (A_i (beta + sum gamma^j col_ij + gamma^w id) - 1)  for every sub-argument,  + alpha_k (B (beta + sum gamma^j t_j) - m).
 * All column arguments are LDE columns; h_a_ldes: 2*n_subarguments pointers (c0, c1); h_alphas: n_subarguments + 1 Fp2. */
BJ_API int32_t bj_quotient_lookup_specialized(bj_ctx* ctx, const uint64_t* const* h_lookup_ldes, uint32_t n_subarguments, uint32_t width,
                                       const uint64_t* d_table_id_lde, const uint64_t* const* h_table_ldes, uint32_t n_table_cols,
                                       const uint64_t* d_multiplicity_lde, const uint64_t* const* h_a_ldes, const uint64_t* d_b_c0,
                                       const uint64_t* d_b_c1, const uint64_t h_beta[2], const uint64_t h_gamma[2],
                                       const uint64_t* h_alphas, uint64_t n_points, uint64_t* d_q_c0, uint64_t* d_q_c1);

/* ---- gate / quotient evaluator over general-purpose columns: the row loop of prove_cpu_basic
 *      (src/cs/implementations/prover.rs:1031-1080; gates over specialised columns, :653-801, use the same call) with GateConstraintEvaluator::evaluate_once (src/cs/traits/evaluator.rs:145-152)
 *      supplied as DATA: the SSA program recorded by the reference's own GPU hook, gpu_synthesizer::GPUDataCapture
 *      (src/gpu_synthesizer/mod.rs:115-133 Index / Relation, :354-443 capture). */
enum { /* Index<F> (gpu_synthesizer/mod.rs:115-121); SHARED = a ConstantPoly listed in row_shared_constants_set */
  BJ_IDX_VARIABLE = 0, BJ_IDX_WITNESS = 1, BJ_IDX_CONSTANT_POLY = 2, BJ_IDX_TEMPORARY = 3, BJ_IDX_CONSTANT_VALUE = 4,
  BJ_IDX_CONSTANT_POLY_SHARED = 5
};
enum { /* Relation<F> (gpu_synthesizer/mod.rs:125-133) */
  BJ_REL_ADD = 0, BJ_REL_DOUBLE = 1, BJ_REL_SUB = 2, BJ_REL_NEGATE = 3, BJ_REL_MUL = 4, BJ_REL_SQUARE = 5, BJ_REL_INVERSE = 6
};
typedef struct bj_gate_index {
  uint32_t kind;   /* BJ_IDX_* */
  uint32_t reserved;
  uint64_t value;  /* column / temporary index, or the field element for BJ_IDX_CONSTANT_VALUE */
} bj_gate_index;
typedef struct bj_gate_relation {
  uint32_t op;            /* BJ_REL_* */
  uint32_t dst_temporary; /* TemporaryValue index this relation defines: programs are SSA (one fresh index per relation, < 2^20);
                           * the library assigns slots by liveness - at most 128 temporaries may be live at once */
  bj_gate_index a, b;     /* b ignored by the unary relations */
} bj_gate_relation;
typedef struct bj_gate_desc {
  const bj_gate_relation* relations; /* GPUDataCapture::relations, in recording order */
  uint32_t n_relations;
  uint32_t n_writes;
  const bj_gate_index* writes;       /* GPUDataCapture::writes_per_repetition (one quotient term each) */
  uint32_t num_repetitions;          /* num_repetitions_on_row */
  uint32_t variables_offset;         /* PerChunkOffset (GatePlacementType::MultipleOnRow) */
  uint32_t witnesses_offset;
  uint32_t constants_offset;
  uint32_t constants_placement_offset; /* first constant column of the gate = selector path length (prover.rs:1000-1013) */
  uint32_t selector_path_len;          /* TreeNode path of the gate; 0 = no selector */
  const uint8_t* selector_path;        /* path[i] != 0: factor const_i, else (1 - const_i) (prover.rs:2775-2916) */
  /* gates on SPECIALISED columns (GatePlacementStrategy::UseSpecializedColumns, prover.rs:653-801): first column of
   * repetition 0 among the variable / witness columns (initial_offset of offsets_for_specialized_evaluators); 0 for
   * general-purpose gates.  Such a gate has selector_path_len = 0, constants_placement_offset = (constants of the
   * general-purpose gates) + initial constants offset, and constants_offset = 0 when share_constants. */
  uint32_t variables_initial_offset;
  uint32_t witnesses_initial_offset;
} bj_gate_desc;
/* For every point t < n_points (= Q * n, the first Q cosets of the LDE, flat coset-major):
 *   q[t] += sum_g selector_g(t) * sum_k alpha_pow[k] * term_k(t),  k running over the terms of all gates in order
 * (gates with zero terms, e.g. NOP, are simply not passed).  Column pointer arrays are HOST arrays of DEVICE pointers,
 * each column flat [coset][row] as produced by bj_lde.  h_alpha_powers: (c0, c1) pairs, one per term. */
BJ_API int32_t bj_quotient_gates_general_purpose(bj_ctx* ctx, const bj_gate_desc* h_gates, uint32_t n_gates,
                                          const uint64_t* const* h_variable_cols, uint32_t n_variables,
                                          const uint64_t* const* h_witness_cols, uint32_t n_witnesses,
                                          const uint64_t* const* h_constant_cols, uint32_t n_constants,
                                          const uint64_t* h_alpha_powers, uint32_t n_alpha_powers,
                                          uint64_t n_points, uint64_t* d_q_c0, uint64_t* d_q_c1);
/* What the evaluator will execute for these gates - pure host code, no device needed.  The recorded programs are validated
 * (SSA, operand ranges against n_variables / n_witnesses / n_constants), rewritten (peephole bits: 1 = x*1, x+0, x*0 become
 * aliases; 2 = a product whose only use is a sum becomes a multiply-add; 4 = trees of single-use sums over products with
 * immediates < 2^28 become one linear combination; 8 = a value read only by push_evaluation_result is pushed by the step that
 * computes it; 15 = what bj_quotient_gates_general_purpose uses) and lowered to 32-byte records = 4 little-endian u64:
 *   word 0: code (bits 0-6) | pushes-its-value flag (bit 7) | destination slot or term index (bits 8-31) |
 *           per-repetition column stride of operand a (bits 32-47) and b (bits 48-63);  word 1: operand a;  word 2: operand b;
 *   word 3: addend slot of a multiply-add.  Operand classes T (slot), L (index into [variables | witnesses | constants] at
 *   repetition 0), I (immediate).  Codes: ADD 0-8, SUB 9-17, MUL 18-26 = base + 3 * class(a) + class(b) with T, L, I = 0, 1, 2;
 *   DOUBLE 27-29, NEGATE 30-32, SQUARE 33-35, INVERSE 36-38, MOVE 39-41 = base + class(a); multiply-add a * b + slot 42-47 =
 *   42 + 3 * class(a) + class(b), class(a) in {T, L}; 48 = linear combination: word 0 stride = variables stride, word 1 = number
 *   of terms n, word 2 = constant term, followed by ceil(n / 4) records of four (u32 ref, u32 k) pairs, ref = slot or
 *   0x80000000 | variable column.
 * h_records may be NULL (sizes only); h_gate_first_record has n_gates + 1 entries.  Used by the CPU test suite, which runs an
 * emulator of this format against the gate evaluators (tests/test_gate_compiler_cpu.py). */
BJ_API int32_t bj_gate_programs_compile(const bj_gate_desc* h_gates, uint32_t n_gates, uint32_t n_variables, uint32_t n_witnesses,
                                 uint32_t n_constants, uint32_t peephole, uint64_t* h_records, uint64_t capacity_records,
                                 uint64_t* n_records, uint32_t* h_gate_first_record, uint32_t* max_live_temporaries);

/* ---- quotient: copy-permutation relations + the z(1) = 1 term (src/cs/implementations/copy_permutation.rs:1000-1249,
 *      src/cs/implementations/prover.rs:1189-1227) over the first 2^log_quotient_degree cosets of the LDE ----
 * All polynomial arguments are LDE columns (flat [L][n], 2^log_lde cosets).  h_partial_ldes: 2*(n_chunks-1) device
 * pointers (c0, c1 per partial product).  h_alphas: (n_chunks + 1) Fp2 challenge powers: the z(1)=1 term first, then one
 * per copy-permutation relation (order of prover.rs:1176-1250).  Accumulates into q. */
BJ_API int32_t bj_quotient_copy_permutation(bj_ctx* ctx, const uint64_t* const* h_variable_ldes, const uint64_t* const* h_sigma_ldes,
                                     uint32_t n_cols, const uint64_t* h_non_residues, const uint64_t* d_z_c0,
                                     const uint64_t* d_z_c1, const uint64_t* const* h_partial_ldes, const uint64_t h_beta[2],
                                     const uint64_t h_gamma[2], const uint64_t* h_alphas, uint32_t log_n, uint32_t log_lde,
                                     uint32_t log_quotient_degree, uint32_t chunk_size, uint64_t* d_q_c0, uint64_t* d_q_c1);
/* divide_by_vanishing_for_bitreversed_coset_enumeration (src/cs/implementations/utils.rs:770-817): q[coset j] *= 1/((7 w^bitrev(j))^n - 1) */
BJ_API int32_t bj_quotient_divide_by_vanishing(bj_ctx* ctx, uint64_t* d_q_c0, uint64_t* d_q_c1, uint32_t log_n,
                                        uint32_t log_quotient_degree);

/* ---- openings: values of n_cols base-field polynomials at the Fp2 point `at`, from their first LDE coset
 *      (barycentric evaluation, src/cs/implementations/utils.rs:907-1243; prover.rs:1519-1802).
 * h_cols: host array of device pointers to LDE columns (only the first 2^log_n values, coset 0, are read).
 * h_out: n_cols (c0, c1) pairs.  `at` must not lie on the coset 7<w_n> (on a sharded context: on the coset of local slot 0).
 * Synchronises. */
BJ_API int32_t bj_barycentric_evaluate(bj_ctx* ctx, const uint64_t* const* h_cols, uint32_t n_cols, uint32_t log_n,
                                const uint64_t h_at[2], uint64_t* h_out);

/* ---- host-buffer convenience entry points (what the Rust shim calls when columns live in host Vecs).
 * They upload, run, download and synchronise; used for the end-to-end measurement. */
BJ_API int32_t bj_ntt_natural_to_bitreversed_host(bj_ctx* ctx, uint64_t* h_data, uint32_t log_n, uint32_t n_cols,
                                           uint64_t coset);
BJ_API int32_t bj_intt_natural_to_natural_host(bj_ctx* ctx, uint64_t* h_data, uint32_t log_n, uint32_t n_cols,
                                        uint64_t coset);

/* ---- host-side Fiat-Shamir (never on the GPU; must be replayed bit-exactly):
 * GoldilocksPoisedon2Transcript = AlgebraicSpongeBasedTranscript<_, 8, 12, 4, Poseidon2, Overwrite>
 * (src/cs/implementations/transcript.rs:62-129), BoolsBuffer::get_bits (:369-417), compute_fri_schedule
 * (src/cs/implementations/prover.rs:2281-2372). */
typedef struct bj_transcript bj_transcript;
BJ_API bj_transcript* bj_transcript_new(void);          /* GoldilocksPoisedon2Transcript */
/* Blake2sTranscript (transcript.rs:155-260; the transcript of sha256_bench_non_recursive): field elements enter as the
 * 8 LE bytes of their reduced value, caps as raw 32-byte digests; a challenge is 8 output bytes reduced mod p; query bits
 * take all 64 bits of 8 challenge bytes (BoolsBuffer, non-algebraic branch). */
BJ_API bj_transcript* bj_transcript_new_blake2s(void);
BJ_API bj_transcript* bj_transcript_new_keccak256(void); /* Keccak256Transcript (transcript.rs:262-367): same scheme, Keccak-256 */
/* GoldilocksPoisedonTranscript (transcript.rs:131-138): the algebraic sponge transcript over the Poseidon (v1) permutation
 * (src/implementations/poseidon_goldilocks_naive.rs) - the TR of run_sha256_prover_recursive_mode and
 * run_sha256_prover_recursive_mode_poseidon2 (src/gadgets/sha256/mod.rs:275-293) */
BJ_API bj_transcript* bj_transcript_new_poseidon(void);
BJ_API void bj_transcript_free(bj_transcript* t);
BJ_API void bj_transcript_witness_field_elements(bj_transcript* t, const uint64_t* els, size_t n);
BJ_API void bj_transcript_witness_merkle_tree_cap(bj_transcript* t, const uint64_t* cap_digests, size_t n_digests);
BJ_API uint64_t bj_transcript_get_challenge(bj_transcript* t);
/* num_bits query-index bits, LSB first, as one integer; each refill keeps the (64 - max_needed) low bits of a challenge */
BJ_API uint64_t bj_transcript_get_index_bits(bj_transcript* t, uint32_t num_bits, uint32_t max_needed);
/* schedule: caller array of >= 32 entries */
BJ_API int32_t bj_compute_fri_schedule(uint32_t security_bits, uint32_t cap_size, uint32_t pow_bits, uint32_t rate_log_two,
                                uint32_t initial_degree_log_two, uint32_t* new_pow_bits, uint32_t* num_queries,
                                uint32_t* schedule, uint32_t* schedule_len, uint32_t* final_degree);

/* ---- FRI commit phase: do_fri (src/cs/implementations/fri/mod.rs:49-357) ----
 * d_c0 / d_c1: the DEEP codeword on the full LDE domain (2^log_full_size values each, LDE layout; borrowed - must stay
 * alive as long as the returned oracles are queried).  schedule: interpolation_log2s_schedule (compute_fri_schedule).
 * Builds the base oracle and every intermediate oracle (Poseidon2 trees with 2^k c0 values then 2^k c1 values per
 * leaf), absorbs each cap into the transcript, draws the two challenge elements, folds, and finally bit-reverses +
 * iNTTs the last vector into the monomial forms which are absorbed as well.  Returns BJ_ERR_INVALID_ARG if the folded
 * codeword is not of low degree (the reference's self-check panics, fri/mod.rs:326-334). */
typedef struct bj_fri_oracles bj_fri_oracles;
BJ_API int32_t bj_do_fri(bj_ctx* ctx, bj_transcript* transcript, const uint64_t* d_c0, const uint64_t* d_c1,
                  uint32_t log_full_size, const uint32_t* schedule, uint32_t n_schedule, uint32_t log_lde,
                  uint32_t cap_size, bj_fri_oracles** out);
#define BJ_HASHER_POSEIDON2 0u
#define BJ_HASHER_BLAKE2S 1u
#define BJ_HASHER_KECCAK256 2u
/* same with the tree hasher chosen (BJ_HASHER_*: GoldilocksPoseidon2Sponge, Blake2s256 or Keccak256, src/cs/oracle/mod.rs:114-313) */
BJ_API int32_t bj_do_fri_with_hasher(bj_ctx* ctx, bj_transcript* transcript, const uint64_t* d_c0, const uint64_t* d_c1,
                              uint32_t log_full_size, const uint32_t* schedule, uint32_t n_schedule, uint32_t log_lde,
                              uint32_t cap_size, uint32_t hasher, bj_fri_oracles** out);
BJ_API void bj_fri_oracles_free(bj_fri_oracles* o);
BJ_API uint32_t bj_fri_oracles_num_oracles(const bj_fri_oracles* o);
BJ_API uint32_t bj_fri_oracles_num_monomials(const bj_fri_oracles* o);
BJ_API int32_t bj_fri_oracles_get_cap(const bj_fri_oracles* o, uint32_t oracle_idx, uint64_t* h_out /* 4*cap u64 */);
BJ_API int32_t bj_fri_oracles_get_monomials(const bj_fri_oracles* o, uint64_t* h_c0, uint64_t* h_c1);
BJ_API int32_t bj_fri_oracles_get_challenges(const bj_fri_oracles* o, uint64_t* h_out /* 2 u64 per oracle */);
/* OracleQuery::construct for one FRI oracle (src/cs/implementations/proof.rs:65-97): leaf elements (2 * 2^k u64:
 * c0 values then c1 values) and the sibling path (path_len digests, bottom-up, cap level excluded). */
BJ_API int32_t bj_fri_oracles_query(bj_fri_oracles* o, uint32_t oracle_idx, uint64_t leaf_index, uint64_t* h_leaf_elements,
                             uint64_t* h_path, uint32_t* path_len);
/* n leaves of one oracle at once (two device round trips): h_leaf_elements [n][2 * 2^k], h_paths [n][*path_len][4] */
BJ_API int32_t bj_fri_oracles_query_batch(bj_fri_oracles* o, uint32_t oracle_idx, const uint64_t* h_leaf_indices, uint32_t n_indices,
                                   uint64_t* h_leaf_elements, uint64_t* h_paths, uint32_t* path_len);
/* Query helpers for the base oracles (witness / stage 2 / quotient / setup): gather the leaf preimages of n_indices
 * leaves (h_out[q][s * elems_per_leaf + e]) and their Merkle paths (h_out[q][depth][4]); both synchronise.  Every source
 * column holds n_leaves * elems_per_leaf elements; an index >= n_leaves is rejected with BJ_ERR_INVALID_ARG (no launch). */
BJ_API int32_t bj_query_leaf_elements(bj_ctx* ctx, const uint64_t* const* h_sources, uint32_t n_sources, uint32_t elems_per_leaf,
                               uint64_t n_leaves, const uint64_t* h_indices, uint32_t n_indices, uint64_t* h_out);
BJ_API int32_t bj_merkle_paths(bj_ctx* ctx, const uint64_t* d_leaf_hashes, const uint64_t* d_nodes, uint64_t n_leaves,
                        uint32_t cap_size, const uint64_t* h_indices, uint32_t n_indices, uint64_t* h_out);

/* ---- proof of work: impl PoWRunner for Blake2s256 (src/cs/implementations/pow.rs:52-147).  Returns the smallest u64
 * `challenge` of the first successful 2^24-candidate batch for which the first 8 bytes (little endian) of
 * Blake2s-256(seed || challenge.to_le_bytes()) have at least pow_bits (<= 32) trailing zero bits.  The seed is the byte
 * string of run_from_field_elements (the LE bytes of the reduced field elements; 5 challenges = 40 bytes in
 * prove_cpu_basic, prover.rs:2109-2126).  Synchronises. */
BJ_API int32_t bj_pow_blake2s(bj_ctx* ctx, const uint8_t* h_seed, uint32_t seed_len, uint32_t pow_bits, uint64_t* h_challenge);
/* impl PoWRunner for Keccak256 (pow.rs:140-230): the same search with Keccak-256 (seed <= 120 bytes) */
BJ_API int32_t bj_pow_keccak256(bj_ctx* ctx, const uint8_t* h_seed, uint32_t seed_len, uint32_t pow_bits, uint64_t* h_challenge);

/* ---- setup / witness materialisation on the device (what feeds bj_setup_create and bj_prove) ----
 * Variable encoding as in the reference (src/cs/mod.rs:44-47, :155-180): a u64 whose bit 63 marks a placeholder and whose
 * low 48 bits are the variable index.
 * bj_materialize_columns: materialize_variables_polynomials_from_dense_hint (src/cs/implementations/witness.rs:325-385):
 *   d_out[c][row] = d_all_values[hint[c][row]] for row < hint_rows (hint laid out [n_cols][hint_rows]); placeholders and
 *   the rows beyond the hint are zero.  Fails if a hint points past n_values.  Synchronises.
 * bj_create_permutation_polys: create_permutation_polys (src/cs/implementations/setup.rs:419-502): d_placement is
 *   copy_permutation_data, [n_cols][2^log_n] variables; d_sigmas [n_cols][2^log_n] receives the sigma columns (identity
 *   k_c * w^row on cells that are never copied, one cycle per variable over its occurrences in column-major order). */
BJ_API int32_t bj_materialize_columns(bj_ctx* ctx, const uint64_t* d_all_values, uint64_t n_values, const uint64_t* d_hint,
                               uint32_t n_cols, uint64_t hint_rows, uint32_t log_n, uint64_t* d_out);
BJ_API int32_t bj_create_permutation_polys(bj_ctx* ctx, const uint64_t* d_placement, uint32_t n_cols, uint32_t log_n,
                                    uint64_t* d_sigmas);

/* ---- the prover entry point: CSReferenceAssembly::prove_cpu_basic (src/cs/implementations/prover.rs:153-2269) and the part of
 *      the setup materialisation it depends on (setup.rs:1093-1255: sigma / constant / lookup-table columns -> LDE -> setup tree).
 * Scope: gates on general-purpose columns (bj_gate_desc programs), copy permutation over all variable columns, optional
 * log-derivative lookup over specialised columns with the table id in a constant column (lookup_width = 0: none), Poseidon2
 * or Blake2s tree hasher and transcript, public inputs, Blake2s proof of work (pow_bits > 0).  Host C++ inside the library: transcript, schedule, query
 * indices and proof assembly never leave the host; every heavy step is one of the entry points above.
 * Column arguments are DEVICE arrays [column][2^log_n] in natural row order.  bj_setup BORROWS d_sigmas / d_constants /
 * d_lookup_tables (stage 2 reads them again): they must outlive the setup.  The gate programs are copied.
 * bj_prove returns BJ_ERR_INVALID_ARG for an unsatisfied circuit (the reference panics, prover.rs:1425-1438). */
typedef struct bj_circuit {
  uint32_t log_n;            /* trace length 2^log_n */
  uint32_t num_variables;    /* columns under the copy permutation (general purpose + specialised lookup columns) */
  uint32_t num_constants;
  uint32_t quotient_degree;  /* power of two; may exceed fri_lde_factor (production: factor 2, degree 8) - columns are then evaluated
                              * at max(fri_lde_factor, quotient_degree) cosets and the oracles commit to the first fri_lde_factor
                              * of them (prover.rs:178-196 used_lde_degree / subset_for_degree) */
  uint32_t fri_lde_factor, merkle_tree_cap_size, security_level, pow_bits; /* ProofConfig (prover.rs:55-73) */
  const bj_gate_desc* gates;
  uint32_t n_gates;
  uint32_t lookup_width;            /* columns per lookup tuple without the table id; 0 = no lookup argument */
  uint32_t lookup_num_repetitions;  /* sub-arguments */
  uint32_t lookup_variables_offset; /* first lookup column among the variables */
  uint32_t lookup_table_id_column;  /* constant column holding the table id */
  /* public inputs: places (variable column, row) whose witness values are published (CSReferenceAssembly::public_inputs;
   * prover.rs:264-266, 1805-1821, 2010-2041) */
  const uint32_t* public_input_columns;
  const uint32_t* public_input_rows;
  uint32_t n_public_inputs;
  uint32_t tree_hasher; /* BJ_HASHER_POSEIDON2 (recursive-mode bench), BJ_HASHER_BLAKE2S (sha256_bench_non_recursive), BJ_HASHER_KECCAK256 */
  uint32_t transcript;  /* 0: Poseidon2 sponge transcript, 1: Blake2sTranscript, 2: Keccak256Transcript, 3: Poseidon (v1) sponge transcript */
} bj_circuit;
typedef struct bj_setup bj_setup;
typedef struct bj_proof bj_proof;
BJ_API int32_t bj_setup_create(bj_ctx* ctx, const bj_circuit* circuit, const uint64_t* d_sigmas, const uint64_t* d_constants,
                        const uint64_t* d_lookup_tables /* [lookup_width + 1][n] or NULL */, bj_setup** out);
BJ_API void bj_setup_free(bj_setup* setup);
BJ_API int32_t bj_setup_get_cap(const bj_setup* setup, uint64_t* h_cap /* 4 * cap_size u64: VerificationKey::setup_merkle_tree_cap */);
BJ_API int32_t bj_prove(bj_ctx* ctx, const bj_setup* setup, const uint64_t* d_variables, const uint64_t* d_multiplicities /* or NULL */,
                 bj_proof** out);
BJ_API void bj_proof_free(bj_proof* proof);
/* the proof in the reference's serde_json shape (Proof<F, H, EXT>, src/cs/implementations/proof.rs:57-143).
 * Call with buf == NULL to learn the size (incl. the terminating 0), then with a buffer of at least that size. */
BJ_API int32_t bj_proof_to_json(const bj_proof* proof, char* buf, size_t capacity, size_t* needed);
/* wall-clock seconds of the six stages (witness, stage 2, quotient, openings, DEEP+FRI, queries), device work included */
BJ_API int32_t bj_proof_stage_seconds(const bj_proof* proof, double out[6]);

/* device self-test: PTX field arithmetic vs the portable C versions on n pseudo-random + edge inputs */
BJ_API int32_t bj_selftest_field(bj_ctx* ctx, uint64_t n, uint64_t seed, uint64_t* h_mismatches);

/* ---- host self-test hooks: the same gl64 source compiled for the host (CPU tests, no device needed) ---- */
BJ_API uint64_t bj_host_gl_mul(uint64_t a, uint64_t b);
BJ_API uint64_t bj_host_gl_add(uint64_t a, uint64_t b);
BJ_API uint64_t bj_host_gl_sub(uint64_t a, uint64_t b);
BJ_API uint64_t bj_host_gl_inv(uint64_t a);
BJ_API uint64_t bj_host_gl_mul_pow2(uint64_t a, uint32_t s);
BJ_API void bj_host_e2_mul(const uint64_t a[2], const uint64_t b[2], uint64_t out[2]);
BJ_API void bj_host_e2_inv(const uint64_t a[2], uint64_t out[2]);
BJ_API void bj_host_poseidon2_permutation(uint64_t state[12]);
BJ_API void bj_host_poseidon_permutation(uint64_t state[12]); /* Poseidon (v1), poseidon_goldilocks_naive.rs:154-165 */
BJ_API void bj_host_keccak256(const uint8_t* data, size_t n, uint8_t out[32]);

#ifdef __cplusplus
}
#endif
#endif /* BOOJUM_B200_H */
