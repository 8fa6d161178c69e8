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
 *   - there is no CPU fallback: without a CUDA device bj_ctx_create fails with BJ_ERR_NO_DEVICE.
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
 * legacy default stream.  The library never creates streams of its own. */
BJ_API int32_t bj_ctx_create(int32_t device, void* stream, bj_ctx** out_ctx);
BJ_API int32_t bj_ctx_destroy(bj_ctx* ctx);
BJ_API int32_t bj_ctx_set_stream(bj_ctx* ctx, void* stream);
BJ_API int32_t bj_ctx_synchronize(bj_ctx* ctx);
BJ_API const char* bj_last_error(const bj_ctx* ctx);
/* number of kernels this library launched through ctx so far (for launch accounting) */
BJ_API uint64_t bj_launch_count(const bj_ctx* ctx);

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

/* ---- host-buffer convenience entry points (what the Rust shim calls when columns live in host Vecs).
 * They upload, run, download and synchronise; used for the end-to-end measurement. */
BJ_API int32_t bj_ntt_natural_to_bitreversed_host(bj_ctx* ctx, uint64_t* h_data, uint32_t log_n, uint32_t n_cols,
                                           uint64_t coset);
BJ_API int32_t bj_intt_natural_to_natural_host(bj_ctx* ctx, uint64_t* h_data, uint32_t log_n, uint32_t n_cols,
                                        uint64_t coset);

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

#ifdef __cplusplus
}
#endif
#endif /* BOOJUM_B200_H */
