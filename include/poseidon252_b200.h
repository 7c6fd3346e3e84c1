/* poseidon252_b200 -- C ABI of the B200-native batched Poseidon/Hades engine.
 *
 * Drop-in boundary for the hot path of dusk-poseidon (reference = /root/reference, a pure-Rust,
 * one-state-at-a-time CPU crate with no FFI of its own).  The reference's seam for this path is
 * the trait pair dusk_safe::Safe<BlsScalar,5> (impl: src/hades/permutation/scalar.rs:24-36) +
 * Hades<BlsScalar> (src/hades/permutation.rs:34-124) under the public surface src/lib.rs:13-31.
 * This header is what a Rust `extern "C"` block for the batch entry points
 * (hades::permute_batch, Hash::digest_batch, encrypt_batch, decrypt_batch, merkle4) binds; the
 * binding itself is in bindings/rust/ and INTEGRATION.md.
 *
 * Conventions
 *   - p252_fr is bit-identical to `BlsScalar.0`: 4 x u64 little-endian limbs of x*R mod p
 *     (Montgomery form, R = 2^256 mod p, value < p).  No conversion happens at the boundary.
 *   - All batch buffers are item-major arrays (the layout of `&[BlsScalar]`, src/hash.rs:94).
 *   - The caller owns every buffer; the library owns only the context (reference borrows inputs,
 *     src/hash.rs:94, and returns fresh Vecs, src/hash.rs:128).
 *   - `flags` says where the buffers live: P252_MEM_HOST (library stages H2D/D2H itself) or
 *     P252_MEM_DEVICE (pointers are device pointers of ctx's GPU, 16-byte aligned; add
 *     P252_ASYNC to return right after enqueueing on the context's stream).
 *   - Every function returns a p252_status; nothing unwinds across the boundary.  Positive codes
 *     mirror dusk_poseidon::Error (src/error.rs:11-32); negative codes are engine failures.
 *   - There is NO CPU fallback: without a usable sm_100 device p252_create fails.
 *   - A context is bound to one device and one stream; calls on one context serialise; separate
 *     contexts are independent (the reference is stateless: ScalarPermutation is a ZST,
 *     src/hades/permutation/scalar.rs:15).
 */
#ifndef POSEIDON252_B200_H
#define POSEIDON252_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define P252_WIDTH 5 /* dusk_poseidon::HADES_WIDTH, src/hades.rs:34 */

typedef struct p252_fr {
    uint64_t l[4];
} p252_fr;

typedef struct p252_ctx p252_ctx;

typedef enum p252_status {
    P252_OK = 0,
    /* dusk_poseidon::Error, src/error.rs:11-32 */
    P252_ERR_IO_PATTERN_VIOLATION = 1,
    P252_ERR_INVALID_IO_PATTERN = 2,
    P252_ERR_TOO_FEW_INPUT_ELEMENTS = 3,
    P252_ERR_ENCRYPTION_FAILED = 4,
    P252_ERR_DECRYPTION_FAILED = 5,
    P252_ERR_INVALID_POINT = 6,
    /* engine */
    P252_ERR_INVALID_ARGUMENT = -1,
    P252_ERR_CUDA = -2,
    P252_ERR_NCCL = -3,
    P252_ERR_NO_DEVICE = -4,
    P252_ERR_OUT_OF_MEMORY = -5
} p252_status;

/* u64::from(Domain), src/hash.rs:38-56 */
typedef enum p252_domain {
    P252_DOMAIN_MERKLE4 = 0,
    P252_DOMAIN_MERKLE2 = 1,
    P252_DOMAIN_ENCRYPTION = 2,
    P252_DOMAIN_OTHER = 3
} p252_domain;

enum { P252_MEM_HOST = 0, P252_MEM_DEVICE = 1, P252_ASYNC = 2 };

/* ---- library / context ------------------------------------------------------------------- */
const char* p252_version(void);
const char* p252_strerror(int status);
int p252_device_count(int* count);

/* Create a context on CUDA device `device` with its own stream.  Fails with P252_ERR_NO_DEVICE
 * when there is no sm_100 GPU (no CPU fallback). */
int p252_create(int device, p252_ctx** out);
/* Same, but enqueue all work on an existing CUDA stream (cudaStream_t passed as void*), e.g. the
 * caller's torch stream, so that the caller's CUDA events bracket the kernels. */
int p252_create_on_stream(int device, void* cuda_stream, p252_ctx** out);
void p252_destroy(p252_ctx* ctx);
int p252_sync(p252_ctx* ctx);
/* Text of the last CUDA/NCCL failure on this context ("" if none). */
const char* p252_last_error(const p252_ctx* ctx);
/* Number of kernels this context has launched since creation. */
uint64_t p252_launch_count(const p252_ctx* ctx);
/* Pinned host memory for P252_MEM_HOST callers that want full PCIe bandwidth. */
int p252_host_alloc(size_t bytes, void** out);
int p252_host_free(void* p);

/* ---- host-side sponge bookkeeping (no GPU needed) ----------------------------------------- */
/* u64::from(Domain), src/hash.rs:43-55 */
int p252_domain_separator(int domain, uint64_t* out);
/* dusk-safe tag input: `calls` are the io-pattern, absorb(len) = 0x80000000|len, squeeze(len) = len
 * (as produced by io_pattern, src/hash.rs:62-85); consecutive calls of one kind aggregate.
 * Writes the byte string hashed into the tag; *out_len in = capacity, out = length. */
int p252_tag_input(const uint32_t* calls, size_t ncalls, uint64_t domain_sep, uint8_t* out, size_t* out_len);
/* BlsScalar::hash_to_scalar (src/hades/permutation/scalar.rs:29-31): BLAKE2b-512 -> mod p. */
int p252_hash_to_scalar(const uint8_t* bytes, size_t len, p252_fr* out);
/* Safe::tag of the pattern: hash_to_scalar(tag_input(calls, domain_sep)). */
int p252_tag(const uint32_t* calls, size_t ncalls, uint64_t domain_sep, p252_fr* tag);
/* io_pattern(domain, [in_len], out_len) + tag (src/hash.rs:62-85,131-137): checks the Merkle
 * arities (-> P252_ERR_IO_PATTERN_VIOLATION) and zero lengths (-> P252_ERR_INVALID_IO_PATTERN). */
int p252_hash_tag(int domain, size_t in_len, size_t out_len, p252_fr* tag);
/* tag of dusk_safe::encrypt/decrypt for message length L (src/encryption.rs:67-73). */
int p252_encryption_tag(size_t L, p252_fr* tag);

/* ---- batch entry points (the GPU path) ----------------------------------------------------- */
/* hades::permute_batch: n independent Safe::permute calls (src/hades/permutation/scalar.rs:25-27
 * -> Hades::perm, src/hades/permutation.rs:105-123).  states: n x 5, in place. */
int p252_permute_batch(p252_ctx* ctx, p252_fr* states, size_t n, int flags);
/* The reference's dense formulation executed on the device (cross-check / cost comparison). */
int p252_permute_batch_dense(p252_ctx* ctx, p252_fr* states, size_t n, int flags);

/* Sponge with a caller-supplied tag: start(tag) -> absorb(in_len) -> squeeze(out_len)
 * (Hash::finalize, src/hash.rs:128-155).  in: n x in_len, out: n x out_len. */
int p252_digest_batch(p252_ctx* ctx, const p252_fr* tag, const p252_fr* in, size_t n, size_t in_len,
                      p252_fr* out, size_t out_len, int flags);
/* Hash::digest_batch: n x Hash::digest(domain, in[i]) with Hash::output_len(out_len)
 * (src/hash.rs:111-115,191-195); tag computed on the host once per batch. */
int p252_hash_batch(p252_ctx* ctx, int domain, const p252_fr* in, size_t n, size_t in_len, p252_fr* out,
                    size_t out_len, int flags);

/* Hash::digest_truncated batch (src/hash.rs:164-183,203-210): every output scalar is taken out of Montgomery
 * form and masked to 250 bits; out_raw receives the raw limbs the reference passes to JubJubScalar::from_raw. */
int p252_hash_batch_truncated(p252_ctx* ctx, int domain, const p252_fr* in, size_t n, size_t in_len, p252_fr* out_raw,
                              size_t out_len, int flags);

/* Wire format (BlsScalar::from_bytes / to_bytes as used at src/hades.rs:94-105,131): n canonical 32-byte
 * little-endian integers <-> BlsScalar.0.  from_bytes: ok[i] = 0 and out[i] = 0 when the value is >= p (the
 * reference returns None); ok may be NULL. */
int p252_scalars_from_bytes(p252_ctx* ctx, const uint8_t* bytes, size_t n, p252_fr* out, uint8_t* ok, int flags);
int p252_scalars_to_bytes(p252_ctx* ctx, const p252_fr* in, size_t n, uint8_t* bytes, int flags);

/* encrypt_batch: n x encrypt(msg[i], (u,v)[i], nonce[i]) (src/encryption.rs:62-74).
 * msg: n x L, secret_uv: n x 2 (JubJubAffine::get_u/get_v), nonce: n, cipher: n x (L+1). */
int p252_encrypt_batch(p252_ctx* ctx, const p252_fr* msg, size_t n, size_t L, const p252_fr* secret_uv,
                       const p252_fr* nonce, p252_fr* cipher, int flags);
/* decrypt_batch (src/encryption.rs:83-95).  cipher: n x (L+1), msg: n x L, ok: n bytes; ok[i] = 0
 * <=> the reference returns Error::DecryptionFailed for item i (its msg is zeroed).  Returns
 * P252_OK even when some items fail; *n_failed (optional) receives their count (HOST flags only). */
int p252_decrypt_batch(p252_ctx* ctx, const p252_fr* cipher, size_t n, size_t L, const p252_fr* secret_uv,
                       const p252_fr* nonce, p252_fr* msg, uint8_t* ok, size_t* n_failed, int flags);

/* One level of an arity-4 tree: parents[i] = Hash::digest(Domain::Merkle4, children[4i..4i+4])
 * (src/hash.rs:22-26). */
int p252_merkle4_level(p252_ctx* ctx, const p252_fr* children, size_t n_parents, p252_fr* parents, int flags);
/* Number of nodes above the leaves of a full arity-4 tree: (n_leaves-1)/3; n_leaves must be 4^k. */
int p252_merkle4_tree_nodes(size_t n_leaves, size_t* n_internal, int* n_levels);
/* Whole tree on one GPU.  nodes_out: all internal levels, bottom-up, concatenated
 * (n_leaves/4 + n_leaves/16 + ... + 1 scalars); the root is the last element. */
int p252_merkle4_build(p252_ctx* ctx, const p252_fr* leaves, size_t n_leaves, p252_fr* nodes_out, int flags);

/* The same for arity 2 or 4 (node = Hash::digest(Domain::Merkle2 | Merkle4, children), src/hash.rs:22-31):
 * internal nodes = (n_leaves - 1) / (arity - 1); n_leaves must be a power of the arity. */
int p252_merkle_tree_nodes(int arity, size_t n_leaves, size_t* n_internal, int* n_levels);
int p252_merkle_build(p252_ctx* ctx, int arity, const p252_fr* leaves, size_t n_leaves, p252_fr* nodes_out, int flags);

/* ---- multi-GPU tree build: one process per GPU, one NCCL all-gather per level ---------------- */
#define P252_NCCL_UNIQUE_ID_BYTES 128
/* rank 0 creates the id and ships it to the other ranks by any means (torch.distributed / MPI) */
int p252_dist_unique_id(uint8_t id[P252_NCCL_UNIQUE_ID_BYTES]);
int p252_dist_init(p252_ctx* ctx, const uint8_t id[P252_NCCL_UNIQUE_ID_BYTES], int rank, int nranks);
int p252_dist_finalize(p252_ctx* ctx);
/* The partition p252_merkle4_build_dist follows (pure host arithmetic, no GPU needed): for every internal
 * level, bottom-up, where it lives in nodes_out, which slice this rank computes, and whether the level is
 * all-gathered (sharded = 1) or computed redundantly by every rank (levels with fewer nodes than ranks). */
typedef struct p252_level_plan {
    uint64_t level_offset; /* first node of the level inside nodes_out            */
    uint64_t level_size;   /* nodes in the level                                   */
    uint64_t my_offset;    /* first node (within the level) this rank computes     */
    uint64_t my_count;     /* how many it computes                                 */
    int32_t sharded;       /* 1: slices + all-gather; 0: every rank computes all   */
    int32_t reserved;
} p252_level_plan;
int p252_merkle4_shard_plan(size_t n_leaves_total, int nranks, int rank, p252_level_plan* levels, int capacity,
                            int* n_levels);
/* leaves_shard: this rank's contiguous n_leaves_total/nranks leaves (DEVICE or HOST per flags).
 * Every level's output is sharded contiguously across ranks, computed, then all-gathered so that
 * each rank ends with the complete level (levels smaller than nranks are computed redundantly).
 * nodes_out (same space as leaves_shard): all internal levels as in p252_merkle4_build. */
int p252_merkle4_build_dist(p252_ctx* ctx, const p252_fr* leaves_shard, size_t n_leaves_total, p252_fr* nodes_out,
                            int flags);

#ifdef __cplusplus
}
#endif
#endif /* POSEIDON252_B200_H */
