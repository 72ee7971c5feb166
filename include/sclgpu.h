/* include/sclgpu.h -- C ABI of libsclgpu.so, the B200 (sm_100a) implementation of
 * SCL's data-parallel hot path: batched scl::math::Fp<61>/Fp<127> arithmetic,
 * util::PRG (AES-128-CTR) expansion, Shamir share / Lagrange reconstruct.
 *
 * SCL (anderspkd/secure-computation-library 0.1.0) has no FFI of its own; its
 * batch seam is the value-semantic template API.  Every
 * entry point below names the reference interface it replaces (paths relative
 * to the reference root).  Conventions:
 *
 *  - Elements are SCL's FF::write bytes (ff.h:295-297): little-endian canonical
 *    residues in [0,p); 8 bytes per Fp<61> element (mersenne61.cc:92-95),
 *    16 bytes per Fp<127> element, low word first (mersenne127.cc:120-123).
 *    Fp<127> buffers are passed as `const void*` / `void*` and must be 16-byte
 *    aligned.  Inputs are expected canonical, as SCL's own containers hold them.
 *  - `seed` is the 16-byte AES key exactly as PRG::create leaves it: the user
 *    seed zero-padded / truncated to 16 bytes (prg.cc:88-101).
 *  - `first_block` makes the PRG seekable: keystream block i is
 *    AES_seed(LE64(i) || LE64(PRG_NONCE)) (prg.cc:82-84, prg.h:34-43) and a call
 *    that SCL would make on a PRG whose counter is c passes first_block = c.
 *    Each function documents how many blocks the equivalent SCL calls consume
 *    so the caller can advance its own counter.
 *  - Host entry points take HOST pointers; the library stages through pinned
 *    memory and does all device work itself.  `_dev` entry points take DEVICE
 *    pointers valid on the context's device, enqueue on the context's stream
 *    (sclgpu_set_stream) and do not synchronise unless they return a status
 *    that depends on device results (recover_d, lagrange).
 *  - Return value: SCLGPU_OK or a negative code.  The mapping to the
 *    reference's exceptions is given per code; the reference's exact what()
 *    string for the last failure is available from sclgpu_last_error().
 *  - One context per device per process; calls on one context are serialised
 *    by the caller (SCL itself is single-threaded).
 *  - There is no CPU fallback: every compute entry point fails with
 *    SCLGPU_ECUDA if no sm_100 device is usable.
 */
#ifndef SCLGPU_H
#define SCLGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SCLGPU_OK 0
#define SCLGPU_EINVAL (-1)  /* std::invalid_argument (vector.h:483, matrix.h:165,425,500) */
#define SCLGPU_ELOGIC (-2)  /* std::logic_error: "not enough shares provided to detect errors"
                               (shamir.h:122-124) or "0 not invertible modulo prime" (small_ff.h:70-72) */
#define SCLGPU_EDETECT (-3) /* std::logic_error("error detected during recovery") (shamir.h:133-135)
                               for at least one secret of the batch; see err[] */
#define SCLGPU_ECUDA (-4)   /* CUDA runtime failure / no usable device */
#define SCLGPU_ENOMEM (-5)  /* device or pinned-host allocation failed */
#define SCLGPU_ECORRECT (-6) /* std::logic_error("could not correct shares") (shamir.h:243-245) for at
                                least one sharing of the batch; see status[] */

/* Layout of a batch of N sharings of n shares each. */
#define SCLGPU_SECRET_MAJOR 0 /* [N][n]: row j = the Vector SCL returns for secret j (shamir.h:60-67) */
#define SCLGPU_PARTY_MAJOR 1  /* [n][N]: row i = party i's share of every secret (device-native SoA) */

typedef struct sclgpu_ctx sclgpu_ctx;

/* ---- context, memory, plumbing (no reference counterpart: SCL is CPU-only) */
int sclgpu_init(int device, sclgpu_ctx** ctx);
void sclgpu_destroy(sclgpu_ctx* ctx);
/* cudaStream_t to enqueue on (NULL = the legacy default stream). */
int sclgpu_set_stream(sclgpu_ctx* ctx, void* cuda_stream);
int sclgpu_sync(sclgpu_ctx* ctx);
const char* sclgpu_last_error(const sclgpu_ctx* ctx);
const char* sclgpu_strerror(int code);
/* number of kernels this context has launched so far (bench.py gpu_launches) */
uint64_t sclgpu_launch_count(const sclgpu_ctx* ctx);
int sclgpu_device_info(const sclgpu_ctx* ctx, int* sm_count, int* cc_major, int* cc_minor,
                       size_t* free_bytes, size_t* total_bytes);
int sclgpu_malloc(sclgpu_ctx* ctx, size_t bytes, void** dptr);
int sclgpu_free(sclgpu_ctx* ctx, void* dptr);
int sclgpu_host_alloc(sclgpu_ctx* ctx, size_t bytes, void** hptr); /* pinned */
int sclgpu_host_free(sclgpu_ctx* ctx, void* hptr);
int sclgpu_memcpy_h2d(sclgpu_ctx* ctx, void* dptr, const void* hptr, size_t bytes);
int sclgpu_memcpy_d2h(sclgpu_ctx* ctx, void* hptr, const void* dptr, size_t bytes);

/* ---- util::PRG ------------------------------------------------------------
 * PRG::next(buf, n) (prg.cc:124-146): ceil(n/16) blocks starting at
 * first_block; the first n bytes are written (the tail of the last block is
 * discarded, as the reference does).  Consumes ceil(n/16) blocks. */
int sclgpu_prg_expand(sclgpu_ctx* ctx, const uint8_t seed[16], uint64_t first_block,
                      uint64_t n_bytes, uint8_t* out);
int sclgpu_prg_expand_dev(sclgpu_ctx* ctx, const uint8_t seed[16], uint64_t first_block,
                          uint64_t n_bytes, uint8_t* d_out);
/* The same keystream (prg.cc:82-84) from the BITSLICED AES-128 kernel (no table lookups: Boyar-Peralta S-box
 * circuit, 32 blocks per thread) -- the measured comparison arm of the T-table kernels; bit-identical output.
 * Whole blocks only: n_bytes a multiple of 16 and d_out 16-byte aligned, else SCLGPU_EINVAL.
 * SCLGPU_PRG_BITSLICED=1 in the environment routes sclgpu_prg_expand[_dev] through it. */
int sclgpu_prg_expand_bitsliced_dev(sclgpu_ctx* ctx, const uint8_t seed[16], uint64_t first_block,
                                    uint64_t n_bytes, uint8_t* d_out);

/* ---- FF::read / Vector::random / FF::random ---------------------------------
 * from_bytes: FF::read = load LE word then "% p" (ff.h:63-67,
 *   mersenne61.cc:87-90, mersenne127.cc:115-118) on n packed elements.
 * random:     Vector<Fp>::random(n, prg) (vector.h:508-519): ONE next() of
 *   n*byteSize bytes; consumes ceil(n*byteSize/16) blocks.
 * ff_random:  FF::random(prg) called n times (ff.h:72-76): one whole AES block
 *   per element, bytes beyond byteSize dropped; consumes n blocks. */
int sclgpu_fp61_from_bytes(sclgpu_ctx* ctx, const uint8_t* bytes, uint64_t n, uint64_t* out);
int sclgpu_fp127_from_bytes(sclgpu_ctx* ctx, const uint8_t* bytes, uint64_t n, void* out);
int sclgpu_fp61_random(sclgpu_ctx* ctx, const uint8_t seed[16], uint64_t first_block, uint64_t n,
                       uint64_t* out);
int sclgpu_fp127_random(sclgpu_ctx* ctx, const uint8_t seed[16], uint64_t first_block, uint64_t n,
                        void* out);
int sclgpu_fp61_ff_random(sclgpu_ctx* ctx, const uint8_t seed[16], uint64_t first_block,
                          uint64_t n, uint64_t* out);
int sclgpu_fp127_ff_random(sclgpu_ctx* ctx, const uint8_t seed[16], uint64_t first_block,
                           uint64_t n, void* out);
int sclgpu_fp61_from_bytes_dev(sclgpu_ctx* ctx, const uint8_t* d_bytes, uint64_t n, uint64_t* d_out);
int sclgpu_fp127_from_bytes_dev(sclgpu_ctx* ctx, const uint8_t* d_bytes, uint64_t n, void* d_out);
int sclgpu_fp61_random_dev(sclgpu_ctx* ctx, const uint8_t seed[16], uint64_t first_block,
                           uint64_t n, uint64_t* d_out);
int sclgpu_fp127_random_dev(sclgpu_ctx* ctx, const uint8_t seed[16], uint64_t first_block,
                            uint64_t n, void* d_out);
int sclgpu_fp61_ff_random_dev(sclgpu_ctx* ctx, const uint8_t seed[16], uint64_t first_block,
                              uint64_t n, uint64_t* d_out);
int sclgpu_fp127_ff_random_dev(sclgpu_ctx* ctx, const uint8_t seed[16], uint64_t first_block,
                               uint64_t n, void* d_out);

/* ---- ss::shamirSecretShare (shamir.h:52-68), called N times on one PRG ------
 * Secret j gets the coefficients SCL would draw when its PRG counter is
 * first_block + j*B, B = ceil((t+1)*byteSize/16) blocks per call: coefficient
 * k (1 <= k <= t) = read(keystream bytes [k*bs,(k+1)*bs)), slot 0 is consumed
 * and replaced by the secret (shamir.h:56-57).  shares[j][i] = f_j(i+1)
 * (poly.h:56-64).  Consumes N*B blocks.  t >= 0, n >= 0; n < 2^31.
 * Host version: `shares` is SCLGPU_SECRET_MAJOR (what SCL returns). */
int sclgpu_fp61_shamir_share(sclgpu_ctx* ctx, const uint64_t* secrets, uint64_t N, uint32_t t,
                             uint32_t n, const uint8_t seed[16], uint64_t first_block,
                             uint64_t* shares);
int sclgpu_fp127_shamir_share(sclgpu_ctx* ctx, const void* secrets, uint64_t N, uint32_t t,
                              uint32_t n, const uint8_t seed[16], uint64_t first_block,
                              void* shares);
int sclgpu_fp61_shamir_share_dev(sclgpu_ctx* ctx, const uint64_t* d_secrets, uint64_t N,
                                 uint32_t t, uint32_t n, const uint8_t seed[16],
                                 uint64_t first_block, uint64_t* d_shares, int layout);
int sclgpu_fp127_shamir_share_dev(sclgpu_ctx* ctx, const void* d_secrets, uint64_t N, uint32_t t,
                                  uint32_t n, const uint8_t seed[16], uint64_t first_block,
                                  void* d_shares, int layout);
/* Sharing from caller-supplied coefficients (Polynomial::create + evaluate,
 * poly.h:56-64,179-198): d_coeffs is [t+1][N] (coefficient k of secret j at
 * k*N + j), coefficient 0 being the secret.  Used to separate the PRG from the
 * evaluation in tests and in the C4 staged pipeline. */
int sclgpu_fp61_shamir_share_coeffs_dev(sclgpu_ctx* ctx, const uint64_t* d_coeffs, uint64_t N,
                                        uint32_t t, uint32_t n, uint64_t* d_shares, int layout);
int sclgpu_fp127_shamir_share_coeffs_dev(sclgpu_ctx* ctx, const void* d_coeffs, uint64_t N,
                                         uint32_t t, uint32_t n, void* d_shares, int layout);

/* ---- shamirRecoverP fused with the all-gather of its result (batch sharded over the
 * GPUs of a box, "gather reconstructed values").  This rank reconstructs its N sharings
 * (party-major planes [n][N]) and the kernel stores secret j into EVERY destination:
 * d_dsts[r][offset + j], r < n_dsts <= 8, where d_dsts[r] is rank r's copy of the gathered
 * vector as addressable from this device (own memory or peer memory over NVLink).  No
 * collective call, no staging buffer: the stores are posted while the planes are read.
 * Completion: when every rank's stream has finished its call (barrier + sclgpu_sync). */
int sclgpu_fp61_recover_p_gather_dev(sclgpu_ctx* ctx, const uint64_t* d_shares, uint64_t N, uint32_t n,
                                     const uint64_t* alphas, const uint64_t* x, uint64_t* const* d_dsts,
                                     uint32_t n_dsts, uint64_t offset);
/* Peer-memory plumbing.  One process per GPU: sclgpu_ipc_export the 64-byte handle of a
 * sclgpu_malloc'ed buffer, hand it to the other processes, sclgpu_ipc_open it there
 * (cudaIpcOpenMemHandle with lazy peer access), sclgpu_ipc_close when done.  One process
 * driving several GPUs: sclgpu_enable_peer(ctx, other_device). */
int sclgpu_memcpy_d2d(sclgpu_ctx* ctx, void* d_dst, const void* d_src, size_t bytes); /* stream-ordered, peer memory included */
int sclgpu_ipc_export(sclgpu_ctx* ctx, void* d_ptr, uint8_t handle[64]);
int sclgpu_ipc_open(sclgpu_ctx* ctx, const uint8_t handle[64], void** d_ptr);
int sclgpu_ipc_close(sclgpu_ctx* ctx, void* d_ptr);
int sclgpu_enable_peer(sclgpu_ctx* ctx, int peer_device);

/* ---- the C2 step in one launch: N x { ss::shamirSecretShare (shamir.h:52-68),
 * ss::shamirRecoverP (shamir.h:82-104) } on party-major planes ([n][N]; plane i = what
 * party i holds).  The share groups of the tcgen05 share kernel (limited by the
 * shared-memory pipe of the fused AES-CTR) and reconstruction warps (FMA pipe + memory)
 * run in the same persistent CTAs.
 *   d_rec_shares == d_shares : the sharings produced by THIS call are reconstructed, each
 *       128-secret tile as soon as it is stored (d_out[j] == d_secrets[j] afterwards);
 *   otherwise d_rec_shares is ANOTHER batch of N sharings ([n][N]) -- e.g. the batch a
 *       previous call produced and the parties returned -- reconstructed concurrently.
 * alphas == NULL: nodes 1..n, x = 0 (shamir.h:100-104); else n nodes and x as in
 * sclgpu_fp61_recover_p.  Any (t, n): shapes outside the fused kernel (t > 15 or
 * n > 32) run as sclgpu_fp61_shamir_share_dev followed by sclgpu_fp61_recover_p_dev. */
int sclgpu_fp61_shamir_share_recover_dev(sclgpu_ctx* ctx, const uint64_t* d_secrets, uint64_t N,
                                         uint32_t t, uint32_t n, const uint8_t seed[16],
                                         uint64_t first_block, uint64_t* d_shares,
                                         const uint64_t* d_rec_shares, const uint64_t* alphas,
                                         const uint64_t* x, uint64_t* d_out);

/* The same launch with the all-gather of the reconstructed secrets fused in (see
 * sclgpu_fp61_recover_p_gather_dev): secret j of this rank's slice is stored to
 * d_dsts[r][offset + j] for every r < n_dsts <= 8 by the reconstruction warps, i.e. the NVLink
 * traffic runs under the share groups' work. */
int sclgpu_fp61_shamir_share_recover_gather_dev(sclgpu_ctx* ctx, const uint64_t* d_secrets, uint64_t N,
                                                uint32_t t, uint32_t n, const uint8_t seed[16],
                                                uint64_t first_block, uint64_t* d_shares,
                                                const uint64_t* d_rec_shares, const uint64_t* alphas,
                                                const uint64_t* x, uint64_t* const* d_dsts,
                                                uint32_t n_dsts, uint64_t offset);

/* ---- ss::shamirRecoverC (shamir.h:203-258, Berlekamp-Welch) on N sharings ---------
 * t = (n-1)/3 and the first np = 3t+1 shares of each sharing are used, as in the
 * reference; alphas == NULL means 1..n (shamir.h:256-258).  Per sharing j:
 *   f[j][0..np)    coefficients of ErrorCorrectedSecret::f (the recovered polynomial;
 *                  the secret is f[j][0]), zero padded to np entries;
 *   err[j][0..t]   coefficients of ErrorCorrectedSecret::err (monic, its roots are the
 *                  nodes of the corrupted shares), zero padded;
 *   status[j]      1 where the reference throws "could not correct shares" (f, err = 0).
 * Returns SCLGPU_ECORRECT if any status[j] is set (everything is still written),
 * *n_failed (nullable) = how many.  Any n >= 1 whose (3t+1) x (3t+2) system fits the shared memory
 * of one SM: np <= 32 runs one warp per sharing, larger np one CTA per sharing (Fp61: n <= 166,
 * Fp127: n <= 118; SCLGPU_EINVAL beyond -- the reference itself has no limit, shamir.h:203-246).
 * Up to t corrupted shares per sharing are corrected. */
int sclgpu_fp61_recover_c(sclgpu_ctx* ctx, const uint64_t* shares, uint64_t N, uint32_t n,
                          const uint64_t* alphas, uint64_t* f, uint64_t* err, uint8_t* status,
                          uint64_t* n_failed);
int sclgpu_fp127_recover_c(sclgpu_ctx* ctx, const void* shares, uint64_t N, uint32_t n, const void* alphas,
                           void* f, void* err, uint8_t* status, uint64_t* n_failed);
int sclgpu_fp61_recover_c_dev(sclgpu_ctx* ctx, const uint64_t* d_shares, uint64_t N, uint32_t n, int layout,
                              const uint64_t* alphas, uint64_t* d_f, uint64_t* d_err, uint8_t* d_status,
                              uint64_t* n_failed);
int sclgpu_fp127_recover_c_dev(sclgpu_ctx* ctx, const void* d_shares, uint64_t N, uint32_t n, int layout,
                               const void* alphas, void* d_f, void* d_err, uint8_t* d_status,
                               uint64_t* n_failed);

/* ---- array-valued secrets: ss::shamirSecretShare / shamirRecoverP on math::Array<FF, W> ----
 * The templates of shamir.h:52-68 / :100-104 instantiated with T = math::Array<FF, W>
 * (include/scl/math/array.h:69-), as ss::pedersenSecretShare does with W = 2 for
 * {secret, randomness} (include/scl/ss/pedersen.h:137-138).  Sharing j draws ONE
 * Vector<Array>::random(t+1) = (t+1)*W*byteSize keystream bytes (vector.h:508-519; Array::read
 * takes W consecutive elements, array.h:82-88): component w of coefficient k is stream element
 * k*W + w, the first W elements are consumed and replaced by the secret, and the call uses
 * sclgpu_share_array_blocks(byteSize, W, t) = ceil((t+1)*W*byteSize/16) blocks; Array arithmetic
 * is component-wise, x runs over Array(1), Array(2), ...  secrets: [N][W].
 * SCLGPU_SECRET_MAJOR shares: [N][n][W] (the N Vectors SCL returns, concatenated);
 * SCLGPU_PARTY_MAJOR: [n][N][W] (party i's N Arrays contiguous).  Host versions are secret-major.
 * recover_p_array: alphas 1..n, x = 0 (the one-argument overload), out [N][W].
 * 1 <= W <= 4096; W = 1 equals shamir_share / recover_p. */
uint64_t sclgpu_share_array_blocks(uint32_t element_bytes, uint32_t W, uint32_t t);
int sclgpu_fp61_shamir_share_array(sclgpu_ctx* ctx, const uint64_t* secrets, uint64_t N, uint32_t W,
                                   uint32_t t, uint32_t n, const uint8_t seed[16],
                                   uint64_t first_block, uint64_t* shares);
int sclgpu_fp127_shamir_share_array(sclgpu_ctx* ctx, const void* secrets, uint64_t N, uint32_t W,
                                    uint32_t t, uint32_t n, const uint8_t seed[16],
                                    uint64_t first_block, void* shares);
int sclgpu_fp61_shamir_share_array_dev(sclgpu_ctx* ctx, const uint64_t* d_secrets, uint64_t N,
                                       uint32_t W, uint32_t t, uint32_t n, const uint8_t seed[16],
                                       uint64_t first_block, uint64_t* d_shares, int layout);
int sclgpu_fp127_shamir_share_array_dev(sclgpu_ctx* ctx, const void* d_secrets, uint64_t N,
                                        uint32_t W, uint32_t t, uint32_t n, const uint8_t seed[16],
                                        uint64_t first_block, void* d_shares, int layout);
int sclgpu_fp61_recover_p_array(sclgpu_ctx* ctx, const uint64_t* shares, uint64_t N, uint32_t W,
                                uint32_t n, uint64_t* out);
int sclgpu_fp127_recover_p_array(sclgpu_ctx* ctx, const void* shares, uint64_t N, uint32_t W,
                                 uint32_t n, void* out);
int sclgpu_fp61_recover_p_array_dev(sclgpu_ctx* ctx, const uint64_t* d_shares, uint64_t N, uint32_t W,
                                    uint32_t n, int layout, uint64_t* d_out);
int sclgpu_fp127_recover_p_array_dev(sclgpu_ctx* ctx, const void* d_shares, uint64_t N, uint32_t W,
                                     uint32_t n, int layout, void* d_out);

/* ---- math::Matrix<FF>::hyperInvertible(n, m) (matrix.h:462-475) ----------------
 * Row i = computeLagrangeBasis(range(1, m+1), -i), -i being the field element p - i
 * (lagrange.h:80-82 -> FF(int)).  out: host, row-major n x m.  SCLGPU_EINVAL
 * ("n or m cannot be 0", matrix.h:165) for an empty shape. */
int sclgpu_fp61_hyper_invertible(sclgpu_ctx* ctx, uint32_t n, uint32_t m, uint64_t* out);
int sclgpu_fp127_hyper_invertible(sclgpu_ctx* ctx, uint32_t n, uint32_t m, void* out);

/* ---- per-party packets: Serializer<math::Vector<FF>> wire layout ------------------
 * What a dealer sends to party i after sharing N secrets is a net::Packet holding the
 * math::Vector of party i's N shares, i.e. Serializer<Vector<FF>>::write
 * (vector.h:596-629 -> serializer.h:160-176): a little-endian u32 element count
 * (StlVecSizeType, serializer.h:111) followed by N x FF::write bytes (ff.h:355-391).
 * shamir_share_packets: as shamir_share, but the output is n host buffers,
 *   packets[i] of sclgpu_packet_bytes(byteSize, N) bytes = party i's serialized
 *   Vector (no [N][n] matrix, no transposition).  N < 2^32 (Packet::SizeType).
 * recover_p_packets: shamirRecoverP (alphas/x as in recover_p) from the n packets a
 *   reconstructing party received; SCLGPU_EINVAL ("Vec sizes mismatch") if a
 *   packet's element count differs from N. */
uint64_t sclgpu_packet_bytes(uint32_t element_bytes, uint64_t n_elements);
int sclgpu_fp61_shamir_share_packets(sclgpu_ctx* ctx, const uint64_t* secrets, uint64_t N, uint32_t t,
                                     uint32_t n, const uint8_t seed[16], uint64_t first_block,
                                     uint8_t* const* packets);
int sclgpu_fp127_shamir_share_packets(sclgpu_ctx* ctx, const void* secrets, uint64_t N, uint32_t t,
                                      uint32_t n, const uint8_t seed[16], uint64_t first_block,
                                      uint8_t* const* packets);
int sclgpu_fp61_recover_p_packets(sclgpu_ctx* ctx, const uint8_t* const* packets, uint64_t N, uint32_t n,
                                  const uint64_t* alphas, const uint64_t* x, uint64_t* out);
int sclgpu_fp127_recover_p_packets(sclgpu_ctx* ctx, const uint8_t* const* packets, uint64_t N, uint32_t n,
                                   const void* alphas, const void* x, void* out);

/* ---- ss::additiveShare (include/scl/ss/additive.h:42-53), called N times on one PRG
 * Secret j: n-1 shares drawn with FF::random (ff.h:72-76: ONE whole keystream block
 * per share, bytes beyond byteSize dropped) from blocks
 * [first_block + j*(n-1), first_block + (j+1)*(n-1)), last share = secret - sum.
 * Consumes N*(n-1) blocks.  n >= 1 (the reference's loop bound n-1 is unsigned).
 * additive_recover: the reconstruction the reference documents, shares.sum()
 * (additive.h:38-39, vector.h:262-267), per sharing.
 * Host versions use SCLGPU_SECRET_MAJOR ([N][n], row j = SCL's Vector for secret j). */
int sclgpu_fp61_additive_share(sclgpu_ctx* ctx, const uint64_t* secrets, uint64_t N, uint32_t n,
                               const uint8_t seed[16], uint64_t first_block, uint64_t* shares);
int sclgpu_fp127_additive_share(sclgpu_ctx* ctx, const void* secrets, uint64_t N, uint32_t n,
                                const uint8_t seed[16], uint64_t first_block, void* shares);
int sclgpu_fp61_additive_share_dev(sclgpu_ctx* ctx, const uint64_t* d_secrets, uint64_t N, uint32_t n,
                                   const uint8_t seed[16], uint64_t first_block, uint64_t* d_shares,
                                   int layout);
int sclgpu_fp127_additive_share_dev(sclgpu_ctx* ctx, const void* d_secrets, uint64_t N, uint32_t n,
                                    const uint8_t seed[16], uint64_t first_block, void* d_shares,
                                    int layout);
int sclgpu_fp61_additive_recover(sclgpu_ctx* ctx, const uint64_t* shares, uint64_t N, uint32_t n,
                                 uint64_t* out);
int sclgpu_fp127_additive_recover(sclgpu_ctx* ctx, const void* shares, uint64_t N, uint32_t n, void* out);
int sclgpu_fp61_additive_recover_dev(sclgpu_ctx* ctx, const uint64_t* d_shares, uint64_t N, uint32_t n,
                                     int layout, uint64_t* d_out);
int sclgpu_fp127_additive_recover_dev(sclgpu_ctx* ctx, const void* d_shares, uint64_t N, uint32_t n,
                                      int layout, void* d_out);

/* ---- math::computeLagrangeBasis (lagrange.h:55-71) --------------------------
 * out[i] = prod_{j != i} (x - nodes[j]) / (nodes[i] - nodes[j]).  nodes == NULL
 * means Vector::range(1, n+1) (vector.h:491-505).  Computed on the device.
 * SCLGPU_ELOGIC ("0 not invertible modulo prime") if two nodes coincide. */
int sclgpu_fp61_lagrange_basis(sclgpu_ctx* ctx, const uint64_t* nodes, uint32_t n,
                               const uint64_t* x, uint64_t* out);
int sclgpu_fp127_lagrange_basis(sclgpu_ctx* ctx, const void* nodes, uint32_t n, const void* x,
                                void* out);

/* ---- ss::shamirRecoverP (shamir.h:82-87, 100-104) on N sharings -------------
 * out[j] = <shares_j, basis(alphas, x)> over ALL n shares.  alphas == NULL is
 * the one-argument overload: alphas = 1..n, x = 0 (x is then ignored). */
int sclgpu_fp61_recover_p(sclgpu_ctx* ctx, const uint64_t* shares, uint64_t N, uint32_t n,
                          const uint64_t* alphas, const uint64_t* x, uint64_t* out);
int sclgpu_fp127_recover_p(sclgpu_ctx* ctx, const void* shares, uint64_t N, uint32_t n,
                           const void* alphas, const void* x, void* out);
/* d_shares / d_out are device pointers; alphas / x stay HOST pointers (n values). */
int sclgpu_fp61_recover_p_dev(sclgpu_ctx* ctx, const uint64_t* d_shares, uint64_t N, uint32_t n,
                              int layout, const uint64_t* alphas, const uint64_t* x,
                              uint64_t* d_out);
int sclgpu_fp127_recover_p_dev(sclgpu_ctx* ctx, const void* d_shares, uint64_t N, uint32_t n,
                               int layout, const void* alphas, const void* x, void* d_out);

/* ---- ss::shamirRecoverD (shamir.h:117-140, 152-155) on N sharings -----------
 * alphas == NULL is the (shares, t) overload: alphas = 1..2t+1, d = t, x = 0
 * (n_alphas, d, x ignored).  Otherwise the five-argument form.
 * SCLGPU_ELOGIC if n_given < d+t or n_alphas < d+t ("not enough shares provided
 * to detect errors").  Exactly as the reference, only share indices
 * d+1 .. d+t-1 are checked against the interpolation through shares 0..d.
 * err[j] = 1 where the reference would throw "error detected during recovery"
 * (out[j] is then 0); returns SCLGPU_EDETECT if any err[j] is set (out/err are
 * still fully written), SCLGPU_OK otherwise.  *n_detected (nullable) receives
 * the number of flagged secrets.  n_given = shares per secret in the buffer. */
int sclgpu_fp61_recover_d(sclgpu_ctx* ctx, const uint64_t* shares, uint64_t N, uint32_t n_given,
                          uint32_t t, const uint64_t* alphas, uint32_t n_alphas, uint32_t d,
                          const uint64_t* x, uint64_t* out, uint8_t* err, uint64_t* n_detected);
int sclgpu_fp127_recover_d(sclgpu_ctx* ctx, const void* shares, uint64_t N, uint32_t n_given,
                           uint32_t t, const void* alphas, uint32_t n_alphas, uint32_t d,
                           const void* x, void* out, uint8_t* err, uint64_t* n_detected);
int sclgpu_fp61_recover_d_dev(sclgpu_ctx* ctx, const uint64_t* d_shares, uint64_t N,
                              uint32_t n_given, int layout, uint32_t t, const uint64_t* alphas,
                              uint32_t n_alphas, uint32_t d, const uint64_t* x, uint64_t* d_out,
                              uint8_t* d_err, uint64_t* n_detected);
int sclgpu_fp127_recover_d_dev(sclgpu_ctx* ctx, const void* d_shares, uint64_t N,
                               uint32_t n_given, int layout, uint32_t t, const void* alphas,
                               uint32_t n_alphas, uint32_t d, const void* x, void* d_out,
                               uint8_t* d_err, uint64_t* n_detected);

/* ---- math::Vector entrywise ops (vector.h:192-301, 522-556) -----------------
 * add / sub / mul = add / subtract / multiplyEntryWise; scale = scalarMultiply
 * (scalar points to ONE element); dot (vector.h:252-259, innerProd :45-52) and
 * sum (:262-267) write one element.  Size mismatch is the caller's check (the
 * host mirror raises "Vec sizes mismatch").  muladd is the Beaver-style
 * z = e*b + d*a + c + e*d (test/scl/protocol/beaver.h:57-61, vectorised). */
int sclgpu_fp61_vec_add(sclgpu_ctx* ctx, const uint64_t* a, const uint64_t* b, uint64_t n, uint64_t* out);
int sclgpu_fp61_vec_sub(sclgpu_ctx* ctx, const uint64_t* a, const uint64_t* b, uint64_t n, uint64_t* out);
int sclgpu_fp61_vec_mul(sclgpu_ctx* ctx, const uint64_t* a, const uint64_t* b, uint64_t n, uint64_t* out);
int sclgpu_fp61_vec_scale(sclgpu_ctx* ctx, const uint64_t* a, const uint64_t* scalar, uint64_t n, uint64_t* out);
int sclgpu_fp61_vec_muladd(sclgpu_ctx* ctx, const uint64_t* e, const uint64_t* b, const uint64_t* d,
                           const uint64_t* a, const uint64_t* c, uint64_t n, uint64_t* z);
int sclgpu_fp61_dot(sclgpu_ctx* ctx, const uint64_t* a, const uint64_t* b, uint64_t n, uint64_t* out);
int sclgpu_fp61_sum(sclgpu_ctx* ctx, const uint64_t* a, uint64_t n, uint64_t* out);
int sclgpu_fp127_vec_add(sclgpu_ctx* ctx, const void* a, const void* b, uint64_t n, void* out);
int sclgpu_fp127_vec_sub(sclgpu_ctx* ctx, const void* a, const void* b, uint64_t n, void* out);
int sclgpu_fp127_vec_mul(sclgpu_ctx* ctx, const void* a, const void* b, uint64_t n, void* out);
int sclgpu_fp127_vec_scale(sclgpu_ctx* ctx, const void* a, const void* scalar, uint64_t n, void* out);
int sclgpu_fp127_vec_muladd(sclgpu_ctx* ctx, const void* e, const void* b, const void* d,
                            const void* a, const void* c, uint64_t n, void* z);
int sclgpu_fp127_dot(sclgpu_ctx* ctx, const void* a, const void* b, uint64_t n, void* out);
int sclgpu_fp127_sum(sclgpu_ctx* ctx, const void* a, uint64_t n, void* out);
/* Vector::equals (vector.h:358-375; also Matrix::equals on rows*cols elements): *equal = 1 iff all n
 * elements agree.  Every element is compared (no early exit, as in the reference).  The _dev forms take
 * device pointers and synchronise the stream to return the answer. */
int sclgpu_fp61_vec_equal(sclgpu_ctx* ctx, const uint64_t* a, const uint64_t* b, uint64_t n, int* equal);
int sclgpu_fp127_vec_equal(sclgpu_ctx* ctx, const void* a, const void* b, uint64_t n, int* equal);
int sclgpu_fp61_vec_equal_dev(sclgpu_ctx* ctx, const uint64_t* a, const uint64_t* b, uint64_t n, int* equal);
int sclgpu_fp127_vec_equal_dev(sclgpu_ctx* ctx, const void* a, const void* b, uint64_t n, int* equal);
/* device-pointer forms; `scalar` of vec_scale_dev is a HOST pointer, dot/sum
 * write one element to DEVICE memory */
int sclgpu_fp61_vec_add_dev(sclgpu_ctx* ctx, const uint64_t* a, const uint64_t* b, uint64_t n, uint64_t* out);
int sclgpu_fp61_vec_sub_dev(sclgpu_ctx* ctx, const uint64_t* a, const uint64_t* b, uint64_t n, uint64_t* out);
int sclgpu_fp61_vec_mul_dev(sclgpu_ctx* ctx, const uint64_t* a, const uint64_t* b, uint64_t n, uint64_t* out);
int sclgpu_fp61_vec_scale_dev(sclgpu_ctx* ctx, const uint64_t* a, const uint64_t* scalar, uint64_t n, uint64_t* out);
int sclgpu_fp61_vec_muladd_dev(sclgpu_ctx* ctx, const uint64_t* e, const uint64_t* b, const uint64_t* d,
                               const uint64_t* a, const uint64_t* c, uint64_t n, uint64_t* z);
int sclgpu_fp61_dot_dev(sclgpu_ctx* ctx, const uint64_t* a, const uint64_t* b, uint64_t n, uint64_t* out);
int sclgpu_fp61_sum_dev(sclgpu_ctx* ctx, const uint64_t* a, uint64_t n, uint64_t* out);
int sclgpu_fp127_vec_add_dev(sclgpu_ctx* ctx, const void* a, const void* b, uint64_t n, void* out);
int sclgpu_fp127_vec_sub_dev(sclgpu_ctx* ctx, const void* a, const void* b, uint64_t n, void* out);
int sclgpu_fp127_vec_mul_dev(sclgpu_ctx* ctx, const void* a, const void* b, uint64_t n, void* out);
int sclgpu_fp127_vec_scale_dev(sclgpu_ctx* ctx, const void* a, const void* scalar, uint64_t n, void* out);
int sclgpu_fp127_vec_muladd_dev(sclgpu_ctx* ctx, const void* e, const void* b, const void* d,
                                const void* a, const void* c, uint64_t n, void* z);
int sclgpu_fp127_dot_dev(sclgpu_ctx* ctx, const void* a, const void* b, uint64_t n, void* out);
int sclgpu_fp127_sum_dev(sclgpu_ctx* ctx, const void* a, uint64_t n, void* out);

/* ---- math::Matrix (matrix.h) ------------------------------------------------
 * matvec: Matrix::multiply(Vector) (matrix.h:498-513), A row-major rows x cols
 *   (matrix.h:199-201); SCLGPU_EINVAL when rows or cols is 0 (matrix.h:165).
 * vandermonde: Matrix::vandermonde(n, m) with xs = 1..n (matrix.h:102-104,
 *   445-460): out[i][j] = (i+1)^j, row-major n x m. */
int sclgpu_fp61_matvec(sclgpu_ctx* ctx, const uint64_t* A, uint32_t rows, uint32_t cols,
                       const uint64_t* x, uint64_t* y);
int sclgpu_fp127_matvec(sclgpu_ctx* ctx, const void* A, uint32_t rows, uint32_t cols,
                        const void* x, void* y);
int sclgpu_fp61_matvec_dev(sclgpu_ctx* ctx, const uint64_t* d_A, uint32_t rows, uint32_t cols,
                           const uint64_t* d_x, uint64_t* d_y);
int sclgpu_fp127_matvec_dev(sclgpu_ctx* ctx, const void* d_A, uint32_t rows, uint32_t cols,
                            const void* d_x, void* d_y);
/* matmul: Matrix::multiply(Matrix) (matrix.h:476-495): C (rows x cols) = A (rows x inner) * B (inner x cols),
 *   all row-major; the host mirror raises "matmul: this->cols() != that->rows()" (matrix.h:480) itself, a zero
 *   dimension is SCLGPU_EINVAL ("n or m cannot be 0", matrix.h:165). */
int sclgpu_fp61_matmul(sclgpu_ctx* ctx, const uint64_t* A, uint32_t rows, uint32_t inner, const uint64_t* B,
                       uint32_t cols, uint64_t* C);
int sclgpu_fp127_matmul(sclgpu_ctx* ctx, const void* A, uint32_t rows, uint32_t inner, const void* B, uint32_t cols,
                        void* C);
int sclgpu_fp61_matmul_dev(sclgpu_ctx* ctx, const uint64_t* d_A, uint32_t rows, uint32_t inner, const uint64_t* d_B,
                           uint32_t cols, uint64_t* d_C);
int sclgpu_fp127_matmul_dev(sclgpu_ctx* ctx, const void* d_A, uint32_t rows, uint32_t inner, const void* d_B,
                            uint32_t cols, void* d_C);
int sclgpu_fp61_vandermonde(sclgpu_ctx* ctx, uint32_t n, uint32_t m, uint64_t* out);
int sclgpu_fp127_vandermonde(sclgpu_ctx* ctx, uint32_t n, uint32_t m, void* out);
/* Matrix::vandermonde(n, m, xs) with the caller's nodes (matrix.h:445-460): row i = (1, xs[i], xs[i]^2, ..);
 * n_xs != n -> SCLGPU_EINVAL "|xs| != number of rows". */
int sclgpu_fp61_vandermonde_xs(sclgpu_ctx* ctx, uint32_t n, uint32_t m, const uint64_t* xs, uint32_t n_xs, uint64_t* out);
int sclgpu_fp127_vandermonde_xs(sclgpu_ctx* ctx, uint32_t n, uint32_t m, const void* xs, uint32_t n_xs, void* out);
/* Polynomial::evaluate (poly.h:56-64: Horner from the top coefficient) of N polynomials of degree <= t at n
 * caller-chosen points xs (HOST pointer in both forms).  Host form: coeffs [N][t+1], row j = polynomial j's
 * coefficients, constant term first (Polynomial::coefficients()); out [N][n].  Device form: coefficient planes
 * [t+1][N] as in *_shamir_share_coeffs_dev, out in `layout`. */
int sclgpu_fp61_poly_evaluate(sclgpu_ctx* ctx, const uint64_t* coeffs, uint64_t N, uint32_t t, const uint64_t* xs,
                              uint32_t n, uint64_t* out);
int sclgpu_fp127_poly_evaluate(sclgpu_ctx* ctx, const void* coeffs, uint64_t N, uint32_t t, const void* xs, uint32_t n,
                               void* out);
int sclgpu_fp61_poly_evaluate_dev(sclgpu_ctx* ctx, const uint64_t* d_coeffs, uint64_t N, uint32_t t, const uint64_t* xs,
                                  uint32_t n, uint64_t* d_out, int layout);
int sclgpu_fp127_poly_evaluate_dev(sclgpu_ctx* ctx, const void* d_coeffs, uint64_t N, uint32_t t, const void* xs,
                                   uint32_t n, void* d_out, int layout);
/* Matrix::transpose (matrix.h:344-355) of a host matrix [rows][cols] -> [cols][rows];
 * Matrix::scalarMultiply (matrix.h:325-342) is *_vec_scale on rows*cols elements. */
int sclgpu_fp61_transpose(sclgpu_ctx* ctx, const uint64_t* in, uint64_t rows, uint64_t cols, uint64_t* out);
int sclgpu_fp127_transpose(sclgpu_ctx* ctx, const void* in, uint64_t rows, uint64_t cols, void* out);

/* ---- layout helpers (device) -------------------------------------------------
 * [rows][cols] -> [cols][rows] of 8-byte (fp61) / 16-byte (fp127) elements. */
int sclgpu_fp61_transpose_dev(sclgpu_ctx* ctx, const uint64_t* d_in, uint64_t rows, uint64_t cols,
                              uint64_t* d_out);
int sclgpu_fp127_transpose_dev(sclgpu_ctx* ctx, const void* d_in, uint64_t rows, uint64_t cols,
                               void* d_out);

/* ---- measurement helper -------------------------------------------------------
 * Integer-pipe microbenchmark used for the "int-mul roofline" denominator:
 * runs `iters` dependent-chain-free IMAD (kind 0), IMAD.WIDE.U32 (kind 1),
 * LOP3 (kind 2), IADD3 (kind 3) or LDS.32 (kind 4) warp instructions per warp on every SM and
 * returns the achieved thread-level operations per second in *ops_per_s. */
int sclgpu_pipe_microbench(sclgpu_ctx* ctx, int kind, uint32_t iters, double* ops_per_s);

/* ---- several GPUs behind one handle (SCL is one process, one thread:
 * coro/runtime.h:126-163).  A batch call cuts [0, N) into contiguous slices, one per device,
 * slice g starting its PRG at first_block + lo_g * B with the same seed, so the results are
 * what N calls on ONE scl::util::PRG return; one worker thread per device runs the
 * single-device host pipeline on its slice of the caller's buffers.  devices == NULL means
 * 0..n_devices-1.  Fails as a whole (no partial device set, no CPU fallback). */
typedef struct sclgpu_mctx sclgpu_mctx;
int sclgpu_multi_init(const int* devices, int n_devices, sclgpu_mctx** mctx);
void sclgpu_multi_destroy(sclgpu_mctx* mctx);
int sclgpu_multi_device_count(const sclgpu_mctx* mctx);
/* the slicing rule (pure function, needs no device): units [*lo, *hi) of [0, n_units) go to slice `index` of `parts`;
 * boundaries are multiples of `align` (2 for Fp61 Vector::random and the plane kernels' 128-bit accesses) */
int sclgpu_multi_slice(uint64_t n_units, int parts, int index, uint64_t align, uint64_t* lo, uint64_t* hi);
sclgpu_ctx* sclgpu_multi_context(sclgpu_mctx* mctx, int index); /* the per-device context (borrowed) */
const char* sclgpu_multi_last_error(const sclgpu_mctx* mctx);
/* same contracts as the single-device host entry points of the same name */
int sclgpu_multi_fp61_shamir_share(sclgpu_mctx* mctx, const uint64_t* secrets, uint64_t N, uint32_t t,
                                   uint32_t n, const uint8_t seed[16], uint64_t first_block,
                                   uint64_t* shares);
int sclgpu_multi_fp127_shamir_share(sclgpu_mctx* mctx, const void* secrets, uint64_t N, uint32_t t,
                                    uint32_t n, const uint8_t seed[16], uint64_t first_block,
                                    void* shares);
int sclgpu_multi_fp61_recover_p(sclgpu_mctx* mctx, const uint64_t* shares, uint64_t N, uint32_t n,
                                const uint64_t* alphas, const uint64_t* x, uint64_t* out);
int sclgpu_multi_fp127_recover_p(sclgpu_mctx* mctx, const void* shares, uint64_t N, uint32_t n,
                                 const void* alphas, const void* x, void* out);
int sclgpu_multi_fp61_recover_d(sclgpu_mctx* mctx, const uint64_t* shares, uint64_t N, uint32_t n_given,
                                uint32_t t, const uint64_t* alphas, uint32_t n_alphas, uint32_t d,
                                const uint64_t* x, uint64_t* out, uint8_t* err, uint64_t* n_detected);
int sclgpu_multi_fp127_recover_d(sclgpu_mctx* mctx, const void* shares, uint64_t N, uint32_t n_given,
                                 uint32_t t, const void* alphas, uint32_t n_alphas, uint32_t d,
                                 const void* x, void* out, uint8_t* err, uint64_t* n_detected);
int sclgpu_multi_fp61_random(sclgpu_mctx* mctx, const uint8_t seed[16], uint64_t first_block, uint64_t n,
                             uint64_t* out);

/* ---- asynchronous host calls.  The host entry points return when their last device-to-host
 * copy has landed; these run the same pipeline on a companion context (own streams and
 * scratch, same device) from a worker thread and return at once.  One asynchronous call is in
 * flight per context (a second one first completes the first); sclgpu_wait -- or sclgpu_sync
 * -- completes it and returns its status.  The caller's buffers must stay valid until then.
 * Typical use: share_async(batch k) then recover_p(batch k-1) on the same context: shares
 * travel device-to-host while the previous batch travels host-to-device (PCIe is full duplex). */
int sclgpu_fp61_shamir_share_async(sclgpu_ctx* ctx, const uint64_t* secrets, uint64_t N, uint32_t t,
                                   uint32_t n, const uint8_t seed[16], uint64_t first_block,
                                   uint64_t* shares);
int sclgpu_fp127_shamir_share_async(sclgpu_ctx* ctx, const void* secrets, uint64_t N, uint32_t t,
                                    uint32_t n, const uint8_t seed[16], uint64_t first_block,
                                    void* shares);
int sclgpu_fp61_recover_p_async(sclgpu_ctx* ctx, const uint64_t* shares, uint64_t N, uint32_t n,
                                const uint64_t* alphas, const uint64_t* x, uint64_t* out);
int sclgpu_fp127_recover_p_async(sclgpu_ctx* ctx, const void* shares, uint64_t N, uint32_t n,
                                 const void* alphas, const void* x, void* out);
int sclgpu_wait(sclgpu_ctx* ctx);
/* the CUDA device index of a context; set the message sclgpu_last_error returns (used by the
 * layers above the single-device ABI) */
int sclgpu_device_index(const sclgpu_ctx* ctx, int* device);
void sclgpu_set_error(sclgpu_ctx* ctx, const char* message);

#ifdef __cplusplus
}
#endif
#endif /* SCLGPU_H */
