/*
 * zkb200.h -- C ABI of libzkb200.so: the B200-native (sm_100a) prover hot path of zkt-plonk.
 *
 * The reference (pure Rust, /root/reference) has no FFI today; its de-facto plugin seams are the
 * generic parameters of `ZKTPlonk<F, D, PC, ..>` / `proof_system::prove` (plonk-core/src/plonk.rs:39-46,
 * plonk-core/src/proof_system/prove.rs:59-74).  Each entry point below names the reference interface it
 * replaces; INTEGRATION.md shows the Rust `extern "C"` crate that binds them behind `D` and `PC`.
 *
 * Curves.  The reference is generic over the pairing engine (its own full test runs on Bls12_381 and Bls12_377,
 * plonk-core/src/plonk.rs:226-254; the CLI fixes Bn254, bin/src/instance.rs:7-10) and Rust monomorphises per curve.  This library
 * is compiled once per curve with identical entry points: libzkb200.so (BN254: everything below), libzkb200_bls12_381.so and
 * libzkb200_bls12_377.so (everything but EthereumTranscript, which is bound to Bn254 upstream and answers ZKB_ERR_UNSUPPORTED
 * there; the key files are the same derive(CanonicalSerialize) layouts with 48-byte Fq).  zkb_curve_info tells a caller which one it loaded
 * and the element widths: array sizes written below as [8] / [16] / [80] are the BN254 ones (fq_words = 4); on the BLS12 curves
 * an affine point is 12 words and an XYZZ partial sum 24 (fq_words = 6).  Scalars are 4 words on every curve.
 *
 * Conventions (identical to arkworks 0.3 in-memory forms, so no conversion at the boundary):
 *   - Fr element       = 4 x uint64_t little-endian limbs, MONTGOMERY form (R = 2^256).
 *   - Fq element       = fq_words x uint64_t little-endian limbs, MONTGOMERY form (R = 2^256 on BN254, 2^384 on BLS12-381 / 377).
 *   - MSM scalars      = 4 x uint64_t little-endian limbs, CANONICAL form (what `into_repr()` yields).
 *   - G1 affine point  = x || y (2 x fq_words uint64_t, Montgomery Fq); the point at infinity is (0, 0).
 *   - NTT data is in natural order on input and output.
 *   - Every function returns 0 on success and a negative zkb_status otherwise; nothing throws or aborts.
 *     zkb_last_error(ctx) gives a human-readable reason.  Pointers are caller-owned and not retained.
 *   - "_dev" variants take DEVICE pointers (HBM-resident data, e.g. torch tensors' data_ptr) and enqueue on
 *     the context's stream without synchronising; the plain variants take HOST pointers, copy in and out,
 *     and return when the result is in the caller's buffer.
 *   - There is no CPU fallback: without a CUDA device zkb_ctx_create fails.
 */
#ifndef ZKB200_H
#define ZKB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define ZKB_API __attribute__((visibility("default")))
#else
#define ZKB_API
#endif

typedef struct zkb_ctx zkb_ctx;

typedef enum {
    ZKB_OK = 0,
    ZKB_ERR_INVALID = -1,    /* bad argument (null pointer, size not allowed ...) */
    ZKB_ERR_DOMAIN = -2,     /* log_n > TWO_ADICITY (28 on BN254): Error::InvalidEvalDomainSize, prove.rs:77-81 */
    ZKB_ERR_CUDA = -3,       /* a CUDA runtime call failed */
    ZKB_ERR_NO_SRS = -4,     /* MSM asked for more points than the loaded SRS holds (kzg10 TooManyCoefficients) */
    ZKB_ERR_OOM = -5,
    ZKB_ERR_UNSUPPORTED = -6 /* entry point not compiled into this curve's build of the library (zkb_curve_info) */
} zkb_status;

/* ---- context ------------------------------------------------------------------------------------------------ */
/* One context per GPU and proving thread (prove() is !Send: prove.rs:62).  Owns a stream, the twiddle tables,
 * the resident SRS and all scratch memory; everything is released by zkb_ctx_destroy. */
ZKB_API int zkb_ctx_create(int device, zkb_ctx **out);
ZKB_API void zkb_ctx_destroy(zkb_ctx *ctx);
/* Use an existing cudaStream_t (e.g. torch's current stream) for all subsequent work; NULL = default stream. */
ZKB_API int zkb_ctx_set_stream(zkb_ctx *ctx, void *cuda_stream);
ZKB_API int zkb_ctx_sync(zkb_ctx *ctx);
ZKB_API const char *zkb_last_error(zkb_ctx *ctx);
ZKB_API const char *zkb_version(void);
/* Which curve this shared object was compiled for (the `E: PairingEngine` of the reference's generics): curve_id 0 = BN254,
 * 1 = BLS12-381, 2 = BLS12-377; 64-bit words of a scalar and of a base-field element; bit length of r; has_prover = 1 when
 * the protocol driver (zkb_plonk_setup / zkb_plonk_prove ..) is compiled in: every build (as are the verifier with the curve's own pairing and the
 * key-file readers / writers).  Any pointer may be NULL.  Needs no context and no GPU. */
ZKB_API int zkb_curve_info(int *curve_id, int *fr_words, int *fq_words, int *fr_bits, int *has_prover);
/* The curve's G1 generator (ark-* 0.3 G1_GENERATOR_X / _Y), affine, Montgomery form: 2 x fq_words words. */
ZKB_API int zkb_g1_generator(uint64_t *out_xy);

/* ---- device memory (for HBM-resident pipelines) -------------------------------------------------------------- */
ZKB_API int zkb_dev_alloc(zkb_ctx *ctx, size_t bytes, void **dptr);
ZKB_API int zkb_dev_free(zkb_ctx *ctx, void *dptr);
ZKB_API int zkb_h2d(zkb_ctx *ctx, void *dst_dev, const void *src_host, size_t bytes);   /* synchronous */
ZKB_API int zkb_d2h(zkb_ctx *ctx, void *dst_host, const void *src_dev, size_t bytes);   /* synchronous */

/* ---- NTT: replaces D::{fft,ifft,coset_fft,coset_ifft}_in_place ------------------------------------------------- */
/* Reference: ark-poly 0.3 Radix2EvaluationDomain via plonk-core/src/util.rs:63-140 (poly_from_evals :63-86,
 * poly_from_coset_evals :90-100, evals_from_poly_ref :104-113, coset_evals_from_poly(_ref) :117-140).
 * data holds 2^log_n elements of which the first `len` are input (the rest are treated as zero, as
 * fft_in_place's resize does); inverse: multiplies by n^-1; coset: generator g = Fr::multiplicative_generator() (5 on BN254, 7 on BLS12-381, 22 on BLS12-377; forward: scale coefficient i
 * by g^i first; inverse: scale output i by g^-i). */
ZKB_API int zkb_ntt(zkb_ctx *ctx, uint64_t *data_host, size_t len, unsigned log_n, int inverse, int coset);
ZKB_API int zkb_ntt_dev(zkb_ctx *ctx, uint64_t *data_dev, size_t len, unsigned log_n, int inverse, int coset);
/* `count` independent transforms of the same shape (the quotient round runs 9 coset FFTs on the 4n domain,
 * quotient_poly.rs:52-96); ptrs_host[k] is a DEVICE pointer to transform k's 2^log_n elements. */
ZKB_API int zkb_ntt_batch_dev(zkb_ctx *ctx, uint64_t *const *ptrs_host, size_t count, size_t len, unsigned log_n,
                      int inverse, int coset);

/* Fully expanded inter-pass twiddle / coset tables (N x 32 B each per size and direction, built on first use,
 * transforms of 2^12..2^26 elements) save one field product per element per use.  enable = 0 falls back to the
 * two-level tables (a few KB) when HBM is needed for something else.  Default: enabled. */
ZKB_API int zkb_ntt_set_direct_tables(zkb_ctx *ctx, int enable);
/* Pass kernel: 0 (default) = the radix-4 kernel (compile-time tile size, two stages per barrier, 3 CTAs per SM) on
 * 2048-element tiles, where it measured faster, the generic kernel elsewhere; 1 = the generic kernel everywhere; 2 = the
 * radix-4 kernel on every tile of 256..2048 elements.  Same results for every choice.  ZKB_NTT_KERNEL=<kind> in the
 * environment sets it for contexts created afterwards. */
ZKB_API int zkb_ntt_set_kernel(zkb_ctx *ctx, int kind);

/* ---- MSM: replaces VariableBaseMSM::multi_scalar_mul / kzg10::commit's inner product ------------------------------ */
/* Reference: ark-ec 0.3 msm::VariableBaseMSM::multi_scalar_mul, called from plonk-core/src/commitment.rs:42 and via
 * ark-poly-commit kzg10::{commit,open} from prove.rs:134,179,250,307,374,381,427.
 * The committer key's powers_of_g stay resident in HBM (uploaded once, like `ck` is loaded once by the CLI,
 * bin/src/main.rs:274-281); an MSM then only moves n x 32 B of scalars in and 64 B out. */
ZKB_API int zkb_srs_load_g1(zkb_ctx *ctx, const uint64_t *xy_mont_host, size_t n);
ZKB_API int zkb_srs_load_g1_dev(zkb_ctx *ctx, const uint64_t *xy_mont_dev, size_t n);   /* copies; caller keeps its buffer */
ZKB_API size_t zkb_srs_size(zkb_ctx *ctx);
/* Fixed-base window tables for the resident SRS: rows[w][i] = 2^(c*w) * SRS[i] (W x n x 64 B of HBM, built once per
 * key like PC::trim).  Later MSMs against the SRS feed one shared bucket set and need no window fold.
 * c > 0: that window size; c == 0: cost model (20 at n = 2^20); c < 0: drop the tables.  Loading a new SRS drops them. */
ZKB_API int zkb_srs_precompute(zkb_ctx *ctx, int c);
/* sum_{i<n} scalars[i] * SRS[offset + i]  ->  affine (x, y) Montgomery; *is_inf = 1 and (0,0) for the identity.
 * Precondition (not checked): every scalar is a canonical integer < r, what Fr::into_repr hands VariableBaseMSM (a value
 * >= 2^255 would lose its top digit's carry).  The single-MSM entry points return ZKB_ERR_INVALID while a zkb_commit_push
 * batch is open on the context (they would reuse its staging buffer and result slot). */
ZKB_API int zkb_msm_g1(zkb_ctx *ctx, const uint64_t *scalars_host, size_t offset, size_t n, uint64_t out_xy[8], int *is_inf);
ZKB_API int zkb_msm_g1_dev(zkb_ctx *ctx, const uint64_t *scalars_dev, size_t offset, size_t n, uint64_t out_xy[8], int *is_inf);
/* Same, but returns the un-normalised XYZZ partial sum (X, Y, ZZ, ZZZ: 16 limbs) of one point-range shard. */
ZKB_API int zkb_msm_g1_dev_partial(zkb_ctx *ctx, const uint64_t *scalars_dev, size_t offset, size_t n, uint64_t out_xyzz[16]);
/* Point-range sharded MSM over the ranks of the context's communicator (zkb_comm_init): every rank
 * passes the scalars of ITS resident SRS range, the partial sums are exchanged over NCCL inside the call and every rank
 * returns the same affine point.  Collective: all ranks must call it.  Without a communicator it is zkb_msm_g1_dev. */
ZKB_API int zkb_msm_g1_sharded_dev(zkb_ctx *ctx, const uint64_t *scalars_dev, size_t offset, size_t n, uint64_t out_xy[8], int *is_inf);
/* The same with HOST scalars: the upload of this rank's scalars overlaps their accumulation (the two-range path of zkb_msm_g1). */
ZKB_API int zkb_msm_g1_sharded(zkb_ctx *ctx, const uint64_t *scalars_host, size_t offset, size_t n, uint64_t out_xy[8], int *is_inf);
/* Combine `count` shard results (after the NCCL all-gather of 128 B per rank) into the affine commitment. */
ZKB_API int zkb_g1_sum_partials(const uint64_t *xyzz, size_t count, uint64_t out_xy[8], int *is_inf);
/* Arbitrary bases: drop-in for VariableBaseMSM::multi_scalar_mul(bases, scalars) and
 * HomomorphicCommitment::multi_scalar_mul (commitment.rs:31-46).  Uses min(len) = n pairs. */
ZKB_API int zkb_msm_g1_bases(zkb_ctx *ctx, const uint64_t *points_host, const uint64_t *scalars_host, size_t n,
                     uint64_t out_xy[8], int *is_inf);
/* The same with the bases in HBM too (caller-held points: any committer key, e.g. ipa_pc's ck.comm_key or the folded keys of its
 * opening rounds).  scalars_mont != 0: the scalars are Montgomery-form coefficients (converted like into_repr first). */
ZKB_API int zkb_msm_g1_points_dev(zkb_ctx *ctx, const uint64_t *points_dev, const uint64_t *scalars_dev, size_t n, int scalars_mont,
                          uint64_t out_xy[8], int *is_inf);
/* kzg10::commit for one polynomial resident in HBM (Montgomery coefficients, as produced by the iNTT):
 * into_repr conversion + MSM against SRS[offset .. offset+n).  Reference: ark-poly-commit 0.3 kzg10::commit as
 * reached from prove.rs:133-135,178-180,249-251,306-308,373-375 (the caller skips leading zero coefficients by
 * passing offset/n, as skip_leading_zeros_and_convert_to_bigints does). */
ZKB_API int zkb_commit_dev(zkb_ctx *ctx, const uint64_t *coeffs_mont_dev, size_t offset, size_t n, uint64_t out_xy[8], int *is_inf);
/* kzg10::commit for a batch of polynomials (PolynomialCommitment::commit commits several polynomials per call:
 * prove.rs:133-135 a,b,c; :178-180 t,h1,h2; :249-251 z1,z2; :306-308 q_lo,q_mid,q_hi).  The MSMs are pipelined: the
 * window reduction and download of MSM k overlap the sort and bucket accumulation of MSM k+1.
 * coeffs_mont_dev[k]: device pointer; offsets may be NULL (all 0); out_xy: count x 8 limbs; is_inf: count (or NULL). */
ZKB_API int zkb_commit_batch_dev(zkb_ctx *ctx, const uint64_t *const *coeffs_mont_dev, const size_t *offsets, const size_t *lens,
                         size_t count, uint64_t *out_xy, int *is_inf);
/* The same batch, incrementally: zkb_commit_push enqueues the MSM of one more polynomial and returns at once,
 * zkb_commit_finish waits for the open batch and writes its commitments in push order.  Between pushes the caller
 * may enqueue other work or block in a host copy while the GPU commits (zkb_plonk_prove uploads wire b during the
 * commitment to wire a).  The coefficient buffers must stay untouched until zkb_commit_finish returns. */
ZKB_API int zkb_commit_push(zkb_ctx *ctx, const uint64_t *coeffs_mont_dev, size_t offset, size_t len);
ZKB_API int zkb_commit_finish(zkb_ctx *ctx, uint64_t *out_xy, int *is_inf);
/* Close the open batch WITHOUT the exchange between ranks: this rank's XYZZ partial sum of every pushed commitment (16
 * limbs each, the identity where the rank had no share).  zkb_g1_sum_partials over the ranks gives the commitment. */
ZKB_API int zkb_commit_finish_partials(zkb_ctx *ctx, uint64_t *out_xyzz);
/* out_points_dev[i] = scalars_dev[i] * base: builds [tau^i]G-style SRS / synthetic points directly in HBM
 * (what PC::setup's FixedBaseMSM does once per SRS, plonk.rs:195). */
ZKB_API int zkb_g1_fixed_base_mul_dev(zkb_ctx *ctx, const uint64_t base_xy[8], const uint64_t *scalars_dev, size_t n,
                              uint64_t *out_points_dev);
/* A large single MSM runs as point ranges through ONE set of buckets: the sort phase of range k + 1 (and, for host scalars, its
 * upload) overlaps the bucket accumulation of range k, the bucket reduction is paid once.  dev_parts / host_parts: ranges for
 * scalars in HBM / in host memory (1 = the whole MSM at once); min_log: smallest MSM (log2 points) cut this way.
 * Defaults 1 / 4 / 19 (measured on B200, profiles/r02am: ranges pay for host scalars, 3.24 -> 3.04 ms at 2^20, not for scalars already in
 * HBM).  Results are identical for every setting. */
ZKB_API int zkb_msm_set_parts(zkb_ctx *ctx, int dev_parts, int host_parts, int min_log);
/* Force the window size c (0 = automatic cost model).  For tests and tuning. */
ZKB_API int zkb_msm_set_window(zkb_ctx *ctx, int c);
/* Batched-affine pair rounds in front of the XYZZ bucket accumulation (csrc/msm_pairs.cuh): each round adds
 * entries 2k and 2k + 1 of every bucket in affine coordinates (5M + 1S per addition, one inversion per thread and batch of
 * 32..128 pairs, prefix products in HBM), halving the work left for the 8M + 2S XYZZ additions.  mode: 0 = none, 1..6 =
 * that many rounds, -1 = chosen per MSM from the mean bucket load.  Same results for every mode (affine sums are canonical).
 * ZKB_MSM_MODE=<mode> in the environment sets the mode of contexts created afterwards. */
ZKB_API int zkb_msm_set_mode(zkb_ctx *ctx, int mode);

/* ---- grand products, quotient, polynomial utilities (all device-resident, enqueue on the context's stream) -------------- */
/* compute_z1_poly before its final iFFT (plonk-core/src/permutation/mod.rs:181-254): out[0] = 1,
 * out[i+1] = out[i] * prod_w (w_i + beta*k_w*omega^i + gamma) / (w_i + beta*sigma_w,i + gamma), i < n-1.
 * a,b,c,sigma1..3,out: 2^log_n Fr (Montgomery) in HBM; beta, gamma: host, Montgomery.  Follow with zkb_ntt_dev(inverse). */
ZKB_API int zkb_z1_evals_dev(zkb_ctx *ctx, unsigned log_n, const uint64_t beta[4], const uint64_t gamma[4], const uint64_t *a,
                     const uint64_t *b, const uint64_t *c, const uint64_t *sigma1, const uint64_t *sigma2,
                     const uint64_t *sigma3, uint64_t *out_dev);
/* compute_z2_poly before its final iFFT (plonk-core/src/lookup/mod.rs:25-82). */
ZKB_API int zkb_z2_evals_dev(zkb_ctx *ctx, unsigned log_n, const uint64_t delta[4], const uint64_t epsilon[4], const uint64_t *f,
                     const uint64_t *t, const uint64_t *h1, const uint64_t *h2, uint64_t *out_dev);
/* 1 if the last grand product on this context hit a zero denominator -- the reference panics there
 * (`dominator.inverse().unwrap()`, permutation/mod.rs:242, lookup/mod.rs:73); synchronises the stream. */
ZKB_API int zkb_grand_product_failed(zkb_ctx *ctx);
/* quotient_poly::compute between its coset FFTs and the final coset iFFT (quotient_poly.rs:98-224):
 * out[i] = (arith_i + perm_i + lookup_i) / zh(x_i) for i < 4n on the coset 5*<w_4n>.
 * challenges: alpha, beta, gamma, delta, epsilon (5 x 4 limbs, host, Montgomery).
 * wit[9]:  z1, z2, a, b, c, pi, t, h1, h2 coset evaluations (4n each, HBM).
 * epk[11]: q_m, q_l, q_r, q_o, q_c, q_lookup, q_table, sigma1, sigma2, sigma3, l_1 coset tables (4n each, HBM);
 *          x_coset and zh_coset of the reference's ExtendedProverKey are computed on the fly (zh takes 4 values). */
ZKB_API int zkb_quotient_evals_dev(zkb_ctx *ctx, unsigned log_n, const uint64_t challenges[20], const uint64_t *const wit[9],
                           const uint64_t *const epk[11], uint64_t *out_dev);
/* l_1_coset of extend_prover_key (keys/mod.rs:117-119): 4n values into out_dev. */
ZKB_API int zkb_l1_coset_dev(zkb_ctx *ctx, unsigned log_n, uint64_t *out_dev);
/* DensePolynomial::evaluate (linearization_poly.rs:55-75): out = sum_k coeffs[k] z^k; z, out on the host. Synchronous. */
ZKB_API int zkb_poly_eval_dev(zkb_ctx *ctx, const uint64_t *coeffs_dev, size_t n, const uint64_t z[4], uint64_t out[4]);
/* The same for k <= 16 polynomials in one launch pair and one host round trip (the 12 openings of prove.rs:320-345 share two
 * points): points_host and out_host hold k x 4 words; polynomial j is evaluated at points_host[j]. Synchronous. */
ZKB_API int zkb_poly_eval_many_dev(zkb_ctx *ctx, size_t k, const uint64_t *const *polys_dev, const size_t *lens,
                           const uint64_t *points_host, uint64_t *out_host);
/* out[i] = sum_{j<k} scalars[j] * polys[j][i], i < out_len (k <= 16; shorter polynomials are zero-extended):
 * the axpy chains of linearization_poly.rs:77-111 and SonicKZG10::open's eta-combination. */
ZKB_API int zkb_poly_lincomb_dev(zkb_ctx *ctx, size_t k, const uint64_t *const *polys_dev, const size_t *lens,
                         const uint64_t *scalars_host, uint64_t *out_dev, size_t out_len);
/* kzg10::open's witness polynomial: quot = (p(X) - p(z)) / (X - z) (n-1 coefficients), eval_out = p(z). Synchronous. */
ZKB_API int zkb_poly_divide_linear_dev(zkb_ctx *ctx, const uint64_t *coeffs_dev, size_t n, const uint64_t z[4], uint64_t *quot_dev,
                               uint64_t eval_out[4]);
/* add_blinders_to_poly (prove.rs:472-483): coeffs[len+i] = b_i and coeffs[i] -= b_i for i < k (k <= 16);
 * the buffer must have room for len + k coefficients. */
ZKB_API int zkb_poly_add_blinders_dev(zkb_ctx *ctx, uint64_t *coeffs_dev, size_t len, const uint64_t *blinders_host, size_t k);

/* Number of coefficients left once trailing zeros are dropped, as DensePolynomial::from_coefficients_vec does
 * (blinders are appended at that length: prove.rs:472-483).  Synchronous. */
ZKB_API int zkb_poly_effective_len_dev(zkb_ctx *ctx, const uint64_t *coeffs_dev, size_t n, size_t *out_len);

/* ---- the prover rounds behind two calls ------------------------------------------------------------------------------- */
/* Replaces proof_system::setup (setup.rs:42-166, extend = true) and proof_system::prove (prove.rs:59-470) for
 * BN254 / KZG10 / Merlin.  The committer key must already be resident (zkb_srs_load_g1, >= n + 8 powers; optionally
 * zkb_srs_precompute).  All field inputs are arkworks' in-memory form (Montgomery limbs), exactly what the composer
 * holds after pad_to(n):
 *   selectors[6] = q_m, q_l, q_r, q_o, q_c, q_lookup evaluations (n each); sigma[3] = sigma1..3 evaluations
 *   (compute_all_sigma_evals); table_size = the const generic TABLE_SIZE; pi_positions = sorted public-input rows.
 * The key object owns the selector / sigma polynomials, the 11 coset tables of the ExtendedProverKey, the verifier
 * key commitments and a scratch arena for one proof, all in HBM. */
typedef struct zkb_plonk_pk zkb_plonk_pk;
ZKB_API int zkb_plonk_setup(zkb_ctx *ctx, unsigned log_n, const uint64_t *const selectors[6], const uint64_t *const sigma[3],
                    size_t table_size, const size_t *pi_positions, size_t n_pi, zkb_plonk_pk **out);
ZKB_API void zkb_plonk_pk_destroy(zkb_ctx *ctx, zkb_plonk_pk *pk);
/* The `T: TranscriptProtocol` parameter of ZKTPlonk (plonk.rs:39-46).  kind 0: MerlinTranscript (transcript.rs:49-109,
 * the default binary's); kind 1: EthereumTranscript (gadgets/src/transcript.rs:8-90, bin feature "ethereum-transcript":
 * two chained Keccak-256 states, big-endian items, challenges below 2^253).  Applies to later zkb_plonk_prove calls. */
ZKB_API int zkb_plonk_pk_set_transcript(zkb_plonk_pk *pk, int kind);
/* Round 2's witness plumbing (prove.rs:145-167): LookupTable::into_multiset, f = q_lookup * c and MultiSet::combine_split
 * (multiset.rs:103-146) either sparse on a host thread under round 1 (right while a few percent of the rows are lookup gates) or
 * on the device (csrc/lookup.cu: one hash probe per row, bucket counts by atomics, offsets by one scan, halves by binary
 * search).  mode 0 = the device path when more than n / 8 rows are lookup gates (default), 1 = host, 2 = device.  Same bytes. */
ZKB_API int zkb_plonk_pk_set_lookup_mode(zkb_plonk_pk *pk, int mode);
/* The device path on its own: t, f, h1, h2 (2^log_n Montgomery elements each, device) from the table (HOST pointer, table_len <= n
 * elements in the composer's order), q_lookup's evaluations and the wire c (device).  Enqueued on the context's stream;
 * *status_host (may be NULL) is valid after zkb_ctx_sync: 0 ok, bit 0 ElementNotIndexedInTable, bit 1 halves not n long. */
ZKB_API int zkb_lookup_multisets_dev(zkb_ctx *ctx, unsigned log_n, const uint64_t *table_host, size_t table_len,
                                     const uint64_t *q_lookup_evals_dev, const uint64_t *c_evals_dev, uint64_t *t_dev, uint64_t *f_dev,
                                     uint64_t *h1_dev, uint64_t *h2_dev, int *status_host);
/* Bytes of a serialised Proof on this build's curve (proof.rs:112-154: 11 compressed commitments, two openings, 12
 * evaluations): 802 on BN254 -- the [802] written in the prototypes below -- and 1010 on BLS12-381 / BLS12-377. */
ZKB_API size_t zkb_plonk_proof_bytes(void);
/* VerifierKey commitments in seed_transcript order (keys/mod.rs:264-274): q_m q_l q_r q_o q_c sigma1 sigma2 sigma3
 * q_lookup q_table; 10 x (x || y) Montgomery. */
ZKB_API int zkb_plonk_vk_commitments(const zkb_plonk_pk *pk, uint64_t out_xy[80], int is_inf[10]);
/* a, b, c: wire evaluations (n each, host); table: the lookup table's entries (table_len <= table_size);
 * pi_values: one value per pi_position; blinders: the 19 elements the reference draws with F::rand, in draw order
 * a(2) b(2) c(2) h1(3) h2(2) z1(3) z2(3) b0 b1; proof_out: the 802 bytes of Proof's CanonicalSerialize.
 * timings_ms (optional, 8 floats; non-NULL drains the stream at every boundary): wire upload, round 1, host lookup
 * plumbing, round 2, round 3, round 4, round 5, total.
 * NOT re-entrant per key: although `pk` is const in the signature (it is never changed as a KEY), a proof uses the key's
 * HBM arena, pinned staging and dirty-range bookkeeping as scratch, so one key serves one zkb_plonk_prove at a time (one
 * proving thread per context, as prove() is !Send upstream: prove.rs:62); concurrent proofs need one key object each. */
ZKB_API int zkb_plonk_prove(zkb_ctx *ctx, const zkb_plonk_pk *pk, const uint64_t *a, const uint64_t *b, const uint64_t *c,
                    const uint64_t *table, size_t table_len, const uint64_t *pi_values, const uint64_t *blinders,
                    uint8_t proof_out[802], float timings_ms[8]);
/* ProvingComposer::wire_evals (prove.rs:49-55) on the device.  zkb_plonk_pk_set_wiring gives the key the
 * circuit's wire maps (w_l, w_r, w_o of the padded composer: n variable indices each, 0 = Variable::Zero); zkb_plonk_prove_vars
 * then takes the variable assignment (n_vars Montgomery elements, var_values[0] = 0 for Variable::Zero) instead of the three
 * wire vectors: the assignment crosses PCIe once (typically well under 3n elements) and the wires are gathered in HBM.  Same
 * 802 bytes as zkb_plonk_prove on a = value_of_var(w_l[i]) etc. */
ZKB_API int zkb_plonk_pk_set_wiring(zkb_ctx *ctx, zkb_plonk_pk *pk, const uint32_t *w_l, const uint32_t *w_r, const uint32_t *w_o);
ZKB_API int zkb_plonk_prove_vars(zkb_ctx *ctx, const zkb_plonk_pk *pk, const uint64_t *var_values, size_t n_vars, const uint64_t *table,
                         size_t table_len, const uint64_t *pi_values, const uint64_t *blinders, uint8_t proof_out[802],
                         float timings_ms[8]);

/* ---- verifier: host code, no GPU needed ---------------------------------------------------------------- */
/* Proof::verify (plonk-core/src/proof_system/proof.rs:285-503) for BN254 / KZG10: transcript replay, compute_r0 (:163-217),
 * the 13-point linearisation commitment (:220-282) and PC::check twice (:441-502), each the product of two pairings
 * e(sum eta^i C_i - (sum eta^i v_i) G + z W, h) * e(-W, beta_h) == 1 that SonicKZG10::check evaluates.
 * n, pi_roots (Montgomery Fr), vk_xy / vk_inf: the VerifierKey (zkb_vk_file_read / zkb_plonk_vk_commitments layout);
 * pub_inputs: one Montgomery Fr per root; proof: the zkb_plonk_proof_bytes() bytes of Proof's CanonicalSerialize (802 / 1010);
 * g2_h, g2_beta_h: the G2 half of sonic_pc::VerifierKey, each x.c0 x.c1 y.c0 y.c1 (fq_words limbs each, Montgomery Fq: arkworks'
 * in-memory G2Affine; 16 words on BN254, 24 on the BLS12 curves).  Every build verifies with its own curve's pairing;
 * transcript_kind as zkb_plonk_pk_set_transcript.
 * Returns 0 = accepted, 1 / 2 = Error::ProofVerificationError { step }, negative = malformed input (a point off its
 * curve, a non-canonical integer ...). */
ZKB_API int zkb_plonk_verify(size_t n, const uint64_t *pi_roots_mont, size_t n_pi, const uint64_t vk_xy[80], const int vk_inf[10],
                     const uint64_t *pub_inputs_mont, const uint8_t proof[802], const uint64_t g2_h[16], const uint64_t g2_beta_h[16],
                     int transcript_kind);
/* e(P, Q) on this build's curve (optimal ate; ark-ec 0.3 PairingEngine::pairing): the 12 coefficients of the result in the basis
 * 1, w, .., w^11 of Fq12 = Fq[w] / (w^12 - 2 xi0 w^6 + xi0^2 + b) with w^6 = xi0 + i, i^2 = -b (BN254: w^12 - 18 w^6 + 82; BLS12-381:
 * w^12 - 2 w^6 + 2; BLS12-377: w^12 + 5), canonical integers (fq_words limbs each).  For tests against oracle/pairing.py and
 * oracle/pairing_bls.py; zkb_pairing_product_is_one is what the verifier uses (PairingEngine::product_of_pairings == 1). */
ZKB_API int zkb_pairing(const uint64_t g1_xy[8], const uint64_t g2_xy[16], uint64_t out_canonical[48]);
ZKB_API int zkb_pairing_product_is_one(const uint64_t *g1_xy, const uint64_t *g2_xy, size_t count, int *is_one);
/* out = k * Q on G2 (k canonical, Q and out as x.c0 x.c1 y.c0 y.c1 Montgomery limbs; identity = zeros).  Host code, setup
 * time only: ark-poly-commit 0.3 kzg10::setup's beta_h = beta * h, needed to verify against a synthetic SRS (bench.py). */
ZKB_API int zkb_g2_mul(const uint64_t g2_xy[16], const uint64_t scalar_canonical[4], uint64_t out_xy[16]);

/* ---- inner-product-argument commitments (the reference's second PC) ------------------------------------- */
/* `IPA<G, D>` (plonk-core/src/commitment.rs:49-86) = ark-poly-commit 0.3 ipa_pc::InnerProductArgPC over G1 (the PC of half of
 * plonk-core/src/test.rs's runs, :73,84).  commit = the MSM above over ck.comm_key.  open folds three vectors of length n = d + 1
 * (coefficients c, powers of the point z, key G) log2 n times; per round
 *     L = <c_r, G_l> + <c_r, z_l> h',  R = <c_l, G_r> + <c_l, z_r> h',  x = H(x_prev, L, R),
 *     c_l += x^-1 c_r,  z_l += x z_r,  G_l += x G_r.
 * zkb_ipa_round_lr_dev returns L, R and the two inner products of a round (h_prime_xy: h' = x_0 h as an affine Montgomery point
 * in host memory, or NULL for the two MSMs alone; the caller hashes: the transcript stays in the host language);
 * zkb_ipa_round_fold_dev folds in place -- afterwards the first n / 2 entries of
 * each vector are the next round's.  All vectors live in HBM: c, z = n x 4 words Montgomery, G = n affine points.  n: a power
 * of two >= 2.  x, x_inv Montgomery (x * x_inv == 1 is checked).  Results are field / group elements: identical to the CPU's. */
ZKB_API int zkb_ipa_round_lr_dev(zkb_ctx *ctx, const uint64_t *coeffs_dev, const uint64_t *z_dev, const uint64_t *key_dev, size_t n,
                         const uint64_t *h_prime_xy, uint64_t l_xy[8], int *l_inf, uint64_t r_xy[8], int *r_inf, uint64_t ip_l[4],
                         uint64_t ip_r[4]);
ZKB_API int zkb_ipa_round_fold_dev(zkb_ctx *ctx, uint64_t *coeffs_dev, uint64_t *z_dev, uint64_t *key_dev, size_t n, const uint64_t x[4],
                           const uint64_t x_inv[4]);
/* The verifier's linear-time step (ipa_pc::check after succinct_check): <h, G> over the whole key, h(X) = prod_j (1 + x_j X^(2^(k-1-j)))
 * expanded in HBM (SuccinctCheckPolynomial::compute_coeffs); equals the proof's final_comm_key for an honest proof.
 * challenges_mont: the k = log2 n round challenges in proof order, Montgomery form. */
ZKB_API int zkb_ipa_final_key_dev(zkb_ctx *ctx, const uint64_t *key_dev, size_t n, const uint64_t *challenges_mont, uint64_t out_xy[8],
                          int *is_inf);

/* ---- the reference CLI's key files ---------------------------------------------------------------------- */
/* `compile` writes ck / cvk / pk / epk / vk with ark-serialize 0.3 serialize_unchecked (bin/src/parser.rs:16-29,
 * main.rs:96-113); `prove-withdraw` reads them back before every proof (parser.rs:5-14, main.rs:274-281).  Byte
 * layout: csrc/keyfile.cu's header.  File side: canonical little-endian integers, identity = (0, 1) + flag bit;
 * memory side: this library's forms (Montgomery limbs, identity = (0, 0)).  The *_file_* functions are host code
 * and need no GPU.  The epk file is never read: its coset tables are rebuilt in HBM (10 coset NTTs).
 * ck = sonic_pc::CommitterKey<Bn254> as PC::trim(pp, 4n, 0, None) leaves it (plonk.rs:79-85): powers_of_g,
 * powers_of_gamma_g, three `None`s, max_degree.  On the BLS12 builds the same layouts with 48-byte base-field elements (a G1
 * point is 96 file bytes and 12 words in memory, a G2 point 192 bytes and 24 words): the key types are generic over the pairing
 * engine (keys/mod.rs:29-40,180-203).  Array sizes below are BN254's. */
ZKB_API int zkb_ck_file_info(const char *path, size_t *n_powers, size_t *max_degree);
ZKB_API int zkb_ck_file_read(const char *path, size_t first, size_t count, uint64_t *xy_mont_out);
ZKB_API int zkb_ck_file_write(const char *path, const uint64_t *xy_mont, size_t n_powers, const uint64_t *gamma_xy_mont, size_t n_gamma,
                      size_t max_degree);
/* cvk = sonic_pc::VerifierKey<Bn254>: only its first four fields are read -- g, gamma_g (G1; either may be NULL) and
 * h, beta_h (G2: x.c0 x.c1 y.c0 y.c1, Montgomery), which is what zkb_plonk_verify takes. */
ZKB_API int zkb_cvk_file_read(const char *path, uint64_t g_xy[8], uint64_t gamma_g_xy[8], uint64_t h_xy[16], uint64_t beta_h_xy[16]);
/* deserialize_from_file::<CommitterKey>(ck_path) + keeping powers_of_g resident: the first min(max_points, all) powers
 * become the context's SRS (max_points = 0: all). */
ZKB_API int zkb_srs_load_ck_file(zkb_ctx *ctx, const char *path, size_t max_points);
/* ProverKey<Fr> (keys/mod.rs:29-40): ten LabeledPolynomials in the order q_m q_l q_r q_o q_c sigma1 sigma2 sigma3
 * q_lookup q_table, labels as setup.rs:93-102, no degree or hiding bounds.  lens: coefficients per polynomial. */
ZKB_API int zkb_pk_file_info(const char *path, size_t lens[10]);
ZKB_API int zkb_pk_file_read(const char *path, uint64_t *const coeffs_mont_out[10], const size_t caps[10], size_t lens[10]);
ZKB_API int zkb_pk_file_write(const char *path, const uint64_t *const coeffs_mont[10], const size_t lens[10]);
/* VerifierKey<Fr, KZG10<Bn254>> (keys/mod.rs:180-203): n, pi_roots, ten commitments in seed_transcript order.
 * pi_roots_mont may be NULL (only *n_roots is reported); at most cap_roots roots are stored. */
ZKB_API int zkb_vk_file_read(const char *path, size_t *n, uint64_t *pi_roots_mont, size_t cap_roots, size_t *n_roots,
                     uint64_t commits_xy[80], int is_inf[10]);
ZKB_API int zkb_vk_file_write(const char *path, size_t n, const uint64_t *pi_roots_mont, size_t n_roots, const uint64_t commits_xy[80],
                      const int is_inf[10]);
/* A proving key from the ProverKey's coefficient-form polynomials (host, Montgomery, pk-file order) instead of the
 * composer's evaluation columns; vk_xy (10 x 8 limbs, may be NULL: the commitments are recomputed) and vk_inf (may be
 * NULL) are the VerifierKey's commitments.  Proves byte-identically to the key zkb_plonk_setup builds. */
ZKB_API int zkb_plonk_pk_from_polys(zkb_ctx *ctx, unsigned log_n, const uint64_t *const polys[10], const size_t lens[10],
                            size_t table_size, const size_t *pi_positions, size_t n_pi, const uint64_t *vk_xy, const int *vk_inf,
                            zkb_plonk_pk **out);
/* pk + vk files -> proving key (public-input rows are recovered from vk.pi_roots = omega^row, setup.rs:123);
 * proving key -> pk + vk files the reference CLI can read (either path may be NULL). */
ZKB_API int zkb_plonk_load_keys(zkb_ctx *ctx, const char *pk_path, const char *vk_path, size_t table_size, zkb_plonk_pk **out);
ZKB_API int zkb_plonk_save_keys(zkb_ctx *ctx, const zkb_plonk_pk *pk, const char *pk_path, const char *vk_path);

/* ---- multi-GPU: one process per GPU, commitments sharded by point range ----------------------------------- */
/* The reference is single-process (rayon threads only), so these have no reference counterpart; they carry the
 * exchange step of a sharded kzg10::commit.  Every rank runs the same prover (SPMD) on the same witness; a rank keeps
 * one contiguous range of powers_of_g resident (zkb_srs_load_g1* with its range, then zkb_srs_set_range), runs the
 * bucket method on the matching slice of every polynomial, and the XYZZ partial sums (128 B per rank and
 * commitment) are all-gathered over NCCL (NVLink / NVSwitch) and added identically on every rank, so all ranks
 * produce the same transcript and the same proof bytes.  NCCL is dlopen'ed by zkb_comm_init.
 *   rank 0: zkb_comm_unique_id(id); the host layer broadcasts the 128 bytes (torch.distributed, MPI, a socket ...);
 *   all:    zkb_comm_init(ctx, id, rank, world)  (collective). */
ZKB_API int zkb_comm_unique_id(uint8_t out[128]);
ZKB_API int zkb_comm_init(zkb_ctx *ctx, const uint8_t id[128], int rank, int world);
ZKB_API int zkb_comm_destroy(zkb_ctx *ctx);
ZKB_API int zkb_comm_rank(zkb_ctx *ctx);
ZKB_API int zkb_comm_world(zkb_ctx *ctx);
/* recv_host[r * bytes ..] = rank r's send_host[0 .. bytes) on every rank (collective, synchronous). */
ZKB_API int zkb_comm_allgather_host(zkb_ctx *ctx, const void *send_host, size_t bytes, void *recv_host);
/* Declare the resident SRS to be the range [global_lo, global_lo + resident points) of a committer key of global_n
 * powers.  zkb_commit_dev / zkb_commit_batch_dev / zkb_plonk_setup / zkb_plonk_prove then take GLOBAL offsets and
 * lengths; zkb_msm_g1* keep addressing the resident range.  zkb_srs_size reports global_n. */
ZKB_API int zkb_srs_set_range(zkb_ctx *ctx, size_t global_lo, size_t global_n);
/* The other layout: every rank of the communicator holds the WHOLE committer key (load it on every rank, then call this
 * before zkb_srs_precompute).  Commitments still take global offsets and lengths; per batch the library either cuts every
 * commitment over all ranks or gives each commitment of the batch its own group of ranks (the 3 / 3 /
 * 2 / 3 independent commitments of prove.rs:134,179,250,307), by a cost model fitted to measurements (fanout = -1), or as
 * forced (0 = shard, 1 = fan out).  Results are identical for every layout. */
ZKB_API int zkb_srs_set_replicated(zkb_ctx *ctx, int fanout);
/* How many commitments the next zkb_commit_push batch will hold (zkb_commit_batch_dev knows; an incremental batch does
 * not): needed to fan a batch out over a replicated key.  Without it every pushed commitment is cut over all ranks. */
ZKB_API int zkb_commit_expect(zkb_ctx *ctx, size_t count);

/* ---- test hooks (parity of the device field library against the oracle) ----------------------------------------- */
/* field: 0 = Fr, 1 = Fq;  op: 0 mul, 1 add, 2 sub, 3 sqr(a), 4 inv(a), 5 to_mont(a), 6 from_mont(a).  Host pointers. */
ZKB_API int zkb_test_fp_binop(zkb_ctx *ctx, int field, int op, uint64_t *out, const uint64_t *a, const uint64_t *b, size_t n);

/* Host only (no GPU): the round driver's MultiSet::combine_split (multiset.rs:103-146) specialised to t = table entries
 * + zero padding up to n and f zero outside `rows`.  h1 / h2 (n elements each) are persistent staging kept zero outside
 * the regions recorded in dirty[4] = {h1 prefix end, h1 suffix start, h2 prefix end, h2 suffix start} (start with
 * zeroed buffers and {0, n, 0, n}).  Returns ZKB_ERR_INVALID for ElementNotIndexedInTable. */
ZKB_API int zkb_test_combine_split(const uint64_t *table, size_t table_len, size_t n, const uint64_t *f, const uint32_t *rows,
                           size_t n_rows, uint64_t *h1, uint64_t *h2, size_t dirty[4], size_t out_lens[2]);

/* Test hook: set rank / world of a context without creating a communicator (single-GPU tests of the multi-GPU layouts:
 * zkb_commit_push + zkb_commit_finish_partials only). */
/* the split k = k1 + lambda k2 (mod r) of csrc/ipa.cu's key fold (magnitudes and signs; returns 1 when short and verified) */
ZKB_API int zkb_test_glv_split(const uint64_t k[4], uint64_t m1[4], int *neg1, uint64_t m2[4], int *neg2);
ZKB_API int zkb_test_set_rank_world(zkb_ctx *ctx, int rank, int world);
/* Host only (no GPU): [lo, hi) of commitment k (of a batch of E, `len` coefficients at SRS offset `offset`) that `rank` of
 * `world` computes over a replicated key (zkb_srs_set_replicated); fanout as there.  lo == hi: nothing. */
ZKB_API int zkb_test_replicated_share(int world, int rank, int fanout, size_t E, size_t k, size_t offset, size_t len, size_t out_lo_hi[2]);

/* Host only (no GPU): run a scripted transcript and return its challenges (32 B canonical little-endian each).
 * ops[i]: 0 append_u64(args[9i]), 1 append_scalar(args[9i..9i+4), Montgomery Fr), 2 append_commitment(x = args[9i..],
 * y = args[9i+4..], Montgomery Fq; args[9i+8] != 0: identity), 3 challenge_scalar.  The Merlin transcript is created
 * with the label "test".  Pins the C++ EthereumTranscript against gadgets/src/transcript.rs:100-127. */
ZKB_API int zkb_test_transcript(int kind, const uint8_t *ops, size_t n_ops, const uint64_t *args, uint8_t *challenges_out);

/* Kernels this context has enqueued so far (bench.py's gpu_launches is a difference of two readings). */
ZKB_API uint64_t zkb_launch_count(zkb_ctx *ctx);
/* Device time (CUDA events on the context's stream) of the phases of the last MSM, in ms:
 * out_ms[0] digits + counting sort + task ordering, [1] bucket accumulation (the integer-bound kernel),
 * [2] oversized-bucket combine, [3] window reduction, [4] total.  info[0] = n * windows, info[1] = c, info[2] = windows. */
ZKB_API int zkb_msm_last_timing(zkb_ctx *ctx, float out_ms[5], uint64_t info[3]);
/* The part of out_ms[0] above spent in the batched-affine pair rounds of the last MSM, and how many rounds ran. */
ZKB_API int zkb_msm_last_pair_rounds(zkb_ctx *ctx, float *pairs_ms, int *rounds);

/* ---- measurement: integer-pipe peak (not in MEASURED_PEAKS.json; SURVEY.md 8d asks for it) ------------------------- */
/* mode 0: 32-bit IMAD/s, mode 1: IMAD.WIDE.U32/s (the instruction the Montgomery product is made of),
 * mode 2: Fq Montgomery products/s in a dependency-chained loop, mode 3: FP64 FMA/s, modes 6-8: pipe-sharing probes
 * (IMAD + DFMA, IMAD + ALU, DFMA + ALU in one thread; result = groups/s).  All 148 SMs, best of 3 timed launches. */
ZKB_API int zkb_bench_int(zkb_ctx *ctx, int mode, double *ops_per_sec);

#ifdef __cplusplus
}
#endif
#endif /* ZKB200_H */
