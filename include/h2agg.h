/* h2agg.h -- C ABI of the B200-native prover backend for the halo2 aggregation circuit.
 *
 * This is the drop-in boundary.  The reference
 * (scroll-tech/halo2-snark-aggregator) has no FFI today; its hot path enters the external
 * crate halo2_proofs at
 *     halo2-snark-aggregator-circuit/src/verify_circuit.rs:986-994   create_proof(...)
 *     halo2-snark-aggregator-circuit/src/verify_circuit.rs:974-979   keygen_pk(...)
 *     halo2-snark-aggregator-circuit/src/verify_circuit.rs:760-761   keygen_vk(...)
 *     halo2-snark-aggregator-circuit/src/sample_circuit.rs:75-83     create_proof(...) (inner proofs)
 * and from there reaches the free functions this library replaces (halo2_proofs is swapped
 * with a Cargo [patch], precedent: /root/reference/Cargo.toml:10-11; the Rust binding is in
 * INTEGRATION.md and rust/h2agg-sys/).
 *
 * Data layout everywhere = the Rust in-memory layout of halo2curves 0.2.1:
 *   Fr, Fq      4 x u64 little-endian limbs, Montgomery form (R = 2^256), fully reduced
 *   G1Affine    {x: Fq, y: Fq}            64 bytes, identity = (0, 0)
 *   G1          {x: Fq, y: Fq, z: Fq}     96 bytes Jacobian, identity z = 0
 * so a Rust slice can be passed as a pointer without conversion.
 *
 * Every function returns 0 on success; 1 = invalid argument, 2 = CUDA failure, 3 = no device,
 * 4 = the data violate the caller's constraint system (h2agg_permute_expression_pair only).
 * h2agg_last_error() gives the message.  There is no CPU fallback inside this library: with no
 * usable GPU h2agg_init fails and nothing else can be called.
 *
 * Threading: a context serialises its entry points internally (best_fft is entered from rayon
 * workers in the reference, halo2-snark-aggregator-sdk/src/lib.rs:52-55).  Host-pointer entry
 * points return when the result is in host memory; *_dev entry points enqueue on the context's
 * stream and return immediately.
 */
#ifndef H2AGG_H
#define H2AGG_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct h2agg_ctx h2agg_ctx;

/* ---- life cycle ------------------------------------------------------------------------- */
/* One context per process per GPU (one process per GPU is the multi-GPU model). */
int h2agg_init(int device_id, h2agg_ctx** out);
void h2agg_destroy(h2agg_ctx* ctx);
const char* h2agg_last_error(h2agg_ctx* ctx); /* ctx may be NULL: error of the last failed init */
const char* h2agg_version(void);
/* Run on a caller-owned CUDA stream (cudaStream_t) instead of the context's own.  Switching streams
 * does not synchronise: order them with events, as with any CUDA work.  MSM batches (lanes + their own
 * workspaces) and NTTs (the context's ping-pong buffer) share no scratch, so a caller may run one
 * NTT stream concurrently with MSM batches issued on another stream. */
int h2agg_set_stream(h2agg_ctx* ctx, void* cuda_stream);
int h2agg_synchronize(h2agg_ctx* ctx);
/* Number of kernels this context has launched so far (evidence counter for bench.py). */
uint64_t h2agg_launch_count(h2agg_ctx* ctx);
/* Per-kernel-class device timing with CUDA events on the context's stream (bench.py roofline).
 * classes: 0 msm_accumulate, 1 msm digit/sort kernels, 2 msm bucket+window reduction, 3 ntt pass,
 * 4 whole MSM, 5 witness expansion, 6 evaluate_h.  h2agg_kernel_times synchronises, returns the sums since the last call and resets. */
int h2agg_kernel_timing(h2agg_ctx* ctx, int enable);
int h2agg_kernel_times(h2agg_ctx* ctx, double* ms_per_class, uint64_t* count_per_class, int n_classes);
/* Pin / unpin a host range so H2D copies run at PCIe speed (the SRS, witness columns). */
int h2agg_host_register(h2agg_ctx* ctx, const void* p, size_t bytes);
int h2agg_host_unregister(h2agg_ctx* ctx, const void* p);
/* MSM window width c in bits (0 = automatic). Exposed for sweeps and tests. */
int h2agg_set_msm_window(h2agg_ctx* ctx, int c_bits);
/* Test hook: cap the log2 radix of an NTT pass (default 8) so the multi-pass code paths (up to 4 passes)
 * can be checked against the CPU oracle at small sizes; log_n must stay <= 4 * cap. */
int h2agg_set_ntt_radix_cap(h2agg_ctx* ctx, int log2_radix);
/* Fixed-base tables: when an SRS is registered, precompute 2^(c w) P_i for every window w
 * (W x the SRS size in HBM, c up to 20) so all windows share one bucket set.  Default on. */
int h2agg_set_srs_precompute(h2agg_ctx* ctx, int enable);
/* (table_mode, c, number of windows) MSMs against this SRS use. */
int h2agg_srs_config(h2agg_ctx* ctx, uint64_t srs_id, int* table_mode, int* c_bits, int* n_windows);
/* The (c, number of windows) the MSM uses for n pairs with per-call bases (plain mode). */
int h2agg_msm_config(h2agg_ctx* ctx, size_t n, int* c_bits, int* n_windows);

/* ---- SRS residency -------------------------------------------------------------------------
 * ParamsKZG { g, g_lagrange } are immutable for the life of the params
 * (verify_circuit.rs:701-731 get_params_cached); register them once, keep them in HBM. */
int h2agg_srs_register(h2agg_ctx* ctx, const uint64_t* bases_affine /* n*8 */, size_t n, uint64_t* out_srs_id);
int h2agg_srs_register_dev(h2agg_ctx* ctx, const void* d_bases_affine, size_t n, uint64_t* out_srs_id); /* borrowed */
int h2agg_srs_release(h2agg_ctx* ctx, uint64_t srs_id);

/* ---- K1: multi-scalar multiplication  (replaces halo2_proofs::arithmetic::best_multiexp,
 *      reached via ParamsKZG::commit_lagrange / commit from create_proof, verify_circuit.rs:986)
 * result = sum_i scalars[i] * bases[i].  Bases come from srs_id (first n points) or, when
 * srs_id == 0, from bases_affine.  out_jacobian receives the NORMALISED point (x, y, 1), or
 * (0, 1, 0) for the identity, so that its first 64 bytes are the affine coordinates the
 * transcript writes (halo2-snark-aggregator-api/src/transcript/sha.rs:156-173). */
int h2agg_msm_g1(h2agg_ctx* ctx, uint64_t srs_id, const uint64_t* bases_affine /* n*8 or NULL */,
                 const uint64_t* scalars /* n*4 */, size_t n, uint64_t out_jacobian[12]);
/* One commit round: n_cols columns of n scalars against the same bases -> n_cols affine points. */
int h2agg_msm_g1_batch(h2agg_ctx* ctx, uint64_t srs_id, const uint64_t* const* scalar_cols, size_t n_cols, size_t n,
                       uint64_t* out_affine /* n_cols*8 */);
/* Same with device-resident columns; d_out160s receives n_cols x (affine 64 B + Jacobian 96 B).
 * Internally alternates two streams so the latency-bound tail of one MSM overlaps the bucket
 * accumulation of the next; joined back into the context's stream before returning. */
int h2agg_msm_g1_batch_dev(h2agg_ctx* ctx, uint64_t srs_id, const void* d_bases_affine, const void* const* d_scalar_cols,
                           size_t n_cols, size_t n, void* d_out160s);
/* Device-resident variant: d_scalars (n*32 B) and bases already in HBM; d_out160 receives
 * affine (64 B) followed by the normalised Jacobian (96 B).  Asynchronous. */
int h2agg_msm_g1_dev(h2agg_ctx* ctx, uint64_t srs_id, const void* d_bases_affine, const void* d_scalars, size_t n,
                     void* d_out160);
/* Window-sharded MSM: only windows [win_begin, win_end) of the signed-digit
 * decomposition; the partial results of all shards add up to the full MSM. Host pointers. */
int h2agg_msm_g1_windows(h2agg_ctx* ctx, uint64_t srs_id, const uint64_t* bases_affine, const uint64_t* scalars,
                         size_t n, int win_begin, int win_end, uint64_t out_jacobian[12]);
int h2agg_msm_g1_windows_dev(h2agg_ctx* ctx, uint64_t srs_id, const void* d_bases_affine, const void* d_scalars,
                             size_t n, int win_begin, int win_end, void* d_out160);
int h2agg_msm_g1_batch_windows_dev(h2agg_ctx* ctx, uint64_t srs_id, const void* d_bases_affine,
                                   const void* const* d_scalar_cols, size_t n_cols, size_t n, int win_begin, int win_end,
                                   void* d_out160s);
/* The same with a window range PER COLUMN (win_begins[i], win_ends[i]; win_end < 0 = all windows): everything a rank
 * owes to one commit phase -- whole columns and window shards of others -- goes out in one call, so the latency-bound
 * prologue / epilogue of a shard (digit passes, scans, bucket tree) runs on its lane beside the other columns' bucket
 * accumulation instead of after it. */
int h2agg_msm_g1_batch_ranges_dev(h2agg_ctx* ctx, uint64_t srs_id, const void* d_bases_affine,
                                  const void* const* d_scalar_cols, size_t n_cols, size_t n, const int* win_begins,
                                  const int* win_ends, void* d_out160s);
/* Device-side combine of all-gathered partials: out[j] = sum_{i<m} P(i, j), the Jacobian point P(i, j)
 * living at d_points + i * stride_bytes + j * 160; writes n_out x 160 B (affine + Jacobian). */
int h2agg_g1_sum_dev(h2agg_ctx* ctx, const void* d_points, size_t m, size_t stride_bytes, size_t n_out, void* d_out160s);
/* Sum of m Jacobian points (96 B each; the all-gathered shard partials) -> normalised Jacobian. */
int h2agg_g1_sum(h2agg_ctx* ctx, const uint64_t* points_jacobian /* m*12 */, size_t m, uint64_t out_jacobian[12]);

/* ---- K2: NTT  (replaces halo2_proofs::arithmetic::best_fft(a, omega, log_n)) -----------------
 * In place, natural order in and out. log_n <= 28. */
int h2agg_ntt_fr(h2agg_ctx* ctx, uint64_t* a /* 2^log_n * 4 */, const uint64_t omega[4], uint32_t log_n);
/* EvaluationDomain::ifft / lagrange_to_coeff: best_fft(a, omega_inv) then * n_inv (= ifft_divisor). */
int h2agg_intt_fr(h2agg_ctx* ctx, uint64_t* a, const uint64_t omega_inv[4], const uint64_t n_inv[4], uint32_t log_n);
/* A round of inverse transforms (one per committed column), in place, pipelined over the context's
 * lanes: H2D of column i+1, the passes of column i and D2H of column i-1 overlap. */
int h2agg_intt_fr_batch(h2agg_ctx* ctx, uint64_t* const* cols, size_t n_cols, const uint64_t omega_inv[4],
                        const uint64_t n_inv[4], uint32_t log_n);
/* Device-resident, asynchronous; scale may be NULL. d_a in place. */
int h2agg_ntt_fr_dev(h2agg_ctx* ctx, void* d_a, const uint64_t omega[4], const uint64_t* scale /* 4 or NULL */,
                     uint32_t log_n);

/* ---- K3: coset transforms of EvaluationDomain ----------------------------------------------------
 * coeff_to_extended: out[i] = NTT_{2^ext_k}( zeta^(i mod 3) * coeffs[i], zero padded )   (n = 2^k inputs)
 * extended_to_coeff: a = iNTT_{2^ext_k}(a) * ext_n_inv, then a[i] *= zeta^-(i mod 3); the first
 *                    out_len elements are the result (halo2 truncates to n*(j-1)). */
int h2agg_coeff_to_extended(h2agg_ctx* ctx, const uint64_t* coeffs /* 2^k * 4 */, uint32_t k, uint32_t ext_k,
                            const uint64_t zeta[4], const uint64_t omega_ext[4], uint64_t* out /* 2^ext_k * 4 */);
int h2agg_extended_to_coeff(h2agg_ctx* ctx, uint64_t* a /* 2^ext_k * 4, in place */, uint32_t ext_k,
                            const uint64_t omega_ext_inv[4], const uint64_t ext_n_inv[4], const uint64_t zeta[4],
                            size_t out_len);
int h2agg_coeff_to_extended_batch(h2agg_ctx* ctx, const uint64_t* const* coeff_cols, uint64_t* const* out_cols,
                                  size_t n_cols, uint32_t k, uint32_t ext_k, const uint64_t zeta[4],
                                  const uint64_t omega_ext[4]);
/* One commit round, fused per column and pipelined over the lanes: every Lagrange column is uploaded ONCE,
 * committed against the SRS (commit_lagrange), turned into coefficients (lagrange_to_coeff) and, when ext_out is
 * given, evaluated on the extended coset (coeff_to_extended) -- the three things create_proof does with each
 * committed column.  out_affine: n_cols*8; coeff_out[i]: 2^k*4 (may alias the
 * input) or NULL array; ext_out[i]: 2^ext_k*4 or NULL array / NULL entries. */
int h2agg_commit_round(h2agg_ctx* ctx, uint64_t srs_id, const uint64_t* const* lagrange_cols, size_t n_cols, uint32_t k,
                       const uint64_t omega_inv[4], const uint64_t n_inv[4], uint64_t* out_affine,
                       uint64_t* const* coeff_out, uint32_t ext_k, const uint64_t zeta[4], const uint64_t omega_ext[4],
                       uint64_t* const* ext_out);
/* Resident form of the commit round: columns still arrive from HOST memory (the witness is produced by the caller),
 * but their coefficient and extended-coset forms stay in caller-owned DEVICE buffers (d_coeff_out[i]: 2^k*32 B,
 * d_ext_out[i]: 2^ext_k*32 B or NULL array / NULL entries) where evaluate_h (N1), the evaluation round and the GWC
 * quotients (N2) consume them, so that only the 64-byte commitments travel back over PCIe. */
int h2agg_commit_round_resident(h2agg_ctx* ctx, uint64_t srs_id, const uint64_t* const* lagrange_cols, size_t n_cols,
                                uint32_t k, const uint64_t omega_inv[4], const uint64_t n_inv[4], uint64_t* out_affine,
                                void* const* d_lagrange_out /* NULL, or per column NULL / 2^k*32 B: keep the Lagrange form too */,
                                void* const* d_coeff_out, uint32_t ext_k, const uint64_t zeta[4],
                                const uint64_t omega_ext[4], void* const* d_ext_out);
/* The same round for Lagrange columns that are ALREADY in HBM (columns the device produced itself: the permuted
 * lookup columns and the grand products of the second and third round).  The commitments come back to the host. */
int h2agg_commit_round_dev(h2agg_ctx* ctx, uint64_t srs_id, const void* const* d_lagrange_cols, size_t n_cols, uint32_t k,
                           const uint64_t omega_inv[4], const uint64_t n_inv[4], uint64_t* out_affine,
                           void* const* d_coeff_out, uint32_t ext_k, const uint64_t zeta[4], const uint64_t omega_ext[4],
                           void* const* d_ext_out);
/* Deferred transforms.  create_proof needs a column's COMMITMENT before the next Fiat-Shamir challenge, but its coefficient
 * and extended forms only in the quotient / evaluation stages.  With enable = 1, h2agg_commit_round_resident / _dev put
 * their lagrange_to_coeff / coeff_to_extended passes on the context's low-priority background stream (ordered after the
 * column's upload) and return as soon as the commitments are there, so the pipe-bound NTT passes fill the latency-bound
 * stretches of the following rounds (sorts, scans, MSM tails, host round trips).  h2agg_transforms_join makes the main
 * stream wait for them; h2agg_synchronize and h2agg_memcpy_d2h join implicitly.  Default off. */
int h2agg_set_defer_transforms(h2agg_ctx* ctx, int enable);
int h2agg_transforms_join(h2agg_ctx* ctx);
/* The transform half of a commit round alone -- lagrange_to_coeff (+ coeff_to_extended when d_ext_out), out of place --
 * for columns whose commitment is computed elsewhere (window-sharded over other GPUs).  Background stream when deferred. */
int h2agg_transforms_dev(h2agg_ctx* ctx, const void* const* d_lagrange_cols, size_t n_cols, uint32_t k,
                         const uint64_t omega_inv[4], const uint64_t n_inv[4], void* const* d_coeff_out, uint32_t ext_k,
                         const uint64_t zeta[4], const uint64_t omega_ext[4], void* const* d_ext_out);
int h2agg_coeff_to_extended_dev(h2agg_ctx* ctx, const void* d_coeffs, uint32_t k, uint32_t ext_k,
                                const uint64_t zeta[4], const uint64_t omega_ext[4], void* d_out);
int h2agg_extended_to_coeff_dev(h2agg_ctx* ctx, void* d_a, uint32_t ext_k, const uint64_t omega_ext_inv[4],
                                const uint64_t ext_n_inv[4], const uint64_t zeta[4], size_t out_len);

/* ---- N2 (first "next" row): evaluation round + GWC quotients on the device -------
 * eval_polynomial(poly, point) = sum_i poly[i] * point^i   (halo2_proofs::arithmetic::eval_polynomial)
 * kate_division(a, b): quotient of a(X) by (X - b), n - 1 coefficients, remainder dropped
 *                      (halo2_proofs::arithmetic::kate_division).  The _dev form writes n entries
 *                      (the last one is zero). */
int h2agg_eval_polynomial(h2agg_ctx* ctx, const uint64_t* poly /* n*4 */, size_t n, const uint64_t point[4], uint64_t out[4]);
int h2agg_eval_polynomial_dev(h2agg_ctx* ctx, const void* d_poly, size_t n, const uint64_t point[4], void* d_out32);
/* n_polys polynomials of n coefficients each at ONE point -> d_out[i] = polys[i](point) (n_polys x 32 B): what the
 * evaluation round does per opening point, in five launches for the whole group instead of five per polynomial. */
int h2agg_eval_polynomials_dev(h2agg_ctx* ctx, const void* const* d_polys, size_t n_polys, size_t n, const uint64_t point[4],
                               void* d_out);
int h2agg_kate_division(h2agg_ctx* ctx, const uint64_t* a /* n*4 */, size_t n, const uint64_t b[4], uint64_t* q /* (n-1)*4 */);
int h2agg_kate_division_dev(h2agg_ctx* ctx, const void* d_a, size_t n, const uint64_t b[4], void* d_q /* n*32 B */);

/* ---- N3 (next row): the scan primitives of the permutation / lookup grand products ----------------
 * batch_invert: a[i] <- 1/a[i] in place, zeros stay zero (halo2_proofs BatchInvert semantics).
 * grand_product: z[0] = 1, z[i+1] = z[i] * num[i] / den[i]  (n outputs; the running product columns). */
int h2agg_batch_invert(h2agg_ctx* ctx, uint64_t* a /* n*4, in place */, size_t n);
int h2agg_batch_invert_dev(h2agg_ctx* ctx, void* d_a, size_t n);
int h2agg_grand_product(h2agg_ctx* ctx, const uint64_t* num, const uint64_t* den, size_t n, uint64_t* z /* n*4 */);
int h2agg_grand_product_dev(h2agg_ctx* ctx, const void* d_num, const void* d_den, size_t n, void* d_z);

/* N3, sorting half: halo2_proofs plonk/lookup/prover.rs permute_expression_pair over the usable rows
 * (n - blinding_factors - 1; the caller appends its random blinding rows):
 *   permuted_input = the input sorted by Fr::cmp (numeric order of the canonical value);
 *   permuted_table[row] = permuted_input[row] on every row where the sorted input changes value (one instance of that
 *   value leaves the table multiset); the remaining table values, ascending, fill the rows of repeated inputs taken
 *   from the back.  The verifier's view of the result: halo2-snark-aggregator-api/src/systems/halo2/lookup.rs:58-119.
 * Returns 4 (halo2: Error::ConstraintSystemFailure) when an input value does not occur in the table; the outputs are
 * then unspecified.  Outputs may alias the inputs.  The _dev form synchronises the stream to read that status. */
int h2agg_permute_expression_pair(h2agg_ctx* ctx, const uint64_t* input /* u*4 */, const uint64_t* table /* u*4 */,
                                  size_t usable_rows, uint64_t* permuted_input, uint64_t* permuted_table);
int h2agg_permute_expression_pair_dev(h2agg_ctx* ctx, const void* d_input, const void* d_table, size_t usable_rows,
                                      void* d_permuted_input, void* d_permuted_table);
/* N3, the row-wise steps around the sort and the running products (halo2_proofs plonk/lookup/prover.rs and
 * plonk/permutation/prover.rs; the recurrences are the ones the reference's verifier checks,
 * halo2-snark-aggregator-api/src/systems/halo2/lookup.rs:58-119, permutation.rs:54-136).  Device-resident, asynchronous.
 *   compress_expressions: exprs = { n_exprs, POLY x n_exprs } (POLY as in the quotient plan below) over LAGRANGE columns of
 *       2^k rows; a rotation r reads row (i + r) mod 2^k; out[i] = fold(acc * theta + expr_j(i)).
 *   lookup_product: z[0] = 1, z[i+1] = z[i] (A[i] + beta)(S[i] + gamma) / ((A'[i] + beta)(S'[i] + gamma)), n rows.
 *   permutation_product (one column set of <= 16 columns): z[0] = *d_last_z (1 if NULL),
 *       z[i+1] = z[i] prod_j (v_j[i] + beta delta^(first + j) omega^i + gamma) / (v_j[i] + beta sigma_j[i] + gamma);
 *       beta_delta_start = beta * delta^first, `first` = index of the set's first column in the permutation.
 * All n = 2^k rows are produced; the caller overwrites the last blinding_factors rows with its random values. */
int h2agg_compress_expressions_dev(h2agg_ctx* ctx, const uint32_t* exprs, size_t n_words, const void* const* d_columns,
                                   size_t n_columns, const uint64_t* consts, size_t n_consts, uint32_t k,
                                   const uint64_t theta[4], void* d_out);
int h2agg_lookup_product_dev(h2agg_ctx* ctx, const void* d_input, const void* d_table, const void* d_permuted_input,
                             const void* d_permuted_table, size_t n, const uint64_t beta[4], const uint64_t gamma[4],
                             void* d_z);
/* lookup_product for all lookups of a constraint system in one call (lookup i on lane i mod 8: the latency chains of
 * the grand products overlap); same results as n_lookups calls of h2agg_lookup_product_dev. */
int h2agg_lookup_products_dev(h2agg_ctx* ctx, size_t n_lookups, const void* const* d_inputs, const void* const* d_tables,
                              const void* const* d_permuted_inputs, const void* const* d_permuted_tables, size_t n,
                              const uint64_t beta[4], const uint64_t gamma[4], void* const* d_z);
int h2agg_permutation_product_dev(h2agg_ctx* ctx, const void* const* d_values, const void* const* d_sigmas, size_t n_cols,
                                  uint32_t k, const uint64_t omega[4], const uint64_t beta_delta_start[4],
                                  const uint64_t delta[4], const uint64_t beta[4], const uint64_t gamma[4],
                                  const void* d_last_z, void* d_z);
/* a <- a sorted ascending by Fr::cmp (Montgomery in and out), in place: the sort inside the above, exposed. n <= 2^25. */
int h2agg_sort_fr(h2agg_ctx* ctx, uint64_t* a /* n*4 */, size_t n);
int h2agg_sort_fr_dev(h2agg_ctx* ctx, void* d_a, size_t n);

/* ---- N1 (next row, rank 1): quotient numerator on the extended coset -----------------------------
 * Replaces halo2_proofs plonk/evaluation.rs Evaluator::evaluate_h (+ EvaluationDomain::divide_by_vanishing_poly
 * when t_evaluations is given), reached from create_proof (verify_circuit.rs:986).  The reference's verifier holds
 * the same equations and fixes their order: gates, permutation, lookups
 * (halo2-snark-aggregator-api/src/systems/halo2/params.rs:95-150, permutation.rs:54-136, lookup.rs:58-119),
 * folded as h = h * y + term and divided by X^n - 1 (vanish.rs:28-29).
 *
 * d_columns: device pointers to extended-coset evaluation vectors (2^ext_k x 32 B each) of every polynomial the
 * constraint system touches -- fixed, advice, instance, permutation sigma, l_0, l_last, l_active_row, permutation
 * products z, and per lookup the product z and the permuted input / table columns -- in any order; the plan
 * refers to them by index.  Row idx stands for X = zeta * omega_ext^idx; rotation r reads row
 * (idx + r * 2^(ext_k - k)) mod 2^ext_k.
 *
 * plan (uint32 words):
 *   [0] 0x31485148  [1] number of gate polynomials  [2] number of permutation columns (0 = no argument)
 *   [3] permutation chunk_len (cs.degree() - 2)  [4] (int32) rotation of "last" = -(blinding_factors + 1)
 *   [5] number of lookups  [6] [7] [8] column index of l_0, l_last, l_active_row
 *   gate polynomials, in order, each a POLY
 *   if [2] > 0: [2] x { value column, sigma column }, then ceil([2] / [3]) x { z column }
 *   per lookup: n_input_exprs, POLY x n;  n_table_exprs, POLY x n;  z column, permuted input column, permuted table column
 *   POLY = n_terms, then per term { constant index into consts[] or 0xffffffff (coefficient 1), n_factors,
 *          n_factors x (column | (uint16)(int16 rotation) << 16) }   -- a sum of products of column queries
 * An invalid plan (index out of range, truncated section, trailing words) is rejected with status 1. */
typedef struct h2agg_quotient_args {
  uint32_t k, ext_k;
  const uint32_t* plan;
  size_t n_plan_words;
  const void* const* d_columns;
  size_t n_columns;
  const uint64_t* consts; /* n_consts * 4, Montgomery */
  size_t n_consts;
  const uint64_t* y;     /* challenges, 4 limbs each */
  const uint64_t* beta;
  const uint64_t* gamma;
  const uint64_t* theta;
  const uint64_t* omega_ext; /* generator of the 2^ext_k domain */
  const uint64_t* zeta;      /* coset shift g_coset (Fr::ZETA) */
  const uint64_t* delta;     /* Fr::DELTA: the permutation argument's column separator */
  const uint64_t* t_evaluations; /* t_len x 4: 1 / ((zeta omega_ext^i)^n - 1), or NULL to skip the division */
  size_t t_len;                  /* power of two (halo2: 2^(ext_k - k)) */
} h2agg_quotient_args;
int h2agg_evaluate_h_dev(h2agg_ctx* ctx, const h2agg_quotient_args* args, void* d_out /* 2^ext_k * 32 B */);
/* Row window of the same computation (multi-GPU row sharding of the quotient): rows
 * [row_begin, row_begin + row_count) of the coset go to d_out[0 .. row_count).  col_row0[c] (NULL = all zero) is the
 * global row held by element 0 of column c's buffer, cyclically: a rank only needs each column on its window plus the
 * rotation halo (rot * 2^(ext_k - k) rows either side), not the whole 2^ext_k vector. */
int h2agg_evaluate_h_rows_dev(h2agg_ctx* ctx, const h2agg_quotient_args* args, uint64_t row_begin, uint64_t row_count,
                              const uint64_t* col_row0 /* n_columns or NULL */, void* d_out /* row_count * 32 B */);
/* GWC multi-opening (halo2_proofs poly/kzg/multiopen/gwc/prover.rs: poly_batch = poly_batch * v + poly):
 * out[j] = sum_i polys[i][j] * v^(n_polys-1-i).  d_out must not alias an input. */
int h2agg_poly_fold_dev(h2agg_ctx* ctx, const void* const* d_polys, size_t n_polys, size_t n, const uint64_t v[4],
                        void* d_out);
/* out[j] = sum_i weights[i] * polys[i][j] (weights: n_polys x 4 Montgomery limbs; n_polys = 0 zeroes out).  One rank's
 * share of a fold -- its own polynomials with the powers of v they carry in the full query list -- and, with unit
 * weights, the sum of the gathered shares.  d_out must not alias an input. */
int h2agg_poly_lincomb_dev(h2agg_ctx* ctx, const void* const* d_polys, const uint64_t* weights, size_t n_polys, size_t n,
                           void* d_out);

/* ---- W1-W5: witness synthesis of halo2-ecc-circuit-lib ---------------------------
 * A recording implementation of the reference's chip surface -- ArithFieldChip (ScalarChip), Encode, and ArithEccChip::{add, sub, scalar_mul,
 * scalar_mul_constant, multi_exp, assign_var, assign_const, normalize}
 * (halo2-snark-aggregator-api/src/arith/ecc.rs:5-61, common.rs:3-42) as bound to
 * EccChipOps::{add, sub, mul, constant_mul, shamir, assign_point, assign_constant_point, reduce}
 * by halo2-snark-aggregator-circuit/src/chips/ecc_chip.rs:28-133.  The host walks the sequential op
 * chain and reproduces the exact row layout; the kernel expands every integer op into its advice
 * rows.  Handles are indices; negative return = error (h2agg_wit_error).  Points are affine
 * Montgomery Fq pairs, (0,0) = identity; scalars are CANONICAL 256-bit integers < r. */
typedef struct h2agg_witness h2agg_witness;
h2agg_witness* h2agg_wit_new(void);
/* Host threads a multi_exp records its independent sections on (candidate tables, inner window sums); default
 * min(16, cores) or $H2AGG_WIT_THREADS.  n < 1 only queries.  Returns the previous value.  The recorded layout and
 * values do not depend on it. */
int h2agg_wit_set_threads(int n);
void h2agg_wit_free(h2agg_witness* w);
const char* h2agg_wit_error(h2agg_witness* w);
uint64_t h2agg_wit_rows(h2agg_witness* w); /* current row offset = rows the layout occupies */
uint64_t h2agg_wit_ops(h2agg_witness* w);  /* op records queued for the kernel */
int64_t h2agg_wit_assign_point(h2agg_witness* w, const uint64_t xy[8]);          /* assign_point (on-curve check rows) */
int64_t h2agg_wit_assign_constant_point(h2agg_witness* w, const uint64_t xy[8]); /* assign_constant_point */
int64_t h2agg_wit_assign_scalar(h2agg_witness* w, const uint64_t s_canonical[4]);
int64_t h2agg_wit_ecc_add(h2agg_witness* w, int64_t a, int64_t b);
int64_t h2agg_wit_ecc_sub(h2agg_witness* w, int64_t a, int64_t b);
int64_t h2agg_wit_ecc_double(h2agg_witness* w, int64_t a);
int64_t h2agg_wit_ecc_reduce(h2agg_witness* w, int64_t a);                        /* normalize */
int64_t h2agg_wit_ecc_mul(h2agg_witness* w, int64_t a, int64_t s);                /* scalar_mul */
int64_t h2agg_wit_ecc_shamir(h2agg_witness* w, const int64_t* pts, const int64_t* scalars, size_t n); /* multi_exp */
int64_t h2agg_wit_ecc_constant_mul(h2agg_witness* w, const uint64_t base_xy[8], int64_t s); /* scalar_mul_constant */
int h2agg_wit_point_value(h2agg_witness* w, int64_t h, uint64_t out_xy[8], int* is_identity); /* to_value */
/* The trait-level point operations above take their operands as the reference's adapter does
 * (halo2-snark-aggregator-circuit/src/chips/ecc_chip.rs:34-52, 99-131: add on a.clone() / b.clone(), sub on a.clone(),
 * scalar_mul on rhs.clone(), normalize on v.clone(), multi_exp on a moved Vec): the curvature / native caches an operation
 * fills in die with the clone, a handle keeps the caches it was created with -- that is part of the row layout.
 *
 * ArithFieldChip: ScalarChip over FiveColumnBaseGate (halo2-snark-aggregator-circuit/src/chips/scalar_chip.rs:17-127;
 * traits halo2-snark-aggregator-api/src/arith/field.rs:6-105, common.rs:3-42).  Handles are AssignedValue<Fr>; values
 * cross the ABI as CANONICAL 256-bit integers < r (status -1 otherwise).  The provided trait methods (sum_with_constant,
 * mul_add, mul_add_accumulate, pow_constant, field.rs:37-104) are compositions of these and stay with the caller.
 *   assign_var   = h2agg_wit_assign_scalar  (BaseGateOps::assign)           assign_const / _zero / _one = _field_assign_const
 *   add, sub     = sum_with_constant([(a,1),(b,+-1)], 0)                    mul, square = [a, b, ab]
 *   div          = div_unsafe: [b, a/b, a]; b = 0 is an error (Rust unwraps) mul_add_constant = [a, b, ab + c] */
int64_t h2agg_wit_field_assign_const(h2agg_witness* w, const uint64_t c_canonical[4]);
int64_t h2agg_wit_field_add(h2agg_witness* w, int64_t a, int64_t b);
int64_t h2agg_wit_field_sub(h2agg_witness* w, int64_t a, int64_t b);
int64_t h2agg_wit_field_mul(h2agg_witness* w, int64_t a, int64_t b);
int64_t h2agg_wit_field_square(h2agg_witness* w, int64_t a);
int64_t h2agg_wit_field_div(h2agg_witness* w, int64_t a, int64_t b);
int64_t h2agg_wit_field_sum_with_coeff_and_constant(h2agg_witness* w, const int64_t* elems, const uint64_t* coeffs_canonical /* n*4 */,
                                                    size_t n, const uint64_t constant_canonical[4]);
int64_t h2agg_wit_field_mul_add_constant(h2agg_witness* w, int64_t a, int64_t b, const uint64_t c_canonical[4]);
int h2agg_wit_scalar_value(h2agg_witness* w, int64_t h, uint64_t out_canonical[4]);           /* to_value */
/* AssignedValue.cell: (advice column, row) -- what constrain_instance (verify_circuit.rs:357-367) and copy constraints bind */
int h2agg_wit_scalar_cell(h2agg_witness* w, int64_t h, uint32_t* column, uint32_t* row);
/* Encode: PoseidonEncodeChip::encode_point (halo2-snark-aggregator-circuit/src/chips/encode_chip.rs:18-33) = the natives
 * of x and y, taken on clones of the coordinates; encode_scalar / decode_scalar are the identity (:35-51). */
int h2agg_wit_encode_point(h2agg_witness* w, int64_t point, int64_t out_natives[2]);
/* What Halo2VerifierCircuits::synthesize does around the chip calls:
 *   assign_identity          ArithCommonChip::assign_zero of the EccChip (ecc_chip.rs:54-56)
 *   ecc_assert_equal         the `coherent` commitment pairs (verify_circuit.rs:487-493)
 *   assert_not_identity      base_gate.assert_false(&p.z) (:495-496)
 *   expose_final_pair        second region (:264-344): reduce the coordinates of (w_x, w_g), take the parity bits of the
 *                            y's, pack each point into two 136-bit halves -> 4 cells for constrain_instance rows 0..3 */
int64_t h2agg_wit_ecc_assign_identity(h2agg_witness* w);
int h2agg_wit_ecc_assert_equal(h2agg_witness* w, int64_t a, int64_t b);
int h2agg_wit_assert_not_identity(h2agg_witness* w, int64_t point);
int h2agg_wit_expose_final_pair(h2agg_witness* w, int64_t w_x, int64_t w_g, int64_t out_cells[4]);
/* Expand everything recorded into the 5 advice columns (n_rows Fr each, Montgomery; rows past the
 * recorded offset are zero like unassigned halo2 cells).  Host pointers / device pointers. */
int h2agg_witness_expand(h2agg_ctx* ctx, h2agg_witness* w, uint64_t* const advice_cols[5], size_t n_rows);
int h2agg_witness_expand_dev(h2agg_ctx* ctx, h2agg_witness* w, void* const d_cols[5], size_t n_rows);

/* ---- N4 (data formats either side of the path): PrimeField::to_repr / from_repr in bulk -----------------------
 * The reference's stage files hold scalars as 32-byte little-endian canonical integers
 * (halo2-snark-aggregator-circuit/src/fs.rs:134-146 load_instances, :169-180 write_verify_circuit_instance,
 * :182-197 write_verify_circuit_final_pair).  from_repr = 0: Montgomery limbs -> repr bytes; 1: repr bytes -> Montgomery
 * limbs, status 4 if a value is not < r (from_repr(..).unwrap() would panic).  Host pointers, n x 32 B each way. */
int h2agg_fr_repr(h2agg_ctx* ctx, int from_repr, const void* in, void* out, size_t n);

/* N4: the G1 point codec of the stage files.  ParamsKZG::read / write (halo2_proofs poly/kzg/commitment.rs) hold g and
 * g_lagrange as COMPRESSED points -- halo2curves 0.2.1 G1Affine::to_bytes / from_bytes: 32-byte little-endian x, parity of
 * y in bit 7 of byte 31, identity = 32 zero bytes.  The reference decodes the 2 x 2^k points of verify_circuit.params at the
 * start of every verify_run (halo2-snark-aggregator-circuit/src/fs.rs:109-115) and of HALO2_PARAMS_k in get_params_cached
 * (verify_circuit.rs:701-731); the inner proofs' transcript points use the same encoding
 * (halo2-snark-aggregator-api/src/systems/halo2/transcript.rs:63-65).  decompress: one Fq square root per point; status 4
 * when an encoding is not a curve point or x >= p (from_bytes(..) is None; the reference unwraps). */
int h2agg_g1_decompress(h2agg_ctx* ctx, const uint8_t* in /* n*32 */, uint64_t* out_affine /* n*8, Montgomery */, size_t n);
int h2agg_g1_decompress_dev(h2agg_ctx* ctx, const void* d_in /* n*32 B */, void* d_out_affine /* n*64 B */, size_t n);
int h2agg_g1_compress(h2agg_ctx* ctx, const uint64_t* affine /* n*8 */, uint8_t* out /* n*32 */, size_t n);

/* ---- small helpers used by tests and the host layer (run on the device) ---------------------- */
/* out[i] = a[i] * b[i] in Fr (field = 0) or Fq (field = 1); host pointers; Montgomery form. */
int h2agg_field_mul(h2agg_ctx* ctx, int field, const uint64_t* a, const uint64_t* b, uint64_t* out, size_t n);
/* out[i] = a[i] + b[i] (op 0), a[i] - b[i] (op 1), a[i]^-1 (op 2, b ignored) */
int h2agg_field_op(h2agg_ctx* ctx, int field, int op, const uint64_t* a, const uint64_t* b, uint64_t* out, size_t n);

/* ---- device memory + synthetic inputs (bench / test plumbing; not part of the reference surface) ---
 * Deterministic, counter-based inputs of SURVEY.md 8d generated straight into HBM; bit-identical to
 * oracle_gen_scalars / oracle_gen_bases.  kind: 0 uniform Fr, 1 witness-like a0..a3
 * (70% 17-bit, 10% {0,1}, 20% zero), 2 witness-like a4 (50% 68-bit, 20% full, 30% zero), 3 uniform 17-bit. */
int h2agg_dev_alloc(h2agg_ctx* ctx, size_t bytes, void** out);
int h2agg_dev_free(h2agg_ctx* ctx, void* p);
int h2agg_memcpy_h2d(h2agg_ctx* ctx, void* d_dst, const void* src, size_t bytes);
int h2agg_memcpy_d2h(h2agg_ctx* ctx, void* dst, const void* d_src, size_t bytes);
int h2agg_synth_scalars_dev(h2agg_ctx* ctx, uint64_t seed, int kind, uint64_t first, uint64_t n, void* d_out);
int h2agg_synth_bases_dev(h2agg_ctx* ctx, uint64_t seed, uint64_t first, uint64_t n, void* d_out);

#ifdef __cplusplus
}
#endif
#endif /* H2AGG_H */
