/*
 * qldpc_b200.h -- C ABI of the B200-native BP+OSD decoder (libqldpc_b200.so).
 *
 * This is the drop-in boundary for the decode hot path of michelebanfi/qLDPC.  The reference
 * has no FFI layer -- its boundary is a set of Python call signatures --
 * so each entry point below names the reference function(s) it stands in for; the Python
 * package qldpc_b200 re-exports the reference's own names on top of these calls (see
 * INTEGRATION.md for the ctypes binding a maintainer of the reference would add).
 *
 * Conventions
 *   - plain pointers and sizes only; every function returns 0 on success, non-zero on error
 *     (qldpc_last_error() gives the message).  There is NO CPU fallback: without a CUDA device
 *     every compute call fails with QLDPC_ERR_CUDA.
 *   - "_host" functions take HOST pointers in the reference's dtypes (uint8/int8 0/1 arrays,
 *     float64 LLRs) and do the host<->device copies themselves (synchronous on return).
 *   - "_dev" functions take DEVICE pointers, bit-packed rows of 32-bit little-endian words
 *     (bit j of a row = bit (j & 31) of word (j >> 5)), and a cudaStream_t passed as void*; they
 *     only enqueue work.
 *   - shots are the leading axis of every batched array.
 *   - a qldpc_code owns device workspaces and streams: use one handle per host thread / per stream at a time
 *     (handles are cheap; the Tanner-graph tables are a few tens of KB).
 */
#ifndef QLDPC_B200_H
#define QLDPC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QLDPC_OK 0
#define QLDPC_ERR_ARG 1
#define QLDPC_ERR_CUDA 2
#define QLDPC_ERR_UNSUPPORTED 3

/* BP variants */
#define QLDPC_MIN_SUM 0          /* rework/decoding.py:5    performMinSum_Symmetric                       */
#define QLDPC_SUM_PRODUCT 1      /* decoding/beliefPropagation.py:88 performBeliefPropagationFast (and :6) */
#define QLDPC_SUM_PRODUCT_SYM 2  /* rework/decoding.py:131  performBeliefPropagation_Symmetric            */

/* posterior-LLR output modes */
#define QLDPC_LLR_NONE 0
#define QLDPC_LLR_FAILED 1       /* only for shots whose BP did not converge (what OSD consumes) */
#define QLDPC_LLR_ALL 2

#define QLDPC_NUM_COUNTERS 16
/* indices into the counter vector of qldpc_check_* / qldpc_mc_sweep (see misc_kernels.cuh) */
#define QLDPC_CNT_SHOTS 0
#define QLDPC_CNT_BP_FAILED 1
#define QLDPC_CNT_LOGICAL 2
#define QLDPC_CNT_LOGICAL_OSD 3
#define QLDPC_CNT_DEGENERATE 4
#define QLDPC_CNT_MISCORRECTED 5
#define QLDPC_CNT_INCORRECTABLE 6
#define QLDPC_CNT_INVALID 7
#define QLDPC_CNT_ITER_SUM 8
#define QLDPC_CNT_LOGICAL_BP 9
#define QLDPC_CNT_RESID_WEIGHT 10
#define QLDPC_CNT_ERR_WEIGHT 11

typedef struct qldpc_code qldpc_code; /* opaque: device-resident Tanner graph + workspaces */

typedef struct {
    int32_t variant;    /* QLDPC_MIN_SUM | QLDPC_SUM_PRODUCT | QLDPC_SUM_PRODUCT_SYM                 */
    int32_t precision;  /* 32: float32 messages (production) | 64: float64, the reference's arithmetic */
    int32_t max_iter;   /* maxIter of the reference                                                  */
    int32_t staged;     /* kernel choice.  0: auto -- float32 min-sum on a uniform-row-weight H runs the
                           T-lanes-per-shot shared-memory kernel, anything else the thread-per-shot kernel,
                           in shared memory when the per-shot state fits, else HBM-staged;
                           1: force the HBM-staged kernel; 2: force the thread-per-shot kernel;
                           3: force the warp-per-shot kernel (float32 min-sum on BB-shaped H);
                           4: force the CTA-per-shot kernel (float32 min-sum, space-time shaped H)         */
    int32_t lanes_per_shot; /* 0: auto | 4 | 8 (T-lanes-per-shot kernel) | 32 (warp-per-shot kernel)        */
    int32_t refill_min; /* 0: auto.  Idle shots per warp that trigger a refill from the shot cursor  */
    double alpha;       /* min-sum normalisation / sum-product scaling                               */
    double damping;     /* damping on Q                                                              */
    double clip;        /* clip_llr                                                                  */
} qldpc_bp_config;

const char *qldpc_last_error(void);
int qldpc_version(void);
int qldpc_device_count(void);

/* Tanner graph from host arrays.  CSR of H (row_ptr[m+1], col_idx[E] ascending inside a row),
 * var_ptr[n+1], and for every variable the ids of its edges (CSR numbering) in the order their
 * messages are added into the posterior, for iteration 0 (var_edge0) and iterations >= 1
 * (var_edge1): this reproduces NumPy's summation order of `np.sum(R, axis=0)`
 * (decoding/beliefPropagation.py:129, rework/decoding.py:61) -- SURVEY.md H2.
 * L: k x n dense 0/1 logical operators (codes/<name>.npz key Lx) or NULL with k = 0.
 * Replaces the per-call preprocessing of the reference (csr_matrix(H), mask = H != 0, ...:
 * beliefPropagation.py:93-104, decoding.py:13-19). */
int qldpc_code_create(int32_t m, int32_t n, const int32_t *row_ptr, const int32_t *col_idx,
                      const int32_t *var_ptr, const int32_t *var_edge0, const int32_t *var_edge1,
                      int32_t k, const uint8_t *L, qldpc_code **out);
void qldpc_code_destroy(qldpc_code *code);
/* geometry the launcher picked for a config: shots resident per SM, shared-memory bytes per CTA,
 * 1 if the HBM-staged kernel is used */
int qldpc_bp_geometry(qldpc_code *code, const qldpc_bp_config *cfg, int32_t *shots_per_cta,
                      int32_t *smem_bytes, int32_t *staged);

/* Diagnostic: modelled shared-memory wavefronts per shot-iteration of the variable pass of the T-lanes-per-shot
 * kernel (message rows x2 + summary rows), with variables in natural order (`before`) and in the bank-conflict
 * optimised order the float32 kernel uses (`after`). */
int qldpc_tiled_conflict_model(qldpc_code *code, int32_t lanes_per_shot, double *before, double *after);

/* Warp-per-shot kernels (BB-shaped H): which lane owns which check / variable and in which register slot an edge sits
 * is free, and decides the bank conflicts of the kernel's shared-memory scatter and gather.  qldpc_code_create builds a
 * conflict-free labelling (balanced check slots, degree-bounded placement of the variables, Koenig edge colouring; < 1 ms
 * for the reference's codes).  This call reports it and lets tests switch: `steps` < 0 installs the natural labelling
 * (check c at lane c mod 32, ...), `steps` > 0 rebuilds the constructed one with that search budget, 0 only reports.
 * Results of the kernels do not depend on the labelling (bit-identical).  Synchronises the device.
 *   cost[3] <- modelled scatter + gather wavefronts per shot-iteration: natural labelling, current labelling, floor.
 * QLDPC_ERR_UNSUPPORTED when the kernel does not apply to this code. */
int qldpc_warp_layout_tune(qldpc_code *code, int64_t steps, int32_t *cost);

/* ---------------- host-pointer API (reference dtypes) ---------------- */

/* Batched BP.  Stands in for performMinSum_Symmetric / performBeliefPropagationFast /
 * performBeliefPropagation / performBeliefPropagation_Symmetric (B = 1) and
 * performBeliefPropagationBatch (beliefPropagationGPU.py:81).
 *   synd [B][m] uint8, prior [n] float64 (initialBelief) ->
 *   hard [B][n] int8, conv [B] uint8, iters [B] int32 (0-based exit iteration; may be NULL),
 *   llr [B][n] float64 posterior `values` (may be NULL). */
int qldpc_bp_decode_host(qldpc_code *code, const qldpc_bp_config *cfg, const double *prior, int64_t B,
                         const uint8_t *synd, int8_t *hard, uint8_t *conv, int32_t *iters, double *llr);

/* The alpha_estimation=True return of performMinSum_Symmetric (rework/decoding.py:58-59: R_new / alpha after the
 * first check pass, dump_iter = 0) and of performBeliefPropagation_Symmetric (:168-169: R at currentIter == 10):
 * check-to-variable messages of iteration dump_iter, r_edges [B][E] float64 in CSR edge order.  float64 only. */
int qldpc_bp_messages_host(qldpc_code *code, const qldpc_bp_config *cfg, const double *prior, int64_t B,
                           const uint8_t *synd, int32_t dump_iter, double *r_edges);

/* Batched OSD.  Stands in for performOSD (decoding/OSD.py:3) when order == 0 and
 * performOSD_enhanced (decoding/OSD_enhanced.py:5) otherwise; max_combinations <= 0 means None.
 *   synd [B][m] uint8, llr [B][n] float64, hard [B][n] uint8/int8 -> out [B][n] uint8.
 * Column order: stable ascending |llr| (ties -> lower index).
 * order > 0: the combination sweep (OSD_enhanced.py:66-131) runs on the shots whose OSD-0 solution misses the syndrome,
 * selected on the device; order <= 16.  Check matrices with more than 160 rows: order > 0 is accepted when H has full row
 * rank (every syndrome is then consistent and the sweep is dead code, OSD_enhanced.py:58-60) and refused with
 * QLDPC_ERR_UNSUPPORTED otherwise. */
int qldpc_osd_decode_host(qldpc_code *code, int64_t B, const uint8_t *synd, const double *llr,
                          const uint8_t *hard, int32_t order, int64_t max_combinations, uint8_t *out);

/* Fused BP -> OSD on the BP failures: the per-shot body of the reference's Monte-Carlo loops
 * (paperResults.py:71-77, paperResults_GPU.py:108-123, rework/main.py:80-88).
 * osd_order < 0: BP only; 0: performOSD; > 0: performOSD_enhanced(order = osd_order), i.e. OSD-0 plus the combination
 * sweep on the shots whose OSD-0 solution misses the syndrome (possible only for syndromes outside the column space of H,
 * e.g. with measurement errors); same limits as qldpc_osd_decode_host.
 * synd [B][m] uint8 -> corr [B][n] uint8, conv [B] uint8 (BP flag), iters [B] int32 (may be NULL).
 * This is the call bench.py times end to end. */
int qldpc_bposd_decode_host(qldpc_code *code, const qldpc_bp_config *cfg, const double *prior, int64_t B,
                            const uint8_t *synd, int32_t osd_order, uint8_t *corr, uint8_t *conv,
                            int32_t *iters);

/* Who packs the uint8 rows of qldpc_bposd_decode_host, and what the calls have moved so far.  The byte rows either cross
 * PCIe as they are and are packed / expanded by kernels (host_pack 0: 72 + 149 bytes per [[144,12,12]] shot), or host threads
 * pack / expand them between the caller's arrays and pinned staging buffers and bit-packed rows cross the bus (host_pack 1:
 * 12 + 25 bytes), or each chunk of the pipeline takes whichever side is free (host_pack 2, the default when the host has
 * the threads: a chunk goes to the host threads whenever the byte-row copies already queued keep the bus busy for as long
 * as packing it takes; the split follows the measured speeds of both sides, bus contention between the ranks of one box
 * included).  Decided at the first call per code handle from the measured throughput of the thread pool (host_pack_rate,
 * shots/s; host_pack is -1 before that call); environment: QLDPC_HOST_PACK = 0 | 1 | 2 forces a mode, QLDPC_HOST_THREADS sets
 * the pool size (default: hardware threads / visible GPUs, at most 16).  h2d_bytes / d2h_bytes: cumulative bytes of the
 * host<->device copies issued by qldpc_bposd_decode_host[_packed] on this handle; chunks_host / chunks_device: how many
 * chunks went to either side.  Results never depend on the mode (tested).  Any pointer may be NULL. */
int qldpc_host_transfer_stats(qldpc_code *code, uint64_t *h2d_bytes, uint64_t *d2h_bytes, int32_t *host_pack,
                              double *host_pack_rate, uint64_t *chunks_host, uint64_t *chunks_device);
/* mode 0 / 1 / 2: fix the mode for this handle; -1: measure again at the next call */
int qldpc_set_host_pack(qldpc_code *code, int32_t mode);

/* Same call with bit-packed host rows (synd [B][words_m], corr [B][words_n] uint32): 37 instead of 221 bytes per
 * [[144,12,12]] shot cross PCIe.  Not a reference dtype; for callers that keep syndromes packed. */
int qldpc_bposd_decode_host_packed(qldpc_code *code, const qldpc_bp_config *cfg, const double *prior, int64_t B,
                                   const uint32_t *synd, int32_t osd_order, uint32_t *corr, uint8_t *conv,
                                   int32_t *iters);

/* Per-shot checks of the drivers (paperResults.py:83-100, rework/main.py:90-112):
 * err/corr [B][n] uint8, synd [B][m] uint8 -> flags [B] uint8 (bit0 logical, bit1 valid,
 * bit2 degenerate), weight [B] int32 (residual weight), counters[QLDPC_NUM_COUNTERS] (any may be NULL). */
int qldpc_check_host(qldpc_code *code, int64_t B, const uint8_t *err, const uint8_t *corr,
                     const uint8_t *synd, const uint8_t *conv, const int32_t *iters,
                     int32_t distance, uint8_t *flags, int32_t *weight, uint64_t *counters);

/* syndromes of given errors, synd = err * H^T mod 2 (paperResults.py:65, beliefPropagationGPU.py:198):
 * err [B][n] uint8 -> synd [B][m] uint8 */
int qldpc_syndrome_host(qldpc_code *code, int64_t B, const uint8_t *err, uint8_t *synd);

/* generate_errors_and_syndromes_batch (beliefPropagationGPU.py:181) on the device: Philox4x32-10
 * keyed by (seed, first_shot + i).  err [B][n] uint8, synd [B][m] uint8.  draws = 2 XORs two
 * independent draws (paperResults.py:61-63). */
int qldpc_sample_host(qldpc_code *code, double p, uint64_t seed, uint64_t first_shot, int32_t draws,
                      int64_t B, uint8_t *err, uint8_t *synd);
/* the same with measurement errors: every syndrome bit is then flipped with probability q_meas, the
 * "simple phenomenological error model" of paperResults.py:66-68 (Philox stream "MEAS" of the same
 * (seed, shot id) counter space).  Such syndromes need not lie in the column space of H. */
int qldpc_sample_noisy_host(qldpc_code *code, double p, double q_meas, uint64_t seed, uint64_t first_shot,
                            int32_t draws, int64_t B, uint8_t *err, uint8_t *synd);

/* Whole Monte-Carlo point on the device: sample -> BP -> OSD on failures -> checks -> counters.
 * Shots [first_shot, first_shot + nshots) of the stream `seed`; counters are ADDED into
 * counters[QLDPC_NUM_COUNTERS] (host).  Results depend only on (seed, shot id): sharding the
 * range over ranks and summing the counters gives identical totals. */
int qldpc_mc_sweep(qldpc_code *code, const qldpc_bp_config *cfg, const double *prior, double p,
                   uint64_t seed, uint64_t first_shot, int64_t nshots, int32_t draws, int32_t osd_order,
                   int32_t distance, uint64_t *counters);
/* the same with measurement errors of rate q_meas on the syndrome (see qldpc_sample_noisy_host) */
int qldpc_mc_sweep_noisy(qldpc_code *code, const qldpc_bp_config *cfg, const double *prior, double p,
                         double q_meas, uint64_t seed, uint64_t first_shot, int64_t nshots, int32_t draws,
                         int32_t osd_order, int32_t distance, uint64_t *counters);

/* The statistics behind estimate_alpha_from_code (rework/Alvarado.py:10-66) without materialising messages: with a uniform
 * prior L the message R_new / alpha of the first check pass (rework/decoding.py:58-59) is (1 - 2 s_c) L on every edge of
 * check c, so its two histograms (true bit 0 / 1) reduce to counts[2 * bit + s], ADDED into counts[4].  err [B][n] uint8
 * (the caller's errors, e.g. NumPy's stream as in the reference), or NULL: B device-sampled shots (p, seed, first_shot). */
int qldpc_alpha_counts(qldpc_code *code, int64_t B, const uint8_t *err, double p, uint64_t seed, uint64_t first_shot,
                       uint64_t *counts);

/* Posterior-LLR histograms without returning B*n floats (BP_per_Iteration.py:56,60 and rework/Alvarado.py:159-162 collect
 * every posterior LLR on the host): device-sampled shots as in qldpc_mc_sweep, BP only, hist [3][nbins] uint64 ADDED into:
 * row 0 = LLRs of variables whose true error bit is 0, row 1 = true bit 1, row 2 = all LLRs of BP-failed shots.  Uniform
 * bins over [lo, hi), out-of-range values in the end bins.  bp_failed (may be NULL) is incremented by the BP failures. */
int qldpc_bp_llr_histogram(qldpc_code *code, const qldpc_bp_config *cfg, const double *prior, double p, uint64_t seed,
                           uint64_t first_shot, int64_t nshots, int32_t draws, double lo, double hi, int32_t nbins,
                           uint64_t *hist, uint64_t *bp_failed);

/* ---------------- device-pointer API (bit-packed, asynchronous) ----------------
 * Calls enqueue work on `stream` and return.  Concurrency contract of a handle: ONE host thread, ONE stream at a time -- the
 * control block (shot cursor, failure counter), the device copies of the prior and the internal workspaces (LLR hand-off,
 * failure / invalid lists, staging arrays) are per handle, so two calls in flight on different streams race on them; use one
 * handle per stream (handles are cheap: tables < 100 KB).  Implicit synchronisation: a call with a prior that differs from
 * the previous call's copies it and synchronises `stream`; a call that needs a larger workspace than any call before it
 * frees and allocates device memory (device-wide synchronisation).  After one call with the largest batch and the final
 * prior the calls are pure kernel launches plus 8-byte memsets, and can be stream-captured. */

/* words per packed syndrome / error row */
int qldpc_words_m(const qldpc_code *code);
int qldpc_words_n(const qldpc_code *code);

/* BP on packed device syndromes.  llr is float32 or float64 per cfg->precision.
 * fail_idx/fail_count (may be NULL): compacted ids of the BP failures, for qldpc_osd_decode_dev.
 * iter_total (may be NULL): device uint64 accumulating EXECUTED iterations (roofline accounting): at low error rates the float32 warp
 * kernel retires all-zero syndromes without running the iteration the reference spends on them (exit iteration 0 either way). */
int qldpc_bp_decode_dev(qldpc_code *code, const qldpc_bp_config *cfg, const double *prior_host, int64_t B,
                        const uint32_t *synd, uint32_t *hard, uint8_t *conv, int32_t *iters,
                        void *llr, int32_t llr_mode, int32_t *fail_idx, uint32_t *fail_count,
                        uint64_t *iter_total, void *stream);

/* OSD-0 on the shots listed in idx[0 .. *count_dev) (count taken from the device, no host sync),
 * or on all of [0, count_host) when idx == NULL.  out may alias hard.  llr_f64: 1 if llr is double. */
int qldpc_osd_decode_dev(qldpc_code *code, const int32_t *idx, const uint32_t *count_dev, int64_t count_host,
                         const uint32_t *synd, const void *llr, int32_t llr_f64, const uint32_t *hard,
                         uint32_t *out, uint8_t *valid, void *stream);

/* The OSD-w sweep of performOSD_enhanced (OSD_enhanced.py:66-131) behind a qldpc_osd_decode_dev call that wrote `valid`:
 * the listed shots (as there; with a device-side count, count_host must bound it) whose OSD-0 solution misses the syndrome
 * are compacted on the device and swept; out is updated in place.  hard == NULL: the BP hard decision is llr < 0 (use it
 * when out aliased hard in the OSD-0 call).  order <= 0: no-op.  Limits: see qldpc_osd_decode_host. */
int qldpc_osdw_decode_dev(qldpc_code *code, const int32_t *idx, const uint32_t *count_dev, int64_t count_host,
                          const uint32_t *synd, const void *llr, int32_t llr_f64, const uint32_t *hard,
                          uint32_t *out, const uint8_t *valid, int32_t order, int64_t max_combinations, void *stream);

int qldpc_check_dev(qldpc_code *code, int64_t B, const uint32_t *err, const uint32_t *corr,
                    const uint32_t *synd, const uint8_t *conv, const int32_t *iters, int32_t distance,
                    uint8_t *flags, int32_t *weight, uint64_t *counters_dev, void *stream);

int qldpc_syndrome_dev(qldpc_code *code, int64_t B, const uint32_t *err, uint32_t *synd, void *stream);

/* XORs measurement errors of rate q into packed syndromes synd [B][WM] (paperResults.py:66-68) */
int qldpc_measurement_noise_dev(qldpc_code *code, double q, uint64_t seed, uint64_t first_shot, int64_t B,
                                uint32_t *synd, void *stream);

int qldpc_sample_dev(qldpc_code *code, double p, uint64_t seed, uint64_t first_shot, int32_t draws,
                     int64_t B, uint32_t *err, uint32_t *synd, void *stream);

/* fused BP -> OSD (osd_order as in qldpc_bposd_decode_host) on packed device syndromes; corr [B][words_n] */
int qldpc_bposd_decode_dev(qldpc_code *code, const qldpc_bp_config *cfg, const double *prior_host, int64_t B,
                           const uint32_t *synd, int32_t osd_order, uint32_t *corr, uint8_t *conv,
                           int32_t *iters, uint64_t *iter_total, void *stream);

int qldpc_pack_bits_dev(const uint8_t *in, uint32_t *out, int64_t B, int32_t nbits, void *stream);
int qldpc_unpack_bits_dev(const uint32_t *in, uint8_t *out, int64_t B, int32_t nbits, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* QLDPC_B200_H */
