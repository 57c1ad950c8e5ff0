/*
 * sst_b200.h — C-ABI of the B200 mass-explanation library (libsst_b200.so).
 *
 * The reference (SpectrSeqTools 0.1.2) has no FFI/plugin interface: its hot path is three Python
 * functions and one table builder.  Each entry point below names the reference function it replaces
 * (paths relative to /root/reference/spectrseqtools); spectrseqtools_b200/_cabi.py is the ctypes
 * binding, INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions: plain pointers and sizes only; every function returns an SST_* code (0 = ok) unless
 * noted; sst_last_error(ctx) gives the text of the last failure.  A context owns a main and a side CUDA
 * stream on one device and is not thread-safe; tables belong to the context that made them.  Host pointers may
 * be pageable or pinned (sst_host_alloc gives pinned memory); all calls are synchronous on return except
 * sst_classify_launch / sst_classify_async.
 * Integer masses are in table units (1 mDa for the stock alphabet).  The *_f64 entries and sst_classify take
 * float masses and do the float -> integer conversions of mass_explanation.py:51-58,107-114 on the device with
 * the same IEEE operations (true division, round-half-even, ceil; no FMA contraction), so they match CPython
 * bit for bit (tests/test_host.py, tests/test_gpu_explain.py).
 */
#ifndef SST_B200_H
#define SST_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sst_ctx sst_ctx;
typedef struct sst_table sst_table;

enum {
    SST_OK = 0,
    SST_ERR_CUDA = 1,          /* a CUDA runtime call failed (text in sst_last_error) */
    SST_ERR_NO_DEVICE = 2,     /* no sm_100 GPU / wrong architecture: the product has no CPU fallback */
    SST_ERR_BAD_ARG = 3,       /* -> ValueError */
    SST_ERR_COMPRESSION = 4,   /* unsupported cells-per-word -> ValueError (mass_table.py:285-289) */
    SST_ERR_TOO_MANY_ROWS = 5, /* more than 128 table rows */
    SST_ERR_TOO_DEEP = 6,      /* a composition could exceed 96 nucleotides */
    SST_ERR_NOMEM = 7,         /* result or scratch does not fit in device memory -> MemoryError */
    SST_ERR_MEMO_FULL = 8,     /* first-visit map too small: call again with a larger memo_capacity */
    SST_ERR_STATE = 9,         /* fetch without a preceding run */
    SST_ERR_OUT_OF_TABLE = 10, /* a probed mass lies beyond the table -> NotImplementedError */
    SST_ERR_NAN = 11,          /* a NaN mass: the reference's int(round(nan)) -> ValueError (mass_explanation.py:51,107) */
    SST_ERR_INF = 12           /* an infinite mass or threshold: int(round(inf)) / int(np.ceil(inf)) -> OverflowError */
};

/* per-peak budget modes for sst_explain (see DESIGN.md "Budget semantics") */
enum { SST_MODE_FREE = 0, SST_MODE_EXACT = 1, SST_MODE_MEMO = 2 };
/* per-peak status bits written by sst_explain_fetch */
enum { SST_STATUS_ZERO_IN_WINDOW = 1, SST_STATUS_OUT_OF_TABLE = 2 };
/* sst_is_valid results */
enum { SST_VALID_NO = 0, SST_VALID_YES = 1, SST_VALID_OUT_OF_TABLE = 2 };
#define SST_BUDGET_INF (1 << 30)

/* kernel slots of sst_kernel_ms */
enum {
    SST_K_BUILD = 0, SST_K_TRANSPOSE, SST_K_IS_VALID, SST_K_PHASE_A, SST_K_EXPLAIN_PASS, SST_K_CLASSIFY,
    SST_K_LENGTH_BOUND, SST_K_SPARE, SST_K_COUNT_
};

/* ---- context ---- */
int sst_ctx_create(int device, sst_ctx** out);
void sst_ctx_destroy(sst_ctx* ctx);
const char* sst_last_error(const sst_ctx* ctx);
int sst_device_info(sst_ctx* ctx, int* sm_count, int* cc_major, int* cc_minor, uint64_t* free_bytes, uint64_t* total_bytes);
void* sst_host_alloc(sst_ctx* ctx, size_t bytes); /* pinned host memory, NULL on failure */
void sst_host_free(sst_ctx* ctx, void* p);
/* page-lock / release memory the caller owns (e.g. a POSIX shared-memory segment: results copied there are visible to
 * the other processes of the box without a further copy — the host-side gather of SURVEY §8e) */
int sst_host_register(sst_ctx* ctx, void* p, size_t bytes);
int sst_host_unregister(sst_ctx* ctx, void* p);
/* CUDA-event stopwatch on the context's stream (what bench.py times with) */
int sst_timer_start(sst_ctx* ctx);
int sst_timer_stop(sst_ctx* ctx, float* ms);
/* device time from sst_timer_start to the end of the device work of the last sst_explain_run (its result read-back
 * included, the host's wake-up after it not) */
int sst_timer_stop_at_run(sst_ctx* ctx, float* ms);
/* device time of each kernel family accumulated since the last sst_stats_reset, and launches made */
int sst_stats_reset(sst_ctx* ctx);
int sst_kernel_ms(sst_ctx* ctx, float* ms /* [SST_K_COUNT_] */, uint64_t* launches /* [SST_K_COUNT_] */);
/* blow-up guard: sst_explain fails with SST_ERR_NOMEM when a pass would hold more than `limit` partial
 * compositions in one level (0 = default: whatever fits in device memory; the reference would simply never
 * return on such inputs) */
int sst_set_item_limit(sst_ctx* ctx, uint64_t limit);
/* write `bytes` of device scratch (L2 flush between timed iterations) */
int sst_flush_l2(sst_ctx* ctx, size_t bytes);

/* ---- DP table: replaces set_up_bit_table (mass_table.py:207-248) and load_dp_table (:319-340) ----
 * weights[0] must be 0, the rest strictly ascending, every non-zero weight >= 32; R <= 128.
 * compression must be 32 (uint64 cells); last_col_mask is numpy's value of
 * full << 2*(max_col - (max_mass+1) % max_col) (mass_table.py:246), computed by the caller.
 * with_masks != 0 also builds the mass-major row masks the enumerator needs. */
int sst_table_build(sst_ctx* ctx, const int64_t* weights, int R, int64_t max_mass, int compression,
                    uint64_t last_col_mask, int with_masks, sst_table** out);
/* adopt a table computed elsewhere (row-major R x C uint64, reference layout) */
int sst_table_upload(sst_ctx* ctx, const uint64_t* host_table, const int64_t* weights, int R, int64_t C, sst_table** out);
/* run the build kernels again into the same buffers (benchmarking) */
int sst_table_rebuild(sst_ctx* ctx, sst_table* t);
int sst_table_info(const sst_table* t, int* R, int64_t* C, float* build_ms, float* transpose_ms);
int sst_table_download(sst_ctx* ctx, const sst_table* t, uint64_t* host_out /* R*C */);
int sst_table_download_masks(sst_ctx* ctx, const sst_table* t, int64_t first_mass, int64_t n, uint32_t* host_out /* n*4 */);
void sst_table_destroy(sst_ctx* ctx, sst_table* t);

/* ---- validity: replaces is_valid_mass (mass_explanation.py:45-89) for P (target, threshold) pairs ---- */
int sst_is_valid(sst_ctx* ctx, const sst_table* t, const int64_t* target, const int64_t* thr, int64_t P,
                 uint8_t* out /* P, SST_VALID_* */);
/* the same in three steps (stage = H2D, run = kernel, fetch = D2H) for kernel-only timing */
int sst_valid_stage(sst_ctx* ctx, const int64_t* target, const int64_t* thr, int64_t P);
/* float inputs: the device does the reference's float -> integer conversion (mass_explanation.py:51-58) with
 * the same IEEE operations; thr may be NULL (all relative) and a NaN entry means "threshold None" */
int sst_valid_stage_f64(sst_ctx* ctx, const double* mass, const double* thr, int64_t P, double precision, double tolerance);
int sst_valid_run(sst_ctx* ctx, const sst_table* t);
/* stage + run + fetch in one call with one synchronisation (the reference-shaped is_valid_mass is a batch of one) */
int sst_is_valid_f64(sst_ctx* ctx, const sst_table* t, const double* mass, const double* thr, int64_t P, double precision, double tolerance,
                     uint8_t* out);
int sst_valid_fetch(sst_ctx* ctx, uint8_t* out);

/* ---- fragment classification: replaces the per-(fragment x breakage) map_elements callbacks of classify_fragments
 * (fragment_classification.py:39-82): standard-unit mass = observed[f] - offsets[b] (offsets[b] = breakage weight *
 * precision, multiplied by the caller as the reference does), threshold = tolerance * observed[f], validity probe
 * (is_valid_mass) and singleton test (is_singleton, :104-119).  out[b * F + f] (breakage-major, the reference's
 * concat order): bit 1 valid, bit 2 out-of-table value met before any hit, bit 4 singleton. ---- */
enum { SST_CLASS_VALID = 1, SST_CLASS_OUT_OF_TABLE = 2, SST_CLASS_SINGLETON = 4 };
int sst_classify(sst_ctx* ctx, const sst_table* t, const double* observed, int64_t F, const double* offsets, int B, double precision,
                 double tolerance, uint8_t* out /* B*F */);
int sst_classify_stage(sst_ctx* ctx, const double* observed, int64_t F, const double* offsets, int B);
int sst_classify_run(sst_ctx* ctx, const sst_table* t, double precision, double tolerance);
/* the same without waiting: the kernel is queued on the context's stream, the next synchronous call completes it */
int sst_classify_launch(sst_ctx* ctx, const sst_table* t, double precision, double tolerance);
int sst_classify_fetch(sst_ctx* ctx, uint8_t* out /* B*F */);
/* the whole call on the context's side stream without waiting (inputs and `out` should be pinned and must stay valid
 * until sst_classify_wait returns); an enumeration pass issued in between overlaps it */
int sst_classify_async(sst_ctx* ctx, const sst_table* t, const double* observed, int64_t F, const double* offsets, int B,
                       double precision, double tolerance, uint8_t* out /* B*F */);
/* the same with two flags per byte: fragment f of breakage b in the low (f even) or high nibble of byte
 * (b * Fp + f) / 2, Fp = F rounded up to even; out holds B * Fp / 2 bytes */
int sst_classify_async_packed(sst_ctx* ctx, const sst_table* t, const double* observed, int64_t F, const double* offsets, int B,
                              double precision, double tolerance, uint8_t* out /* B*Fp/2 */);
int sst_classify_wait(sst_ctx* ctx);

/* ---- sequence-length bounds: replaces compute_sequence_length_bound (mass_table.py:343-487), both directions in one
 * walk.  target / thr in table units (round(su_mass / precision), ceil(tolerance * obs_mass / precision)),
 * max_mods = round(modification_rate * max_len), ind / is_mod as for sst_explain.  First-visit memo semantics make
 * the walk sequential (one device thread); memo_capacity: slots of its map (0 = default, SST_ERR_MEMO_FULL = retry
 * with more).  SST_ERR_OUT_OF_TABLE when a window value lies beyond the table. ---- */
int sst_length_bounds(sst_ctx* ctx, const sst_table* t, int64_t target, int64_t thr, int32_t max_mods, int32_t max_len,
                      const int32_t* ind, const uint8_t* is_mod, uint64_t memo_capacity, int64_t* lower, int64_t* upper);

/* ---- enumeration: replaces explain_mass_with_table (mass_explanation.py:92-203) for P peaks ----
 * max_mods[p]: global modification budget (SST_BUDGET_INF = unbounded); mode[p]: SST_MODE_*;
 * ind[r] = round(max_len * rate_r), is_mod[r]: per-row budget data shared by the batch (length R).
 * rec_width: bytes per composition record (multiple of 8, >= longest possible composition).
 * memo_capacity: slots of the first-visit map (0 = default); only used when some mode is MEMO.
 * Results stay on the device until sst_explain_fetch / the next run. */
int sst_explain(sst_ctx* ctx, const sst_table* t, const int64_t* target, const int64_t* thr, const int32_t* max_mods,
                const uint8_t* mode, int64_t P, const int32_t* ind, const uint8_t* is_mod, int rec_width,
                uint64_t memo_capacity, uint64_t* n_roots, uint64_t* n_comps);
/* the same in two halves, so that the kernel-only time can be measured with inputs resident in HBM */
int sst_explain_stage(sst_ctx* ctx, const sst_table* t, const int64_t* target, const int64_t* thr, const int32_t* max_mods,
                      const uint8_t* mode, int64_t P, const int32_t* ind, const uint8_t* is_mod);
/* float inputs: integerises on the host with the reference's float operations (mass_explanation.py:107-114) and
 * picks the per-peak budget mode (FREE when no budget can bind, else MEMO if with_memo else EXACT) */
int sst_explain_stage_f64(sst_ctx* ctx, const sst_table* t, const double* mass, const double* thr, const int32_t* max_mods,
                          int64_t P, const int32_t* ind, const uint8_t* is_mod, double precision, double tolerance, int with_memo);
/* the same with one modification budget for the whole batch (what calculate_explanations passes, common.py:47-65:
 * max_modifications = round(rate * max_len)): no per-peak budget array crosses the bus */
int sst_explain_stage_f64_uniform(sst_ctx* ctx, const sst_table* t, const double* mass, const double* thr, int32_t max_mods,
                                  int64_t P, const int32_t* ind, const uint8_t* is_mod, double precision, double tolerance, int with_memo);
/* rec_width 0 = smallest multiple of 8 that holds the longest possible composition of the staged batch */
int sst_explain_run(sst_ctx* ctx, const sst_table* t, int rec_width, uint64_t memo_capacity, uint64_t* n_roots,
                    uint64_t* n_comps);
int sst_explain_rec_width(const sst_ctx* ctx); /* record width of the last run */
/* Which enumeration pass sst_explain_run uses.  0 = automatic (default): the depth-first ITEM pass (sst_enum.cuh: count ->
 * scan -> fill over partial compositions in peak order, one grid barrier) for batches of ladder differences (at most 16
 * nucleotides per composition, a handful of compositions per peak by the staged batch's cost estimate); the
 * LEVEL-synchronous pass (sst_explain.cuh) for deeper compositions, for batches with thousands of compositions per
 * peak (it spreads single huge subtrees over the machine) and for batches the item pass gives back.  1 = always
 * level-synchronous, 2 = item pass or fail, 3 = DIRECT pass or fail (sst_direct.cuh: per-peak counts looked up in a
 * composition-count table built once per alphabet, scan, output-balanced fill; batches whose budgets cannot bind, at
 * most 16 nucleotides, windows below 2^22 integer masses), -3 = the same as 0.  All passes return the same
 * compositions per peak; direct and item pass also the same record order (depth-first), the level pass another. */
int sst_set_pass(sst_ctx* ctx, int which);
int sst_last_pass(const sst_ctx* ctx); /* 1 = level-synchronous, 2 = depth-first items, 3 = direct produced the last result */
/* diagnostics of the depth-first pass: enable != 0 makes the following runs record %globaltimer (ns) of every CTA at
 * [0] start, [1] first tile's window values counted, [2] its roots written, [3] count phase done, [4] grid barrier
 * passed, [5] fill phase done; out (may be NULL) receives [n_ctas][8] of the last recorded run */
int sst_explain_cta_ns(sst_ctx* ctx, int enable, uint64_t* out, int cap_ctas, int* n_ctas);
/* diagnostics: device timestamps (ns, %globaltimer, CTA 0's clock) at the phase boundaries of the last enumeration
 * pass, in order: [0] start, [1] window values counted, [2] level-0 nodes written, then for every level
 * [counted, grid barrier passed, written]; after the last level [per-level peak totals summed], [placement table
 * written], [records permuted].  Right after a grid barrier that is every CTA's clock.  Unused slots are 0. */
int sst_explain_phase_ns(const sst_ctx* ctx, uint64_t* out /* [32] */);
/* Compositions per call WITHOUT enumerating them: counts_out[p] = what explain_mass_with_table (mass_explanation.py:92-203)
 * would return in number for (mass[p], thr[p]; NaN / NULL = relative) when no modification budget binds — looked up in
 * the composition-count table of the alphabet (built on first use: cnt(r, m) = cnt(r-1, m) + bit1(r, m) * cnt(r, m - w_r),
 * the reference's own UP / LEFT recursion on integers).  ~0 when a window reaches beyond the count table (2^22 integer
 * masses), a count saturates, or an input is not finite.  The host-side partition of a workload over the GPUs of a box
 *  balances blocks by these. */
int sst_count_compositions_f64(sst_ctx* ctx, sst_table* t, const double* mass, const double* thr, int64_t P, double precision,
                               double tolerance, uint64_t* counts_out);
/* ---- N3 + N4: ladder differences and the explanation-based alphabet reduction on a fragment frame that
 * stays on the device between the rounds (kernels: sst_ladder.cuh).
 *
 * sst_ladder_stage: the classified fragments, sorted by standard-unit mass as prediction.py:68-72 leaves them —
 * su[F], observed[F], flags[F] (1 = breakage contains START, 2 = contains END, 4 = is_singleton).  Every fragment
 * starts alive.
 *
 * sst_ladder_round replaces Predictor.collect_diff_explanations_for_su (prediction.py:261-329) for the alive fragments:
 * the two-pointer window of each side (`diff > max_weight` restart, tail behaviour once `end` sits on the last
 * fragment), the l1 error threshold tolerance * (obs1 + obs2) (common.py:37-44), then the singletons (su, tolerance *
 * obs) — generated on the device straight into a staged batch, enumerated by sst_explain_run's machinery, deduplicated
 * by key as the reference's dicts do (the last entering call with a key wins; a pair enters with >= 1 explanation, a
 * singleton always).  mask_out[4]: bit r set = table row r occurs in a winning explanation — what
 * Predictor.filter_by_explanation (:170-202) turns into the reduced alphabet.  Only that mask and two counters cross
 * the bus.  The explanations themselves stay available: sst_explain_fetch (calls in generation order) +
 * sst_ladder_fetch (keys, thresholds, per-call flags: 1 = entered its dict, 2 = is the surviving entry of its key).
 *
 * sst_ladder_revalidate replaces the loop of Predictor._reduce_alphabet (:204-227) after the table has been rebuilt for
 * the reduced alphabet: is_valid_mass(su, tolerance * observed) of every alive fragment; the ones that fail die.
 * SST_ERR_OUT_OF_TABLE as is_valid_mass raises. */
int sst_ladder_stage(sst_ctx* ctx, const double* su, const double* observed, const uint8_t* flags, int64_t F);
int sst_ladder_round(sst_ctx* ctx, const sst_table* t, double max_weight, double precision, double tolerance, int32_t max_mods,
                     const int32_t* ind, const uint8_t* is_mod, int with_memo, uint32_t* mask_out /* [4] */, uint64_t* n_calls,
                     uint64_t* n_comps);
int sst_ladder_revalidate(sst_ctx* ctx, const sst_table* t, double precision, double tolerance, int64_t* n_alive);
int sst_ladder_fetch(sst_ctx* ctx, uint8_t* alive_out /* [F] */, double* key_out /* [n_calls] */, double* thr_out /* [n_calls] */,
                     uint8_t* call_flags_out /* [n_calls] */);
/* diagnostics: host wall time (ns) and number of visits per section of the asynchronous entries since the last call
 * (sections are the marks in sst_cabi.cu: 0-12 sst_explain_submit_f64, 13 the wait in sst_explain_collect, 16-20
 * sst_classify_async); returns the sums, clears them and switches the stopwatch on or off (process-wide) */
int sst_host_profile(int enable, uint64_t* ns_out /* [32] or NULL */, uint64_t* calls_out /* [32] or NULL */);
/* diagnostics: device timeline (ms since a process-wide origin) of the last batch queued by sst_explain_submit_f64 on
 * this context, once it has been collected: [0] first operation, [1] inputs on the device, [2] staged, [3] pass done,
 * [4] results in host memory; -1 = not recorded.  enable switches the recording on or off for the following batches */
int sst_trace_ms(sst_ctx* ctx, int enable, float* out /* [8] or NULL */);
/* The whole call — replaces a loop of calculate_explanations (common.py:47-65) — WITHOUT WAITING: the copy of the inputs,
 * staging, the enumeration pass and ONE copy of the results are queued on the context's stream; sst_explain_collect
 * waits for them.  One modification budget for the batch (what calculate_explanations passes).  All host pointers must
 * stay valid until sst_explain_collect returns.  Results arrive in out_block (pinned: sst_host_alloc), laid out as
 * sst_explain_block_layout(P) says: a 512-byte header the library uses, status[P], peak offsets as uint32[P + 1],
 * records; block_bytes - recs_off bytes are available for records.  The copy carries as many records as the previous
 * batch had (+ 2 %, environment SST_SPEC_MARGIN_PCT); sst_explain_collect fetches the rest if this batch is larger.
 *
 * Nothing looks at the host arrays: the batch is queued on the assumption that it is like the previous one (no
 * modification budget binds, compositions fit the record width of last time, a handful of compositions per peak), and
 * the pass checks the staged batch's summary ON THE DEVICE before it does anything.  A batch that breaks an assumption
 * comes back untouched and is carried out synchronously inside sst_explain_collect (so are batches after it, until
 * one of them would have fitted again).  Non-finite masses / thresholds: SST_ERR_NAN / SST_ERR_INF from
 * sst_explain_collect.  SST_ERR_NOMEM from sst_explain_collect with *n_comps set: the block is too small for the
 * records — the result is still on the device, fetch it with sst_explain_fetch.  Several contexts on one device give
 * several batches in flight: the copies of one overlap the kernels of another (three keep the copy engine busy). */
int sst_explain_submit_f64(sst_ctx* ctx, const sst_table* t, const double* mass, const double* thr, int32_t max_mods, int64_t P,
                           const int32_t* ind, const uint8_t* is_mod, double precision, double tolerance, int with_memo,
                           uint8_t* out_block, uint64_t block_bytes);
int sst_explain_block_layout(int64_t P, uint64_t* status_off, uint64_t* off32_off, uint64_t* recs_off);
/* Split records for the following submissions on this context (off by default): a batch whose compositions fit 8-byte
 * records comes back as PLANES instead of whole records — uint32 lo[cap_n] at recs_off (the first four nucleotides of
 * record i in lo[i], byte 0 = smallest row), then hi_planes byte planes of cap_n bytes each (plane k: nucleotide 5 + k)
 * — with hi_planes = the previous batch's longest composition - 4: a batch of <= 5-nt ladder differences crosses the
 * bus with 5 bytes per composition instead of 8.  cap_n = (block_bytes - recs_off) / (4 + hi_planes), rounded down to a
 * multiple of 16.  sst_explain_rec_layout tells how the last collected submission's records are laid out (split = 0:
 * whole records of rec_width bytes at recs_off, e.g. after a submission that was redone synchronously). */
int sst_set_record_split(sst_ctx* ctx, int enable);
int sst_explain_rec_layout(const sst_ctx* ctx, int* split, uint64_t* cap_n, int* hi_planes);
/* bytes the last collected submission copied device -> host (block header, status, offsets, records incl. the margin) */
uint64_t sst_explain_d2h_bytes(const sst_ctx* ctx);
int sst_explain_collect(sst_ctx* ctx, const sst_table* t, uint64_t* n_comps, int* rec_width);
/* status[P]; peak_off[P+1] (compositions of peak p are records peak_off[p] .. peak_off[p+1]);
 * recs[n_comps * rec_width]: row indices in ascending order, 0-padded.  Any pointer may be NULL. */
int sst_explain_fetch(sst_ctx* ctx, uint8_t* status, uint64_t* peak_off, uint8_t* recs);

#ifdef __cplusplus
}
#endif
#endif /* SST_B200_H */
