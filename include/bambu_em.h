/*
 * bambu_em.h -- C ABI of the B200-native quantification EM (libbambu_em.so).
 *
 * This library replaces, for one whole sample (or many samples) at a time,
 * what the reference bambu 3.3.5 does per gene group through
 *
 *     abundance_quantification()  R/bambu-quantify_utilityFunctions.R:379-391
 *       run_parallel()            R/bambu-quantify_utilityFunctions.R:401-425
 *         getAMat() x3            R/bambu-quantify_utilityFunctions.R:427-431
 *         emWithL1()              R/RcppExports.R:12-14 -> src/RcppExports.cpp:31-46
 *           em_theta()            src/em.cpp:9-71
 *           3 weighted row sums   src/em.cpp:97-102
 *
 * i.e. one call here == the reference's serial lapply over all gene groups,
 * each of which is one .Call("_bambu_emWithL1", A, A_full, A_unique, Y, K,
 * maxiter, minvalue, conv) (R/RcppExports.R:13).
 *
 * Conventions
 *   - plain C, no C++/torch types; caller owns every buffer passed in; the
 *     library never keeps a caller pointer after the call returns (except the
 *     *_bind_device call, which says so);
 *   - every function returns BAMBU_EM_OK (0) or a negative error code and
 *     never throws or longjmps (the reference turns C++ exceptions into R
 *     errors with BEGIN_RCPP/END_RCPP, src/RcppExports.cpp:32,45; an R shim
 *     turns a non-zero code into stop(bambu_em_last_error()));
 *   - there is NO CPU fallback: without a usable sm_100 device the calls fail
 *     with BAMBU_EM_ENODEVICE;
 *   - one host thread per process may drive a plan at a time (R is single
 *     threaded; BiocParallel workers are processes, R/bambu_utilityFunctions.R:6-15);
 *     the one-shot entry points (bambu_em_batched*, bambu_emWithL1, bambu_quantDT*) share a
 *     cached device workspace and are NOT reentrant: call them from one thread at a time;
 *   - txid values must fit R's integer (int32): the arena carries them so; wider ids are
 *     rejected with BAMBU_EM_EINVAL;
 *   - row_nnz_ptr has grp_row_ptr[n_groups] + 1 elements.
 *
 * Arena (segmented CSR; see bambu_b200/arena.py for the prose version)
 *   group g owns rows    [grp_row_ptr[g], grp_row_ptr[g+1])   (transcripts, txid ascending)
 *           and columns  [grp_col_ptr[g], grp_col_ptr[g+1])   (read classes, (gene_sid, eqClassId) ascending)
 *   row r   owns entries [row_nnz_ptr[r], row_nnz_ptr[r+1])   col_idx local to the group,
 *                                                             strictly increasing inside a row
 *   aval      = the "aval" column of the long table              (A      of src/em.cpp:77)
 *   nz_flags  bit0 = equal        => A_full   = aval * equal     (R/...utilityFunctions.R:337)
 *   col_flags bit0 = multi_align  => A_unique = aval * !multi    (R/...utilityFunctions.R:336)
 *   Y = n.obs per column, K = K per column                       (R/...utilityFunctions.R:359-371)
 *
 * Outputs
 *   est[3 * total_rows]  est[k*total_rows + r]: k = 0 counts, 1 fullLengthCounts,
 *                        2 uniqueCounts  (rows of emWithL1's "theta", src/em.cpp:104-107)
 *   iters[n_groups]      EM updates done for the group (the `iter` em_theta computes at
 *                        src/em.cpp:45,69 and emWithL1 drops at :93-94); 0 for
 *                        single-transcript groups (closed form, R/...utilityFunctions.R:410-414)
 *   status[n_groups]     bit0: stopped by maxiter, bit1: non-finite estimate in the group
 */
#ifndef BAMBU_EM_H
#define BAMBU_EM_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BAMBU_EM_OK            0
#define BAMBU_EM_EINVAL      (-1)  /* malformed arena / arguments                    */
#define BAMBU_EM_ENODEVICE   (-2)  /* no usable sm_100 CUDA device                   */
#define BAMBU_EM_ECUDA       (-3)  /* CUDA runtime error (see bambu_em_last_error)   */
#define BAMBU_EM_ENOMEM      (-4)  /* host or device allocation failed               */
#define BAMBU_EM_ESTATE      (-5)  /* plan used out of order                         */

#define BAMBU_EM_STATUS_MAXITER    1
#define BAMBU_EM_STATUS_NONFINITE  2

#define BAMBU_NZ_EQUAL   1
#define BAMBU_COL_MULTI  1

/* ---- library-level ---------------------------------------------------------- */
const char *bambu_em_version(void);
/* message of the last failing call on this thread ("" if none) */
const char *bambu_em_last_error(void);
/* number of usable devices (sm_100); <0 on error */
int bambu_em_device_count(void);
/* select the device used by subsequent calls of this process (default 0) */
int bambu_em_set_device(int device);
/* free the cached workspace of bambu_em_batched (optional; R: .onUnload,
 * R/bambu-quantify_utilityFunctions.R:505-507) */
int bambu_em_shutdown(void);

/* ---- one-shot, host buffers: the drop-in for abundance_quantification ----------
 * Replaces R/bambu-quantify.R:64-66.  All pointers are HOST pointers.  Copies the
 * arena to the device (pipelined in chunks of groups), runs the kernels, copies
 * est/iters/status back, returns when the outputs are complete. */
int bambu_em_batched(int64_t n_groups,
                     const int64_t *grp_row_ptr, const int64_t *grp_col_ptr,
                     const int64_t *row_nnz_ptr, const int32_t *col_idx, const double *aval,
                     const uint8_t *nz_flags, const double *Y, const double *K,
                     const uint8_t *col_flags,
                     int32_t maxiter, double minvalue, double conv,
                     double *est, int32_t *iters, int32_t *status);

/* ---- the same call fed with the wire form v2 ("live arena") ----------------------------------------
 * What bounds a cohort end to end is the host->device copy of the arena, so v2 carries only what the EM
 * reads (about 37 % of the v1 bytes on the synthetic cohorts).  The packer applies the two exact
 * eliminations the device applies to a v1 arena -- columns with Y == 0 and entries with aval == 0 cannot
 * change a result bit (src/em.cpp:48-49, 98-99) -- and therefore also evaluates the one quantity the dropped
 * entries still feed, rs_i = sum_j X_ij (src/em.cpp:33; two accumulators by column parity, added at the end).
 *
 *   grp_row_ptr[n_groups+1]  rows of group g (as in v1)
 *   grp_col_ptr[n_groups+1]  LIVE columns of group g (Y != 0), in the v1 column order
 *   grp_ent_ptr[n_groups+1]  LIVE entries of group g (aval != 0 in a live column), row by row
 *   row_end[rows]            end of the row's entries relative to its group's first entry (u32)
 *   rs[rows]                 sum of ALL aval of the row (f64)
 *   aval[entries]            f64
 *   cidx                     byte blob; group g's block starts at the next multiple of 4 after group g-1's
 *                            and holds, per entry, (live column index << 1) | (ORIGINAL column index & 1):
 *                            uint16 while the group has at most 32767 live columns, uint32 otherwise;
 *                            strictly increasing inside a row
 *   eq_bits                  bit k (word k/32, bit k%32): entry k is `equal`
 *   col_mode 1               colY = int32 nobs, colK = int32 K per live column; the device forms
 *                            Y = nobs / K exactly as R does (R/bambu-quantify_utilityFunctions.R:221-228)
 *   col_mode 0               colY = f64 Y, colK = f64 K per live column
 *   multi_bits               bit j: live column j is multi_align
 *
 * Outputs as in bambu_em_batched.  K must be finite.  The v1 entry point stays; both give identical doubles. */
int bambu_em_batched_v2(int64_t n_groups, const int64_t *grp_row_ptr, const int64_t *grp_col_ptr,
                        const int64_t *grp_ent_ptr, const uint32_t *row_end, const double *rs, const double *aval,
                        const void *cidx, const uint32_t *eq_bits, int32_t col_mode, const void *colY,
                        const void *colK, const uint32_t *multi_bits,
                        int32_t maxiter, double minvalue, double conv,
                        double *est, int32_t *iters, int32_t *status);

/* v1 arena -> wire form v2 (native host code, csrc/wire.cpp).  Borrowed pointers, valid until bambu_wire_free.
 * nnz_nonzero[n_groups] = entries with aval != 0 per group incl. those of dropped columns (the nnz_g of the
 * nnz-iterations metric).  total_bytes = what bambu_em_batched_v2 copies to the device. */
typedef struct bambu_wire bambu_wire;
const char *bambu_wire_last_error(void);
int bambu_wire_create(bambu_wire **wire, int64_t n_groups, const int64_t *grp_row_ptr, const int64_t *grp_col_ptr,
                      const int64_t *row_nnz_ptr, const int32_t *col_idx, const double *aval, const uint8_t *nz_flags,
                      const double *Y, const double *K, const uint8_t *col_flags);
int bambu_wire_free(bambu_wire *wire);
int bambu_wire_sizes(const bambu_wire *wire, int64_t *n_groups, int64_t *n_rows, int64_t *n_cols, int64_t *n_entries,
                     int64_t *cidx_bytes, int32_t *col_mode, int64_t *total_bytes);
int bambu_wire_arrays(const bambu_wire *wire, const int64_t **grp_row_ptr, const int64_t **grp_col_ptr,
                      const int64_t **grp_ent_ptr, const uint32_t **row_end, const double **rs, const double **aval,
                      const void **cidx, const uint32_t **eq_bits, const void **colY, const void **colK,
                      const uint32_t **multi_bits, const int64_t **nnz_nonzero);

/* ---- one dense group: the drop-in for one emWithL1 call ------------------------
 * Replaces .Call("_bambu_emWithL1", ...) (R/RcppExports.R:13, src/em.cpp:77-110).
 * A, A_full, A_unique: M x N column-major (R matrices); Y, K: N.  est: 3 x M
 * column-major, i.e. est[3*i + k] -- the layout of the "theta" matrix emWithL1
 * returns.  iters_out may be NULL. */
int bambu_emWithL1(const double *A, const double *A_full, const double *A_unique,
                   const double *Y, const double *K, int32_t M, int32_t N,
                   int32_t maxiter, double minvalue, double conv,
                   double *est, int32_t *iters_out);

/* ---- plan API: arena resident in HBM, repeated runs, caller-visible stream ------- */
typedef struct bambu_em_plan bambu_em_plan;

/* Analyse the group structure (HOST pointer arrays), build the size buckets and
 * work queues, allocate device memory for the arena, the packed group images
 * and the outputs. */
int bambu_em_plan_create(bambu_em_plan **plan, int64_t n_groups,
                         const int64_t *grp_row_ptr, const int64_t *grp_col_ptr,
                         const int64_t *row_nnz_ptr);
int bambu_em_plan_destroy(bambu_em_plan *plan);
/* Re-analyse a NEW group structure in an existing plan, keeping its device memory (buffers only grow): the
 * cheap way to push differently shaped arenas (one per sample) through one plan.  The payload has to be
 * uploaded / bound again. */
int bambu_em_plan_reset(bambu_em_plan *plan, int64_t n_groups, const int64_t *grp_row_ptr,
                        const int64_t *grp_col_ptr, const int64_t *row_nnz_ptr);
/* stream all work of this plan is ordered on (a cudaStream_t passed as void*);
 * NULL = the plan's own stream.  Lets a harness bracket runs with its own events. */
int bambu_em_plan_set_stream(bambu_em_plan *plan, void *cuda_stream);
/* asynchronous H2D of the arena payload from HOST buffers (pinned => truly async) */
int bambu_em_plan_upload(bambu_em_plan *plan, const int32_t *col_idx, const double *aval,
                         const uint8_t *nz_flags, const double *Y, const double *K,
                         const uint8_t *col_flags);
/* use DEVICE buffers owned by the caller as the arena payload (no copy; the
 * pointers must stay valid until the plan is destroyed or re-bound) */
int bambu_em_plan_bind_device(bambu_em_plan *plan, const int32_t *d_col_idx, const double *d_aval,
                              const uint8_t *d_nz_flags, const double *d_Y, const double *d_K,
                              const uint8_t *d_col_flags);
/* the plan API on the wire form v2: structure and payload as bambu_em_batched_v2 takes them.  For such a plan
 * nnz_nonzero (bambu_em_plan_download) counts the live entries. */
int bambu_em_plan_create_v2(bambu_em_plan **plan, int64_t n_groups, const int64_t *grp_row_ptr,
                            const int64_t *grp_col_ptr, const int64_t *grp_ent_ptr);
int bambu_em_plan_upload_v2(bambu_em_plan *plan, const uint32_t *row_end, const double *rs, const double *aval,
                            const void *cidx, const uint32_t *eq_bits, int32_t col_mode, const void *colY,
                            const void *colK, const uint32_t *multi_bits);
/* enqueue pack + EM kernels (asynchronous) */
int bambu_em_plan_run(bambu_em_plan *plan, int32_t maxiter, double minvalue, double conv);
/* asynchronous D2H of the outputs into HOST buffers, then wait for completion.
 * Any of the pointers may be NULL.  nnz_nonzero[n_groups] receives, per group,
 * the number of entries with aval != 0 (the nnz_g of the nnz-iterations metric). */
int bambu_em_plan_download(bambu_em_plan *plan, double *est, int32_t *iters, int32_t *status,
                           int64_t *nnz_nonzero);
/* what the pack kernels kept per group: live columns (Y != 0), the entries with
 * aval != 0 inside them -- the part of the group the EM loop actually iterates on --
 * the shared memory its EM needs and the number of 32-lane slots of its image.
 * Any pointer may be NULL. */
int bambu_em_plan_download_live(bambu_em_plan *plan, int32_t *live_cols, int32_t *live_nnz,
                                int32_t *smem_bytes, int32_t *slots);
int bambu_em_plan_sync(bambu_em_plan *plan);
/* device pointers of the outputs (est: 3*total_rows f64; iters, status: n_groups i32) */
int bambu_em_plan_device_outputs(bambu_em_plan *plan, double **d_est, int32_t **d_iters,
                                 int32_t **d_status);
/* device time of the last completed run, from CUDA events on the launch stream:
 * pack + SELL kernels (start -> all images built), EM kernels (-> end), whole run.
 * Waits for the run. */
int bambu_em_plan_last_run_ms(bambu_em_plan *plan, float *pack_ms, float *em_ms, float *total_ms);
/* kernels launched by the last run */
int bambu_em_plan_last_run_launches(bambu_em_plan *plan, int32_t *n_launches);
/* bytes of device memory held by the plan */
int bambu_em_plan_device_bytes(bambu_em_plan *plan, int64_t *bytes);

/* ---- the stage around the EM: bambu.quantDT (R/bambu-quantify.R:53-74), native host code ----------
 * Long table: one row per (equivalence class, transcript) pair with the columns GENEID (any integer
 * coding, e.g. match(GENEID, unique(GENEID))), txid, eqClassId, nobs, equal, rcWidth, txlen
 * (R/bambu-quantify_utilityFunctions.R:199-203).  equal / rcWidth / txlen may be NULL (FALSE / 0 / 1).
 *
 * bambu_quant_pack_create = addAval ... removeUnObservedGenes, initialiseOutput, filterTxRc,
 * assignGroups, getInputList (R/...utilityFunctions.R:196-371): the arena for bambu_em_batched plus the
 * bookkeeping the merge needs.  bambu_quant_merge = modifyQuantOut + removeDuplicates (:438-455).
 * bambu_quantDT = pack -> bambu_em_batched -> merge: the drop-in for bambu.quantDT. */
typedef struct bambu_quant_pack bambu_quant_pack;
const char *bambu_quant_last_error(void);
int bambu_quant_pack_create(bambu_quant_pack **pack, int64_t n_rows, const int64_t *gene_code, const int64_t *txid,
                            const int64_t *eqClassId, const double *nobs, const uint8_t *equal, const double *rcWidth,
                            const double *txlen, int32_t degradation_bias);
int bambu_quant_pack_free(bambu_quant_pack *pack);
int bambu_quant_pack_sizes(const bambu_quant_pack *pack, int64_t *n_groups, int64_t *n_rows, int64_t *n_cols,
                           int64_t *nnz, int64_t *n_out, int64_t *n_unobserved, double *d_rate);
/* borrowed pointers, valid until bambu_quant_pack_free; any argument may be NULL */
int bambu_quant_pack_arena(const bambu_quant_pack *pack, const int64_t **grp_row_ptr, const int64_t **grp_col_ptr,
                           const int64_t **row_nnz_ptr, const int32_t **col_idx, const double **aval,
                           const uint8_t **nz_flags, const double **Y, const double **K, const uint8_t **col_flags,
                           const int32_t **txid);
int bambu_quant_pack_bookkeeping(const bambu_quant_pack *pack, const int64_t **out_txid,
                                 const int64_t **unobserved_txid, const int64_t **gene_grp_id);
/* est: 3 x n_rows as returned by bambu_em_batched; outputs hold up to n_unobserved + n_out rows */
int bambu_quant_merge(const bambu_quant_pack *pack, const double *est, int64_t *txid_out, double *counts,
                      double *full_length, double *unique_counts, int64_t *n_result);
int bambu_quantDT(int64_t n_rows, const int64_t *gene_code, const int64_t *txid, const int64_t *eqClassId,
                  const double *nobs, const uint8_t *equal, const double *rcWidth, const double *txlen,
                  int32_t degradation_bias, int32_t maxiter, double minvalue, double conv, int64_t *txid_out,
                  double *counts, double *full_length, double *unique_counts, int64_t *n_result, double *d_rate);
/* 1 if the last bambu_quantDT / bambu_quantDT_dotC call of this thread packed its table on the GPU, 0 if the table went
 * through the host packer (same result, ~25x slower: see bambu_quantDT_hostpack), -1 before any call. */
int bambu_quant_last_used_device(void);
/* The same with the table packed by the host code of bambu_quant_pack_create (bambu_quantDT packs on the GPU and
 * uses this for tables its device formulation does not cover: non-integer or negative nobs, negative or
 * non-finite widths / lengths, a degradation rate that is not a finite number >= 0). */
int bambu_quantDT_hostpack(int64_t n_rows, const int64_t *gene_code, const int64_t *txid, const int64_t *eqClassId,
                  const double *nobs, const uint8_t *equal, const double *rcWidth, const double *txlen,
                  int32_t degradation_bias, int32_t maxiter, double minvalue, double conv, int64_t *txid_out,
                  double *counts, double *full_length, double *unique_counts, int64_t *n_result, double *d_rate);
/* bambu_quantDT for R's .C() interface (no compiled shim needed): every argument points to one of R's own vector
 * types -- integer (int), numeric (double), logical (int; TRUE = 1) -- scalars are length-1 vectors, and the return
 * code comes back through rc.  Output vectors must hold as many elements as there are distinct txids (<= n_rows). */
void bambu_quantDT_dotC(const int *n_rows, const int *gene_code, const int *txid, const int *eqClassId,
                        const double *nobs, const int *equal, const double *rcWidth, const double *txlen,
                        const int *degradation_bias, const int *maxiter, const double *minvalue, const double *conv,
                        int *txid_out, double *counts, double *full_length, double *unique_counts, int *n_result,
                        double *d_rate, int *rc);
/* ---- equivalence read classes from the distance table: genEquiRCs (R/bambu-quantify_utilityFunctions.R:46-192), native
 * host code (csrc/equirc.cpp) -- the producer of the long table bambu_quantDT takes.  All ids are integers: the caller
 * codes GENEID (any coding, consistent across the three tables) and refers to a read class by its row in the read-class
 * table.  Annotation: per transcript txid, gene, its minimal equivalent class eqClassById (eqcls_val[eqcls_ptr[i] ..
 * eqcls_ptr[i+1]), sorted txids) and its length (sum of exon widths).  Read classes (rowData order): gene, width of the
 * exon with exon_rank 1, total width.  Distance table: read class row, txid, readCount, equal, gene, and a flag for the
 * rows whose annotationTxId is "<gene>_unidentified" (dropped, :84).  Output rows in the reference's order: gene, txid,
 * eqClassId (cur_group_id, 1-based), nobs, equal, rcWidth, txlen, minRC.  Borrowed pointers, valid until _free. */
typedef struct bambu_equirc bambu_equirc;
const char *bambu_equirc_last_error(void);
int bambu_equirc_create(bambu_equirc **out, int64_t n_tx, const int64_t *ann_txid, const int64_t *ann_gene,
                        const int64_t *eqcls_ptr, const int64_t *eqcls_val, const double *ann_txlen, int64_t n_rc,
                        const int64_t *rc_gene, const double *rc_first_exon_width, const double *rc_total_width,
                        int64_t n_dt, const int64_t *dt_rc, const int64_t *dt_txid, const double *dt_read_count,
                        const uint8_t *dt_equal, const int64_t *dt_gene, const uint8_t *dt_unidentified);
int bambu_equirc_free(bambu_equirc *p);
int bambu_equirc_size(const bambu_equirc *p, int64_t *n_rows);
int bambu_equirc_table(const bambu_equirc *p, const int64_t **gene, const int64_t **txid, const int64_t **eqClassId,
                       const double **nobs, const uint8_t **equal, const double **rcWidth, const double **txlen,
                       const int64_t **minRC);

/* ---- the quantification stage of a whole cohort: bplapply(readClassList, bambu.quantify) + combineCountSes ------
 * (R/bambu.R:187-192, R/bambu-quantify.R:22-38, R/bambu_utilityFunctions.R:215-241) for the part after genEquiRCs.
 * Per sample s: the long table (n_rows[s] rows; column pointers [s], as bambu_quantDT takes them; equal / rcWidth /
 * txlen may be NULL arrays) is copied to the GPU and packed there, the EM runs on the arena in HBM, and the epilogue --
 * modifyQuantOut, removeDuplicates, calculateCPM (denominator = sum(counts) + incompatible_total[s], R's long double
 * sum, result-row order), counts[match(annotation_txid, txid)], round(., sig_digit) of counts / CPM / fullLengthCounts
 * -- runs on the device and writes the sample's column of the four assays.  Samples are spread over a few worker
 * threads (BAMBU_COHORT_WORKERS, default 4) so that copies, packing and EM of different samples overlap.
 *
 *   assays         HOST, 4 * n_samples * n_tx doubles (may be NULL): assays[(k * n_samples + s) * n_tx + a], k = 0 counts,
 *                  1 CPM, 2 fullLengthCounts, 3 uniqueCounts -- i.e. four column-major n_tx x n_samples R matrices;
 *                  annotation transcripts the sample's table never mentions are NA (NaN), as match() leaves them
 *   assays_device  the same block in DEVICE memory (may be NULL): what a multi-GPU run gathers
 *   d_rate         [n_samples] degradation rate of each sample (may be NULL)
 *   used_device    [n_samples] 1, or 0 for a sample that went through the host packer (see bambu_quantDT_hostpack)
 * Table buffers may be pageable; page-locked ones (cudaHostRegister / cudaMallocHost) copy about 5x faster. */
int bambu_quantDT_cohort(int32_t n_samples, const int64_t *n_rows, const int64_t *const *gene_code,
                         const int64_t *const *txid, const int64_t *const *eqClassId, const double *const *nobs,
                         const uint8_t *const *equal, const double *const *rcWidth, const double *const *txlen,
                         int32_t degradation_bias, int32_t maxiter, double minvalue, double conv, int64_t n_tx,
                         const int64_t *annotation_txid, const double *incompatible_total, int32_t sig_digit,
                         double *assays, double *assays_device, double *d_rate, int32_t *used_device);
/* bambu_quant_pack_create done by the GPU (csrc/quant_dev.cu), the arena copied back into a host pack object.
 * used_device (may be NULL) receives 1, or 0 when the table went through the host packer (see above). */
int bambu_quant_pack_create_gpu(bambu_quant_pack **pack, int64_t n_rows, const int64_t *gene_code, const int64_t *txid,
                                const int64_t *eqClassId, const double *nobs, const uint8_t *equal, const double *rcWidth,
                                const double *txlen, int32_t degradation_bias, int32_t *used_device);

#ifdef __cplusplus
}
#endif
#endif /* BAMBU_EM_H */
