// quant_cohort.inl -- the quantification stage of a whole cohort in one call (included by quant_dev.cu).
//
// Reference: bambu() runs bambu.quantify once per sample under bplapply (R/bambu.R:187-191) and cbinds the
// per-sample SummarizedExperiments (combineCountSes, R/bambu_utilityFunctions.R:215-241).  Here one call takes the
// long tables of all samples and fills the four n_tx x n_samples assays:
//
//   per sample, on a worker thread with its own device workspace, streams and EM plan (so that the table copy and
//   the device pack of one sample overlap the EM of another):
//     bambu.quantDT      table -> GPU -> device pack (quant_dev.inl) -> EM on the arena in HBM (bambu_em_plan_*)
//     modifyQuantOut + rbind(unobserved) + removeDuplicates  (R/bambu-quantify_utilityFunctions.R:438-455)
//                        device: txid -> result-row table, the LAST arena row of a txid wins, unobserved genes' rows are 0
//     calculateCPM       (:493-497) denominator = sum(counts) + sum(incompatible): R's sum() is a long double
//                        accumulation in result-row order, so the counts column (one D2H of n_out doubles) is summed by
//                        the worker in that order on the host FPU -- the only host arithmetic of the epilogue
//     counts[match(annotation txid, txid)], round(., sig.digit) except uniqueCounts  (R/bambu-quantify.R:27-37)
//                        device: one kernel over the annotation rows writes the sample's column of all four assays
//   Transcripts of the annotation that the sample's table never mentions come out NA (NaN), as match() leaves them.
#include <mutex>

namespace cohort {

using qdcuda::CudaError;

struct DBuf {          // grow-only device buffer
    void *p = nullptr;
    size_t cap = 0;
    void ensure(size_t bytes)
    {
        if (bytes <= cap) return;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        const size_t want = bytes + bytes / 4 + 256;
        if (cudaMalloc(&p, want) != cudaSuccess) { cudaGetLastError(); throw std::bad_alloc(); }
        cap = want;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T *as() const { return static_cast<T *>(p); }
};

struct Worker {
    qdcuda::Workspace ws;
    bambu_em_plan *plan = nullptr;
    DBuf code, winner, vals, d_out, d_unobs, cols;      // merge scratch
    double *h_v0 = nullptr;                             // pinned: the counts column in result-row order
    size_t h_v0_cap = 0;
    void ensure_host(size_t n)
    {
        if (n <= h_v0_cap) return;
        if (h_v0) cudaFreeHost(h_v0);
        h_v0 = nullptr; h_v0_cap = 0;
        if (cudaMallocHost((void **)&h_v0, (n + n / 4 + 64) * sizeof(double)) != cudaSuccess) { cudaGetLastError(); throw std::bad_alloc(); }
        h_v0_cap = n + n / 4 + 64;
    }
    void release()
    {
        if (plan) bambu_em_plan_destroy(plan);
        plan = nullptr;
        for (DBuf *b : {&code, &winner, &vals, &d_out, &d_unobs, &cols}) b->release();
        if (h_v0) cudaFreeHost(h_v0);
        h_v0 = nullptr; h_v0_cap = 0;
        ws.release();
    }
};

static std::vector<Worker *> g_workers;
static DBuf g_ann;                 // annotation txids on the device
static std::mutex g_fallback_mutex;

// ---- kernels ------------------------------------------------------------------------------------------------------
// code table over [lo, lo + range): 0 = txid not in the sample's result, 1 = a row of an unobserved gene (zeros),
// 2 + q = result row q of initialiseOutput's order
__global__ void __launch_bounds__(256) k_code_unobs(int32_t *code, const int64_t *unobs, size_t nu, int64_t lo)
{
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < nu; i += (size_t)gridDim.x * 256) code[unobs[i] - lo] = 1;
}
__global__ void __launch_bounds__(256) k_code_out(int32_t *code, const int64_t *out, size_t no, int64_t lo, int32_t *winner)
{
    for (size_t q = (size_t)blockIdx.x * 256 + threadIdx.x; q < no; q += (size_t)gridDim.x * 256) {
        code[out[q] - lo] = 2 + (int32_t)q;
        winner[q] = -1;
    }
}
// modifyQuantOut: outIni[match(txid_est, txid)] := est, row by row -- a txid estimated in several groups keeps the
// value of its LAST arena row
__global__ void __launch_bounds__(256) k_winner(const int32_t *code, const int32_t *row_txid, size_t rows, int64_t lo, int32_t *winner)
{
    for (size_t r = (size_t)blockIdx.x * 256 + threadIdx.x; r < rows; r += (size_t)gridDim.x * 256) {
        const int32_t c = code[(int64_t)row_txid[r] - lo];
        if (c >= 2) atomicMax(&winner[c - 2], (int32_t)r);
    }
}
__global__ void __launch_bounds__(256) k_values(const int32_t *winner, size_t no, const double *est, size_t rows, double *vals)
{
    for (size_t q = (size_t)blockIdx.x * 256 + threadIdx.x; q < no; q += (size_t)gridDim.x * 256) {
        const int32_t r = winner[q];
        vals[q] = r >= 0 ? est[r] : 0.0;
        vals[no + q] = r >= 0 ? est[rows + r] : 0.0;
        vals[2 * no + q] = r >= 0 ? est[2 * rows + r] : 0.0;
    }
}
// the sample's column of the four assays, annotation order (R/bambu-quantify.R:27-37)
__global__ void __launch_bounds__(256) k_assays(const int32_t *code, int64_t lo, int64_t range, const int64_t *ann_txid, size_t n_tx,
                                                const double *vals, size_t no, double total, int digits, double *c_counts,
                                                double *c_cpm, double *c_full, double *c_unique)
{
    const double na = __longlong_as_double(0x7ff8000000000000ll);
    for (size_t a = (size_t)blockIdx.x * 256 + threadIdx.x; a < n_tx; a += (size_t)gridDim.x * 256) {
        const int64_t t = ann_txid[a] - lo;
        const int32_t c = (t >= 0 && t < range) ? code[t] : 0;
        double v0 = na, v1 = na, v2 = na, cpm = na;
        if (c >= 1) {
            v0 = v1 = v2 = 0.0;
            if (c >= 2) { v0 = vals[c - 2]; v1 = vals[no + (c - 2)]; v2 = vals[2 * no + (c - 2)]; }
            cpm = __dmul_rn(__ddiv_rn(v0, total), 1000000.0);       // counts / totalCount * (10^6)
        }
        c_counts[a] = qd::r_round(v0, digits);
        c_cpm[a] = qd::r_round(cpm, digits);
        c_full[a] = qd::r_round(v1, digits);
        c_unique[a] = v2;                                            // not rounded (:36-37)
    }
}

static inline int grid_for(size_t n, int sms) { const size_t b = (n + 255) / 256; return (int)std::max<size_t>(1, std::min<size_t>(b, (size_t)sms * 8)); }

struct Args {
    int32_t n_samples;
    const int64_t *n_rows;
    const int64_t *const *gene, *const *txid, *const *eqc;
    const double *const *nobs;
    const uint8_t *const *equal;
    const double *const *rcw, *const *txl;
    int32_t bias, maxiter;
    double minvalue, conv;
    int64_t n_tx;
    const int64_t *ann_txid;
    const double *incompatible_total;
    int32_t sig_digit;
    double *assays_host, *assays_dev, *d_rate;
    int32_t *used_device;
};

// the host arithmetic of the epilogue: R's sum() -- long double accumulation in row order, rounded to double at the end
static double r_sum(const double *v, size_t n)
{
    long double s = 0.0L;
    for (size_t i = 0; i < n; ++i) s += (long double)v[i];
    return (double)s;
}

// a sample the device formulation does not cover (non-integer counts ...): host packer + EM + merge, host epilogue
static int sample_on_host(const Args &A, int s, double *col /* 4 x n_tx, host */, std::string &err)
{
    std::lock_guard<std::mutex> lock(g_fallback_mutex);       // bambu_em_batched's cached pipeline is not reentrant
    const size_t n = (size_t)A.n_rows[s];
    std::vector<int64_t> t_out(std::max<size_t>(n, 1));
    std::vector<double> c0(std::max<size_t>(n, 1)), c1(c0.size()), c2(c0.size());
    int64_t k = 0;
    double dr = NAN;
    int rc = bambu_quantDT_hostpack((int64_t)n, A.gene[s], A.txid[s], A.eqc[s], A.nobs[s], A.equal ? A.equal[s] : nullptr,
                                    A.rcw ? A.rcw[s] : nullptr, A.txl ? A.txl[s] : nullptr, A.bias, A.maxiter, A.minvalue, A.conv,
                                    t_out.data(), c0.data(), c1.data(), c2.data(), &k, &dr);
    if (rc) { err = bambu_quant_last_error(); return rc; }
    if (A.d_rate) A.d_rate[s] = dr;
    const double total = r_sum(c0.data(), (size_t)k) + (A.incompatible_total ? A.incompatible_total[s] : 0.0);
    std::unordered_map<int64_t, size_t> at;
    at.reserve((size_t)k * 2 + 16);
    for (int64_t q = 0; q < k; ++q) at.emplace(t_out[(size_t)q], (size_t)q);
    const size_t T = (size_t)A.n_tx;
    for (size_t a = 0; a < T; ++a) {
        auto it = at.find(A.ann_txid[a]);
        if (it == at.end()) { col[a] = col[T + a] = col[2 * T + a] = col[3 * T + a] = NAN; continue; }
        const size_t q = it->second;
        col[a] = qd::r_round(c0[q], A.sig_digit);
        col[T + a] = qd::r_round(c0[q] / total * 1000000.0, A.sig_digit);
        col[2 * T + a] = qd::r_round(c1[q], A.sig_digit);
        col[3 * T + a] = c2[q];
    }
    return BAMBU_EM_OK;
}

static int sample_on_device(Worker &w, const Args &A, int s, int64_t alo, int64_t ahi, std::string &err, bool &fallback)
{
    fallback = false;
    const int64_t n = A.n_rows[s];
    if (!w.ws.ensure(qdcuda::workspace_bytes((size_t)n))) { err = "not enough device memory for the packer workspace"; return BAMBU_EM_ENOMEM; }
    qdcuda::CudaExec x(w.ws.st, w.ws.pool, w.ws.cap, w.ws.sms);
    qd::Result R;
    int rc = run_device_pack(x, n, A.gene[s], A.txid[s], A.eqc[s], A.nobs[s], A.equal ? A.equal[s] : (const uint8_t *)nullptr,
                             A.rcw ? A.rcw[s] : nullptr, A.txl ? A.txl[s] : nullptr, A.bias != 0, R, w.ws);
    if (rc == 1) { fallback = true; return BAMBU_EM_OK; }
    if (rc) { err = bambu_quant_last_error(); return rc; }
    if (A.d_rate) A.d_rate[s] = R.d_rate;
    cudaStream_t st = w.ws.st;
    const size_t T = (size_t)A.n_tx, no = R.out_txid.size(), nu = R.unobserved_txid.size(), rows = (size_t)R.n_rows;
    // ---- EM on the arena where the packer left it ----------------------------------------------------------------
    const double *d_est = nullptr;
    if (R.n_groups) {
        std::vector<int64_t> grp, gcp, rnp;
        fetch(x, grp, R.grp_row_ptr, (size_t)R.n_groups + 1);
        fetch(x, gcp, R.grp_col_ptr, (size_t)R.n_groups + 1);
        fetch(x, rnp, R.row_nnz_ptr, rows + 1);
        rc = w.plan ? bambu_em_plan_reset(w.plan, R.n_groups, grp.data(), gcp.data(), rnp.data())
                    : bambu_em_plan_create(&w.plan, R.n_groups, grp.data(), gcp.data(), rnp.data());
        if (!rc) rc = bambu_em_plan_set_stream(w.plan, (void *)st);
        if (!rc) rc = bambu_em_plan_bind_device(w.plan, R.col_idx, R.aval, R.nz_flags, R.Y, R.K, R.col_flags);
        if (!rc) rc = bambu_em_plan_run(w.plan, A.maxiter, A.minvalue, A.conv);
        if (!rc) rc = bambu_em_plan_sync(w.plan);
        double *e = nullptr;
        if (!rc) rc = bambu_em_plan_device_outputs(w.plan, &e, nullptr, nullptr);
        if (rc) { err = bambu_em_last_error(); if (w.plan) bambu_em_plan_destroy(w.plan); w.plan = nullptr; return rc; }
        d_est = e;
    }
    // ---- merge by txid on the device -------------------------------------------------------------------------------
    int64_t lo = alo, hi = ahi;
    for (int64_t t : R.out_txid) { lo = std::min(lo, t); hi = std::max(hi, t); }
    for (int64_t t : R.unobserved_txid) { lo = std::min(lo, t); hi = std::max(hi, t); }
    if (hi < lo) { lo = 0; hi = 0; }
    const uint64_t range = (uint64_t)hi - (uint64_t)lo + 1;
    if (range > ((uint64_t)1 << 28)) { fallback = true; return BAMBU_EM_OK; }      // sparse txid space: the host merge handles it
    w.code.ensure(4 * (size_t)range);
    w.winner.ensure(4 * std::max<size_t>(no, 1));
    w.vals.ensure(8 * 3 * std::max<size_t>(no, 1));
    w.d_out.ensure(8 * std::max<size_t>(no, 1));
    w.d_unobs.ensure(8 * std::max<size_t>(nu, 1));
    w.ensure_host(std::max<size_t>(no, 1));
    QD_CUDA(cudaMemsetAsync(w.code.p, 0, 4 * (size_t)range, st));
    if (nu) {
        QD_CUDA(cudaMemcpyAsync(w.d_unobs.p, R.unobserved_txid.data(), 8 * nu, cudaMemcpyHostToDevice, st));
        k_code_unobs<<<grid_for(nu, w.ws.sms), 256, 0, st>>>(w.code.as<int32_t>(), w.d_unobs.as<int64_t>(), nu, lo);
    }
    if (no) {
        QD_CUDA(cudaMemcpyAsync(w.d_out.p, R.out_txid.data(), 8 * no, cudaMemcpyHostToDevice, st));
        k_code_out<<<grid_for(no, w.ws.sms), 256, 0, st>>>(w.code.as<int32_t>(), w.d_out.as<int64_t>(), no, lo, w.winner.as<int32_t>());
        if (rows) k_winner<<<grid_for(rows, w.ws.sms), 256, 0, st>>>(w.code.as<int32_t>(), R.txid, rows, lo, w.winner.as<int32_t>());
        k_values<<<grid_for(no, w.ws.sms), 256, 0, st>>>(w.winner.as<int32_t>(), no, d_est, rows, w.vals.as<double>());
        QD_CUDA(cudaMemcpyAsync(w.h_v0, w.vals.p, 8 * no, cudaMemcpyDeviceToHost, st));
    }
    QD_CUDA(cudaStreamSynchronize(st));
    // calculateCPM's denominator (R/bambu-quantify_utilityFunctions.R:494): the unobserved rows come first and are zeros
    const double total = r_sum(w.h_v0, no) + (A.incompatible_total ? A.incompatible_total[s] : 0.0);
    double *col;
    if (A.assays_dev) col = nullptr;
    else { w.cols.ensure(8 * 4 * std::max<size_t>(T, 1)); col = w.cols.as<double>(); }
    const size_t S = (size_t)A.n_samples;
    double *c[4];
    for (int k = 0; k < 4; ++k) c[k] = A.assays_dev ? A.assays_dev + ((size_t)k * S + (size_t)s) * T : col + (size_t)k * T;
    if (T) {
        k_assays<<<grid_for(T, w.ws.sms), 256, 0, st>>>(w.code.as<int32_t>(), lo, (int64_t)range, g_ann.as<int64_t>(), T, w.vals.as<double>(),
                                                         no, total, A.sig_digit, c[0], c[1], c[2], c[3]);
        QD_CUDA(cudaGetLastError());
        if (A.assays_host)
            for (int k = 0; k < 4; ++k)
                QD_CUDA(cudaMemcpyAsync(A.assays_host + ((size_t)k * S + (size_t)s) * T, c[k], 8 * T, cudaMemcpyDeviceToHost, st));
    }
    QD_CUDA(cudaStreamSynchronize(st));       // the workspace (arena, row txids) is reused by this worker's next sample
    if (A.used_device) A.used_device[s] = 1;
    return BAMBU_EM_OK;
}

}  // namespace cohort

static void cohort_release()
{
    for (cohort::Worker *w : cohort::g_workers) { w->release(); delete w; }
    cohort::g_workers.clear();
    cohort::g_ann.release();
}

extern "C" int bambu_quantDT_cohort(int32_t n_samples, const int64_t *n_rows, const int64_t *const *gene_code,
                                    const int64_t *const *txid, const int64_t *const *eqClassId, const double *const *nobs,
                                    const uint8_t *const *equal, const double *const *rcWidth, const double *const *txlen,
                                    int32_t degradation_bias, int32_t maxiter, double minvalue, double conv, int64_t n_tx,
                                    const int64_t *annotation_txid, const double *incompatible_total, int32_t sig_digit,
                                    double *assays, double *assays_device, double *d_rate, int32_t *used_device)
{
    using namespace cohort;
    if (n_samples < 0 || n_tx < 0 || (n_samples && (!n_rows || !gene_code || !txid || !eqClassId || !nobs)) || (n_tx && !annotation_txid) ||
        (!assays && !assays_device && n_samples && n_tx)) {
        bambu_quant_set_error("bambu_quantDT_cohort: NULL or negative argument");
        return BAMBU_EM_EINVAL;
    }
    for (int s = 0; s < n_samples; ++s) {
        if (n_rows[s] < 0 || n_rows[s] > (int64_t)UINT32_MAX - 8) { bambu_quant_set_error("row count out of range"); return BAMBU_EM_EINVAL; }
        if (n_rows[s] && (!gene_code[s] || !txid[s] || !eqClassId[s] || !nobs[s])) {
            bambu_quant_set_error("Columns GENEID, txid, eqClassId, nobs, are missing from object.");
            return BAMBU_EM_EINVAL;
        }
    }
    int rc = bambu_em_internal_ensure_device();
    if (rc) { bambu_quant_set_error(bambu_em_last_error()); return rc; }
    int device = 0;
    cudaGetDevice(&device);
    Args A = {n_samples, n_rows, gene_code, txid, eqClassId, nobs, equal, rcWidth, txlen, degradation_bias, maxiter, minvalue, conv,
              n_tx, annotation_txid, incompatible_total, sig_digit, assays, assays_device, d_rate, used_device};
    int n_workers = 4;
    if (const char *e = getenv("BAMBU_COHORT_WORKERS")) n_workers = atoi(e);
    n_workers = std::max(1, std::min(std::min(n_workers, 8), std::max(n_samples, 1)));
    std::string first_error;
    int first_rc = BAMBU_EM_OK;
    try {
        while ((int)g_workers.size() < n_workers) g_workers.push_back(new Worker());
        int64_t alo = INT64_MAX, ahi = INT64_MIN;
        for (int64_t a = 0; a < n_tx; ++a) { alo = std::min(alo, annotation_txid[a]); ahi = std::max(ahi, annotation_txid[a]); }
        if (n_tx) {
            g_ann.ensure(8 * (size_t)n_tx);
            QD_CUDA(cudaMemcpy(g_ann.p, annotation_txid, 8 * (size_t)n_tx, cudaMemcpyHostToDevice));
        }
        {   // kernel attributes and other process-wide state are set up once, before the workers start
            const int64_t z[1] = {0};
            bambu_em_plan *warm = nullptr;
            rc = bambu_em_plan_create(&warm, 0, z, z, z);
            if (warm) bambu_em_plan_destroy(warm);
            if (rc) { bambu_quant_set_error(bambu_em_last_error()); return rc; }
        }
        std::atomic<int> next{0};
        std::mutex err_mutex;
        auto body = [&](int wi) {
            Worker &w = *g_workers[(size_t)wi];
            cudaSetDevice(device);
            std::vector<double> host_col;
            for (;;) {
                const int s = next.fetch_add(1);
                if (s >= n_samples) break;
                { std::lock_guard<std::mutex> l(err_mutex); if (first_rc) break; }
                std::string err;
                int r = BAMBU_EM_OK;
                bool fallback = false;
                if (used_device) used_device[s] = 0;
                try {
                    r = sample_on_device(w, A, s, alo, ahi, err, fallback);
                    if (!r && fallback) {
                        const size_t T = (size_t)n_tx, S = (size_t)n_samples;
                        host_col.assign(4 * std::max<size_t>(T, 1), 0.0);
                        r = sample_on_host(A, s, host_col.data(), err);
                        for (int k = 0; k < 4 && !r && T; ++k) {
                            if (assays) memcpy(assays + ((size_t)k * S + (size_t)s) * T, host_col.data() + (size_t)k * T, 8 * T);
                            if (assays_device) {
                                cudaError_t ce = cudaMemcpy(assays_device + ((size_t)k * S + (size_t)s) * T, host_col.data() + (size_t)k * T, 8 * T, cudaMemcpyHostToDevice);
                                if (ce != cudaSuccess) { r = BAMBU_EM_ECUDA; err = cudaGetErrorString(ce); }
                            }
                        }
                    }
                } catch (const CudaError &e) {
                    r = BAMBU_EM_ECUDA;
                    err = std::string("CUDA error in the cohort stage: ") + cudaGetErrorString(e.e) + " (" + e.what + ")";
                } catch (const std::bad_alloc &) {
                    r = BAMBU_EM_ENOMEM;
                    err = "out of memory in the cohort stage";
                } catch (const std::exception &e) {
                    r = BAMBU_EM_ECUDA;
                    err = std::string("cohort stage: ") + e.what();
                } catch (...) {
                    r = BAMBU_EM_ECUDA;
                    err = "cohort stage: unknown exception";
                }
                if (r) {
                    std::lock_guard<std::mutex> l(err_mutex);
                    if (!first_rc) { first_rc = r; first_error = "sample " + std::to_string(s) + ": " + err; }
                    break;
                }
            }
        };
        std::vector<std::thread> threads;
        for (int wi = 1; wi < n_workers; ++wi) {
            try { threads.emplace_back(body, wi); } catch (const std::system_error &) { break; }      // fewer workers, same result
        }
        body(0);
        for (auto &t : threads) t.join();
    } catch (const CudaError &e) {
        first_rc = BAMBU_EM_ECUDA;
        first_error = std::string("CUDA error in the cohort stage: ") + cudaGetErrorString(e.e) + " (" + e.what + ")";
    } catch (const std::bad_alloc &) {
        first_rc = BAMBU_EM_ENOMEM;
        first_error = "out of memory in the cohort stage";
    } catch (...) {
        first_rc = BAMBU_EM_ECUDA;
        first_error = "cohort stage: unknown exception";
    }
    if (first_rc) bambu_quant_set_error(first_error);
    return first_rc;
}
