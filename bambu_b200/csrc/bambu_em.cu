// bambu_em.cu -- host side of libbambu_em.so: the C ABI of include/bambu_em.h,
// the size-bucket scheduler and the kernel launches.  No torch, no CPU fallback.
//
// Replaces the serial lapply over gene groups of abundance_quantification()
// (R/bambu-quantify_utilityFunctions.R:379-391) by, per size bucket, one pack
// launch + one persistent EM launch on its own stream.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/bambu_em.h"
#include "common.cuh"
#include "em_resident.cuh"
#include "pack_kernel.cuh"

using namespace bambu;

// ------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------
static thread_local std::string g_last_error;

static int fail(int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

#define CUDA_TRY(expr)                                                                        \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess)                                                                \
            return fail(_e == cudaErrorMemoryAllocation ? BAMBU_EM_ENOMEM : BAMBU_EM_ECUDA,   \
                        "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,     \
                        __LINE__);                                                            \
    } while (0)

static int g_device = 0;
static bool g_device_checked = false;
static int g_num_sms = 0;

static int ensure_device()
{
    if (g_device_checked) {
        CUDA_TRY(cudaSetDevice(g_device));
        return BAMBU_EM_OK;
    }
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0)
        return fail(BAMBU_EM_ENODEVICE, "no CUDA device available (%s); libbambu_em has no CPU path",
                    e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    if (g_device >= n) return fail(BAMBU_EM_ENODEVICE, "device %d requested, %d present", g_device, n);
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, g_device));
    if (prop.major != 10)
        return fail(BAMBU_EM_ENODEVICE, "device %d is sm_%d%d; this library is built for sm_100a only",
                    g_device, prop.major, prop.minor);
    CUDA_TRY(cudaSetDevice(g_device));
    g_num_sms = prop.multiProcessorCount;
    g_device_checked = true;
    return BAMBU_EM_OK;
}

// ------------------------------------------------------------------------------
// size buckets
// ------------------------------------------------------------------------------
// A bucket = groups whose worst-case shared-memory footprint falls in one class.
// The class fixes the CTA size and the dynamic shared memory of both kernels, and
// with it how many groups are resident per SM (227 KB / footprint).
struct BucketClass {
    uint32_t max_bytes;   // upper bound of em_smem_bytes for the class
    int threads;
};
static const BucketClass kClasses[] = {
    {7u * 1024u, 64},     // 32 CTAs/SM (the CTA limit)
    {14u * 1024u, 64},    // 16
    {28u * 1024u, 128},   // 8
    {56u * 1024u, 128},   // 4
    {113u * 1024u, 256},  // 2
    {kSmemMax, 512},      // 1
};
static const int kNumClasses = (int)(sizeof(kClasses) / sizeof(kClasses[0]));

struct Bucket {
    int cls = 0;
    int threads = 0;
    uint32_t smem_em = 0, smem_pack = 0;
    std::vector<int32_t> work;     // group ids, largest first
    int32_t *d_work = nullptr;     // view into plan->d_work_all
    int *d_ticket = nullptr;
    int grid_em = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_pack0 = nullptr, ev_pack1 = nullptr, ev_em1 = nullptr;
};

struct bambu_em_plan {
    int64_t n_groups = 0, total_rows = 0, total_cols = 0, nnz = 0;
    // device: structure
    int64_t *d_grp_row_ptr = nullptr, *d_grp_col_ptr = nullptr, *d_row_nnz_ptr = nullptr;
    // device: payload (owned unless bound)
    bool payload_owned = false, payload_ready = false;
    void *d_payload = nullptr;
    const int32_t *d_col_idx = nullptr;
    const double *d_aval = nullptr;
    const uint8_t *d_nz_flags = nullptr;
    const double *d_Y = nullptr, *d_K = nullptr;
    const uint8_t *d_col_flags = nullptr;
    // device: images + outputs
    uint8_t *d_img = nullptr;
    int64_t img_bytes = 0;
    int64_t *d_img_off = nullptr;
    PackInfo *d_info = nullptr;
    int64_t *d_nnz_nonzero = nullptr;
    double *d_est = nullptr;
    int32_t *d_iters = nullptr, *d_status = nullptr;
    int32_t *d_work_all = nullptr;
    int *d_ctrl = nullptr;   // [0] error word, [1..] tickets
    int n_ctrl = 0;
    std::vector<Bucket> buckets;
    std::vector<int32_t> single_tx;   // M == 1 groups: closed form in the pack kernel only
    cudaStream_t own_stream = nullptr, stream = nullptr;
    cudaEvent_t ev_start = nullptr, ev_fork = nullptr, ev_end = nullptr;
    bool ran = false;
    bool single_tx_shortcut = true;   // R/bambu-quantify_utilityFunctions.R:410-414 (run_parallel, not emWithL1)
    int last_launches = 0;
    int64_t device_bytes = 0;
};

template <typename Tp>
static int dev_alloc(bambu_em_plan *p, Tp **ptr, size_t count)
{
    size_t bytes = std::max<size_t>(count * sizeof(Tp), 16);
    CUDA_TRY(cudaMalloc((void **)ptr, bytes));
    p->device_bytes += (int64_t)bytes;
    return BAMBU_EM_OK;
}

// ------------------------------------------------------------------------------
// kernel dispatch by CTA size
// ------------------------------------------------------------------------------
template <int T>
static int launch_bucket_T(bambu_em_plan *p, Bucket &b, const ArenaDev &A, const EmParams &P)
{
    auto pack = pack_resident_kernel<T>;
    auto em = em_resident_kernel<T>;
    CUDA_TRY(cudaFuncSetAttribute(pack, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b.smem_pack));
    CUDA_TRY(cudaFuncSetAttribute(em, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b.smem_em));
    if (b.grid_em == 0) {
        int per_sm = 0;
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, em, T, b.smem_em));
        if (per_sm < 1) return fail(BAMBU_EM_ECUDA, "EM kernel does not fit an SM (smem %u)", b.smem_em);
        b.grid_em = (int)std::min<int64_t>((int64_t)b.work.size(), (int64_t)per_sm * g_num_sms);
    }
    const int n_work = (int)b.work.size();
    CUDA_TRY(cudaEventRecord(b.ev_pack0, b.stream));
    pack<<<n_work, T, b.smem_pack, b.stream>>>(A, b.d_work, n_work, p->d_img_off, p->d_img, p->d_info,
                                                p->d_nnz_nonzero, p->d_est, p->d_status, p->d_ctrl,
                                                p->single_tx_shortcut ? 1 : 0);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(b.ev_pack1, b.stream));
    em<<<b.grid_em, T, b.smem_em, b.stream>>>(b.d_work, n_work, b.d_ticket, p->d_img_off, p->d_img, p->d_info,
                                              p->d_grp_row_ptr, p->total_rows, P, p->d_est, p->d_iters,
                                              p->d_status);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(b.ev_em1, b.stream));
    p->last_launches += 2;
    return BAMBU_EM_OK;
}

static int launch_bucket(bambu_em_plan *p, Bucket &b, const ArenaDev &A, const EmParams &P)
{
    switch (b.threads) {
    case 64: return launch_bucket_T<64>(p, b, A, P);
    case 128: return launch_bucket_T<128>(p, b, A, P);
    case 256: return launch_bucket_T<256>(p, b, A, P);
    case 512: return launch_bucket_T<512>(p, b, A, P);
    }
    return fail(BAMBU_EM_EINVAL, "no kernel instance for %d threads", b.threads);
}

// ------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------
extern "C" {

const char *bambu_em_version(void) { return "bambu_em_b200 0.1.0 (sm_100a)"; }
const char *bambu_em_last_error(void) { return g_last_error.c_str(); }

int bambu_em_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    int ok = 0;
    for (int d = 0; d < n; ++d) {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, d) == cudaSuccess && prop.major == 10) ++ok;
    }
    return ok;
}

int bambu_em_set_device(int device)
{
    if (device < 0) return fail(BAMBU_EM_EINVAL, "negative device index");
    g_device = device;
    g_device_checked = false;
    return ensure_device();
}

int bambu_em_plan_destroy(bambu_em_plan *p)
{
    if (!p) return BAMBU_EM_OK;
    cudaSetDevice(g_device);
    if (p->stream) cudaStreamSynchronize(p->stream);
    for (auto &b : p->buckets) {
        if (b.stream) { cudaStreamSynchronize(b.stream); cudaStreamDestroy(b.stream); }
        if (b.ev_pack0) cudaEventDestroy(b.ev_pack0);
        if (b.ev_pack1) cudaEventDestroy(b.ev_pack1);
        if (b.ev_em1) cudaEventDestroy(b.ev_em1);
    }
    if (p->ev_start) cudaEventDestroy(p->ev_start);
    if (p->ev_fork) cudaEventDestroy(p->ev_fork);
    if (p->ev_end) cudaEventDestroy(p->ev_end);
    if (p->own_stream) cudaStreamDestroy(p->own_stream);
    cudaFree(p->d_grp_row_ptr); cudaFree(p->d_grp_col_ptr); cudaFree(p->d_row_nnz_ptr);
    if (p->payload_owned) cudaFree(p->d_payload);
    cudaFree(p->d_img); cudaFree(p->d_img_off); cudaFree(p->d_info); cudaFree(p->d_nnz_nonzero);
    cudaFree(p->d_est); cudaFree(p->d_iters); cudaFree(p->d_status); cudaFree(p->d_work_all);
    cudaFree(p->d_ctrl);
    delete p;
    return BAMBU_EM_OK;
}

static int plan_create_impl(bambu_em_plan **out, int64_t n_groups, const int64_t *grp_row_ptr,
                            const int64_t *grp_col_ptr, const int64_t *row_nnz_ptr, bool single_tx_shortcut)
{
    if (!out) return fail(BAMBU_EM_EINVAL, "plan pointer is NULL");
    *out = nullptr;
    if (n_groups < 0 || n_groups > INT32_MAX) return fail(BAMBU_EM_EINVAL, "n_groups out of range");
    if (!grp_row_ptr || !grp_col_ptr || !row_nnz_ptr) return fail(BAMBU_EM_EINVAL, "NULL pointer array");
    if (grp_row_ptr[0] != 0 || grp_col_ptr[0] != 0 || row_nnz_ptr[0] != 0)
        return fail(BAMBU_EM_EINVAL, "pointer arrays must start at 0");
    int rc = ensure_device();
    if (rc) return rc;

    bambu_em_plan *p = new (std::nothrow) bambu_em_plan;
    if (!p) return fail(BAMBU_EM_ENOMEM, "out of host memory");
    p->n_groups = n_groups;
    p->single_tx_shortcut = single_tx_shortcut;
    p->total_rows = grp_row_ptr[n_groups];
    p->total_cols = grp_col_ptr[n_groups];
    for (int64_t r = 0; r < p->total_rows; ++r)
        if (row_nnz_ptr[r + 1] < row_nnz_ptr[r]) { delete p; return fail(BAMBU_EM_EINVAL, "row_nnz_ptr decreases at row %lld", (long long)r); }
    p->nnz = row_nnz_ptr[p->total_rows];

    // ---- classify groups, lay out the images ----------------------------------
    std::vector<int64_t> img_off((size_t)n_groups + 1, 0);
    std::vector<std::vector<int32_t>> per_class(kNumClasses);
    std::vector<uint32_t> ub_em((size_t)n_groups, 0), ub_pack((size_t)n_groups, 0);
    std::vector<int64_t> gnnz((size_t)n_groups, 0);
    int64_t off = 0;
    for (int64_t g = 0; g < n_groups; ++g) {
        const int64_t M = grp_row_ptr[g + 1] - grp_row_ptr[g], N = grp_col_ptr[g + 1] - grp_col_ptr[g];
        if (M < 0 || N < 0) { delete p; return fail(BAMBU_EM_EINVAL, "group %lld has negative extent", (long long)g); }
        const int64_t e = row_nnz_ptr[grp_row_ptr[g + 1]] - row_nnz_ptr[grp_row_ptr[g]];
        gnnz[g] = e;
        img_off[g] = off;
        if (M == 0) continue;
        if (M == 1 && single_tx_shortcut) { p->single_tx.push_back((int32_t)g); continue; }
        const double ub = 20.0 * (double)e + 18.0 * (double)(M + N) + 64.0;
        if (M > kMaxIndex || N > kMaxIndex || ub > (double)kSmemMax) {
            delete p;
            return fail(BAMBU_EM_EINVAL,
                        "group %lld (M=%lld N=%lld entries=%lld) exceeds the shared-memory resident tier",
                        (long long)g, (long long)M, (long long)N, (long long)e);
        }
        const uint32_t bytes = em_smem_bytes((uint32_t)M, (uint32_t)N, (uint32_t)e);
        ub_em[g] = bytes;
        ub_pack[g] = pack_smem_bytes((uint32_t)M, (uint32_t)N);
        int c = 0;
        while (c < kNumClasses - 1 && bytes > kClasses[c].max_bytes) ++c;
        per_class[c].push_back((int32_t)g);
        off += img_layout((uint32_t)M, (uint32_t)N, (uint32_t)e).total;
    }
    img_off[n_groups] = off;
    p->img_bytes = off;

    // single-transcript groups ride in a bucket of their own (pack kernel only)
    auto fail_destroy = [&](int code) { bambu_em_plan_destroy(p); return code; };
#define PLAN_TRY(expr) do { int _rc = (expr); if (_rc) return fail_destroy(_rc); } while (0)
#define PLAN_CUDA(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { fail(_e == cudaErrorMemoryAllocation ? BAMBU_EM_ENOMEM : BAMBU_EM_ECUDA, "%s failed: %s", #expr, cudaGetErrorString(_e)); return fail_destroy(_e == cudaErrorMemoryAllocation ? BAMBU_EM_ENOMEM : BAMBU_EM_ECUDA); } } while (0)

    size_t n_work_total = p->single_tx.size();
    for (int c = 0; c < kNumClasses; ++c) {
        if (per_class[c].empty()) continue;
        Bucket b;
        b.cls = c;
        b.threads = kClasses[c].threads;
        b.work = std::move(per_class[c]);
        std::stable_sort(b.work.begin(), b.work.end(), [&](int32_t x, int32_t y) { return gnnz[x] > gnnz[y]; });
        for (int32_t g : b.work) {
            b.smem_em = std::max(b.smem_em, ub_em[g]);
            b.smem_pack = std::max(b.smem_pack, ub_pack[g]);
        }
        n_work_total += b.work.size();
        p->buckets.push_back(std::move(b));
    }

    PLAN_CUDA(cudaStreamCreateWithFlags(&p->own_stream, cudaStreamNonBlocking));
    p->stream = p->own_stream;
    PLAN_CUDA(cudaEventCreate(&p->ev_start));
    PLAN_CUDA(cudaEventCreateWithFlags(&p->ev_fork, cudaEventDisableTiming));
    PLAN_CUDA(cudaEventCreate(&p->ev_end));

    PLAN_TRY(dev_alloc(p, &p->d_grp_row_ptr, (size_t)n_groups + 1));
    PLAN_TRY(dev_alloc(p, &p->d_grp_col_ptr, (size_t)n_groups + 1));
    PLAN_TRY(dev_alloc(p, &p->d_row_nnz_ptr, (size_t)p->total_rows + 1));
    PLAN_TRY(dev_alloc(p, &p->d_img, (size_t)p->img_bytes));
    PLAN_TRY(dev_alloc(p, &p->d_img_off, (size_t)n_groups + 1));
    PLAN_TRY(dev_alloc(p, &p->d_info, (size_t)n_groups));
    PLAN_TRY(dev_alloc(p, &p->d_nnz_nonzero, (size_t)n_groups));
    PLAN_TRY(dev_alloc(p, &p->d_est, (size_t)3 * p->total_rows));
    PLAN_TRY(dev_alloc(p, &p->d_iters, (size_t)n_groups));
    PLAN_TRY(dev_alloc(p, &p->d_status, (size_t)n_groups));
    PLAN_TRY(dev_alloc(p, &p->d_work_all, n_work_total));
    p->n_ctrl = 1 + (int)p->buckets.size();
    PLAN_TRY(dev_alloc(p, &p->d_ctrl, (size_t)p->n_ctrl));

    cudaStream_t s = p->stream;
    PLAN_CUDA(cudaMemcpyAsync(p->d_grp_row_ptr, grp_row_ptr, sizeof(int64_t) * (n_groups + 1), cudaMemcpyHostToDevice, s));
    PLAN_CUDA(cudaMemcpyAsync(p->d_grp_col_ptr, grp_col_ptr, sizeof(int64_t) * (n_groups + 1), cudaMemcpyHostToDevice, s));
    PLAN_CUDA(cudaMemcpyAsync(p->d_row_nnz_ptr, row_nnz_ptr, sizeof(int64_t) * (p->total_rows + 1), cudaMemcpyHostToDevice, s));
    PLAN_CUDA(cudaMemcpyAsync(p->d_img_off, img_off.data(), sizeof(int64_t) * (n_groups + 1), cudaMemcpyHostToDevice, s));
    size_t woff = 0;
    int bi = 0;
    for (auto &b : p->buckets) {
        b.d_work = p->d_work_all + woff;
        b.d_ticket = p->d_ctrl + 1 + bi++;
        PLAN_CUDA(cudaMemcpyAsync(b.d_work, b.work.data(), sizeof(int32_t) * b.work.size(), cudaMemcpyHostToDevice, s));
        woff += b.work.size();
        PLAN_CUDA(cudaStreamCreateWithFlags(&b.stream, cudaStreamNonBlocking));
        PLAN_CUDA(cudaEventCreate(&b.ev_pack0));
        PLAN_CUDA(cudaEventCreate(&b.ev_pack1));
        PLAN_CUDA(cudaEventCreate(&b.ev_em1));
    }
    if (!p->single_tx.empty())
        PLAN_CUDA(cudaMemcpyAsync(p->d_work_all + woff, p->single_tx.data(), sizeof(int32_t) * p->single_tx.size(), cudaMemcpyHostToDevice, s));
    PLAN_CUDA(cudaStreamSynchronize(s));   // the host vectors above go out of scope
#undef PLAN_TRY
#undef PLAN_CUDA
    *out = p;
    return BAMBU_EM_OK;
}

int bambu_em_plan_create(bambu_em_plan **out, int64_t n_groups, const int64_t *grp_row_ptr,
                         const int64_t *grp_col_ptr, const int64_t *row_nnz_ptr)
{
    return plan_create_impl(out, n_groups, grp_row_ptr, grp_col_ptr, row_nnz_ptr, true);
}

int bambu_em_plan_set_stream(bambu_em_plan *p, void *cuda_stream)
{
    if (!p) return fail(BAMBU_EM_EINVAL, "plan is NULL");
    CUDA_TRY(cudaStreamSynchronize(p->stream));
    p->stream = cuda_stream ? (cudaStream_t)cuda_stream : p->own_stream;
    return BAMBU_EM_OK;
}

int bambu_em_plan_upload(bambu_em_plan *p, const int32_t *col_idx, const double *aval, const uint8_t *nz_flags,
                         const double *Y, const double *K, const uint8_t *col_flags)
{
    if (!p) return fail(BAMBU_EM_EINVAL, "plan is NULL");
    if ((p->nnz && (!col_idx || !aval || !nz_flags)) || (p->total_cols && (!Y || !K || !col_flags)))
        return fail(BAMBU_EM_EINVAL, "NULL payload pointer");
    int rc = ensure_device();
    if (rc) return rc;
    if (!p->payload_owned) {
        // one allocation: aval | Y | K | col_idx | nz_flags | col_flags  (8-byte items first)
        const size_t nnz = (size_t)p->nnz, nc = (size_t)p->total_cols;
        const size_t bytes = 8 * nnz + 16 * nc + 4 * nnz + nnz + nc + 64;
        CUDA_TRY(cudaMalloc(&p->d_payload, bytes));
        p->device_bytes += (int64_t)bytes;
        uint8_t *q = (uint8_t *)p->d_payload;
        p->d_aval = (double *)q; q += 8 * nnz;
        p->d_Y = (double *)q; q += 8 * nc;
        p->d_K = (double *)q; q += 8 * nc;
        p->d_col_idx = (int32_t *)q; q += 4 * nnz;
        p->d_nz_flags = q; q += nnz;
        p->d_col_flags = q;
        p->payload_owned = true;
    }
    cudaStream_t s = p->stream;
    CUDA_TRY(cudaMemcpyAsync((void *)p->d_aval, aval, 8 * (size_t)p->nnz, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync((void *)p->d_col_idx, col_idx, 4 * (size_t)p->nnz, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync((void *)p->d_nz_flags, nz_flags, (size_t)p->nnz, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync((void *)p->d_Y, Y, 8 * (size_t)p->total_cols, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync((void *)p->d_K, K, 8 * (size_t)p->total_cols, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync((void *)p->d_col_flags, col_flags, (size_t)p->total_cols, cudaMemcpyHostToDevice, s));
    p->payload_ready = true;
    return BAMBU_EM_OK;
}

int bambu_em_plan_bind_device(bambu_em_plan *p, const int32_t *d_col_idx, const double *d_aval,
                              const uint8_t *d_nz_flags, const double *d_Y, const double *d_K,
                              const uint8_t *d_col_flags)
{
    if (!p) return fail(BAMBU_EM_EINVAL, "plan is NULL");
    if (p->payload_owned) {
        CUDA_TRY(cudaStreamSynchronize(p->stream));
        cudaFree(p->d_payload);
        p->d_payload = nullptr;
        p->payload_owned = false;
    }
    p->d_col_idx = d_col_idx; p->d_aval = d_aval; p->d_nz_flags = d_nz_flags;
    p->d_Y = d_Y; p->d_K = d_K; p->d_col_flags = d_col_flags;
    p->payload_ready = true;
    return BAMBU_EM_OK;
}

int bambu_em_plan_run(bambu_em_plan *p, int32_t maxiter, double minvalue, double conv)
{
    if (!p) return fail(BAMBU_EM_EINVAL, "plan is NULL");
    if (!p->payload_ready) return fail(BAMBU_EM_ESTATE, "run before upload/bind_device");
    if (conv != conv || minvalue != minvalue) return fail(BAMBU_EM_EINVAL, "conv/minvalue is NaN");
    int rc = ensure_device();
    if (rc) return rc;

    EmParams P;
    P.maxiter = maxiter;
    P.minvalue = minvalue;
    P.conv = conv;
    if (conv > 1e-290 && std::isfinite(conv)) {
        P.conv_hi = conv * (1.0 + std::ldexp(1.0, -40));
        P.conv_lo = conv * (1.0 - std::ldexp(1.0, -40));
    } else {   // no screen: always take the exact division
        P.conv_hi = INFINITY;
        P.conv_lo = 0.0;
    }
    ArenaDev A;
    A.grp_row_ptr = p->d_grp_row_ptr; A.grp_col_ptr = p->d_grp_col_ptr; A.row_nnz_ptr = p->d_row_nnz_ptr;
    A.col_idx = p->d_col_idx; A.aval = p->d_aval; A.nz_flags = p->d_nz_flags;
    A.Y = p->d_Y; A.K = p->d_K; A.col_flags = p->d_col_flags;
    A.n_groups = p->n_groups; A.total_rows = p->total_rows;

    cudaStream_t s = p->stream;
    p->last_launches = 0;
    CUDA_TRY(cudaEventRecord(p->ev_start, s));
    CUDA_TRY(cudaMemsetAsync(p->d_ctrl, 0, sizeof(int) * p->n_ctrl, s));
    CUDA_TRY(cudaMemsetAsync(p->d_iters, 0, sizeof(int32_t) * std::max<int64_t>(p->n_groups, 1), s));
    CUDA_TRY(cudaMemsetAsync(p->d_status, 0, sizeof(int32_t) * std::max<int64_t>(p->n_groups, 1), s));
    CUDA_TRY(cudaMemsetAsync(p->d_est, 0, sizeof(double) * std::max<int64_t>(3 * p->total_rows, 1), s));
    CUDA_TRY(cudaMemsetAsync(p->d_info, 0, sizeof(PackInfo) * std::max<int64_t>(p->n_groups, 1), s));
    CUDA_TRY(cudaMemsetAsync(p->d_nnz_nonzero, 0, sizeof(int64_t) * std::max<int64_t>(p->n_groups, 1), s));
    CUDA_TRY(cudaEventRecord(p->ev_fork, s));
    for (auto &b : p->buckets) {
        CUDA_TRY(cudaStreamWaitEvent(b.stream, p->ev_fork, 0));
        rc = launch_bucket(p, b, A, P);
        if (rc) return rc;
        CUDA_TRY(cudaStreamWaitEvent(s, b.ev_em1, 0));
    }
    if (!p->single_tx.empty()) {
        size_t woff = 0;
        for (auto &b : p->buckets) woff += b.work.size();
        const int n = (int)p->single_tx.size();
        pack_resident_kernel<64><<<n, 64, 0, s>>>(A, p->d_work_all + woff, n, p->d_img_off, p->d_img, p->d_info,
                                                   p->d_nnz_nonzero, p->d_est, p->d_status, p->d_ctrl, 1);
        CUDA_TRY(cudaGetLastError());
        p->last_launches += 1;
    }
    CUDA_TRY(cudaEventRecord(p->ev_end, s));
    p->ran = true;
    return BAMBU_EM_OK;
}

static int check_device_error(bambu_em_plan *p)
{
    int err = 0;
    CUDA_TRY(cudaMemcpyAsync(&err, p->d_ctrl, sizeof(int), cudaMemcpyDeviceToHost, p->stream));
    CUDA_TRY(cudaStreamSynchronize(p->stream));
    if (err == kErrColRange) return fail(BAMBU_EM_EINVAL, "col_idx outside its group's column range");
    if (err == kErrColOrder) return fail(BAMBU_EM_EINVAL, "col_idx not strictly increasing inside a row");
    if (err) return fail(BAMBU_EM_ECUDA, "device error word %d", err);
    return BAMBU_EM_OK;
}

int bambu_em_plan_sync(bambu_em_plan *p)
{
    if (!p) return fail(BAMBU_EM_EINVAL, "plan is NULL");
    CUDA_TRY(cudaStreamSynchronize(p->stream));
    if (p->ran) return check_device_error(p);
    return BAMBU_EM_OK;
}

int bambu_em_plan_download(bambu_em_plan *p, double *est, int32_t *iters, int32_t *status, int64_t *nnz_nonzero)
{
    if (!p) return fail(BAMBU_EM_EINVAL, "plan is NULL");
    if (!p->ran) return fail(BAMBU_EM_ESTATE, "download before run");
    cudaStream_t s = p->stream;
    if (est && p->total_rows)
        CUDA_TRY(cudaMemcpyAsync(est, p->d_est, sizeof(double) * 3 * p->total_rows, cudaMemcpyDeviceToHost, s));
    if (iters && p->n_groups)
        CUDA_TRY(cudaMemcpyAsync(iters, p->d_iters, sizeof(int32_t) * p->n_groups, cudaMemcpyDeviceToHost, s));
    if (status && p->n_groups)
        CUDA_TRY(cudaMemcpyAsync(status, p->d_status, sizeof(int32_t) * p->n_groups, cudaMemcpyDeviceToHost, s));
    if (nnz_nonzero && p->n_groups)
        CUDA_TRY(cudaMemcpyAsync(nnz_nonzero, p->d_nnz_nonzero, sizeof(int64_t) * p->n_groups, cudaMemcpyDeviceToHost, s));
    return check_device_error(p);
}

int bambu_em_plan_download_live(bambu_em_plan *p, int32_t *live_cols, int32_t *live_nnz)
{
    if (!p) return fail(BAMBU_EM_EINVAL, "plan is NULL");
    if (!p->ran) return fail(BAMBU_EM_ESTATE, "download before run");
    std::vector<PackInfo> h((size_t)p->n_groups);
    if (p->n_groups) {
        CUDA_TRY(cudaMemcpyAsync(h.data(), p->d_info, sizeof(PackInfo) * p->n_groups, cudaMemcpyDeviceToHost, p->stream));
        CUDA_TRY(cudaStreamSynchronize(p->stream));
    }
    for (int64_t g = 0; g < p->n_groups; ++g) {
        if (live_cols) live_cols[g] = h[g].NL;
        if (live_nnz) live_nnz[g] = h[g].nnzL;
    }
    return BAMBU_EM_OK;
}

int bambu_em_plan_device_outputs(bambu_em_plan *p, double **d_est, int32_t **d_iters, int32_t **d_status)
{
    if (!p) return fail(BAMBU_EM_EINVAL, "plan is NULL");
    if (d_est) *d_est = p->d_est;
    if (d_iters) *d_iters = p->d_iters;
    if (d_status) *d_status = p->d_status;
    return BAMBU_EM_OK;
}

int bambu_em_plan_last_run_ms(bambu_em_plan *p, float *pack_ms, float *em_ms, float *total_ms)
{
    if (!p) return fail(BAMBU_EM_EINVAL, "plan is NULL");
    if (!p->ran) return fail(BAMBU_EM_ESTATE, "no run yet");
    CUDA_TRY(cudaEventSynchronize(p->ev_end));
    float pk = 0.f, em = 0.f, tot = 0.f;
    for (auto &b : p->buckets) {
        float a = 0.f, c = 0.f;
        CUDA_TRY(cudaEventSynchronize(b.ev_em1));
        CUDA_TRY(cudaEventElapsedTime(&a, b.ev_pack0, b.ev_pack1));
        CUDA_TRY(cudaEventElapsedTime(&c, b.ev_pack1, b.ev_em1));
        pk += a;
        em += c;
    }
    CUDA_TRY(cudaEventElapsedTime(&tot, p->ev_start, p->ev_end));
    if (pack_ms) *pack_ms = pk;
    if (em_ms) *em_ms = em;
    if (total_ms) *total_ms = tot;
    return BAMBU_EM_OK;
}

int bambu_em_plan_last_run_launches(bambu_em_plan *p, int32_t *n)
{
    if (!p || !n) return fail(BAMBU_EM_EINVAL, "NULL argument");
    *n = p->last_launches;
    return BAMBU_EM_OK;
}

int bambu_em_plan_device_bytes(bambu_em_plan *p, int64_t *bytes)
{
    if (!p || !bytes) return fail(BAMBU_EM_EINVAL, "NULL argument");
    *bytes = p->device_bytes;
    return BAMBU_EM_OK;
}

int bambu_em_shutdown(void) { return BAMBU_EM_OK; }

static int batched_impl(int64_t n_groups, const int64_t *grp_row_ptr, const int64_t *grp_col_ptr,
                        const int64_t *row_nnz_ptr, const int32_t *col_idx, const double *aval,
                        const uint8_t *nz_flags, const double *Y, const double *K, const uint8_t *col_flags,
                        int32_t maxiter, double minvalue, double conv, double *est, int32_t *iters, int32_t *status,
                        bool single_tx_shortcut)
{
    if (!est || !iters || !status) return fail(BAMBU_EM_EINVAL, "NULL output pointer");
    bambu_em_plan *p = nullptr;
    int rc = plan_create_impl(&p, n_groups, grp_row_ptr, grp_col_ptr, row_nnz_ptr, single_tx_shortcut);
    if (rc) return rc;
    rc = bambu_em_plan_upload(p, col_idx, aval, nz_flags, Y, K, col_flags);
    if (!rc) rc = bambu_em_plan_run(p, maxiter, minvalue, conv);
    if (!rc) rc = bambu_em_plan_download(p, est, iters, status, nullptr);
    bambu_em_plan_destroy(p);
    return rc;
}

int bambu_em_batched(int64_t n_groups, const int64_t *grp_row_ptr, const int64_t *grp_col_ptr,
                     const int64_t *row_nnz_ptr, const int32_t *col_idx, const double *aval,
                     const uint8_t *nz_flags, const double *Y, const double *K, const uint8_t *col_flags,
                     int32_t maxiter, double minvalue, double conv, double *est, int32_t *iters, int32_t *status)
{
    return batched_impl(n_groups, grp_row_ptr, grp_col_ptr, row_nnz_ptr, col_idx, aval, nz_flags, Y, K, col_flags,
                        maxiter, minvalue, conv, est, iters, status, true);
}

int bambu_emWithL1(const double *A, const double *A_full, const double *A_unique, const double *Y, const double *K,
                   int32_t M, int32_t N, int32_t maxiter, double minvalue, double conv, double *est,
                   int32_t *iters_out)
{
    if (!A || !A_full || !A_unique || !Y || !K || !est) return fail(BAMBU_EM_EINVAL, "NULL argument");
    if (M < 1 || N < 0) return fail(BAMBU_EM_EINVAL, "bad matrix extent");
    // dense -> one-group arena.  A_full / A_unique must be aval*{0,1} as filterTxRc builds them
    // (R/bambu-quantify_utilityFunctions.R:336-337); anything else cannot be expressed as flags.
    std::vector<int64_t> grp_row{0, M}, grp_col{0, N}, rnp((size_t)M + 1, 0);
    std::vector<int32_t> ci;
    std::vector<double> av;
    std::vector<uint8_t> nf, cf((size_t)N, 0);
    std::vector<int8_t> col_seen((size_t)N, -1);   // -1 unknown, 0 unique kept, 1 multi
    for (int j = 0; j < N; ++j)
        for (int i = 0; i < M; ++i) {
            const double a = A[i + (size_t)j * M], u = A_unique[i + (size_t)j * M];
            if (a != 0.0) {
                const int8_t m = (u == 0.0) ? 1 : 0;
                if (u != 0.0 && u != a) return fail(BAMBU_EM_EINVAL, "A_unique is not aval*{0,1} at (%d,%d)", i, j);
                if (col_seen[j] >= 0 && col_seen[j] != m)
                    return fail(BAMBU_EM_EINVAL, "A_unique is not constant along column %d", j);
                col_seen[j] = m;
            } else if (u != 0.0) return fail(BAMBU_EM_EINVAL, "A_unique non-zero where A is zero at (%d,%d)", i, j);
        }
    for (int j = 0; j < N; ++j) cf[j] = (col_seen[j] == 1) ? BAMBU_COL_MULTI : 0;
    for (int i = 0; i < M; ++i) {
        for (int j = 0; j < N; ++j) {
            const double a = A[i + (size_t)j * M], f = A_full[i + (size_t)j * M];
            if (a == 0.0) {
                if (f != 0.0) return fail(BAMBU_EM_EINVAL, "A_full non-zero where A is zero at (%d,%d)", i, j);
                continue;
            }
            if (f != 0.0 && f != a) return fail(BAMBU_EM_EINVAL, "A_full is not aval*{0,1} at (%d,%d)", i, j);
            ci.push_back(j);
            av.push_back(a);
            nf.push_back(f != 0.0 ? BAMBU_NZ_EQUAL : 0);
        }
        rnp[i + 1] = (int64_t)ci.size();
    }
    std::vector<double> e3((size_t)3 * M);
    int32_t it = 0, st = 0;
    // keep data() non-NULL for empty payloads
    if (ci.empty()) { ci.push_back(0); av.push_back(0.0); nf.push_back(0); }
    // emWithL1 itself has no single-transcript shortcut (that lives in run_parallel)
    int rc = batched_impl(1, grp_row.data(), grp_col.data(), rnp.data(), ci.data(), av.data(), nf.data(), Y, K,
                          cf.data(), maxiter, minvalue, conv, e3.data(), &it, &st, false);
    if (rc) return rc;
    for (int i = 0; i < M; ++i)
        for (int k = 0; k < 3; ++k) est[3 * (size_t)i + k] = e3[(size_t)k * M + i];
    if (iters_out) *iters_out = it;
    return BAMBU_EM_OK;
}

}  // extern "C"
