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
#include "em_sell.cuh"
#include "pack_kernel.cuh"
#include "sell_kernel.cuh"

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
// size classes
// ------------------------------------------------------------------------------
// Pack buckets: groups whose WORST-CASE footprint (all columns live, no zero entries;
// known on the host from the pointer arrays) falls in one class; the class fixes the
// CTA size and scratch shared memory of the pack + SELL kernels.
// EM classes: the sell kernel measures the ACTUAL shared memory a group's EM needs and
// queues the group on the device for the EM kernel of that class; the class fixes the
// dynamic shared memory and with it how many groups are resident per SM.
struct PackClass {
    uint32_t max_bytes;
    int threads;
};
static const PackClass kPackClasses[] = {
    {14u * 1024u, 64}, {56u * 1024u, 128}, {113u * 1024u, 256}, {kSmemMax, 512},
};
static const int kNumPackClasses = (int)(sizeof(kPackClasses) / sizeof(kPackClasses[0]));

struct EmClass {
    uint32_t max_bytes;   // dynamic shared memory of the class
    int threads;
};
static const EmClass kEmClasses[kNumClasses] = {
    {7u * 1024u, 64},      // 32 CTAs/SM (the CTA limit)
    {14u * 1024u, 64},     // 16
    {28u * 1024u, 128},    // 8
    {56u * 1024u, 256},    // 4
    {113u * 1024u, 512},   // 2
    {kSmemMax, 1024},      // 1
};

struct Bucket {   // pack bucket
    int threads = 0;
    uint32_t smem_pack = 0, smem_sell = 0;
    std::vector<int32_t> work;     // group ids, largest first
    int32_t *d_work = nullptr;     // view into plan->d_work_all
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
};

struct Ctrl {   // device control block, zeroed at the start of every run
    int err;
    int pad;
    unsigned long long pool_ctr;
    int qcount[8];
    int ticket[8];
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
    uint8_t *d_img = nullptr;        // stage-1 images (compact CSR + CSC), worst-case offsets
    int64_t img_bytes = 0;
    int64_t *d_img_off = nullptr;
    PackInfo *d_info = nullptr;
    uint8_t *d_pool = nullptr;       // SELL images, bump-allocated by the sell kernel
    int64_t pool_bytes = 0;
    GroupImg *d_ginfo = nullptr;
    int32_t *d_queue = nullptr;      // kNumClasses x n_groups
    int64_t *d_nnz_nonzero = nullptr;
    double *d_est = nullptr;
    int32_t *d_iters = nullptr, *d_status = nullptr;
    int32_t *d_work_all = nullptr;
    Ctrl *d_ctrl = nullptr;
    std::vector<Bucket> buckets;
    std::vector<int32_t> single_tx;   // M == 1 groups: closed form in the pack kernel only
    cudaStream_t own_stream = nullptr, stream = nullptr;
    cudaStream_t em_stream[kNumClasses] = {};
    cudaEvent_t em_done[kNumClasses] = {};
    int em_grid[kNumClasses] = {};
    cudaEvent_t ev_start = nullptr, ev_fork = nullptr, ev_packed = nullptr, ev_end = nullptr;
    bool ran = false, finished = false;
    bool single_tx_shortcut = true;   // R/bambu-quantify_utilityFunctions.R:410-414 (run_parallel, not emWithL1)
    int32_t last_maxiter = 0;
    double last_minvalue = 0, last_conv = 0;
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
static ClassBounds class_bounds()
{
    ClassBounds cb;
    for (int c = 0; c < kNumClasses; ++c) cb.max_bytes[c] = kEmClasses[c].max_bytes;
    return cb;
}

template <int T>
static int launch_pack_T(bambu_em_plan *p, Bucket &b, const ArenaDev &A)
{
    auto pack = pack_resident_kernel<T>;
    auto sell = sell_kernel<T>;
    CUDA_TRY(cudaFuncSetAttribute(pack, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b.smem_pack));
    CUDA_TRY(cudaFuncSetAttribute(sell, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b.smem_sell));
    const int n_work = (int)b.work.size();
    CUDA_TRY(cudaEventRecord(b.ev0, b.stream));
    pack<<<n_work, T, b.smem_pack, b.stream>>>(A, b.d_work, n_work, p->d_img_off, p->d_img, p->d_info,
                                                p->d_nnz_nonzero, p->d_est, p->d_status, &p->d_ctrl->err,
                                                p->single_tx_shortcut ? 1 : 0);
    CUDA_TRY(cudaGetLastError());
    sell<<<n_work, T, b.smem_sell, b.stream>>>(b.d_work, n_work, p->d_grp_row_ptr, p->d_img_off, p->d_img, p->d_info,
                                                p->d_pool, (unsigned long long)p->pool_bytes, &p->d_ctrl->pool_ctr,
                                                p->d_ginfo, class_bounds(), p->d_queue, p->n_groups,
                                                p->d_ctrl->qcount, &p->d_ctrl->err);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(b.ev1, b.stream));
    p->last_launches += 2;
    return BAMBU_EM_OK;
}

static int launch_pack(bambu_em_plan *p, Bucket &b, const ArenaDev &A)
{
    switch (b.threads) {
    case 64: return launch_pack_T<64>(p, b, A);
    case 128: return launch_pack_T<128>(p, b, A);
    case 256: return launch_pack_T<256>(p, b, A);
    case 512: return launch_pack_T<512>(p, b, A);
    }
    return fail(BAMBU_EM_EINVAL, "no pack kernel instance for %d threads", b.threads);
}

template <int T>
static int em_config_T(int c, int *grid)
{
    auto em = em_sell_kernel<T>;
    const uint32_t smem = kEmClasses[c].max_bytes;
    CUDA_TRY(cudaFuncSetAttribute(em, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, em, T, smem));
    if (per_sm < 1) return fail(BAMBU_EM_ECUDA, "EM kernel class %d does not fit an SM", c);
    *grid = per_sm * g_num_sms;
    return BAMBU_EM_OK;
}

template <int T>
static int launch_em_T(bambu_em_plan *p, int c, const EmParams &P)
{
    em_sell_kernel<T><<<p->em_grid[c], T, kEmClasses[c].max_bytes, p->em_stream[c]>>>(
        p->d_queue + (int64_t)c * p->n_groups, p->d_ctrl->qcount + c, p->d_ctrl->ticket + c, p->d_ginfo, p->d_pool,
        p->d_grp_row_ptr, p->total_rows, P, p->d_est, p->d_iters, p->d_status);
    CUDA_TRY(cudaGetLastError());
    p->last_launches += 1;
    return BAMBU_EM_OK;
}

static int em_config(int c, int *grid)
{
    switch (kEmClasses[c].threads) {
    case 64: return em_config_T<64>(c, grid);
    case 128: return em_config_T<128>(c, grid);
    case 256: return em_config_T<256>(c, grid);
    case 512: return em_config_T<512>(c, grid);
    case 1024: return em_config_T<1024>(c, grid);
    }
    return fail(BAMBU_EM_EINVAL, "no EM kernel instance for %d threads", kEmClasses[c].threads);
}

static int launch_em(bambu_em_plan *p, int c, const EmParams &P)
{
    switch (kEmClasses[c].threads) {
    case 64: return launch_em_T<64>(p, c, P);
    case 128: return launch_em_T<128>(p, c, P);
    case 256: return launch_em_T<256>(p, c, P);
    case 512: return launch_em_T<512>(p, c, P);
    case 1024: return launch_em_T<1024>(p, c, P);
    }
    return fail(BAMBU_EM_EINVAL, "no EM kernel instance for %d threads", kEmClasses[c].threads);
}

// ------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------
extern "C" {

const char *bambu_em_version(void) { return "bambu_em_b200 0.2.0 (sm_100a)"; }
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
        if (b.ev0) cudaEventDestroy(b.ev0);
        if (b.ev1) cudaEventDestroy(b.ev1);
    }
    for (int c = 0; c < kNumClasses; ++c) {
        if (p->em_stream[c]) { cudaStreamSynchronize(p->em_stream[c]); cudaStreamDestroy(p->em_stream[c]); }
        if (p->em_done[c]) cudaEventDestroy(p->em_done[c]);
    }
    if (p->ev_start) cudaEventDestroy(p->ev_start);
    if (p->ev_fork) cudaEventDestroy(p->ev_fork);
    if (p->ev_packed) cudaEventDestroy(p->ev_packed);
    if (p->ev_end) cudaEventDestroy(p->ev_end);
    if (p->own_stream) cudaStreamDestroy(p->own_stream);
    cudaFree(p->d_grp_row_ptr); cudaFree(p->d_grp_col_ptr); cudaFree(p->d_row_nnz_ptr);
    if (p->payload_owned) cudaFree(p->d_payload);
    cudaFree(p->d_img); cudaFree(p->d_img_off); cudaFree(p->d_info); cudaFree(p->d_nnz_nonzero);
    cudaFree(p->d_pool); cudaFree(p->d_ginfo); cudaFree(p->d_queue);
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

    // ---- classify groups by worst-case footprint, lay out the stage-1 images ------
    std::vector<int64_t> img_off((size_t)n_groups + 1, 0);
    std::vector<std::vector<int32_t>> per_class(kNumPackClasses);
    std::vector<uint32_t> ub_pack((size_t)n_groups, 0), ub_sell((size_t)n_groups, 0);
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
        ub_pack[g] = pack_smem_bytes((uint32_t)M, (uint32_t)N);
        ub_sell[g] = sell_scratch_bytes((uint32_t)M, (uint32_t)N);
        int c = 0;
        while (c < kNumPackClasses - 1 && bytes > kPackClasses[c].max_bytes) ++c;
        per_class[c].push_back((int32_t)g);
        off += img_layout((uint32_t)M, (uint32_t)N, (uint32_t)e).total;
    }
    img_off[n_groups] = off;
    p->img_bytes = off;
    // SELL pool: typical need is ~25 B per live entry + ~40 B per row/live column; grown on demand
    p->pool_bytes = (int64_t)(24.0 * (double)p->nnz + 48.0 * (double)(p->total_rows + p->total_cols)) + (4 << 20);
    p->pool_bytes = (p->pool_bytes + 15) / 16 * 16;

    auto fail_destroy = [&](int code) { bambu_em_plan_destroy(p); return code; };
#define PLAN_TRY(expr) do { int _rc = (expr); if (_rc) return fail_destroy(_rc); } while (0)
#define PLAN_CUDA(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { fail(_e == cudaErrorMemoryAllocation ? BAMBU_EM_ENOMEM : BAMBU_EM_ECUDA, "%s failed: %s", #expr, cudaGetErrorString(_e)); return fail_destroy(_e == cudaErrorMemoryAllocation ? BAMBU_EM_ENOMEM : BAMBU_EM_ECUDA); } } while (0)

    size_t n_work_total = p->single_tx.size();
    for (int c = 0; c < kNumPackClasses; ++c) {
        if (per_class[c].empty()) continue;
        Bucket b;
        b.threads = kPackClasses[c].threads;
        b.work = std::move(per_class[c]);
        std::stable_sort(b.work.begin(), b.work.end(), [&](int32_t x, int32_t y) { return gnnz[x] > gnnz[y]; });
        for (int32_t g : b.work) {
            b.smem_pack = std::max(b.smem_pack, ub_pack[g]);
            b.smem_sell = std::max(b.smem_sell, ub_sell[g]);
        }
        n_work_total += b.work.size();
        p->buckets.push_back(std::move(b));
    }

    PLAN_CUDA(cudaStreamCreateWithFlags(&p->own_stream, cudaStreamNonBlocking));
    p->stream = p->own_stream;
    PLAN_CUDA(cudaEventCreate(&p->ev_start));
    PLAN_CUDA(cudaEventCreateWithFlags(&p->ev_fork, cudaEventDisableTiming));
    PLAN_CUDA(cudaEventCreate(&p->ev_packed));
    PLAN_CUDA(cudaEventCreate(&p->ev_end));
    for (int c = 0; c < kNumClasses; ++c) {
        PLAN_CUDA(cudaStreamCreateWithFlags(&p->em_stream[c], cudaStreamNonBlocking));
        PLAN_CUDA(cudaEventCreateWithFlags(&p->em_done[c], cudaEventDisableTiming));
        PLAN_TRY(em_config(c, &p->em_grid[c]));
    }

    PLAN_TRY(dev_alloc(p, &p->d_grp_row_ptr, (size_t)n_groups + 1));
    PLAN_TRY(dev_alloc(p, &p->d_grp_col_ptr, (size_t)n_groups + 1));
    PLAN_TRY(dev_alloc(p, &p->d_row_nnz_ptr, (size_t)p->total_rows + 1));
    PLAN_TRY(dev_alloc(p, &p->d_img, (size_t)p->img_bytes));
    PLAN_TRY(dev_alloc(p, &p->d_pool, (size_t)p->pool_bytes));
    PLAN_TRY(dev_alloc(p, &p->d_img_off, (size_t)n_groups + 1));
    PLAN_TRY(dev_alloc(p, &p->d_info, (size_t)n_groups));
    PLAN_TRY(dev_alloc(p, &p->d_ginfo, (size_t)n_groups));
    PLAN_TRY(dev_alloc(p, &p->d_queue, (size_t)kNumClasses * (size_t)std::max<int64_t>(n_groups, 1)));
    PLAN_TRY(dev_alloc(p, &p->d_nnz_nonzero, (size_t)n_groups));
    PLAN_TRY(dev_alloc(p, &p->d_est, (size_t)3 * p->total_rows));
    PLAN_TRY(dev_alloc(p, &p->d_iters, (size_t)n_groups));
    PLAN_TRY(dev_alloc(p, &p->d_status, (size_t)n_groups));
    PLAN_TRY(dev_alloc(p, &p->d_work_all, n_work_total));
    PLAN_TRY(dev_alloc(p, &p->d_ctrl, 1));

    cudaStream_t s = p->stream;
    PLAN_CUDA(cudaMemcpyAsync(p->d_grp_row_ptr, grp_row_ptr, sizeof(int64_t) * (n_groups + 1), cudaMemcpyHostToDevice, s));
    PLAN_CUDA(cudaMemcpyAsync(p->d_grp_col_ptr, grp_col_ptr, sizeof(int64_t) * (n_groups + 1), cudaMemcpyHostToDevice, s));
    PLAN_CUDA(cudaMemcpyAsync(p->d_row_nnz_ptr, row_nnz_ptr, sizeof(int64_t) * (p->total_rows + 1), cudaMemcpyHostToDevice, s));
    PLAN_CUDA(cudaMemcpyAsync(p->d_img_off, img_off.data(), sizeof(int64_t) * (n_groups + 1), cudaMemcpyHostToDevice, s));
    size_t woff = 0;
    for (auto &b : p->buckets) {
        b.d_work = p->d_work_all + woff;
        PLAN_CUDA(cudaMemcpyAsync(b.d_work, b.work.data(), sizeof(int32_t) * b.work.size(), cudaMemcpyHostToDevice, s));
        woff += b.work.size();
        PLAN_CUDA(cudaStreamCreateWithFlags(&b.stream, cudaStreamNonBlocking));
        PLAN_CUDA(cudaEventCreate(&b.ev0));
        PLAN_CUDA(cudaEventCreate(&b.ev1));
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

static int enqueue_run(bambu_em_plan *p)
{
    EmParams P;
    const double conv = p->last_conv;
    P.maxiter = p->last_maxiter;
    P.minvalue = p->last_minvalue;
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
    const int64_t ng = std::max<int64_t>(p->n_groups, 1);
    p->last_launches = 0;
    CUDA_TRY(cudaEventRecord(p->ev_start, s));
    CUDA_TRY(cudaMemsetAsync(p->d_ctrl, 0, sizeof(Ctrl), s));
    CUDA_TRY(cudaMemsetAsync(p->d_iters, 0, sizeof(int32_t) * ng, s));
    CUDA_TRY(cudaMemsetAsync(p->d_status, 0, sizeof(int32_t) * ng, s));
    CUDA_TRY(cudaMemsetAsync(p->d_est, 0, sizeof(double) * std::max<int64_t>(3 * p->total_rows, 1), s));
    CUDA_TRY(cudaMemsetAsync(p->d_info, 0, sizeof(PackInfo) * ng, s));
    CUDA_TRY(cudaMemsetAsync(p->d_ginfo, 0, sizeof(GroupImg) * ng, s));
    CUDA_TRY(cudaMemsetAsync(p->d_nnz_nonzero, 0, sizeof(int64_t) * ng, s));
    CUDA_TRY(cudaEventRecord(p->ev_fork, s));
    // pack + SELL, one bucket per stream
    for (auto &b : p->buckets) {
        CUDA_TRY(cudaStreamWaitEvent(b.stream, p->ev_fork, 0));
        int rc = launch_pack(p, b, A);
        if (rc) return rc;
        CUDA_TRY(cudaStreamWaitEvent(s, b.ev1, 0));
    }
    if (!p->single_tx.empty()) {
        size_t woff = 0;
        for (auto &b : p->buckets) woff += b.work.size();
        const int n = (int)p->single_tx.size();
        pack_resident_kernel<64><<<n, 64, 0, s>>>(A, p->d_work_all + woff, n, p->d_img_off, p->d_img, p->d_info,
                                                   p->d_nnz_nonzero, p->d_est, p->d_status, &p->d_ctrl->err, 1);
        CUDA_TRY(cudaGetLastError());
        p->last_launches += 1;
    }
    CUDA_TRY(cudaEventRecord(p->ev_packed, s));
    // EM, one class per stream; the queues were filled on the device
    if (!p->buckets.empty()) {
        for (int c = kNumClasses - 1; c >= 0; --c) {
            CUDA_TRY(cudaStreamWaitEvent(p->em_stream[c], p->ev_packed, 0));
            int rc = launch_em(p, c, P);
            if (rc) return rc;
            CUDA_TRY(cudaEventRecord(p->em_done[c], p->em_stream[c]));
            CUDA_TRY(cudaStreamWaitEvent(s, p->em_done[c], 0));
        }
    }
    CUDA_TRY(cudaEventRecord(p->ev_end, s));
    p->ran = true;
    p->finished = false;
    return BAMBU_EM_OK;
}

int bambu_em_plan_run(bambu_em_plan *p, int32_t maxiter, double minvalue, double conv)
{
    if (!p) return fail(BAMBU_EM_EINVAL, "plan is NULL");
    if (!p->payload_ready) return fail(BAMBU_EM_ESTATE, "run before upload/bind_device");
    if (conv != conv || minvalue != minvalue) return fail(BAMBU_EM_EINVAL, "conv/minvalue is NaN");
    int rc = ensure_device();
    if (rc) return rc;
    p->last_maxiter = maxiter;
    p->last_minvalue = minvalue;
    p->last_conv = conv;
    return enqueue_run(p);
}

// wait for the run; if the SELL pool was too small, grow it and run again
static int finish_run(bambu_em_plan *p)
{
    for (int attempt = 0; attempt < 8; ++attempt) {
        int err = 0;
        CUDA_TRY(cudaMemcpyAsync(&err, &p->d_ctrl->err, sizeof(int), cudaMemcpyDeviceToHost, p->stream));
        CUDA_TRY(cudaStreamSynchronize(p->stream));
        if (err == 0) { p->finished = true; return BAMBU_EM_OK; }
        if (err == kErrColRange) return fail(BAMBU_EM_EINVAL, "col_idx outside its group's column range");
        if (err == kErrColOrder) return fail(BAMBU_EM_EINVAL, "col_idx not strictly increasing inside a row");
        if (err == kErrTooBig) return fail(BAMBU_EM_EINVAL, "a gene group's image exceeds the shared-memory resident tier");
        if (err != kErrPool) return fail(BAMBU_EM_ECUDA, "device error word %d", err);
        cudaFree(p->d_pool);
        p->d_pool = nullptr;
        p->device_bytes -= p->pool_bytes;
        p->pool_bytes *= 2;
        int rc = dev_alloc(p, &p->d_pool, (size_t)p->pool_bytes);
        if (rc) return rc;
        rc = enqueue_run(p);
        if (rc) return rc;
    }
    return fail(BAMBU_EM_ENOMEM, "SELL image pool kept overflowing");
}

int bambu_em_plan_sync(bambu_em_plan *p)
{
    if (!p) return fail(BAMBU_EM_EINVAL, "plan is NULL");
    if (p->ran && !p->finished) return finish_run(p);
    CUDA_TRY(cudaStreamSynchronize(p->stream));
    return BAMBU_EM_OK;
}

int bambu_em_plan_download(bambu_em_plan *p, double *est, int32_t *iters, int32_t *status, int64_t *nnz_nonzero)
{
    if (!p) return fail(BAMBU_EM_EINVAL, "plan is NULL");
    if (!p->ran) return fail(BAMBU_EM_ESTATE, "download before run");
    if (!p->finished) { int rc = finish_run(p); if (rc) return rc; }
    cudaStream_t s = p->stream;
    if (est && p->total_rows)
        CUDA_TRY(cudaMemcpyAsync(est, p->d_est, sizeof(double) * 3 * p->total_rows, cudaMemcpyDeviceToHost, s));
    if (iters && p->n_groups)
        CUDA_TRY(cudaMemcpyAsync(iters, p->d_iters, sizeof(int32_t) * p->n_groups, cudaMemcpyDeviceToHost, s));
    if (status && p->n_groups)
        CUDA_TRY(cudaMemcpyAsync(status, p->d_status, sizeof(int32_t) * p->n_groups, cudaMemcpyDeviceToHost, s));
    if (nnz_nonzero && p->n_groups)
        CUDA_TRY(cudaMemcpyAsync(nnz_nonzero, p->d_nnz_nonzero, sizeof(int64_t) * p->n_groups, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    return BAMBU_EM_OK;
}

int bambu_em_plan_download_live(bambu_em_plan *p, int32_t *live_cols, int32_t *live_nnz, int32_t *smem_bytes,
                                int32_t *slots)
{
    if (!p) return fail(BAMBU_EM_EINVAL, "plan is NULL");
    if (!p->ran) return fail(BAMBU_EM_ESTATE, "download before run");
    if (!p->finished) { int rc = finish_run(p); if (rc) return rc; }
    std::vector<PackInfo> h((size_t)p->n_groups);
    std::vector<GroupImg> gi((size_t)p->n_groups);
    if (p->n_groups) {
        CUDA_TRY(cudaMemcpyAsync(h.data(), p->d_info, sizeof(PackInfo) * p->n_groups, cudaMemcpyDeviceToHost, p->stream));
        CUDA_TRY(cudaMemcpyAsync(gi.data(), p->d_ginfo, sizeof(GroupImg) * p->n_groups, cudaMemcpyDeviceToHost, p->stream));
        CUDA_TRY(cudaStreamSynchronize(p->stream));
    }
    for (int64_t g = 0; g < p->n_groups; ++g) {
        if (live_cols) live_cols[g] = h[g].NL;
        if (live_nnz) live_nnz[g] = h[g].nnzL;
        if (smem_bytes) smem_bytes[g] = (int32_t)gi[g].smem_bytes;
        if (slots) slots[g] = (int32_t)(gi[g].slotsR + gi[g].slotsC);
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
    if (!p->finished) { int rc = finish_run(p); if (rc) return rc; }
    CUDA_TRY(cudaEventSynchronize(p->ev_end));
    float pk = 0.f, em = 0.f, tot = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&pk, p->ev_start, p->ev_packed));
    CUDA_TRY(cudaEventElapsedTime(&em, p->ev_packed, p->ev_end));
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
