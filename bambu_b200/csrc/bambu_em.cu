// bambu_em.cu -- host side of libbambu_em.so: the C ABI of include/bambu_em.h,
// the size-class scheduler, the chunked H2D / compute / D2H pipeline and the kernel
// launches.  No torch, no CPU fallback.
//
// Replaces the serial lapply over gene groups of abundance_quantification()
// (R/bambu-quantify_utilityFunctions.R:379-391).  A *slot* processes one contiguous
// range of gene groups (a chunk): per pack class one pack + one SELL launch, then per
// EM class one persistent EM launch, each on its own stream.  The plan API is one
// slot over the whole arena; bambu_em_batched() streams the arena through three
// cached slots so that copies and kernels of neighbouring chunks overlap.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/bambu_em.h"
#include "common.cuh"
#include "em_sell.cuh"
#include "giant.cuh"
#include "giant_pack.cuh"
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
#define TRY(expr) do { int _rc = (expr); if (_rc) return _rc; } while (0)

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
    if (prop.major != 10 || prop.minor != 0)     // an arch-specific ("a") cubin runs on exactly this compute capability
        return fail(BAMBU_EM_ENODEVICE, "device %d is sm_%d%d; this library is built for sm_100a only",
                    g_device, prop.major, prop.minor);
    CUDA_TRY(cudaSetDevice(g_device));
    g_num_sms = prop.multiProcessorCount;
    g_device_checked = true;
    return BAMBU_EM_OK;
}

// for quant_dev.cu (the device packer shares the device selection and its checks)
int bambu_em_internal_ensure_device() { return ensure_device(); }
void bambu_quant_release_workspace();

// ------------------------------------------------------------------------------
// size classes
// ------------------------------------------------------------------------------
// Pack classes: groups whose WORST-CASE footprint (all columns live, no zero entries;
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
// An SM has 228 KB (233472 B) of shared memory; a resident CTA costs its dynamic size plus the kernel's static
// shared memory (16 B), rounded up to 128 B, plus 1 KB of system use.  k CTAs fit when the class size is at most
// floor128(233472 / k - 1024 - 128).  Registers cap the 64-thread kernel at 16 CTAs (64 registers) and the
// 128-thread kernel at 9 (56 registers).  Measured on a config-4 sample, weighted by EM updates: ~77 % of the work is
// in groups of at most 13 KB, ~19 % between that and 24 KB (tools/image_sizes.py).  bench.py / ncu show the result as
// launch__occupancy_limit_shared_mem.
static const EmClass kEmClasses[kNumClasses] = {
    {6144u, 64},           // 32 would fit (the CTA limit); the register file stops at 16
    {13440u, 64},          // 16 CTAs/SM
    {24704u, 128},         // 9
    {57216u, 256},         // 4
    {115584u, 512},        // 2
    {kSmemMax, 1024},      // 1
};
static int g_em_grid[kNumClasses] = {};
static int g_em_grid_lat[kNumClasses] = {};     // latency instance (256 threads, 4 slots in flight) on the classes below 256 threads
constexpr int kLatThreads = 256, kLatUnroll = 4;
constexpr int kLatClasses = 3;                  // classes 0..2 (64 / 64 / 128 threads in throughput mode)
static bool g_em_configured = false;

struct Ctrl {   // device control block, zeroed at the start of every run
    int err;
    int pad;
    unsigned long long pool_ctr;
    int qcount[8];
    int ticket[8];
    unsigned int gticket;        // streamed tier: next locus to hand to a team that has finished one
    unsigned int gpad[3];
    unsigned int gflags[64 * 16]; // streamed tier, per team of CTAs: [0..7] hazard / stopping flags, [32..33] barrier, [40] next locus
};

// ------------------------------------------------------------------------------
// grow-only buffers
// ------------------------------------------------------------------------------
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes, int64_t *account)
    {
        bytes = std::max<size_t>(bytes, 256);
        if (bytes <= cap) return BAMBU_EM_OK;
        if (p) { cudaFree(p); if (account) *account -= (int64_t)cap; }
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 8;          // head-room: chunks differ a little in size
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { want = bytes; e = cudaMalloc(&p, want); }
        if (e != cudaSuccess) return fail(BAMBU_EM_ENOMEM, "cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
        cap = want;
        if (account) *account += (int64_t)cap;
        return BAMBU_EM_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <typename Tp> Tp *as() const { return reinterpret_cast<Tp *>(p); }
};

struct PinBuf {   // pinned host staging
    void *p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes)
    {
        bytes = std::max<size_t>(bytes, 256);
        if (bytes <= cap) return BAMBU_EM_OK;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 8;
        cudaError_t e = cudaMallocHost(&p, want);
        if (e != cudaSuccess) return fail(BAMBU_EM_ENOMEM, "cudaMallocHost(%zu) failed: %s", want, cudaGetErrorString(e));
        cap = want;
        return BAMBU_EM_OK;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
    template <typename Tp> Tp *as() const { return reinterpret_cast<Tp *>(p); }
};

// ------------------------------------------------------------------------------
// slot: everything needed to take one range of gene groups through the kernels
// ------------------------------------------------------------------------------
struct Bucket {   // pack bucket of the current chunk
    int n = 0;                 // groups
    int32_t *d_work = nullptr; // view into the slot's work buffer
    uint32_t smem_pack = 0, smem_sell = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;
};

struct Slot {
    // the range
    int64_t g0 = 0, g1 = 0, r0 = 0, r1 = 0, c0 = 0, c1 = 0, e0 = 0, e1 = 0;
    int64_t ng() const { return g1 - g0; }
    int64_t rows() const { return r1 - r0; }
    int64_t cols() const { return c1 - c0; }
    int64_t nnz() const { return e1 - e0; }
    bool single_tx_shortcut = true;   // R/bambu-quantify_utilityFunctions.R:410-414 (run_parallel, not emWithL1)
    // wire form v2 (bambu_em_batched_v2): columns = live columns, entries = live entries, [ib0, ib1) = the range's
    // bytes of the caller's index blob
    bool v2 = false;
    int64_t ib0 = 0, ib1 = 0;
    int32_t col_mode = 0;
    // device buffers (grow-only)
    DevBuf grp_row_ptr, grp_col_ptr, row_nnz_ptr, payload, img, img_off, info, pool, ginfo, queue, nnz_nonzero, est,
        iters, status, work, ctrl, gdesc, gbuf, grp_idx_off, ghist;
    // per-group extents of the current chunk (host copy: the caller's arrays may be gone by the time a
    // group has to be moved to the streamed tier)
    std::vector<int32_t> hM, hN;
    std::vector<int64_t> hE;
    std::vector<GiantDesc> giants;   // streamed-tier groups of the current chunk (+ spilled ones)
    int64_t giant_bytes = 0, giant_bytes_prepared = 0;
    int n_giant_prepared = 0;
    // streamed tier (giant loci) of the current chunk
    int n_giant = 0;
    uint32_t giant_smem = 0;
    PinBuf h_gdesc;
    cudaStream_t giant_stream = nullptr;
    cudaEvent_t giant_packed = nullptr;
    int64_t pool_bytes = 0;
    // payload views (device, already shifted so that absolute arena indices work)
    bool payload_bound = false, payload_ready = false;
    const int32_t *v_col_idx = nullptr;
    const double *v_aval = nullptr;
    const uint8_t *v_nz_flags = nullptr;
    const double *v_Y = nullptr, *v_K = nullptr;
    const uint8_t *v_col_flags = nullptr;
    // the same for the wire form v2
    const uint32_t *w_row_end = nullptr, *w_eq_bits = nullptr, *w_multi_bits = nullptr;
    const double *w_rs = nullptr, *w_aval = nullptr;
    const uint8_t *w_cidx = nullptr, *w_colY = nullptr, *w_colK = nullptr;
    // host staging (pinned): img_off + work lists of the current chunk
    PinBuf h_img_off, h_work, h_idx_off;
    Bucket buckets[8];
    int n_single = 0;
    int32_t *d_single = nullptr;
    // streams / events
    cudaStream_t own_stream = nullptr, stream = nullptr;
    cudaStream_t em_stream[kNumClasses] = {};
    cudaEvent_t em_done[kNumClasses] = {};
    cudaEvent_t ev_start = nullptr, ev_fork = nullptr, ev_packed = nullptr, ev_end = nullptr, ev_free = nullptr;
    bool inited = false, prepared = false, ran = false, finished = false, busy = false;
    EmParams P = {};
    int launches = 0;
    int64_t device_bytes = 0;
    // pending download of a pipelined chunk (batched path)
    double *dl_est = nullptr;
    int32_t *dl_iters = nullptr, *dl_status = nullptr;
    int64_t dl_total_rows = 0;
};

static int slot_init(Slot &s)
{
    if (s.inited) return BAMBU_EM_OK;
    CUDA_TRY(cudaStreamCreateWithFlags(&s.own_stream, cudaStreamNonBlocking));
    s.stream = s.own_stream;
    CUDA_TRY(cudaEventCreate(&s.ev_start));
    CUDA_TRY(cudaEventCreateWithFlags(&s.ev_fork, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreate(&s.ev_packed));
    CUDA_TRY(cudaEventCreate(&s.ev_end));
    CUDA_TRY(cudaEventCreateWithFlags(&s.ev_free, cudaEventDisableTiming));
    for (int c = 0; c < kNumClasses; ++c) {
        CUDA_TRY(cudaStreamCreateWithFlags(&s.em_stream[c], cudaStreamNonBlocking));
        CUDA_TRY(cudaEventCreateWithFlags(&s.em_done[c], cudaEventDisableTiming));
    }
    for (int b = 0; b < kNumPackClasses; ++b) {
        CUDA_TRY(cudaStreamCreateWithFlags(&s.buckets[b].stream, cudaStreamNonBlocking));
        CUDA_TRY(cudaEventCreateWithFlags(&s.buckets[b].done, cudaEventDisableTiming));
    }
    CUDA_TRY(cudaStreamCreateWithFlags(&s.giant_stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreateWithFlags(&s.giant_packed, cudaEventDisableTiming));
    s.inited = true;
    return BAMBU_EM_OK;
}

static void slot_destroy(Slot &s)
{
    if (s.stream) cudaStreamSynchronize(s.stream);
    for (int b = 0; b < kNumPackClasses; ++b) {
        if (s.buckets[b].stream) { cudaStreamSynchronize(s.buckets[b].stream); cudaStreamDestroy(s.buckets[b].stream); }
        if (s.buckets[b].done) cudaEventDestroy(s.buckets[b].done);
    }
    for (int c = 0; c < kNumClasses; ++c) {
        if (s.em_stream[c]) { cudaStreamSynchronize(s.em_stream[c]); cudaStreamDestroy(s.em_stream[c]); }
        if (s.em_done[c]) cudaEventDestroy(s.em_done[c]);
    }
    if (s.giant_stream) { cudaStreamSynchronize(s.giant_stream); cudaStreamDestroy(s.giant_stream); }
    for (cudaEvent_t e : {s.ev_start, s.ev_fork, s.ev_packed, s.ev_end, s.ev_free, s.giant_packed})
        if (e) cudaEventDestroy(e);
    if (s.own_stream) cudaStreamDestroy(s.own_stream);
    for (DevBuf *b : {&s.grp_row_ptr, &s.grp_col_ptr, &s.row_nnz_ptr, &s.payload, &s.img, &s.img_off, &s.info, &s.pool,
                      &s.ginfo, &s.queue, &s.nnz_nonzero, &s.est, &s.iters, &s.status, &s.work, &s.ctrl, &s.gdesc, &s.gbuf,
                      &s.grp_idx_off, &s.ghist})
        b->release();
    s.h_idx_off.release();
    s.h_gdesc.release();
    s.h_img_off.release();
    s.h_work.release();
    s = Slot();
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
static int em_config_T(int c)
{
    auto em = em_sell_kernel<T>;
    const uint32_t smem = kEmClasses[c].max_bytes;
    CUDA_TRY(cudaFuncSetAttribute(em, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, em, T, smem));
    if (per_sm < 1) return fail(BAMBU_EM_ECUDA, "EM kernel class %d does not fit an SM", c);
    g_em_grid[c] = per_sm * g_num_sms;
    return BAMBU_EM_OK;
}

static int em_config_latency()
{
    auto em = em_sell_kernel<kLatThreads, kLatUnroll>;
    CUDA_TRY(cudaFuncSetAttribute(em, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kEmClasses[kLatClasses - 1].max_bytes));
    for (int c = 0; c < kLatClasses; ++c) {
        int per_sm = 0;
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, em, kLatThreads, kEmClasses[c].max_bytes));
        if (per_sm < 1) return fail(BAMBU_EM_ECUDA, "latency EM kernel does not fit an SM (class %d)", c);
        g_em_grid_lat[c] = per_sm * g_num_sms;
    }
    return BAMBU_EM_OK;
}

template <int T>
static int pack_config_T()
{
    CUDA_TRY(cudaFuncSetAttribute(pack_resident_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax));
    CUDA_TRY(cudaFuncSetAttribute(pack_resident_wire_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax));
    CUDA_TRY(cudaFuncSetAttribute(sell_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax));
    return BAMBU_EM_OK;
}

constexpr int kGiantThreads = 1024;

static bool force_giant()
{
    const char *e = getenv("BAMBU_EM_FORCE_GIANT");
    return e && e[0] == '1';
}

static int configure_kernels()
{
    if (g_em_configured) return BAMBU_EM_OK;
    CUDA_TRY(cudaFuncSetAttribute(em_giant_kernel<kGiantThreads>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax));
    for (int c = 0; c < kNumClasses; ++c) {
        switch (kEmClasses[c].threads) {
        case 64: TRY(em_config_T<64>(c)); break;
        case 128: TRY(em_config_T<128>(c)); break;
        case 256: TRY(em_config_T<256>(c)); break;
        case 512: TRY(em_config_T<512>(c)); break;
        case 1024: TRY(em_config_T<1024>(c)); break;
        default: return fail(BAMBU_EM_EINVAL, "no EM kernel instance for %d threads", kEmClasses[c].threads);
        }
    }
    TRY(em_config_latency());
    TRY(pack_config_T<64>());
    TRY(pack_config_T<128>());
    TRY(pack_config_T<256>());
    TRY(pack_config_T<512>());
    g_em_configured = true;
    return BAMBU_EM_OK;
}

struct SlotViews {   // device pointers shifted so that absolute group / row / column / entry indices work
    ArenaDev A;
    WireDev W;
    const int64_t *img_off;
    PackInfo *info;
    GroupImg *ginfo;
    int64_t *nnz_nonzero;
    double *est;
    int32_t *iters, *status;
    Ctrl *ctrl;
};

static SlotViews slot_views(const Slot &s)
{
    SlotViews v;
    v.A.grp_row_ptr = s.grp_row_ptr.as<int64_t>() - s.g0;
    v.A.grp_col_ptr = s.grp_col_ptr.as<int64_t>() - s.g0;
    v.A.row_nnz_ptr = s.row_nnz_ptr.as<int64_t>() - s.r0;
    v.A.col_idx = s.v_col_idx; v.A.aval = s.v_aval; v.A.nz_flags = s.v_nz_flags;
    v.A.Y = s.v_Y; v.A.K = s.v_K; v.A.col_flags = s.v_col_flags;
    v.A.n_groups = s.ng();
    v.A.total_rows = s.rows();
    v.W.grp_row_ptr = v.A.grp_row_ptr;
    v.W.grp_col_ptr = v.A.grp_col_ptr;
    v.W.grp_ent_ptr = s.row_nnz_ptr.as<int64_t>() - s.g0;       // v2 keeps the per-group entry pointers in this buffer
    v.W.grp_idx_off = s.grp_idx_off.as<int64_t>() - s.g0;
    v.W.row_end = s.w_row_end; v.W.rs = s.w_rs; v.W.aval = s.w_aval; v.W.cidx = s.w_cidx; v.W.eq_bits = s.w_eq_bits;
    v.W.colY = s.w_colY; v.W.colK = s.w_colK; v.W.multi_bits = s.w_multi_bits; v.W.col_mode = s.col_mode;
    v.W.n_groups = s.ng();
    v.W.total_rows = s.rows();
    v.img_off = s.img_off.as<int64_t>() - s.g0;
    v.info = s.info.as<PackInfo>() - s.g0;
    v.ginfo = s.ginfo.as<GroupImg>() - s.g0;
    v.nnz_nonzero = s.nnz_nonzero.as<int64_t>() - s.g0;
    v.est = s.est.as<double>() - s.r0;     // est[k*rows + (row - r0)]
    v.iters = s.iters.as<int32_t>() - s.g0;
    v.status = s.status.as<int32_t>() - s.g0;
    v.ctrl = s.ctrl.as<Ctrl>();
    return v;
}

template <int T>
static int launch_pack_T(Slot &s, Bucket &b, const SlotViews &v)
{
    if (s.v2)
        pack_resident_wire_kernel<T><<<b.n, T, b.smem_pack, b.stream>>>(v.W, b.d_work, b.n, v.img_off, s.img.as<uint8_t>(),
                                                                         v.info, v.nnz_nonzero, v.est, v.status, &v.ctrl->err,
                                                                         s.single_tx_shortcut ? 1 : 0);
    else
        pack_resident_kernel<T><<<b.n, T, b.smem_pack, b.stream>>>(v.A, b.d_work, b.n, v.img_off, s.img.as<uint8_t>(), v.info,
                                                                    v.nnz_nonzero, v.est, v.status, &v.ctrl->err,
                                                                    s.single_tx_shortcut ? 1 : 0);
    CUDA_TRY(cudaGetLastError());
    sell_kernel<T><<<b.n, T, b.smem_sell, b.stream>>>(b.d_work, b.n, v.A.grp_row_ptr, v.img_off, s.img.as<uint8_t>(), v.info,
                                                       s.pool.as<uint8_t>(), (unsigned long long)s.pool_bytes,
                                                       &v.ctrl->pool_ctr, v.ginfo, class_bounds(), s.queue.as<int32_t>(),
                                                       std::max<int64_t>(s.ng(), 1), v.ctrl->qcount, &v.ctrl->err);
    CUDA_TRY(cudaGetLastError());
    s.launches += 2;
    return BAMBU_EM_OK;
}

template <int T>
static int launch_em_T(Slot &s, int c, const SlotViews &v)
{
    em_sell_kernel<T><<<g_em_grid[c], T, kEmClasses[c].max_bytes, s.em_stream[c]>>>(
        s.queue.as<int32_t>() + (int64_t)c * std::max<int64_t>(s.ng(), 1), v.ctrl->qcount + c, v.ctrl->ticket + c, v.ginfo,
        s.pool.as<uint8_t>(), v.A.grp_row_ptr, s.rows(), s.P, v.est, v.iters, v.status);
    CUDA_TRY(cudaGetLastError());
    s.launches += 1;
    return BAMBU_EM_OK;
}

static int launch_pack(Slot &s, int bi, const SlotViews &v)
{
    Bucket &b = s.buckets[bi];
    switch (kPackClasses[bi].threads) {
    case 64: return launch_pack_T<64>(s, b, v);
    case 128: return launch_pack_T<128>(s, b, v);
    case 256: return launch_pack_T<256>(s, b, v);
    case 512: return launch_pack_T<512>(s, b, v);
    }
    return fail(BAMBU_EM_EINVAL, "no pack kernel instance for %d threads", kPackClasses[bi].threads);
}

// A small arena (a single sample, a few samples) is bound by its longest-running groups, not by residency: its small
// classes run on the latency instance of the kernel.  BAMBU_EM_LATENCY_GROUPS moves the threshold (0 = never).
static bool latency_mode(const Slot &s)
{
    const char *e = getenv("BAMBU_EM_LATENCY_GROUPS");
    // measured on config-4 samples (3.6k groups each), device ms in latency / throughput mode: 1 sample 3.2 / 3.8,
    // 2 samples 5.3 / 4.3, 4 samples 9.3 / 6.8, 13 samples 26.6 / 16.9 -- the latency instance only pays while the
    // whole arena is resident at once
    return s.ng() <= (e ? atoll(e) : (int64_t)6000);
}

static int launch_em(Slot &s, int c, const SlotViews &v)
{
    if (c < kLatClasses && latency_mode(s)) {
        em_sell_kernel<kLatThreads, kLatUnroll><<<g_em_grid_lat[c], kLatThreads, kEmClasses[c].max_bytes, s.em_stream[c]>>>(
            s.queue.as<int32_t>() + (int64_t)c * std::max<int64_t>(s.ng(), 1), v.ctrl->qcount + c, v.ctrl->ticket + c, v.ginfo,
            s.pool.as<uint8_t>(), v.A.grp_row_ptr, s.rows(), s.P, v.est, v.iters, v.status);
        CUDA_TRY(cudaGetLastError());
        s.launches += 1;
        return BAMBU_EM_OK;
    }
    switch (kEmClasses[c].threads) {
    case 64: return launch_em_T<64>(s, c, v);
    case 128: return launch_em_T<128>(s, c, v);
    case 256: return launch_em_T<256>(s, c, v);
    case 512: return launch_em_T<512>(s, c, v);
    case 1024: return launch_em_T<1024>(s, c, v);
    }
    return fail(BAMBU_EM_EINVAL, "no EM kernel instance for %d threads", kEmClasses[c].threads);
}

// ------------------------------------------------------------------------------
// slot_prepare: host-side analysis of the range [g0, g1) + upload of its structure
// ------------------------------------------------------------------------------
// grp_row_ptr / grp_col_ptr / row_nnz_ptr are the caller's HOST arrays of the whole arena.  Wire form v2
// (grp_ent_ptr != NULL): columns and entries are the live ones, row_nnz_ptr is not used, idx_byte0 is the offset of
// group g0's index block in the caller's blob.
static int slot_prepare(Slot &s, const int64_t *grp_row_ptr, const int64_t *grp_col_ptr, const int64_t *row_nnz_ptr,
                        int64_t g0, int64_t g1, const int64_t *grp_ent_ptr = nullptr, int64_t idx_byte0 = 0)
{
    TRY(slot_init(s));
    TRY(configure_kernels());
    s.v2 = (grp_ent_ptr != nullptr);
    s.g0 = g0; s.g1 = g1;
    s.r0 = grp_row_ptr[g0]; s.r1 = grp_row_ptr[g1];
    s.c0 = grp_col_ptr[g0]; s.c1 = grp_col_ptr[g1];
    if (s.v2) { s.e0 = grp_ent_ptr[g0]; s.e1 = grp_ent_ptr[g1]; }
    else { s.e0 = row_nnz_ptr[s.r0]; s.e1 = row_nnz_ptr[s.r1]; }
    const int64_t ng = s.ng();
    if (ng > INT32_MAX / kNumClasses) return fail(BAMBU_EM_EINVAL, "too many groups in one chunk");
    if (!s.v2)
        for (int64_t r = s.r0; r < s.r1; ++r)
            if (row_nnz_ptr[r + 1] < row_nnz_ptr[r]) return fail(BAMBU_EM_EINVAL, "row_nnz_ptr decreases at row %lld", (long long)r);
    int64_t *idx_off = nullptr;
    int64_t ib = idx_byte0;
    if (s.v2) {
        TRY(s.h_idx_off.ensure(sizeof(int64_t) * (size_t)(ng + 1)));
        idx_off = s.h_idx_off.as<int64_t>();
    }

    // ---- classify by worst-case footprint, lay out the stage-1 images ---------------
    TRY(s.h_img_off.ensure(sizeof(int64_t) * (size_t)(ng + 1)));
    TRY(s.h_work.ensure(sizeof(int32_t) * (size_t)std::max<int64_t>(ng, 1)));
    int64_t *img_off = s.h_img_off.as<int64_t>();
    std::vector<int32_t> per_class[8];
    std::vector<int32_t> single;
    std::vector<int64_t> gnnz((size_t)ng, 0);
    uint32_t smem_pack[8] = {}, smem_sell[8] = {};
    int64_t off = 0, goff = 0;
    std::vector<GiantDesc> &giants = s.giants;
    giants.clear();
    s.hM.assign((size_t)ng, 0); s.hN.assign((size_t)ng, 0); s.hE.assign((size_t)ng, 0);
    uint32_t gsmem = 0;
    const bool forced_giant = force_giant();
    for (int64_t g = g0; g < g1; ++g) {
        const int64_t M = grp_row_ptr[g + 1] - grp_row_ptr[g], N = grp_col_ptr[g + 1] - grp_col_ptr[g];
        if (M < 0 || N < 0) return fail(BAMBU_EM_EINVAL, "group %lld has negative extent", (long long)g);
        const int64_t e = s.v2 ? grp_ent_ptr[g + 1] - grp_ent_ptr[g]
                               : row_nnz_ptr[grp_row_ptr[g + 1]] - row_nnz_ptr[grp_row_ptr[g]];
        if (e < 0) return fail(BAMBU_EM_EINVAL, "group %lld has a negative entry count", (long long)g);
        if (s.v2) {     // index blocks: 4-byte aligned, u16 per entry while the live column index fits 15 bits
            ib = (ib + 3) / 4 * 4;
            idx_off[g - g0] = ib;
            ib += e * wire_idx_width(N);
        }
        gnnz[g - g0] = e;
        img_off[g - g0] = off;
        s.hM[g - g0] = (int32_t)std::min<int64_t>(M, INT32_MAX); s.hN[g - g0] = (int32_t)std::min<int64_t>(N, INT32_MAX); s.hE[g - g0] = e;
        if (M == 0) continue;
        if (M == 1 && s.single_tx_shortcut) { single.push_back((int32_t)g); continue; }
        const double ub = 20.0 * (double)e + 18.0 * (double)(M + N) + 64.0;
        if (M > kMaxIndex || N > kMaxIndex || M + N > kMaxVecElems || ub > (double)kSmemMax || forced_giant) {
            // streamed tier: image in global memory, all SMs iterate on this one group
            if (8 * (M + 1) + 16 + (int64_t)giant_stage_bytes(1) * (kGiantThreads / 32) > (int64_t)kSmemMax || M > INT32_MAX / 8 || N > INT32_MAX / 2 || e > (int64_t)UINT32_MAX - 64)
                return fail(BAMBU_EM_EINVAL, "group %lld (M=%lld N=%lld entries=%lld) exceeds the streamed tier",
                            (long long)g, (long long)M, (long long)N, (long long)e);
            GiantDesc d;
            memset(&d, 0, sizeof d);
            d.g = (int32_t)g; d.M = (int32_t)M; d.N = (int32_t)N;
            giant_layout(d, M, N, e);
            d.base = goff;
            goff += d.total;
            giants.push_back(d);
            gsmem = std::max<uint32_t>(gsmem, (uint32_t)(8 * (M + 1) + 16));      // theta of the widest locus; staging is added at launch
            continue;
        }
        const uint32_t bytes = em_smem_bytes((uint32_t)M, (uint32_t)N, (uint32_t)e);
        int c = 0;
        while (c < kNumPackClasses - 1 && bytes > kPackClasses[c].max_bytes) ++c;
        per_class[c].push_back((int32_t)g);
        smem_pack[c] = std::max(smem_pack[c], pack_smem_bytes((uint32_t)M, (uint32_t)N));
        smem_sell[c] = std::max(smem_sell[c], sell_scratch_bytes((uint32_t)M, (uint32_t)N));
        off += img_layout((uint32_t)M, (uint32_t)N, (uint32_t)e).total;
    }
    img_off[ng] = off;
    if (s.v2) { idx_off[ng] = ib; s.ib0 = idx_byte0; s.ib1 = ib; }

    // work lists: per pack class, largest groups first (they are queued for the EM first)
    int32_t *work = s.h_work.as<int32_t>();
    size_t woff = 0;
    TRY(s.work.ensure(sizeof(int32_t) * (size_t)std::max<int64_t>(ng, 1), &s.device_bytes));
    for (int c = 0; c < kNumPackClasses; ++c) {
        auto &w = per_class[c];
        std::stable_sort(w.begin(), w.end(), [&](int32_t x, int32_t y) { return gnnz[x - g0] > gnnz[y - g0]; });
        std::copy(w.begin(), w.end(), work + woff);
        s.buckets[c].n = (int)w.size();
        s.buckets[c].d_work = s.work.as<int32_t>() + woff;
        s.buckets[c].smem_pack = smem_pack[c];
        s.buckets[c].smem_sell = smem_sell[c];
        woff += w.size();
    }
    std::copy(single.begin(), single.end(), work + woff);
    s.n_single = (int)single.size();
    s.d_single = s.work.as<int32_t>() + woff;
    woff += single.size();

    // ---- device buffers ------------------------------------------------------------------
    const int64_t rows = s.rows(), nnz = s.nnz(), cols = s.cols();
    // SELL pool: typical need is ~25 B per live entry + ~40 B per row / live column; grown on demand
    int64_t pool = (int64_t)(24.0 * (double)nnz + 48.0 * (double)(rows + cols)) + (4 << 20);
    if (const char *e = getenv("BAMBU_EM_POOL_BYTES")) pool = std::max<int64_t>(atoll(e), 4096);   // tests: force the grow-and-rerun path
    pool = (pool + 15) / 16 * 16;
    s.pool_bytes = getenv("BAMBU_EM_POOL_BYTES") ? pool : std::max<int64_t>(pool, (int64_t)s.pool.cap / 16 * 16);
    TRY(s.grp_row_ptr.ensure(8 * (size_t)(ng + 1), &s.device_bytes));
    TRY(s.grp_col_ptr.ensure(8 * (size_t)(ng + 1), &s.device_bytes));
    TRY(s.row_nnz_ptr.ensure(8 * (size_t)((s.v2 ? ng : rows) + 1), &s.device_bytes));
    if (s.v2) TRY(s.grp_idx_off.ensure(8 * (size_t)(ng + 1), &s.device_bytes));
    TRY(s.img.ensure((size_t)off, &s.device_bytes));
    TRY(s.pool.ensure((size_t)s.pool_bytes, &s.device_bytes));
    TRY(s.img_off.ensure(8 * (size_t)(ng + 1), &s.device_bytes));
    TRY(s.info.ensure(sizeof(PackInfo) * (size_t)ng, &s.device_bytes));
    TRY(s.ginfo.ensure(sizeof(GroupImg) * (size_t)ng, &s.device_bytes));
    TRY(s.queue.ensure(4 * (size_t)kNumClasses * (size_t)std::max<int64_t>(ng, 1), &s.device_bytes));
    TRY(s.nnz_nonzero.ensure(8 * (size_t)ng, &s.device_bytes));
    TRY(s.est.ensure(8 * 3 * (size_t)rows, &s.device_bytes));
    TRY(s.iters.ensure(4 * (size_t)ng, &s.device_bytes));
    TRY(s.status.ensure(4 * (size_t)ng, &s.device_bytes));
    TRY(s.ctrl.ensure(sizeof(Ctrl), &s.device_bytes));
    s.n_giant = (int)giants.size();
    s.giant_smem = gsmem;
    s.giant_bytes = s.giant_bytes_prepared = goff;
    s.n_giant_prepared = s.n_giant;
    if (s.n_giant) {
        if (const char *e = getenv("BAMBU_EM_GIANT_SORT")) {       // experiments: 1 = largest first, 2 = smallest first
            if (e[0] == '1') std::stable_sort(giants.begin(), giants.end(), [&](const GiantDesc &a, const GiantDesc &b) { return s.hE[a.g - g0] > s.hE[b.g - g0]; });
            if (e[0] == '2') std::stable_sort(giants.begin(), giants.end(), [&](const GiantDesc &a, const GiantDesc &b) { return s.hE[a.g - g0] < s.hE[b.g - g0]; });
        }
        TRY(s.h_gdesc.ensure(sizeof(GiantDesc) * giants.size()));
        memcpy(s.h_gdesc.p, giants.data(), sizeof(GiantDesc) * giants.size());
        TRY(s.gdesc.ensure(sizeof(GiantDesc) * giants.size(), &s.device_bytes));
        TRY(s.gbuf.ensure((size_t)goff, &s.device_bytes));
    }

    cudaStream_t st = s.stream;
    if (s.n_giant)
        CUDA_TRY(cudaMemcpyAsync(s.gdesc.p, s.h_gdesc.p, sizeof(GiantDesc) * giants.size(), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(s.grp_row_ptr.p, grp_row_ptr + g0, 8 * (size_t)(ng + 1), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(s.grp_col_ptr.p, grp_col_ptr + g0, 8 * (size_t)(ng + 1), cudaMemcpyHostToDevice, st));
    if (s.v2) {
        CUDA_TRY(cudaMemcpyAsync(s.row_nnz_ptr.p, grp_ent_ptr + g0, 8 * (size_t)(ng + 1), cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(s.grp_idx_off.p, idx_off, 8 * (size_t)(ng + 1), cudaMemcpyHostToDevice, st));
    } else {
        CUDA_TRY(cudaMemcpyAsync(s.row_nnz_ptr.p, row_nnz_ptr + s.r0, 8 * (size_t)(rows + 1), cudaMemcpyHostToDevice, st));
    }
    CUDA_TRY(cudaMemcpyAsync(s.img_off.p, img_off, 8 * (size_t)(ng + 1), cudaMemcpyHostToDevice, st));
    if (woff) CUDA_TRY(cudaMemcpyAsync(s.work.p, work, 4 * woff, cudaMemcpyHostToDevice, st));
    s.prepared = true;
    s.ran = s.finished = false;
    return BAMBU_EM_OK;
}

// H2D of the range's payload from the caller's HOST arrays of the whole arena
static int slot_upload(Slot &s, const int32_t *col_idx, const double *aval, const uint8_t *nz_flags, const double *Y,
                       const double *K, const uint8_t *col_flags)
{
    const size_t nnz = (size_t)s.nnz(), nc = (size_t)s.cols();
    if ((nnz && (!col_idx || !aval || !nz_flags)) || (nc && (!Y || !K || !col_flags)))
        return fail(BAMBU_EM_EINVAL, "NULL payload pointer");
    // one allocation: aval | Y | K | col_idx | nz_flags | col_flags  (8-byte items first)
    TRY(s.payload.ensure(8 * nnz + 16 * nc + 4 * nnz + nnz + nc + 64, &s.device_bytes));
    uint8_t *q = s.payload.as<uint8_t>();
    double *d_aval = (double *)q; q += 8 * nnz;
    double *d_Y = (double *)q; q += 8 * nc;
    double *d_K = (double *)q; q += 8 * nc;
    int32_t *d_ci = (int32_t *)q; q += 4 * nnz;
    uint8_t *d_nf = q; q += nnz;
    uint8_t *d_cf = q;
    cudaStream_t st = s.stream;
    if (nnz) {
        CUDA_TRY(cudaMemcpyAsync(d_aval, aval + s.e0, 8 * nnz, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(d_ci, col_idx + s.e0, 4 * nnz, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(d_nf, nz_flags + s.e0, nnz, cudaMemcpyHostToDevice, st));
    }
    if (nc) {
        CUDA_TRY(cudaMemcpyAsync(d_Y, Y + s.c0, 8 * nc, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(d_K, K + s.c0, 8 * nc, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(d_cf, col_flags + s.c0, nc, cudaMemcpyHostToDevice, st));
    }
    s.v_aval = d_aval - s.e0; s.v_col_idx = d_ci - s.e0; s.v_nz_flags = d_nf - s.e0;
    s.v_Y = d_Y - s.c0; s.v_K = d_K - s.c0; s.v_col_flags = d_cf - s.c0;
    s.payload_bound = false;
    s.payload_ready = true;
    return BAMBU_EM_OK;
}

// H2D of the range's payload in the wire form v2 from the caller's HOST arrays (include/bambu_em.h)
struct WireHost {
    const uint32_t *row_end;
    const double *rs, *aval;
    const uint8_t *cidx;
    const uint32_t *eq_bits;
    int32_t col_mode;
    const void *colY, *colK;
    const uint32_t *multi_bits;
};

static int slot_upload_v2(Slot &s, const WireHost &h)
{
    const size_t ne = (size_t)s.nnz(), nc = (size_t)s.cols(), nr = (size_t)s.rows();
    if ((ne && (!h.aval || !h.cidx || !h.eq_bits)) || (nc && (!h.colY || !h.colK || !h.multi_bits)) || (nr && (!h.row_end || !h.rs)))
        return fail(BAMBU_EM_EINVAL, "NULL payload pointer");
    if (h.col_mode != 0 && h.col_mode != 1) return fail(BAMBU_EM_EINVAL, "col_mode must be 0 (Y, K as doubles) or 1 (nobs, K as int32)");
    const size_t cw = h.col_mode == 0 ? 8 : 4;
    const size_t ibytes = (size_t)(s.ib1 - s.ib0);
    const int64_t ew0 = s.e0 >> 5, ew1 = (s.e1 + 31) >> 5, cw0 = s.c0 >> 5, cw1 = (s.c1 + 31) >> 5;
    // one allocation, 8-byte items first: aval | rs | colY | colK | row_end | eq words | multi words | index blob
    const size_t need = 8 * ne + 8 * nr + 2 * cw * nc + 4 * nr + 4 * (size_t)(ew1 - ew0) + 4 * (size_t)(cw1 - cw0) + ibytes + 64;
    TRY(s.payload.ensure(need, &s.device_bytes));
    uint8_t *q = s.payload.as<uint8_t>();
    double *d_aval = (double *)q; q += 8 * ne;
    double *d_rs = (double *)q; q += 8 * nr;
    uint8_t *d_cY = q; q += cw * nc;
    uint8_t *d_cK = q; q += cw * nc;
    uint32_t *d_re = (uint32_t *)q; q += 4 * nr;
    uint32_t *d_eq = (uint32_t *)q; q += 4 * (size_t)(ew1 - ew0);
    uint32_t *d_mb = (uint32_t *)q; q += 4 * (size_t)(cw1 - cw0);
    uint8_t *d_ix = q;
    cudaStream_t st = s.stream;
    if (ne) {
        CUDA_TRY(cudaMemcpyAsync(d_aval, h.aval + s.e0, 8 * ne, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(d_eq, h.eq_bits + ew0, 4 * (size_t)(ew1 - ew0), cudaMemcpyHostToDevice, st));
        if (ibytes) CUDA_TRY(cudaMemcpyAsync(d_ix, h.cidx + s.ib0, ibytes, cudaMemcpyHostToDevice, st));
    }
    if (nr) {
        CUDA_TRY(cudaMemcpyAsync(d_rs, h.rs + s.r0, 8 * nr, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(d_re, h.row_end + s.r0, 4 * nr, cudaMemcpyHostToDevice, st));
    }
    if (nc) {
        CUDA_TRY(cudaMemcpyAsync(d_cY, (const uint8_t *)h.colY + cw * (size_t)s.c0, cw * nc, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(d_cK, (const uint8_t *)h.colK + cw * (size_t)s.c0, cw * nc, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(d_mb, h.multi_bits + cw0, 4 * (size_t)(cw1 - cw0), cudaMemcpyHostToDevice, st));
    }
    s.w_aval = d_aval - s.e0; s.w_rs = d_rs - s.r0; s.w_row_end = d_re - s.r0;
    s.w_colY = d_cY - cw * (size_t)s.c0; s.w_colK = d_cK - cw * (size_t)s.c0;
    s.w_eq_bits = d_eq - ew0; s.w_multi_bits = d_mb - cw0;
    s.w_cidx = d_ix - s.ib0;
    s.col_mode = h.col_mode;
    s.payload_bound = false;
    s.payload_ready = true;
    return BAMBU_EM_OK;
}

// ---- streamed tier launches over descs [first, first + count) ---------------------------------
static int launch_giant_pack(Slot &s, const SlotViews &v, int first, int count, cudaStream_t st)
{
    if (s.v2) {
        // wire form v2: every phase is a flat grid over (locus, tile) -- the whole GPU packs (giant_pack.cuh)
        GiantDesc *d = s.gdesc.as<GiantDesc>() + first;
        uint8_t *gb = s.gbuf.as<uint8_t>();
        TRY(s.ghist.ensure(4 * 1024 * (size_t)count, &s.device_bytes));
        uint32_t *hist = s.ghist.as<uint32_t>();
        CUDA_TRY(cudaMemsetAsync(hist, 0, 4 * 1024 * (size_t)count, st));
        const int tiles = std::max(1, std::min(1024, (8 * g_num_sms + count - 1) / count));
        const dim3 flat((unsigned)tiles, (unsigned)count), one(1, (unsigned)count);
        gp_cols<<<flat, kGpThreads, 0, st>>>(v.W, d, gb);
        gp_rowcount<<<flat, kGpThreads, 0, st>>>(v.W, d, gb, &v.ctrl->err);
        gp_scan_rows<<<one, 1024, 0, st>>>(d, gb, v.nnz_nonzero, v.W);
        gp_rowfill<<<flat, kGpThreads, 0, st>>>(v.W, d, gb, &v.ctrl->err);
        gp_scan_cols<<<one, 1024, 0, st>>>(d, gb);
        gp_scatter<<<flat, kGpThreads, 0, st>>>(d, gb);
        gp_sortcols<<<flat, kGpThreads, 0, st>>>(d, gb, hist);
        gp_sell_scan_hist<<<one, 1024, 0, st>>>(hist);
        gp_sell_place<<<flat, kGpThreads, 0, st>>>(d, gb, hist);
        gp_sell_slices<<<flat, kGpThreads, 0, st>>>(d, gb);
        gp_sell_scan_slices<<<one, 1024, 0, st>>>(d, gb);
        gp_sell_fill<<<flat, kGpThreads, 0, st>>>(d, gb);
        CUDA_TRY(cudaGetLastError());
        s.launches += 12;
        return BAMBU_EM_OK;
    }
    pack_giant_kernel<1024><<<count, 1024, 0, st>>>(v.A, s.gdesc.as<GiantDesc>() + first, count, s.gbuf.as<uint8_t>(),
                                                     v.nnz_nonzero, &v.ctrl->err);
    CUDA_TRY(cudaGetLastError());
    s.launches += 1;
    return BAMBU_EM_OK;
}

static int launch_giant_em(Slot &s, const SlotViews &v, int first, int count, cudaStream_t st)
{
    // cooperative launch: every CTA must be resident at once, so this runs after the resident
    // classes have drained (ordered on the main stream)
    const GiantDesc *descs = s.gdesc.as<GiantDesc>() + first;
    int n_giant = count;
    const uint8_t *gbuf_c = s.gbuf.as<uint8_t>();
    uint8_t *gbuf = s.gbuf.as<uint8_t>();
    const int64_t *grp = v.A.grp_row_ptr;
    int64_t rows = s.rows();
    EmParams P = s.P;
    double *est = v.est;
    int32_t *iters = v.iters, *status = v.status;
    unsigned int *gflags = v.ctrl->gflags;
    unsigned int *gticket = &v.ctrl->gticket;
    // Loci are iterated by teams of CTAs (giant.cuh): an update of one locus is bound by barrier round trips and add
    // chains, not by bandwidth, so several loci in flight fill the GPU better.  Measured on config 3 (50 loci of 2-3 M
    // entries): 1 team 421 ms, 2: 288, 4: 250-278, 8: 249, 10: 234-242, 12: 241, 16: 240-246; on a 6-locus shard:
    // 2 teams 35.0, 4: 39-41, 6 and more: 30.5 ms.  A team that finishes a locus takes the next one from a ticket
    // (iteration counts -- 300 to 1600 updates per locus -- are not known in advance, so no static split is safe;
    // sorting the loci by size in either direction changed nothing).
    int n_teams = 1;
    if (const char *e = getenv("BAMBU_EM_GIANT_TEAMS")) n_teams = atoi(e);
    else n_teams = 12;
    n_teams = std::max(1, std::min(std::min(n_teams, 16), n_giant));
    // rows per warp block in the M-phase (giant.cuh): 2 rows = 4 parity streams per warp, the loads of the next batch in
    // flight while the current one is multiplied and added.  Config 3 (50 loci): one row per warp 238 ms (round-1 code,
    // 10 teams), blocks of 2 / 4 / 8 rows without the prefetch 221 / 221 / 233, 2 rows with it 203 (10 teams), 199 (12).
    int rb = 2;
    if (const char *e = getenv("BAMBU_EM_GIANT_RB")) rb = atoi(e) >= 2 ? 2 : 1;
    while (rb > 1 && (size_t)s.giant_smem + (size_t)giant_stage_bytes(rb) * (kGiantThreads / 32) > (size_t)kSmemMax) rb /= 2;
    void *args[] = {&descs, &n_giant, &gbuf_c, &gbuf, &grp, &rows, &P, &est, &iters, &status, &gflags, &n_teams, &gticket, &rb};
    // dynamic shared memory: one private copy of theta per CTA + the per-warp staging
    const size_t smem = std::max<size_t>((size_t)s.giant_smem + (size_t)giant_stage_bytes(rb) * (kGiantThreads / 32), 1024);
    int per_sm = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, em_giant_kernel<kGiantThreads>, kGiantThreads, smem));
    if (per_sm < 1) return fail(BAMBU_EM_ECUDA, "streamed EM kernel does not fit an SM");
    CUDA_TRY(cudaLaunchCooperativeKernel((const void *)em_giant_kernel<kGiantThreads>, dim3(g_num_sms), dim3(kGiantThreads),
                                         args, smem, st));
    s.launches += 1;
    return BAMBU_EM_OK;
}

static int slot_enqueue(Slot &s)
{
    const SlotViews v = slot_views(s);
    cudaStream_t st = s.stream;
    const size_t ng = (size_t)std::max<int64_t>(s.ng(), 1);
    s.launches = 0;
    CUDA_TRY(cudaEventRecord(s.ev_start, st));
    CUDA_TRY(cudaMemsetAsync(s.ctrl.p, 0, sizeof(Ctrl), st));
    CUDA_TRY(cudaMemsetAsync(s.iters.p, 0, 4 * ng, st));
    CUDA_TRY(cudaMemsetAsync(s.status.p, 0, 4 * ng, st));
    CUDA_TRY(cudaMemsetAsync(s.est.p, 0, 8 * (size_t)std::max<int64_t>(3 * s.rows(), 1), st));
    CUDA_TRY(cudaMemsetAsync(s.info.p, 0, sizeof(PackInfo) * ng, st));
    CUDA_TRY(cudaMemsetAsync(s.ginfo.p, 0, sizeof(GroupImg) * ng, st));
    CUDA_TRY(cudaMemsetAsync(s.nnz_nonzero.p, 0, 8 * ng, st));
    CUDA_TRY(cudaEventRecord(s.ev_fork, st));
    bool any = false;
    for (int b = 0; b < kNumPackClasses; ++b) {
        if (!s.buckets[b].n) continue;
        any = true;
        CUDA_TRY(cudaStreamWaitEvent(s.buckets[b].stream, s.ev_fork, 0));
        TRY(launch_pack(s, b, v));
        CUDA_TRY(cudaEventRecord(s.buckets[b].done, s.buckets[b].stream));
        CUDA_TRY(cudaStreamWaitEvent(st, s.buckets[b].done, 0));
    }
    if (s.n_single) {
        if (s.v2)
            pack_resident_wire_kernel<64><<<s.n_single, 64, 0, st>>>(v.W, s.d_single, s.n_single, v.img_off, s.img.as<uint8_t>(),
                                                                     v.info, v.nnz_nonzero, v.est, v.status, &v.ctrl->err, 1);
        else
            pack_resident_kernel<64><<<s.n_single, 64, 0, st>>>(v.A, s.d_single, s.n_single, v.img_off, s.img.as<uint8_t>(),
                                                                v.info, v.nnz_nonzero, v.est, v.status, &v.ctrl->err, 1);
        CUDA_TRY(cudaGetLastError());
        s.launches += 1;
    }
    if (s.n_giant) {   // streamed tier: pack now (one CTA per giant group), iterate after the resident classes
        CUDA_TRY(cudaStreamWaitEvent(s.giant_stream, s.ev_fork, 0));
        TRY(launch_giant_pack(s, v, 0, s.n_giant, s.giant_stream));
        CUDA_TRY(cudaEventRecord(s.giant_packed, s.giant_stream));
        CUDA_TRY(cudaStreamWaitEvent(st, s.giant_packed, 0));
    }
    CUDA_TRY(cudaEventRecord(s.ev_packed, st));
    if (any) {   // EM, one class per stream; the queues were filled on the device
        for (int c = kNumClasses - 1; c >= 0; --c) {
            CUDA_TRY(cudaStreamWaitEvent(s.em_stream[c], s.ev_packed, 0));
            TRY(launch_em(s, c, v));
            CUDA_TRY(cudaEventRecord(s.em_done[c], s.em_stream[c]));
            CUDA_TRY(cudaStreamWaitEvent(st, s.em_done[c], 0));
        }
    }
    if (s.n_giant) TRY(launch_giant_em(s, v, 0, s.n_giant, st));
    CUDA_TRY(cudaEventRecord(s.ev_end, st));
    s.ran = true;
    s.finished = false;
    return BAMBU_EM_OK;
}

static int slot_run(Slot &s, int32_t maxiter, double minvalue, double conv)
{
    if (!s.prepared) return fail(BAMBU_EM_ESTATE, "run before the plan was prepared");
    if (!s.payload_ready) return fail(BAMBU_EM_ESTATE, "run before upload/bind_device");
    if (conv != conv || minvalue != minvalue) return fail(BAMBU_EM_EINVAL, "conv/minvalue is NaN");
    s.P.maxiter = maxiter;
    s.P.minvalue = minvalue;
    s.P.conv = conv;
    if (conv > 1e-290 && std::isfinite(conv)) {
        s.P.conv_hi = conv * (1.0 + std::ldexp(1.0, -40));
        s.P.conv_lo = conv * (1.0 - std::ldexp(1.0, -40));
    } else {   // no screen: always take the exact division
        s.P.conv_hi = INFINITY;
        s.P.conv_lo = 0.0;
    }
    return slot_enqueue(s);
}

// Groups the host admitted to the resident tier by their entry count can still need more than one
// SM's shared memory once padded into slices (one very long line is enough).  The sell kernel
// flags them instead of queueing them; here they are packed and iterated by the streamed tier.
static int slot_spill_to_streamed(Slot &s)
{
    const size_t ng = (size_t)s.ng();
    std::vector<GroupImg> gi(ng);
    CUDA_TRY(cudaMemcpyAsync(gi.data(), s.ginfo.p, sizeof(GroupImg) * ng, cudaMemcpyDeviceToHost, s.stream));
    CUDA_TRY(cudaStreamSynchronize(s.stream));
    const int first = (int)s.giants.size();
    for (size_t k = 0; k < ng; ++k) {
        if (gi[k].cls != -1) continue;
        const int64_t M = s.hM[k], N = s.hN[k], e = s.hE[k];
        if (8 * (M + 1) + 16 + (int64_t)giant_stage_bytes(1) * (kGiantThreads / 32) > (int64_t)kSmemMax)
            return fail(BAMBU_EM_EINVAL, "group %lld (M=%lld) exceeds the streamed tier", (long long)(s.g0 + (int64_t)k), (long long)M);
        GiantDesc d;
        memset(&d, 0, sizeof d);
        d.g = (int32_t)(s.g0 + (int64_t)k); d.M = (int32_t)M; d.N = (int32_t)N;
        giant_layout(d, M, N, e);
        d.base = s.giant_bytes;
        s.giant_bytes += d.total;
        s.giants.push_back(d);
        s.giant_smem = std::max<uint32_t>(s.giant_smem, (uint32_t)(8 * (M + 1) + 16));
    }
    const int count = (int)s.giants.size() - first;
    if (count == 0) return fail(BAMBU_EM_ECUDA, "device reported an oversized group but none is marked");
    // the earlier streamed-tier groups are finished: their images may be dropped when the buffer grows
    TRY(s.h_gdesc.ensure(sizeof(GiantDesc) * s.giants.size()));
    memcpy(s.h_gdesc.p, s.giants.data(), sizeof(GiantDesc) * s.giants.size());
    TRY(s.gdesc.ensure(sizeof(GiantDesc) * s.giants.size(), &s.device_bytes));
    TRY(s.gbuf.ensure((size_t)s.giant_bytes, &s.device_bytes));
    cudaStream_t st = s.stream;
    CUDA_TRY(cudaMemcpyAsync(s.gdesc.p, s.h_gdesc.p, sizeof(GiantDesc) * s.giants.size(), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemsetAsync(s.ctrl.p, 0, sizeof(Ctrl), st));      // error word, flags and barrier words
    const SlotViews v = slot_views(s);
    TRY(launch_giant_pack(s, v, first, count, st));
    TRY(launch_giant_em(s, v, first, count, st));
    CUDA_TRY(cudaEventRecord(s.ev_end, st));
    s.n_giant = (int)s.giants.size();
    return BAMBU_EM_OK;
}

// wait for the run; if the SELL pool was too small, grow it and run again; move groups whose padded
// image exceeds shared memory to the streamed tier
static int slot_finish(Slot &s)
{
    if (!s.ran || s.finished) return BAMBU_EM_OK;
    for (int attempt = 0; attempt < 12; ++attempt) {
        int err = 0;
        CUDA_TRY(cudaMemcpyAsync(&err, &s.ctrl.as<Ctrl>()->err, sizeof(int), cudaMemcpyDeviceToHost, s.stream));
        CUDA_TRY(cudaStreamSynchronize(s.stream));
        if (err == 0) { s.finished = true; return BAMBU_EM_OK; }
        if (err & kErrColRange) return fail(BAMBU_EM_EINVAL, "col_idx outside its group's column range");
        if (err & kErrColOrder) return fail(BAMBU_EM_EINVAL, "col_idx not strictly increasing inside a row");
        if (err & kErrRowPtr) return fail(BAMBU_EM_EINVAL, "row_end decreases or runs past its group's entries");
        if (err & kErrParity) return fail(BAMBU_EM_EINVAL, "a live column carries both index parities");
        if (err & kErrPool) {
            s.pool_bytes *= 2;
            TRY(s.pool.ensure((size_t)s.pool_bytes, &s.device_bytes));
            // a rerun starts from the chunk's own streamed-tier list again
            s.giants.resize((size_t)s.n_giant_prepared);
            s.n_giant = s.n_giant_prepared;
            s.giant_bytes = s.giant_bytes_prepared;
            TRY(slot_enqueue(s));
            continue;
        }
        if (err & kErrTooBig) { TRY(slot_spill_to_streamed(s)); continue; }
        return fail(BAMBU_EM_ECUDA, "device error word %d", err);
    }
    return fail(BAMBU_EM_ENOMEM, "SELL image pool kept overflowing");
}

// D2H of the range's outputs into the caller's HOST arrays of the whole arena (asynchronous)
static int slot_download_async(Slot &s, double *est, int64_t total_rows, int32_t *iters, int32_t *status, int64_t *nnz)
{
    cudaStream_t st = s.stream;
    const size_t rows = (size_t)s.rows(), ng = (size_t)s.ng();
    if (est && rows)
        for (int k = 0; k < 3; ++k)
            CUDA_TRY(cudaMemcpyAsync(est + (size_t)k * total_rows + s.r0, s.est.as<double>() + (size_t)k * rows, 8 * rows,
                                     cudaMemcpyDeviceToHost, st));
    if (iters && ng) CUDA_TRY(cudaMemcpyAsync(iters + s.g0, s.iters.p, 4 * ng, cudaMemcpyDeviceToHost, st));
    if (status && ng) CUDA_TRY(cudaMemcpyAsync(status + s.g0, s.status.p, 4 * ng, cudaMemcpyDeviceToHost, st));
    if (nnz && ng) CUDA_TRY(cudaMemcpyAsync(nnz + s.g0, s.nnz_nonzero.p, 8 * ng, cudaMemcpyDeviceToHost, st));
    return BAMBU_EM_OK;
}

// ------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------
struct bambu_em_plan {
    Slot slot;
    int64_t n_groups = 0;
};

constexpr int kPipeSlots = 3;
static Slot g_pipe[kPipeSlots];   // cached slots of bambu_em_batched
static bool g_pipe_used = false;

extern "C" {

const char *bambu_em_version(void) { return "bambu_em_b200 0.3.0 (sm_100a)"; }
const char *bambu_em_last_error(void) { return g_last_error.c_str(); }

int bambu_em_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    int ok = 0;
    for (int d = 0; d < n; ++d) {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, d) == cudaSuccess && prop.major == 10 && prop.minor == 0) ++ok;
    }
    return ok;
}

int bambu_em_shutdown(void)
{
    if (g_pipe_used) {
        cudaSetDevice(g_device);
        for (Slot &s : g_pipe) slot_destroy(s);
        g_pipe_used = false;
    }
    bambu_quant_release_workspace();
    return BAMBU_EM_OK;
}

int bambu_em_set_device(int device)
{
    if (device < 0) return fail(BAMBU_EM_EINVAL, "negative device index");
    if (device != g_device) {
        bambu_em_shutdown();      // the cached workspace lives on the old device
        g_em_configured = false;
    }
    g_device = device;
    g_device_checked = false;
    return ensure_device();
}

static int check_arrays(int64_t n_groups, const int64_t *grp_row_ptr, const int64_t *grp_col_ptr,
                        const int64_t *row_nnz_ptr)
{
    if (n_groups < 0 || n_groups > INT32_MAX) return fail(BAMBU_EM_EINVAL, "n_groups out of range");
    if (!grp_row_ptr || !grp_col_ptr || !row_nnz_ptr) return fail(BAMBU_EM_EINVAL, "NULL pointer array");
    if (grp_row_ptr[0] != 0 || grp_col_ptr[0] != 0 || row_nnz_ptr[0] != 0)
        return fail(BAMBU_EM_EINVAL, "pointer arrays must start at 0");
    // one O(G) pass before any of these is used as an index into another array
    for (int64_t g = 0; g < n_groups; ++g)
        if (grp_row_ptr[g + 1] < grp_row_ptr[g] || grp_col_ptr[g + 1] < grp_col_ptr[g])
            return fail(BAMBU_EM_EINVAL, "group pointer arrays must be non-decreasing (group %lld)", (long long)g);
    return BAMBU_EM_OK;
}

// the same for the wire form v2: all three arrays are per group
static int check_arrays_v2(int64_t n_groups, const int64_t *grp_row_ptr, const int64_t *grp_col_ptr,
                           const int64_t *grp_ent_ptr)
{
    TRY(check_arrays(n_groups, grp_row_ptr, grp_col_ptr, grp_ent_ptr));
    for (int64_t g = 0; g < n_groups; ++g)
        if (grp_ent_ptr[g + 1] < grp_ent_ptr[g])
            return fail(BAMBU_EM_EINVAL, "grp_ent_ptr must be non-decreasing (group %lld)", (long long)g);
    return BAMBU_EM_OK;
}

int bambu_em_plan_destroy(bambu_em_plan *p)
{
    if (!p) return BAMBU_EM_OK;
    cudaSetDevice(g_device);
    slot_destroy(p->slot);
    delete p;
    return BAMBU_EM_OK;
}

static int plan_create_impl(bambu_em_plan **out, int64_t n_groups, const int64_t *grp_row_ptr,
                            const int64_t *grp_col_ptr, const int64_t *row_nnz_ptr, bool single_tx_shortcut)
{
    if (!out) return fail(BAMBU_EM_EINVAL, "plan pointer is NULL");
    *out = nullptr;
    TRY(check_arrays(n_groups, grp_row_ptr, grp_col_ptr, row_nnz_ptr));
    TRY(ensure_device());
    bambu_em_plan *p = new (std::nothrow) bambu_em_plan;
    if (!p) return fail(BAMBU_EM_ENOMEM, "out of host memory");
    p->n_groups = n_groups;
    p->slot.single_tx_shortcut = single_tx_shortcut;
    int rc = slot_prepare(p->slot, grp_row_ptr, grp_col_ptr, row_nnz_ptr, 0, n_groups);
    if (!rc && cudaStreamSynchronize(p->slot.stream) != cudaSuccess) rc = fail(BAMBU_EM_ECUDA, "structure upload failed");
    if (rc) { bambu_em_plan_destroy(p); return rc; }
    *out = p;
    return BAMBU_EM_OK;
}

int bambu_em_plan_create(bambu_em_plan **out, int64_t n_groups, const int64_t *grp_row_ptr,
                         const int64_t *grp_col_ptr, const int64_t *row_nnz_ptr)
{
    return plan_create_impl(out, n_groups, grp_row_ptr, grp_col_ptr, row_nnz_ptr, true);
}

int bambu_em_plan_reset(bambu_em_plan *p, int64_t n_groups, const int64_t *grp_row_ptr, const int64_t *grp_col_ptr,
                        const int64_t *row_nnz_ptr)
{
    if (!p) return fail(BAMBU_EM_EINVAL, "plan is NULL");
    TRY(check_arrays(n_groups, grp_row_ptr, grp_col_ptr, row_nnz_ptr));
    TRY(ensure_device());
    Slot &s = p->slot;
    CUDA_TRY(cudaStreamSynchronize(s.stream));
    s.payload_bound = s.payload_ready = false;
    p->n_groups = n_groups;
    TRY(slot_prepare(s, grp_row_ptr, grp_col_ptr, row_nnz_ptr, 0, n_groups));
    if (cudaStreamSynchronize(s.stream) != cudaSuccess) return fail(BAMBU_EM_ECUDA, "structure upload failed");
    return BAMBU_EM_OK;
}

int bambu_em_plan_set_stream(bambu_em_plan *p, void *cuda_stream)
{
    if (!p) return fail(BAMBU_EM_EINVAL, "plan is NULL");
    CUDA_TRY(cudaStreamSynchronize(p->slot.stream));
    p->slot.stream = cuda_stream ? (cudaStream_t)cuda_stream : p->slot.own_stream;
    return BAMBU_EM_OK;
}

int bambu_em_plan_upload(bambu_em_plan *p, const int32_t *col_idx, const double *aval, const uint8_t *nz_flags,
                         const double *Y, const double *K, const uint8_t *col_flags)
{
    if (!p) return fail(BAMBU_EM_EINVAL, "plan is NULL");
    if (p->slot.v2) return fail(BAMBU_EM_ESTATE, "the plan was created for the wire form v2 (use bambu_em_plan_upload_v2)");
    TRY(ensure_device());
    return slot_upload(p->slot, col_idx, aval, nz_flags, Y, K, col_flags);
}

int bambu_em_plan_bind_device(bambu_em_plan *p, const int32_t *d_col_idx, const double *d_aval,
                              const uint8_t *d_nz_flags, const double *d_Y, const double *d_K,
                              const uint8_t *d_col_flags)
{
    if (!p) return fail(BAMBU_EM_EINVAL, "plan is NULL");
    Slot &s = p->slot;
    if (s.v2) return fail(BAMBU_EM_ESTATE, "the plan was created for the wire form v2");
    CUDA_TRY(cudaStreamSynchronize(s.stream));
    s.v_col_idx = d_col_idx; s.v_aval = d_aval; s.v_nz_flags = d_nz_flags;
    s.v_Y = d_Y; s.v_K = d_K; s.v_col_flags = d_col_flags;
    s.payload_bound = true;
    s.payload_ready = true;
    return BAMBU_EM_OK;
}

int bambu_em_plan_run(bambu_em_plan *p, int32_t maxiter, double minvalue, double conv)
{
    if (!p) return fail(BAMBU_EM_EINVAL, "plan is NULL");
    TRY(ensure_device());
    return slot_run(p->slot, maxiter, minvalue, conv);
}

int bambu_em_plan_sync(bambu_em_plan *p)
{
    if (!p) return fail(BAMBU_EM_EINVAL, "plan is NULL");
    if (p->slot.ran && !p->slot.finished) return slot_finish(p->slot);
    CUDA_TRY(cudaStreamSynchronize(p->slot.stream));
    return BAMBU_EM_OK;
}

int bambu_em_plan_download(bambu_em_plan *p, double *est, int32_t *iters, int32_t *status, int64_t *nnz_nonzero)
{
    if (!p) return fail(BAMBU_EM_EINVAL, "plan is NULL");
    Slot &s = p->slot;
    if (!s.ran) return fail(BAMBU_EM_ESTATE, "download before run");
    TRY(slot_finish(s));
    TRY(slot_download_async(s, est, s.rows(), iters, status, nnz_nonzero));
    CUDA_TRY(cudaStreamSynchronize(s.stream));
    return BAMBU_EM_OK;
}

int bambu_em_plan_download_live(bambu_em_plan *p, int32_t *live_cols, int32_t *live_nnz, int32_t *smem_bytes,
                                int32_t *slots)
{
    if (!p) return fail(BAMBU_EM_EINVAL, "plan is NULL");
    Slot &s = p->slot;
    if (!s.ran) return fail(BAMBU_EM_ESTATE, "download before run");
    TRY(slot_finish(s));
    const size_t ng = (size_t)s.ng();
    std::vector<PackInfo> h(ng);
    std::vector<GroupImg> gi(ng);
    if (ng) {
        CUDA_TRY(cudaMemcpyAsync(h.data(), s.info.p, sizeof(PackInfo) * ng, cudaMemcpyDeviceToHost, s.stream));
        CUDA_TRY(cudaMemcpyAsync(gi.data(), s.ginfo.p, sizeof(GroupImg) * ng, cudaMemcpyDeviceToHost, s.stream));
        CUDA_TRY(cudaStreamSynchronize(s.stream));
    }
    for (size_t g = 0; g < ng; ++g) {
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
    if (d_est) *d_est = p->slot.est.as<double>();
    if (d_iters) *d_iters = p->slot.iters.as<int32_t>();
    if (d_status) *d_status = p->slot.status.as<int32_t>();
    return BAMBU_EM_OK;
}

int bambu_em_plan_last_run_ms(bambu_em_plan *p, float *pack_ms, float *em_ms, float *total_ms)
{
    if (!p) return fail(BAMBU_EM_EINVAL, "plan is NULL");
    Slot &s = p->slot;
    if (!s.ran) return fail(BAMBU_EM_ESTATE, "no run yet");
    TRY(slot_finish(s));
    CUDA_TRY(cudaEventSynchronize(s.ev_end));
    float pk = 0.f, em = 0.f, tot = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&pk, s.ev_start, s.ev_packed));
    CUDA_TRY(cudaEventElapsedTime(&em, s.ev_packed, s.ev_end));
    CUDA_TRY(cudaEventElapsedTime(&tot, s.ev_start, s.ev_end));
    if (pack_ms) *pack_ms = pk;
    if (em_ms) *em_ms = em;
    if (total_ms) *total_ms = tot;
    return BAMBU_EM_OK;
}

int bambu_em_plan_last_run_launches(bambu_em_plan *p, int32_t *n)
{
    if (!p || !n) return fail(BAMBU_EM_EINVAL, "NULL argument");
    *n = p->slot.launches;
    return BAMBU_EM_OK;
}

int bambu_em_plan_device_bytes(bambu_em_plan *p, int64_t *bytes)
{
    if (!p || !bytes) return fail(BAMBU_EM_EINVAL, "NULL argument");
    *bytes = p->slot.device_bytes;
    return BAMBU_EM_OK;
}

// ---- one-shot: stream the arena through the cached three-slot pipeline ---------------
static int64_t chunk_target_bytes()
{
    const char *e = getenv("BAMBU_EM_CHUNK_MB");
    double mb = e ? atof(e) : 256.0;      // measured on the 100-sample cohort: 64 MB 219 ms, 128 MB 161, 256 MB 153, 512 MB 163
    if (!(mb > 0.0)) mb = 256.0;
    return (int64_t)(mb * 1048576.0);
}

static int64_t chunk_target_bytes_v2()
{
    const char *e = getenv("BAMBU_EM_CHUNK_MB");
    double mb = e ? atof(e) : 96.0;       // ~ the same number of groups per chunk as 256 MB of the v1 arena
    if (!(mb > 0.0)) mb = 96.0;
    return (int64_t)(mb * 1048576.0);
}

static int collect(Slot &s)   // finish a pipelined chunk: wait, (retry), download
{
    if (!s.busy) return BAMBU_EM_OK;
    s.busy = false;
    const bool had_to_wait = !s.finished;
    TRY(slot_finish(s));
    (void)had_to_wait;
    TRY(slot_download_async(s, s.dl_est, s.dl_total_rows, s.dl_iters, s.dl_status, nullptr));
    return BAMBU_EM_OK;
}

static int batched_impl(int64_t n_groups, const int64_t *grp_row_ptr, const int64_t *grp_col_ptr,
                        const int64_t *row_nnz_ptr, const int32_t *col_idx, const double *aval,
                        const uint8_t *nz_flags, const double *Y, const double *K, const uint8_t *col_flags,
                        int32_t maxiter, double minvalue, double conv, double *est, int32_t *iters, int32_t *status,
                        bool single_tx_shortcut)
{
    if (!est || !iters || !status) return fail(BAMBU_EM_EINVAL, "NULL output pointer");
    TRY(check_arrays(n_groups, grp_row_ptr, grp_col_ptr, row_nnz_ptr));
    TRY(ensure_device());
    g_pipe_used = true;
    const int64_t total_rows = grp_row_ptr[n_groups];
    const int64_t target_max = chunk_target_bytes();
    int rc = BAMBU_EM_OK;
    int64_t g0 = 0;
    int k = 0;
    while (g0 < n_groups && !rc) {
        // next chunk: groups until ~target bytes of arena.  The first chunks are small (32 MB, doubling) so the
        // GPU starts after a short copy; later ones are large enough to amortise the per-chunk host work.
        const int64_t ramp = (int64_t)32 << (20 + (k < 6 ? k : 6));
        const int64_t target = ramp < target_max ? ramp : target_max;
        int64_t g1 = g0;
        int64_t bytes = 0;
        while (g1 < n_groups && (g1 == g0 || bytes < target)) {
            const int64_t rows = grp_row_ptr[g1 + 1] - grp_row_ptr[g1], cols = grp_col_ptr[g1 + 1] - grp_col_ptr[g1];
            const int64_t e = row_nnz_ptr[grp_row_ptr[g1 + 1]] - row_nnz_ptr[grp_row_ptr[g1]];
            bytes += 13 * e + 17 * cols + 8 * rows + 16;
            ++g1;
        }
        Slot &s = g_pipe[k % kPipeSlots];
        ++k;
        rc = collect(s);                        // the slot's previous chunk (its D2H is then ordered on its stream)
        if (rc) break;
        s.single_tx_shortcut = single_tx_shortcut;
        rc = slot_prepare(s, grp_row_ptr, grp_col_ptr, row_nnz_ptr, g0, g1);
        if (!rc) rc = slot_upload(s, col_idx, aval, nz_flags, Y, K, col_flags);
        if (!rc) rc = slot_run(s, maxiter, minvalue, conv);
        if (!rc) {
            s.busy = true;
            s.dl_est = est; s.dl_iters = iters; s.dl_status = status; s.dl_total_rows = total_rows;
        }
        g0 = g1;
    }
    for (Slot &s : g_pipe) {
        int rc2 = collect(s);
        if (!rc) rc = rc2;
    }
    for (Slot &s : g_pipe)
        if (s.inited && cudaStreamSynchronize(s.stream) != cudaSuccess && !rc)
            rc = fail(BAMBU_EM_ECUDA, "pipeline stream failed: %s", cudaGetErrorString(cudaGetLastError()));
    if (n_groups == 0) return BAMBU_EM_OK;
    return rc;
}

// the same pipeline fed with the wire form v2
static int batched_impl_v2(int64_t n_groups, const int64_t *grp_row_ptr, const int64_t *grp_col_ptr,
                           const int64_t *grp_ent_ptr, const WireHost &h, int32_t maxiter, double minvalue, double conv,
                           double *est, int32_t *iters, int32_t *status)
{
    if (!est || !iters || !status) return fail(BAMBU_EM_EINVAL, "NULL output pointer");
    TRY(check_arrays_v2(n_groups, grp_row_ptr, grp_col_ptr, grp_ent_ptr));
    TRY(ensure_device());
    g_pipe_used = true;
    const int64_t total_rows = grp_row_ptr[n_groups];
    const int64_t colb = h.col_mode == 0 ? 16 : 8;
    // Every chunk pays a fixed price (host analysis, ~15 launches, and above all the tail of its longest-running groups
    // on a GPU that is emptying), so a small call is cut into two chunks (the second one's copy overlaps the first one's
    // EM).  Measured on 13 samples (171 MB, device time 16.8 ms): 1 chunk 21.7, 2 chunks 21.4, ramp 12-24-48-87 MB 25.2,
    // 8 chunks 30.6, 16 chunks 39.8 ms.  Calls beyond twice the chunk size are not affected.
    const int64_t total_bytes = 10 * grp_ent_ptr[n_groups] + colb * grp_col_ptr[n_groups] + 12 * grp_row_ptr[n_groups];
    const char *frac_env = getenv("BAMBU_EM_CHUNK_MIN_FRACTION");      // experiments
    const int64_t min_frac = frac_env ? std::max<int64_t>(1, atoll(frac_env)) : 2;
    const int64_t target_min = total_bytes / min_frac;
    const int64_t target_max = std::max<int64_t>(chunk_target_bytes_v2(), 1);
    int rc = BAMBU_EM_OK;
    int64_t g0 = 0, ib0 = 0;
    int k = 0;
    // BAMBU_EM_TIMING=1: where the host thread of this call spends its time (stderr, one line per call)
    const bool timing = getenv("BAMBU_EM_TIMING") != nullptr;
    double t_collect = 0, t_prepare = 0, t_upload = 0, t_run = 0;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto since = [&](std::chrono::steady_clock::time_point t) { return std::chrono::duration<double, std::milli>(now() - t).count(); };
    const auto t_call = now();
    while (g0 < n_groups && !rc) {
        const int64_t ramp = (int64_t)12 << (20 + (k < 6 ? k : 6));       // 12 MB doubling: the wire form is ~1/3 of v1
        const int64_t target = std::min<int64_t>(std::max<int64_t>(ramp, target_min), target_max);
        int64_t g1 = g0, bytes = 0;
        while (g1 < n_groups && (g1 == g0 || bytes < target)) {
            const int64_t rows = grp_row_ptr[g1 + 1] - grp_row_ptr[g1], cols = grp_col_ptr[g1 + 1] - grp_col_ptr[g1];
            const int64_t e = grp_ent_ptr[g1 + 1] - grp_ent_ptr[g1];
            const int64_t gb = (8 + wire_idx_width(cols)) * e + colb * cols + 12 * rows + 32;
            // giant loci iterate for long (their copy is small next to their EM) and a chunk's cooperative launch ends
            // with its slowest team: count them at a quarter of their size so that more of them share a chunk
            bytes += (e > 200000) ? gb / 4 : gb;
            ++g1;
        }
        Slot &s = g_pipe[k % kPipeSlots];
        ++k;
        auto t0 = now();
        rc = collect(s);
        t_collect += since(t0);
        if (rc) break;
        s.single_tx_shortcut = true;
        t0 = now();
        rc = slot_prepare(s, grp_row_ptr, grp_col_ptr, nullptr, g0, g1, grp_ent_ptr, ib0);
        t_prepare += since(t0);
        t0 = now();
        if (!rc) rc = slot_upload_v2(s, h);
        t_upload += since(t0);
        t0 = now();
        if (!rc) rc = slot_run(s, maxiter, minvalue, conv);
        t_run += since(t0);
        if (!rc) {
            s.busy = true;
            s.dl_est = est; s.dl_iters = iters; s.dl_status = status; s.dl_total_rows = total_rows;
            ib0 = s.ib1;
        }
        g0 = g1;
    }
    const auto t_drain = now();
    for (Slot &s : g_pipe) {
        int rc2 = collect(s);
        if (!rc) rc = rc2;
    }
    for (Slot &s : g_pipe)
        if (s.inited && cudaStreamSynchronize(s.stream) != cudaSuccess && !rc)
            rc = fail(BAMBU_EM_ECUDA, "pipeline stream failed: %s", cudaGetErrorString(cudaGetLastError()));
    if (timing)
        fprintf(stderr, "[bambu_em_batched_v2] %d chunks, %.1f ms: wait for a free slot %.1f, analyse %.1f, enqueue copies %.1f, "
                        "enqueue kernels %.1f, drain %.1f\n", k, since(t_call), t_collect, t_prepare, t_upload, t_run, since(t_drain));
    if (n_groups == 0) return BAMBU_EM_OK;
    return rc;
}

int bambu_em_batched_v2(int64_t n_groups, const int64_t *grp_row_ptr, const int64_t *grp_col_ptr,
                        const int64_t *grp_ent_ptr, const uint32_t *row_end, const double *rs, const double *aval,
                        const void *cidx, const uint32_t *eq_bits, int32_t col_mode, const void *colY, const void *colK,
                        const uint32_t *multi_bits, int32_t maxiter, double minvalue, double conv, double *est,
                        int32_t *iters, int32_t *status)
{
    const WireHost h = {row_end, rs, aval, (const uint8_t *)cidx, eq_bits, col_mode, colY, colK, multi_bits};
    return batched_impl_v2(n_groups, grp_row_ptr, grp_col_ptr, grp_ent_ptr, h, maxiter, minvalue, conv, est, iters, status);
}

int bambu_em_plan_create_v2(bambu_em_plan **out, int64_t n_groups, const int64_t *grp_row_ptr,
                            const int64_t *grp_col_ptr, const int64_t *grp_ent_ptr)
{
    if (!out) return fail(BAMBU_EM_EINVAL, "plan pointer is NULL");
    *out = nullptr;
    TRY(check_arrays_v2(n_groups, grp_row_ptr, grp_col_ptr, grp_ent_ptr));
    TRY(ensure_device());
    bambu_em_plan *p = new (std::nothrow) bambu_em_plan;
    if (!p) return fail(BAMBU_EM_ENOMEM, "out of host memory");
    p->n_groups = n_groups;
    int rc = slot_prepare(p->slot, grp_row_ptr, grp_col_ptr, nullptr, 0, n_groups, grp_ent_ptr, 0);
    if (!rc && cudaStreamSynchronize(p->slot.stream) != cudaSuccess) rc = fail(BAMBU_EM_ECUDA, "structure upload failed");
    if (rc) { bambu_em_plan_destroy(p); return rc; }
    *out = p;
    return BAMBU_EM_OK;
}

int bambu_em_plan_upload_v2(bambu_em_plan *p, const uint32_t *row_end, const double *rs, const double *aval,
                            const void *cidx, const uint32_t *eq_bits, int32_t col_mode, const void *colY,
                            const void *colK, const uint32_t *multi_bits)
{
    if (!p) return fail(BAMBU_EM_EINVAL, "plan is NULL");
    if (!p->slot.v2) return fail(BAMBU_EM_ESTATE, "the plan was created for the v1 arena");
    TRY(ensure_device());
    const WireHost h = {row_end, rs, aval, (const uint8_t *)cidx, eq_bits, col_mode, colY, colK, multi_bits};
    return slot_upload_v2(p->slot, h);
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
    std::vector<uint8_t> nf, cf((size_t)N + 1, 0);
    std::vector<int8_t> col_seen((size_t)N + 1, -1);   // -1 unknown, 0 unique kept, 1 multi
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
