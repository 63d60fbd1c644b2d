// quant_dev.cu -- the device packer: bambu.quantDT's long table -> arena on the GPU (quant_dev.inl run
// by CUDA kernels), and bambu_quantDT's fast path that keeps the arena in HBM for the EM.
//
// Everything here is HBM-bound integer work over N-row columns: flat grid-stride loops, a stable LSD
// radix sort (8-bit digits; per-tile digit histograms -> one scan -> stable scatter with warp-level
// match ranking) and a three-kernel exclusive scan.  Grids are sized to the SM count.
#include <cuda_runtime.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <new>
#include <algorithm>
#include <atomic>
#include <string>
#include <system_error>
#include <thread>
#include <type_traits>
#include <unordered_map>
#include <vector>

#include "../../include/bambu_em.h"
#include "quant_pack.h"

#define QD_FN __host__ __device__ __forceinline__
#define QD_LAMBDA [=] __device__

__device__ __forceinline__ void qd_or(uint32_t *p, uint32_t v) { atomicOr(p, v); }
__device__ __forceinline__ void qd_inc(uint32_t *p) { atomicAdd(p, 1u); }
__device__ __forceinline__ void qd_add(double *p, double v) { atomicAdd(p, v); }
__device__ __forceinline__ void qd_min(uint64_t *p, uint64_t v) { atomicMin(reinterpret_cast<unsigned long long *>(p), (unsigned long long)v); }
__device__ __forceinline__ void qd_max(uint64_t *p, uint64_t v) { atomicMax(reinterpret_cast<unsigned long long *>(p), (unsigned long long)v); }

#include "quant_dev.inl"
#include "r_round.h"

namespace qdcuda {

struct CudaError {
    cudaError_t e;
    const char *what;
};
#define QD_CUDA(call)                                                \
    do {                                                             \
        cudaError_t e_ = (call);                                     \
        if (e_ != cudaSuccess) throw qdcuda::CudaError{e_, #call};           \
    } while (0)

constexpr int kThreads = 256;
constexpr int kSortItems = 16;                       // keys per thread in a sort tile
constexpr int kSortTile = kThreads * kSortItems;     // 4096
constexpr int kScanItems = 8;
constexpr int kScanTile = kThreads * kScanItems;     // 2048

template <class F>
__global__ void __launch_bounds__(kThreads) pfor_kernel(size_t n, F f)
{
    for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (size_t)gridDim.x * kThreads) f(i);
}

__global__ void __launch_bounds__(kThreads) minmax_kernel(const int64_t *x, size_t n, long long *out)
{
    long long lo = INT64_MAX, hi = INT64_MIN;
    for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (size_t)gridDim.x * kThreads) {
        const long long v = x[i];
        lo = v < lo ? v : lo;
        hi = v > hi ? v : hi;
    }
    for (int o = 16; o; o >>= 1) {
        const long long a = __shfl_xor_sync(0xffffffffu, lo, o), b = __shfl_xor_sync(0xffffffffu, hi, o);
        lo = a < lo ? a : lo;
        hi = b > hi ? b : hi;
    }
    if ((threadIdx.x & 31) == 0) { atomicMin(out, lo); atomicMax(out + 1, hi); }
}
__global__ void minmax_init_kernel(long long *out) { out[0] = INT64_MAX; out[1] = INT64_MIN; }

// ---- exclusive scan of u32 ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t block_excl_scan_u32(uint32_t v, uint32_t *total, uint32_t *s_warp /*[9]*/)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t inc = v;
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_warp[w] = inc;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t run = 0;
        for (int q = 0; q < kThreads / 32; ++q) { const uint32_t t = s_warp[q]; s_warp[q] = run; run += t; }
        s_warp[kThreads / 32] = run;
    }
    __syncthreads();
    *total = s_warp[kThreads / 32];
    return s_warp[w] + inc - v;
}

__global__ void __launch_bounds__(kThreads) scan_local_kernel(const uint32_t *in, uint32_t *out, size_t n, uint32_t *bsum)
{
    __shared__ uint32_t s_warp[kThreads / 32 + 1];
    const size_t base = (size_t)blockIdx.x * kScanTile + (size_t)threadIdx.x * kScanItems;
    uint32_t v[kScanItems], sum = 0;
#pragma unroll
    for (int q = 0; q < kScanItems; ++q) { v[q] = base + q < n ? in[base + q] : 0u; sum += v[q]; }
    uint32_t total;
    uint32_t run = block_excl_scan_u32(sum, &total, s_warp);
#pragma unroll
    for (int q = 0; q < kScanItems; ++q) { if (base + q < n) out[base + q] = run; run += v[q]; }
    if (threadIdx.x == 0) bsum[blockIdx.x] = total;
}
// one block: exclusive scan of the block totals in place, grand total to bsum[nb]
__global__ void __launch_bounds__(kThreads) scan_bsum_kernel(uint32_t *bsum, size_t nb)
{
    __shared__ uint32_t s_warp[kThreads / 32 + 1];
    uint32_t carry = 0;
    for (size_t base = 0; base < nb; base += kThreads) {
        const size_t i = base + threadIdx.x;
        const uint32_t v = i < nb ? bsum[i] : 0u;
        uint32_t total;
        const uint32_t ex = block_excl_scan_u32(v, &total, s_warp);
        if (i < nb) bsum[i] = carry + ex;
        carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) bsum[nb] = carry;
}
__global__ void __launch_bounds__(kThreads) scan_add_kernel(uint32_t *out, size_t n, const uint32_t *bsum)
{
    const uint32_t add = bsum[blockIdx.x];
    const size_t base = (size_t)blockIdx.x * kScanTile;
    for (int q = threadIdx.x; q < kScanTile; q += kThreads)
        if (base + q < n) out[base + q] += add;
}

// ---- LSD radix sort pass (8-bit digit) -----------------------------------------------------------------
// hist[d * nb + b] = keys of tile b with digit d; scanned as one array it gives every (digit, tile) its base.
__global__ void __launch_bounds__(kThreads) sort_hist_kernel(const uint64_t *k, size_t n, int shift, uint32_t *hist, uint32_t nb)
{
    __shared__ uint32_t s_h[256];
    s_h[threadIdx.x] = 0;
    __syncthreads();
    const size_t base = (size_t)blockIdx.x * kSortTile;
    for (int q = threadIdx.x; q < kSortTile; q += kThreads)
        if (base + q < n) atomicAdd(&s_h[(uint32_t)(k[base + q] >> shift) & 255u], 1u);
    __syncthreads();
    hist[(size_t)threadIdx.x * nb + blockIdx.x] = s_h[threadIdx.x];
}
// Stable scatter: warp w of the tile owns 512 consecutive keys and walks them 32 at a time in order; a key's
// rank among equal digits = (digit count of the warp so far) + (equal digits in lower lanes of this round).
__global__ void __launch_bounds__(kThreads) sort_scatter_kernel(const uint64_t *k, const uint32_t *v, uint64_t *k2, uint32_t *v2,
                                                                size_t n, int shift, const uint32_t *base_of, uint32_t nb)
{
    __shared__ uint32_t s_cnt[kThreads / 32][256];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int q = threadIdx.x; q < (kThreads / 32) * 256; q += kThreads) (&s_cnt[0][0])[q] = 0;
    __syncthreads();
    const size_t wbase = (size_t)blockIdx.x * kSortTile + (size_t)w * (32 * kSortItems);
    uint64_t key[kSortItems];
    uint32_t rank[kSortItems];
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
        const size_t i = wbase + (size_t)r * 32 + lane;
        const bool ok = i < n;
        key[r] = ok ? k[i] : 0;
        const uint32_t d = ok ? ((uint32_t)(key[r] >> shift) & 255u) : (256u + lane);   // idle lanes match nobody
        const uint32_t peers = __match_any_sync(0xffffffffu, d);
        const int leader = __ffs(peers) - 1;
        uint32_t before = 0;
        if (ok && lane == leader) { before = s_cnt[w][d]; s_cnt[w][d] = before + __popc(peers); }
        before = __shfl_sync(0xffffffffu, before, leader);
        rank[r] = before + __popc(peers & ((1u << lane) - 1u));
        __syncwarp();
    }
    __syncthreads();
    {   // digit threadIdx.x: global base of this tile, then running offsets for the warps in order
        uint32_t run = base_of[(size_t)threadIdx.x * nb + blockIdx.x];
        for (int q = 0; q < kThreads / 32; ++q) { const uint32_t t = s_cnt[q][threadIdx.x]; s_cnt[q][threadIdx.x] = run; run += t; }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
        const size_t i = wbase + (size_t)r * 32 + lane;
        if (i < n) {
            const uint32_t d = (uint32_t)(key[r] >> shift) & 255u;
            const size_t pos = (size_t)s_cnt[w][d] + rank[r];
            k2[pos] = key[r];
            v2[pos] = v[i];
        }
    }
}

// The table columns come from pageable host memory, where a copy call returns only once the data is staged
// (one host thread moves ~6 GB/s that way): one helper thread per column feeds the copy engines on its own
// stream, while the calling thread already launches the passes over the columns that have arrived.
struct Uploader {
    struct Col { void *dst; const void *src; size_t bytes; };
    Col cols[qd::kNumCols] = {};
    cudaEvent_t ev[qd::kNumCols] = {};
    cudaStream_t st[qd::kNumCols] = {};
    std::atomic<int> issued[qd::kNumCols];
    std::atomic<int> error{0};
    std::thread th[qd::kNumCols];
    double done_ms[qd::kNumCols] = {};
    std::chrono::steady_clock::time_point t0;
    int device = 0;
    Uploader() { for (auto &i : issued) i.store(0); }
    void start()
    {
        for (int c = 0; c < qd::kNumCols; ++c) QD_CUDA(cudaEventCreateWithFlags(&ev[c], cudaEventDisableTiming));
        t0 = std::chrono::steady_clock::now();
        for (int c = 0; c < qd::kNumCols; ++c) {
            if (!cols[c].bytes) {        // nothing to copy: the event stays unrecorded, waiting on it is a no-op
                issued[c].store(1, std::memory_order_release);
                continue;
            }
            auto copy_column = [this, c] {
                cudaError_t e = cudaSetDevice(device);
                if (e == cudaSuccess) e = cudaMemcpyAsync(cols[c].dst, cols[c].src, cols[c].bytes, cudaMemcpyHostToDevice, st[c]);
                if (e == cudaSuccess) e = cudaEventRecord(ev[c], st[c]);
                if (e != cudaSuccess) error.store((int)e);
                done_ms[c] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
                issued[c].store(1, std::memory_order_release);
            };
            try {
                th[c] = std::thread(copy_column);
            } catch (const std::system_error &) {     // no thread to be had: copy this column from the calling thread
                copy_column();
            }
        }
    }
    void finish()
    {
        for (int c = 0; c < qd::kNumCols; ++c)
            if (th[c].joinable()) { th[c].join(); cudaStreamSynchronize(st[c]); }
        for (int c = 0; c < qd::kNumCols; ++c) if (ev[c]) { cudaEventDestroy(ev[c]); ev[c] = nullptr; }
    }
    ~Uploader() { finish(); }
};

struct CudaExec {
    Uploader *up = nullptr;
    bool have[qd::kNumCols] = {};
    std::function<void()> on_arrival[qd::kNumCols];      // e.g. widening of an int32 column, queued behind its copy
    cudaStream_t st = nullptr;
    uint8_t *pool = nullptr;
    size_t cap = 0, top = 0;
    int grid_cap = 148 * 8;
    long launches = 0;

    CudaExec(cudaStream_t s, uint8_t *p, size_t bytes, int sms) : st(s), pool(p), cap(bytes), grid_cap(sms * 8) {}

    template <class T>
    T *alloc(size_t n)
    {
        const size_t bytes = (n * sizeof(T) + 255) & ~(size_t)255;
        if (top + bytes > cap) throw std::bad_alloc();
        T *p = reinterpret_cast<T *>(pool + top);
        top += bytes;
        return p;
    }
    // order everything launched from here on behind the arrival of table column c
    void need(int c)
    {
        if (!up || have[c]) return;
        while (!up->issued[c].load(std::memory_order_acquire)) std::this_thread::yield();
        if (up->error.load()) throw CudaError{(cudaError_t)up->error.load(), "copy of the table to the device"};
        QD_CUDA(cudaStreamWaitEvent(st, up->ev[c], 0));
        have[c] = true;
        if (on_arrival[c]) on_arrival[c]();
    }
    // BAMBU_PACK_TIMING: wall time of each stage of the pipeline (drains the stream, so it perturbs the total)
    bool timing = false;
    std::chrono::steady_clock::time_point t_lap;
    void lap(const char *what)
    {
        if (!timing) return;
        cudaStreamSynchronize(st);
        const auto t = std::chrono::steady_clock::now();
        fprintf(stderr, "[device pack] %-30s %7.2f ms\n", what, std::chrono::duration<double, std::milli>(t - t_lap).count());
        t_lap = t;
    }
    size_t mark() const { return top; }
    void release(size_t m) { top = m; }
    void fill(void *p, int byte, size_t bytes) { if (bytes) QD_CUDA(cudaMemsetAsync(p, byte, bytes, st)); }
    int grid_for(size_t n, size_t per_block) const
    {
        const size_t b = (n + per_block - 1) / per_block;
        return (int)(b < (size_t)grid_cap ? (b ? b : 1) : (size_t)grid_cap);
    }
    template <class F>
    void pfor(size_t n, F f)
    {
        if (!n) return;
        pfor_kernel<<<grid_for(n, kThreads), kThreads, 0, st>>>(n, f);
        ++launches;
    }
    void minmax(const int64_t *x, size_t n, int64_t *d_out)
    {
        long long *o = reinterpret_cast<long long *>(d_out);
        minmax_init_kernel<<<1, 1, 0, st>>>(o);
        minmax_kernel<<<grid_for(n, kThreads * 8), kThreads, 0, st>>>(x, n, o);
        launches += 2;
    }
    // exclusive scan; the total is left at the end of the block-sum array (device) and optionally fetched
    void scan_device(const uint32_t *in, uint32_t *out, size_t n, uint32_t *bsum)
    {
        const size_t nb = (n + kScanTile - 1) / kScanTile;
        scan_local_kernel<<<(unsigned)nb, kThreads, 0, st>>>(in, out, n, bsum);
        scan_bsum_kernel<<<1, kThreads, 0, st>>>(bsum, nb);
        scan_add_kernel<<<(unsigned)nb, kThreads, 0, st>>>(out, n, bsum);
        launches += 3;
    }
    uint32_t scan(const uint32_t *in, uint32_t *out, size_t n)
    {
        if (!n) return 0;
        const size_t mk = mark();
        const size_t nb = (n + kScanTile - 1) / kScanTile;
        uint32_t *bsum = alloc<uint32_t>(nb + 1);
        scan_device(in, out, n, bsum);
        uint32_t total;
        d2h(&total, bsum + nb, 1);
        release(mk);
        return total;
    }
    void sort_pairs(uint64_t *&k, uint32_t *&v, uint64_t *&k2, uint32_t *&v2, size_t n, int bits)
    {
        if (n < 2) return;
        const size_t mk = mark();
        const uint32_t nb = (uint32_t)((n + kSortTile - 1) / kSortTile);
        const size_t hn = (size_t)256 * nb;
        uint32_t *hist = alloc<uint32_t>(hn), *base_of = alloc<uint32_t>(hn);
        uint32_t *bsum = alloc<uint32_t>((hn + kScanTile - 1) / kScanTile + 1);
        for (int shift = 0; shift < bits; shift += 8) {
            sort_hist_kernel<<<nb, kThreads, 0, st>>>(k, n, shift, hist, nb);
            scan_device(hist, base_of, hn, bsum);
            sort_scatter_kernel<<<nb, kThreads, 0, st>>>(k, v, k2, v2, n, shift, base_of, nb);
            launches += 2;
            uint64_t *tk = k; k = k2; k2 = tk;
            uint32_t *tv = v; v = v2; v2 = tv;
        }
        release(mk);
    }
    template <class T>
    void d2h(T *dst, const T *src, size_t n)
    {
        if (!n) return;
        QD_CUDA(cudaMemcpyAsync(dst, src, n * sizeof(T), cudaMemcpyDeviceToHost, st));
        QD_CUDA(cudaStreamSynchronize(st));
    }
    template <class T>
    void h2d(T *dst, const T *src, size_t n)
    {
        if (!n) return;
        QD_CUDA(cudaMemcpyAsync(dst, src, n * sizeof(T), cudaMemcpyHostToDevice, st));
        QD_CUDA(cudaStreamSynchronize(st));      // the source is a host temporary
    }
};

// grow-only device workspace and staging stream of the packer (per process)
struct Workspace {
    uint8_t *pool = nullptr;
    size_t cap = 0;
    cudaStream_t st = nullptr, up_st[qd::kNumCols] = {};
    int sms = 148;
    int device = -1;
    bool ensure(size_t bytes)
    {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return false;
        if (dev != device) {
            release();
            device = dev;
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        }
        if (!st && cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) return false;
        for (auto &u : up_st)
            if (!u && cudaStreamCreateWithFlags(&u, cudaStreamNonBlocking) != cudaSuccess) return false;
        if (bytes <= cap) return true;
        if (pool) cudaFree(pool);
        pool = nullptr;
        cap = 0;
        if (cudaMalloc(&pool, bytes) != cudaSuccess) { cudaGetLastError(); return false; }
        cap = bytes;
        return true;
    }
    void release()
    {
        if (pool) cudaFree(pool);
        if (st) cudaStreamDestroy(st);
        for (auto &u : up_st) { if (u) cudaStreamDestroy(u); u = nullptr; }
        pool = nullptr; cap = 0; st = nullptr; device = -1;
    }
};
Workspace g_ws;
bambu_em_plan *g_plan = nullptr;      // bambu_quantDT's cached plan

size_t workspace_bytes(size_t n) { return n * 288 + (64u << 20); }

}  // namespace qdcuda

void bambu_quant_set_error(const std::string &s);      // quant_pack.cpp
int bambu_em_internal_ensure_device();                 // bambu_em.cu

namespace {

struct DevTable {          // the long table in device memory
    qd::Input in;
};

// copies the table into the workspace and runs the pipeline.  Returns BAMBU_EM_* or +1 for "use the host packer".
// IdT: int64_t, or int (R's integer): copied as they are and widened on the device.  EqT: uint8_t, or int (R's logical).
template <class IdT, class EqT>
int run_device_pack(qdcuda::CudaExec &x, int64_t n, const IdT *gene, const IdT *tx, const IdT *eq,
                    const double *nobs, const EqT *equal, const double *rcw, const double *txl, bool d_mode,
                    qd::Result &R, qdcuda::Workspace &ws = qdcuda::g_ws)
{
    const size_t N = (size_t)n;
    qd::Input in;
    in.n = N;
    in.d_mode = d_mode;
    if (getenv("BAMBU_PACK_POISON"))      // debugging aid: nothing may depend on what the workspace held before
        QD_CUDA(cudaMemsetAsync(x.pool, 0xA5, x.cap, x.st));
    qdcuda::Uploader up;
    for (int c = 0; c < qd::kNumCols; ++c) up.st[c] = ws.up_st[c];
    QD_CUDA(cudaGetDevice(&up.device));
    auto col = [&](int c, auto *src, bool wanted) -> decltype(src) {
        typedef typename std::remove_const<typename std::remove_pointer<decltype(src)>::type>::type T;
        if (!src || !wanted || !N) return nullptr;
        T *d = x.alloc<T>(N);
        up.cols[c] = qdcuda::Uploader::Col{d, src, N * sizeof(T)};
        return d;
    };
    auto id_col = [&](int c, const IdT *src) -> const int64_t * {
        const IdT *raw = col(c, src, true);
        if (std::is_same<IdT, int64_t>::value || !raw) return reinterpret_cast<const int64_t *>(raw);
        int64_t *wide = x.alloc<int64_t>(N);
        x.on_arrival[c] = [&x, raw, wide, N] { x.pfor(N, [=] __device__(size_t i) { wide[i] = (int64_t)raw[i]; }); };
        return wide;
    };
    in.gene = id_col(qd::kColGene, gene); in.tx = id_col(qd::kColTx, tx); in.eq = id_col(qd::kColEq, eq);
    in.nobs = col(qd::kColNobs, nobs, true);
    {
        const EqT *raw = col(qd::kColEqual, equal, true);
        if (std::is_same<EqT, uint8_t>::value || !raw) in.equal = reinterpret_cast<const uint8_t *>(raw);
        else {
            uint8_t *flag = x.alloc<uint8_t>(N);
            x.on_arrival[qd::kColEqual] = [&x, raw, flag, N] { x.pfor(N, [=] __device__(size_t i) { flag[i] = raw[i] == 1 ? 1 : 0; }); };
            in.equal = flag;                              // NA_LOGICAL (INT_MIN) counts as FALSE
        }
    }
    in.txl = col(qd::kColTxlen, txl, d_mode); in.rcw = col(qd::kColRcWidth, rcw, d_mode);
    up.start();
    x.up = &up;
    x.timing = getenv("BAMBU_PACK_TIMING") != nullptr;
    x.t_lap = std::chrono::steady_clock::now();
    struct Detach {      // the thread is joined (Uploader's destructor) before the executor forgets it
        qdcuda::CudaExec &x;
        ~Detach() { x.up = nullptr; }
    } detach{x};
    std::string err;
    const int rc = qd::pack(x, in, R, err);
    const auto t_pack = std::chrono::steady_clock::now();
    up.finish();
    for (auto &f : x.on_arrival) f = nullptr;
    if (getenv("BAMBU_PACK_TIMING")) {
        fprintf(stderr, "[device pack] pipeline done at %.2f ms; column copies staged at", std::chrono::duration<double, std::milli>(t_pack - up.t0).count());
        for (int c = 0; c < qd::kNumCols; ++c) fprintf(stderr, " %.2f", up.done_ms[c]);
        fprintf(stderr, " ms\n");
    }
    if (up.error.load()) throw qdcuda::CudaError{(cudaError_t)up.error.load(), "copy of the table to the device"};
    if (rc == qd::kUnsupported) return 1;
    if (rc == qd::kInvalid) { bambu_quant_set_error(err); return BAMBU_EM_EINVAL; }
    return BAMBU_EM_OK;
}

template <class T>
void fetch(qdcuda::CudaExec &x, std::vector<T> &dst, const T *src, size_t n)
{
    dst.resize(n);
    x.d2h(dst.data(), src, n);
}

template <class IdT>
int check_args(bambu_quant_pack **out, int64_t n, const IdT *gene_code, const IdT *txid_in, const IdT *eq_in,
               const double *nobs)
{
    if (!out) { bambu_quant_set_error("NULL output"); return BAMBU_EM_EINVAL; }
    *out = nullptr;
    if (n < 0 || n > (int64_t)UINT32_MAX - 8) { bambu_quant_set_error("row count out of range"); return BAMBU_EM_EINVAL; }
    if (n && (!gene_code || !txid_in || !eq_in || !nobs)) {
        bambu_quant_set_error("Columns GENEID, txid, eqClassId, nobs, are missing from object.");
        return BAMBU_EM_EINVAL;
    }
    return BAMBU_EM_OK;
}

}  // namespace

// the host-packer path for either input flavour (widens R's int columns first)
inline int hostpack_any(int64_t n, const int64_t *g, const int64_t *t, const int64_t *e, const double *nobs, const uint8_t *equal,
                        const double *rcw, const double *txl, int32_t bias, int32_t maxiter, double minvalue, double conv,
                        int64_t *txid_out, double *counts, double *full_length, double *unique_counts, int64_t *n_result,
                        double *d_rate)
{
    return bambu_quantDT_hostpack(n, g, t, e, nobs, equal, rcw, txl, bias, maxiter, minvalue, conv, txid_out, counts, full_length,
                                  unique_counts, n_result, d_rate);
}
inline int hostpack_any(int64_t n, const int *g, const int *t, const int *e, const double *nobs, const int *equal,
                        const double *rcw, const double *txl, int32_t bias, int32_t maxiter, double minvalue, double conv,
                        int64_t *txid_out, double *counts, double *full_length, double *unique_counts, int64_t *n_result,
                        double *d_rate)
{
    const size_t N = (size_t)n;
    std::vector<int64_t> g64(N), t64(N), e64(N);
    std::vector<uint8_t> eqf(N);
    for (size_t i = 0; i < N; ++i) { g64[i] = g[i]; t64[i] = t[i]; e64[i] = e[i]; }
    if (equal) for (size_t i = 0; i < N; ++i) eqf[i] = equal[i] == 1;
    return bambu_quantDT_hostpack(n, g64.data(), t64.data(), e64.data(), nobs, equal ? eqf.data() : nullptr, rcw, txl, bias, maxiter,
                                  minvalue, conv, txid_out, counts, full_length, unique_counts, n_result, d_rate);
}

// 1 when the last bambu_quantDT / bambu_quantDT_dotC call of this thread packed its table on the GPU, 0 when it went
// through the host packer (a ~25x slower path: non-integer counts etc., see quant_dev.inl), -1 before any call
static thread_local int g_last_used_device = -1;

template <class IdT, class EqT>
int quantdt_core(int64_t n, const IdT *gene_code, const IdT *txid, const IdT *eqClassId, const double *nobs,
                  const EqT *equal, const double *rcWidth, const double *txlen, int32_t degradation_bias,
                  int32_t maxiter, double minvalue, double conv, int64_t *txid_out, double *counts, double *full_length,
                  double *unique_counts, int64_t *n_result, double *d_rate_out)
{
    g_last_used_device = 0;
    const char *force_host = getenv("BAMBU_QUANT_HOST_PACK");
    if (force_host && force_host[0] == '1')
        return hostpack_any(n, gene_code, txid, eqClassId, nobs, equal, rcWidth, txlen, degradation_bias, maxiter,
                            minvalue, conv, txid_out, counts, full_length, unique_counts, n_result, d_rate_out);
    bambu_quant_pack *dummy = nullptr;
    int rc = check_args(&dummy, n, gene_code, txid, eqClassId, nobs);
    if (rc) return rc;
    if (!txid_out || !counts || !full_length || !unique_counts || !n_result) { bambu_quant_set_error("NULL output"); return BAMBU_EM_EINVAL; }
    rc = bambu_em_internal_ensure_device();
    if (rc) { bambu_quant_set_error(bambu_em_last_error()); return rc; }
    if (!qdcuda::g_ws.ensure(qdcuda::workspace_bytes((size_t)n))) {
        bambu_quant_set_error("not enough device memory for the packer workspace");
        return BAMBU_EM_ENOMEM;
    }
    const bool timing = getenv("BAMBU_PACK_TIMING") != nullptr;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
        return std::chrono::duration<double, std::milli>(b - a).count();
    };
    const auto t0 = now();
    bambu_quant_pack P;
    try {
        qdcuda::CudaExec x(qdcuda::g_ws.st, qdcuda::g_ws.pool, qdcuda::g_ws.cap, qdcuda::g_ws.sms);
        qd::Result R;
        rc = run_device_pack(x, n, gene_code, txid, eqClassId, nobs, equal, rcWidth, txlen, degradation_bias != 0, R);
        if (rc == 1)
            return hostpack_any(n, gene_code, txid, eqClassId, nobs, equal, rcWidth, txlen, degradation_bias, maxiter,
                                minvalue, conv, txid_out, counts, full_length, unique_counts, n_result, d_rate_out);
        if (rc) return rc;
        fetch(x, P.grp_row_ptr, R.grp_row_ptr, (size_t)R.n_groups + 1);
        fetch(x, P.grp_col_ptr, R.grp_col_ptr, (size_t)R.n_groups + 1);
        fetch(x, P.row_nnz_ptr, R.row_nnz_ptr, (size_t)R.n_rows + 1);
        fetch(x, P.txid, R.txid, (size_t)R.n_rows);
        P.out_txid.swap(R.out_txid);
        P.unobserved_txid.swap(R.unobserved_txid);
        P.d_rate = R.d_rate;
        QD_CUDA(cudaGetLastError());
        const auto t1 = now();
        std::vector<double> est((size_t)std::max<int64_t>(3 * (int64_t)R.n_rows, 1));
        if (R.n_groups) {
            // one cached plan: its device buffers only grow, so repeated calls allocate nothing
            bambu_em_plan *&plan = qdcuda::g_plan;
            rc = plan ? bambu_em_plan_reset(plan, R.n_groups, P.grp_row_ptr.data(), P.grp_col_ptr.data(), P.row_nnz_ptr.data())
                      : bambu_em_plan_create(&plan, R.n_groups, P.grp_row_ptr.data(), P.grp_col_ptr.data(), P.row_nnz_ptr.data());
            if (!rc) rc = bambu_em_plan_bind_device(plan, R.col_idx, R.aval, R.nz_flags, R.Y, R.K, R.col_flags);
            if (!rc) rc = bambu_em_plan_run(plan, maxiter, minvalue, conv);
            if (!rc) rc = bambu_em_plan_download(plan, est.data(), nullptr, nullptr, nullptr);
            if (rc) {
                bambu_quant_set_error(bambu_em_last_error());
                if (plan) bambu_em_plan_destroy(plan);
                plan = nullptr;
                return rc;
            }
        }
        const auto t2 = now();
        rc = bambu_quant_merge(&P, est.data(), txid_out, counts, full_length, unique_counts, n_result);
        if (!rc && d_rate_out) *d_rate_out = P.d_rate;
        if (!rc) g_last_used_device = 1;
        if (timing)
            fprintf(stderr, "[bambu_quantDT] rows %lld: H2D + device pack %.2f ms (%ld launches), EM %.2f ms, merge %.2f ms\n",
                    (long long)n, ms(t0, t1), x.launches, ms(t1, t2), ms(t2, now()));
        return rc;
    } catch (const qdcuda::CudaError &e) {
        bambu_quant_set_error(std::string("CUDA error in the device packer: ") + cudaGetErrorString(e.e) + " (" + e.what + ")");
        return BAMBU_EM_ECUDA;
    } catch (const std::bad_alloc &) {
        bambu_quant_set_error("out of memory in the device packer");
        return BAMBU_EM_ENOMEM;
    } catch (const std::exception &e) {       // nothing may unwind through the C ABI (thread creation, std::function, ...)
        bambu_quant_set_error(std::string("device packer: ") + e.what());
        return BAMBU_EM_ECUDA;
    } catch (...) {
        bambu_quant_set_error("device packer: unknown exception");
        return BAMBU_EM_ECUDA;
    }
}

extern "C" {

// bambu_quant_pack_create done by the GPU; the result is copied back into a host pack object (for callers
// that want the arena on the host, and for the tests that compare it with the host packer byte for byte).
// Tables the device formulation does not cover (see quant_dev.inl) go through the host packer.
int bambu_quant_pack_create_gpu(bambu_quant_pack **out, int64_t n, const int64_t *gene_code, const int64_t *txid_in,
                                const int64_t *eq_in, const double *nobs, const uint8_t *equal, const double *rcWidth,
                                const double *txlen, int32_t degradation_bias, int32_t *used_device)
{
    int rc = check_args(out, n, gene_code, txid_in, eq_in, nobs);
    if (rc) return rc;
    if (used_device) *used_device = 0;
    rc = bambu_em_internal_ensure_device();
    if (rc) { bambu_quant_set_error(bambu_em_last_error()); return rc; }
    if (!qdcuda::g_ws.ensure(qdcuda::workspace_bytes((size_t)n))) {
        bambu_quant_set_error("no CUDA device or not enough device memory for the packer workspace");
        return BAMBU_EM_ECUDA;
    }
    bambu_quant_pack *P = new (std::nothrow) bambu_quant_pack;
    if (!P) { bambu_quant_set_error("out of memory"); return BAMBU_EM_ENOMEM; }
    try {
        qdcuda::CudaExec x(qdcuda::g_ws.st, qdcuda::g_ws.pool, qdcuda::g_ws.cap, qdcuda::g_ws.sms);
        qd::Result R;
        rc = run_device_pack(x, n, gene_code, txid_in, eq_in, nobs, equal, rcWidth, txlen, degradation_bias != 0, R);
        if (rc == 1) {
            delete P;
            return bambu_quant_pack_create(out, n, gene_code, txid_in, eq_in, nobs, equal, rcWidth, txlen, degradation_bias);
        }
        if (rc) { delete P; return rc; }
        fetch(x, P->grp_row_ptr, R.grp_row_ptr, (size_t)R.n_groups + 1);
        fetch(x, P->grp_col_ptr, R.grp_col_ptr, (size_t)R.n_groups + 1);
        fetch(x, P->row_nnz_ptr, R.row_nnz_ptr, (size_t)R.n_rows + 1);
        fetch(x, P->col_idx, R.col_idx, R.n_entries);
        fetch(x, P->aval, R.aval, R.n_entries);
        fetch(x, P->nz_flags, R.nz_flags, R.n_entries);
        fetch(x, P->Y, R.Y, (size_t)R.n_cols);
        fetch(x, P->K, R.K, (size_t)R.n_cols);
        fetch(x, P->col_flags, R.col_flags, (size_t)R.n_cols);
        fetch(x, P->txid, R.txid, (size_t)R.n_rows);
        P->out_txid.swap(R.out_txid);
        P->unobserved_txid.swap(R.unobserved_txid);
        P->gene_grp_id.swap(R.gene_grp_id);
        P->d_rate = R.d_rate;
        QD_CUDA(cudaGetLastError());
    } catch (const qdcuda::CudaError &e) {
        bambu_quant_set_error(std::string("CUDA error in the device packer: ") + cudaGetErrorString(e.e) + " (" + e.what + ")");
        delete P;
        return BAMBU_EM_ECUDA;
    } catch (const std::bad_alloc &) {
        bambu_quant_set_error("out of memory in the device packer");
        delete P;
        return BAMBU_EM_ENOMEM;
    } catch (const std::exception &e) {
        bambu_quant_set_error(std::string("device packer: ") + e.what());
        delete P;
        return BAMBU_EM_ECUDA;
    } catch (...) {
        bambu_quant_set_error("device packer: unknown exception");
        delete P;
        return BAMBU_EM_ECUDA;
    }
    if (used_device) *used_device = 1;
    *out = P;
    return BAMBU_EM_OK;
}


// bambu.quantDT in one call (R/bambu-quantify.R:53-74): the long table is copied to the GPU once, packed
// there, the EM runs on the arena where it lies in HBM, and only the estimates come back for the merge
// by txid.  Output arrays must hold at least as many rows as there are distinct txids in the input (<= n).
int bambu_quantDT(int64_t n, const int64_t *gene_code, const int64_t *txid, const int64_t *eqClassId, const double *nobs,
                  const uint8_t *equal, const double *rcWidth, const double *txlen, int32_t degradation_bias,
                  int32_t maxiter, double minvalue, double conv, int64_t *txid_out, double *counts, double *full_length,
                  double *unique_counts, int64_t *n_result, double *d_rate_out)
{
    return quantdt_core(n, gene_code, txid, eqClassId, nobs, equal, rcWidth, txlen, degradation_bias, maxiter, minvalue, conv,
                        txid_out, counts, full_length, unique_counts, n_result, d_rate_out);
}

int bambu_quant_last_used_device(void) { return g_last_used_device; }

// bambu_quantDT for R's .C() interface, which needs no compiled shim: every argument is a pointer to one of R's
// own vector types (integer = int, numeric = double, logical = int), scalars are length-1 vectors, the return code
// travels in the last argument.  dyn.load("libbambu_em.so"); .C("bambu_quantDT_dotC", ...) -- see INTEGRATION.md.
void bambu_quantDT_dotC(const int *n_rows, const int *gene_code, const int *txid, const int *eqClassId, const double *nobs,
                        const int *equal, const double *rcWidth, const double *txlen, const int *degradation_bias,
                        const int *maxiter, const double *minvalue, const double *conv, int *txid_out, double *counts,
                        double *full_length, double *unique_counts, int *n_result, double *d_rate, int *rc)
{
    if (!rc) return;
    if (!n_rows || !degradation_bias || !maxiter || !minvalue || !conv || !n_result || *n_rows < 0) {
        bambu_quant_set_error("NULL or negative scalar argument");
        *rc = BAMBU_EM_EINVAL;
        return;
    }
    const size_t N = (size_t)*n_rows;
    try {
        std::vector<int64_t> t_out(std::max<size_t>(N, 1));
        int64_t k = 0;
        *rc = quantdt_core((int64_t)N, gene_code, txid, eqClassId, nobs, equal, rcWidth, txlen, *degradation_bias, *maxiter,
                           *minvalue, *conv, t_out.data(), counts, full_length, unique_counts, &k, d_rate);
        if (*rc == BAMBU_EM_OK) {
            if (txid_out) for (int64_t q = 0; q < k; ++q) txid_out[q] = (int)t_out[(size_t)q];
            *n_result = (int)k;
        }
    } catch (const std::bad_alloc &) {
        bambu_quant_set_error("out of memory");
        *rc = BAMBU_EM_ENOMEM;
    }
}

}  // extern "C"

#include "quant_cohort.inl"

// releases the packer's cached device workspace (bambu_em_shutdown calls it)
void bambu_quant_release_workspace()
{
    cohort_release();
    if (qdcuda::g_plan) bambu_em_plan_destroy(qdcuda::g_plan);
    qdcuda::g_plan = nullptr;
    qdcuda::g_ws.release();
}
