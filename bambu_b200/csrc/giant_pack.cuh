// giant_pack.cuh -- image of the streamed tier's groups from the wire form v2, built by the WHOLE GPU.
//
// pack_giant_kernel (giant.cuh) gives a locus one CTA: fine for the few medium groups that spill out of the resident
// tier, but a 3 M-entry locus then takes ~25 ms (L2 latency of one SM's worth of threads).  Here every phase is a flat
// grid over (locus = blockIdx.y, tile = blockIdx.x): warps own rows, threads own columns; the only per-locus
// serial steps are three block scans.  Same image (GiantDesc), same order inside every line, so em_giant_kernel and
// its results do not change.
//
//   gp_cols      per live column: Y (= nobs / K), K*Y, multi_align; zero the histogram
//   gp_rowcount  warp per row: indices in range / strictly increasing, entries per parity stream
//   gp_scan_rows block scan of the 2M stream lengths -> hp
//   gp_rowfill   warp per row: entries into their stream (ballot prefix), column histogram, one parity per column
//   gp_scan_cols block scan -> cp, cursor
//   gp_scatter   warp per row: entries into their column (atomic cursor; arbitrary order inside a column)
//   gp_sortcols  thread per column: rows ascending (the order arma::sum adds them) + length histogram
//   gp_sell_*    columns by length into slices of 32, slot-major (the E-phase reads coalesced lines)
#pragma once
#include "giant.cuh"

namespace bambu {

constexpr int kGpThreads = 256;

struct GpWire {           // per-locus views of the wire form
    const uint32_t *rend;
    const uint16_t *c16;
    const uint32_t *c32;
    const double *av;
    bool wide;
    int64_t row0, col0, ent0;
    uint32_t e;
};

__device__ __forceinline__ GpWire gp_wire(const WireDev &W, const GiantDesc &D)
{
    GpWire w;
    const int g = D.g;
    w.row0 = W.grp_row_ptr[g]; w.col0 = W.grp_col_ptr[g]; w.ent0 = W.grp_ent_ptr[g];
    w.e = (uint32_t)(W.grp_ent_ptr[g + 1] - w.ent0);
    w.rend = W.row_end + w.row0;
    w.c16 = reinterpret_cast<const uint16_t *>(W.cidx + W.grp_idx_off[g]);
    w.c32 = reinterpret_cast<const uint32_t *>(W.cidx + W.grp_idx_off[g]);
    w.wide = wire_idx_width(D.N) == 4;
    w.av = W.aval + w.ent0;
    return w;
}

__global__ void __launch_bounds__(kGpThreads) gp_cols(WireDev W, GiantDesc *descs, uint8_t *gbuf)
{
    GiantDesc &D = descs[blockIdx.y];
    uint8_t *img = gbuf + D.base;
    const int N = D.N;
    const int64_t col0 = W.grp_col_ptr[D.g];
    double *Yl = reinterpret_cast<double *>(img + D.Yl), *KY = reinterpret_cast<double *>(img + D.KY);
    uint8_t *cm = img + D.cm;
    uint32_t *colmap = reinterpret_cast<uint32_t *>(img + D.colmap), *colcnt = reinterpret_cast<uint32_t *>(img + D.colcnt);
    for (int j = blockIdx.x * kGpThreads + threadIdx.x; j <= N; j += gridDim.x * kGpThreads) {
        colcnt[j] = 0;
        if (j == N) { cm[N] = 0; break; }
        double y, kk;
        wire_col(W, col0 + j, y, kk);
        Yl[j] = y;
        KY[j] = __dmul_rn(kk, y);
        cm[j] = (uint8_t)wire_bit(W.multi_bits, col0 + j);
        colmap[j] = 0;              // parities seen
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { D.NL = N; }
}

__global__ void __launch_bounds__(kGpThreads) gp_rowcount(WireDev W, GiantDesc *descs, uint8_t *gbuf, int *err_flag)
{
    const GiantDesc &D = descs[blockIdx.y];
    const GpWire w = gp_wire(W, D);
    uint8_t *img = gbuf + D.base;
    const int M = D.M, lane = threadIdx.x & 31;
    const uint32_t NL = (uint32_t)D.N;
    double *rs = reinterpret_cast<double *>(img + D.rs);
    uint32_t *hcnt = reinterpret_cast<uint32_t *>(img + D.hcnt);
    const int wpb = kGpThreads / 32;
    if (blockIdx.x == 0 && threadIdx.x == 0 && M > 0 && w.rend[M - 1] != w.e) atomicOr(err_flag, kErrRowPtr);
    for (int i = blockIdx.x * wpb + (threadIdx.x >> 5); i < M; i += gridDim.x * wpb) {
        const uint32_t lo = i ? w.rend[i - 1] : 0u, hi = w.rend[i];
        uint32_t ne = 0, no = 0;
        if (hi < lo || hi > w.e) { if (lane == 0) atomicOr(err_flag, kErrRowPtr); }
        else {
            for (uint32_t base = lo; base < hi; base += 32) {
                const uint32_t k = base + lane;
                const bool in = k < hi;
                const uint32_t c = in ? (w.wide ? w.c32[k] : (uint32_t)w.c16[k]) : 0u;
                const uint32_t prev = (in && k > lo) ? (w.wide ? w.c32[k - 1] : (uint32_t)w.c16[k - 1]) : 0u;
                const bool ok = in && (c >> 1) < NL;
                if (in && !ok) atomicOr(err_flag, kErrColRange);
                if (in && k > lo && (c >> 1) <= (prev >> 1)) atomicOr(err_flag, kErrColOrder);
                const uint32_t mo = __ballot_sync(0xffffffffu, ok && (c & 1u)), me = __ballot_sync(0xffffffffu, ok && !(c & 1u));
                no += __popc(mo);
                ne += __popc(me);
            }
        }
        if (lane == 0) {
            rs[i] = W.rs[w.row0 + i];
            hcnt[2 * i] = ne;
            hcnt[2 * i + 1] = no;
        }
    }
}

// exclusive block scan of a[0..n) in place by one CTA of 1024 threads; total returned to every thread
__global__ void __launch_bounds__(1024) gp_scan_rows(GiantDesc *descs, uint8_t *gbuf, int64_t *nnz_nonzero, WireDev W)
{
    __shared__ uint32_t s_warp[34];
    GiantDesc &D = descs[blockIdx.y];
    uint8_t *img = gbuf + D.base;
    uint32_t *hcnt = reinterpret_cast<uint32_t *>(img + D.hcnt), *hp = reinterpret_cast<uint32_t *>(img + D.hp);
    const int n = 2 * D.M;
    const uint32_t total = block_excl_scan<1024>(hcnt, n, s_warp);
    for (int t = threadIdx.x; t < n; t += 1024) hp[t] = hcnt[t];
    if (threadIdx.x == 0) {
        hp[n] = total;
        D.nnzL = (int32_t)total;
        nnz_nonzero[D.g] = (int64_t)(W.grp_ent_ptr[D.g + 1] - W.grp_ent_ptr[D.g]);
    }
}

__global__ void __launch_bounds__(kGpThreads) gp_rowfill(WireDev W, GiantDesc *descs, uint8_t *gbuf, int *err_flag)
{
    const GiantDesc &D = descs[blockIdx.y];
    const GpWire w = gp_wire(W, D);
    uint8_t *img = gbuf + D.base;
    const int M = D.M, lane = threadIdx.x & 31;
    const uint32_t NL = (uint32_t)D.N;
    const uint32_t *hp = reinterpret_cast<const uint32_t *>(img + D.hp);
    uint32_t *cidx = reinterpret_cast<uint32_t *>(img + D.cidx);
    double *cval = reinterpret_cast<double *>(img + D.cval);
    uint8_t *eq = img + D.eq;
    uint32_t *colcnt = reinterpret_cast<uint32_t *>(img + D.colcnt), *colmap = reinterpret_cast<uint32_t *>(img + D.colmap);
    const int wpb = kGpThreads / 32;
    const uint32_t lt = (1u << lane) - 1u;
    for (int i = blockIdx.x * wpb + (threadIdx.x >> 5); i < M; i += gridDim.x * wpb) {
        const uint32_t lo = i ? w.rend[i - 1] : 0u, hi = w.rend[i];
        if (hi < lo || hi > w.e) continue;
        uint32_t pe = hp[2 * i], po = hp[2 * i + 1];
        for (uint32_t base = lo; base < hi; base += 32) {
            const uint32_t k = base + lane;
            const bool in = k < hi;
            const uint32_t c = in ? (w.wide ? w.c32[k] : (uint32_t)w.c16[k]) : 0u;
            const bool ok = in && (c >> 1) < NL;
            const bool odd = (c & 1u) != 0;
            const uint32_t mo = __ballot_sync(0xffffffffu, ok && odd), me = __ballot_sync(0xffffffffu, ok && !odd);
            if (ok) {
                const uint32_t lj = c >> 1;
                const uint32_t pos = odd ? po + __popc(mo & lt) : pe + __popc(me & lt);
                cidx[pos] = lj;
                cval[pos] = w.av[k];
                eq[pos] = (uint8_t)wire_bit(W.eq_bits, w.ent0 + k);
                atomicAdd(&colcnt[lj], 1u);
                const uint32_t bit = 1u << (c & 1u);
                if (!(colmap[lj] & bit)) { if (atomicOr(&colmap[lj], bit) & (bit ^ 3u)) atomicOr(err_flag, kErrParity); }
            }
            po += __popc(mo);
            pe += __popc(me);
        }
    }
}

__global__ void __launch_bounds__(1024) gp_scan_cols(GiantDesc *descs, uint8_t *gbuf)
{
    __shared__ uint32_t s_warp[34];
    const GiantDesc &D = descs[blockIdx.y];
    uint8_t *img = gbuf + D.base;
    uint32_t *colcnt = reinterpret_cast<uint32_t *>(img + D.colcnt), *cursor = reinterpret_cast<uint32_t *>(img + D.cursor);
    uint32_t *cp = reinterpret_cast<uint32_t *>(img + D.cp);
    const int NL = D.N;
    const uint32_t total = block_excl_scan<1024>(colcnt, NL, s_warp);
    for (int j = threadIdx.x; j < NL; j += 1024) { cp[j] = colcnt[j]; cursor[j] = colcnt[j]; }
    if (threadIdx.x == 0) cp[NL] = total;
}

__global__ void __launch_bounds__(kGpThreads) gp_scatter(GiantDesc *descs, uint8_t *gbuf)
{
    const GiantDesc &D = descs[blockIdx.y];
    uint8_t *img = gbuf + D.base;
    const int M = D.M, lane = threadIdx.x & 31;
    const uint32_t *hp = reinterpret_cast<const uint32_t *>(img + D.hp);
    const uint32_t *cidx = reinterpret_cast<const uint32_t *>(img + D.cidx);
    const double *cval = reinterpret_cast<const double *>(img + D.cval);
    uint32_t *cursor = reinterpret_cast<uint32_t *>(img + D.cursor);
    uint32_t *ridx = reinterpret_cast<uint32_t *>(img + D.ridx);
    double *rval = reinterpret_cast<double *>(img + D.rval);
    const int wpb = kGpThreads / 32;
    for (int i = blockIdx.x * wpb + (threadIdx.x >> 5); i < M; i += gridDim.x * wpb) {
        const uint32_t lo = hp[2 * i], hi = hp[2 * i + 2];
        for (uint32_t p = lo + lane; p < hi; p += 32) {
            const uint32_t q = atomicAdd(&cursor[cidx[p]], 1u);
            ridx[q] = ((uint32_t)i << 1) | (uint32_t)(i & 1);
            rval[q] = cval[p];
        }
    }
}

// rows ascending inside every column; histogram of the column lengths (reversed bins: ascending scan = descending length)
__global__ void __launch_bounds__(kGpThreads) gp_sortcols(GiantDesc *descs, uint8_t *gbuf, uint32_t *hist_all)
{
    const GiantDesc &D = descs[blockIdx.y];
    uint8_t *img = gbuf + D.base;
    const int NL = D.N;
    const uint32_t *cp = reinterpret_cast<const uint32_t *>(img + D.cp);
    uint32_t *ridx = reinterpret_cast<uint32_t *>(img + D.ridx);
    double *rval = reinterpret_cast<double *>(img + D.rval);
    uint32_t *hist = hist_all + 1024 * (size_t)blockIdx.y;
    for (int j = blockIdx.x * kGpThreads + threadIdx.x; j < NL; j += gridDim.x * kGpThreads) {
        const uint32_t lo = cp[j], hi = cp[j + 1];
        sort_column(ridx, rval, lo, hi);
        atomicAdd(&hist[1023 - min(hi - lo, 1023u)], 1u);
    }
}

__global__ void __launch_bounds__(1024) gp_sell_scan_hist(uint32_t *hist_all)
{
    __shared__ uint32_t s_warp[34];
    block_excl_scan<1024>(hist_all + 1024 * (size_t)blockIdx.y, 1024, s_warp);
}

__global__ void __launch_bounds__(kGpThreads) gp_sell_place(GiantDesc *descs, uint8_t *gbuf, uint32_t *hist_all)
{
    const GiantDesc &D = descs[blockIdx.y];
    uint8_t *img = gbuf + D.base;
    const int NL = D.N, nsl = (NL + 31) / 32;
    const uint32_t *cp = reinterpret_cast<const uint32_t *>(img + D.cp);
    uint32_t *ccol = reinterpret_cast<uint32_t *>(img + D.ccol);
    uint32_t *hist = hist_all + 1024 * (size_t)blockIdx.y;
    for (int j = blockIdx.x * kGpThreads + threadIdx.x; j < nsl * 32; j += gridDim.x * kGpThreads) {
        if (j < NL) ccol[atomicAdd(&hist[1023 - min(cp[j + 1] - cp[j], 1023u)], 1u)] = (uint32_t)j;
        else ccol[j] = (uint32_t)NL;          // padding positions of the last slice (never claimed by a column)
    }
}

// slice length = longest column of the slice (warp per slice)
__global__ void __launch_bounds__(kGpThreads) gp_sell_slices(GiantDesc *descs, uint8_t *gbuf)
{
    const GiantDesc &D = descs[blockIdx.y];
    uint8_t *img = gbuf + D.base;
    const int NL = D.N, nsl = (NL + 31) / 32, lane = threadIdx.x & 31;
    const uint32_t *cp = reinterpret_cast<const uint32_t *>(img + D.cp);
    const uint32_t *ccol = reinterpret_cast<const uint32_t *>(img + D.ccol);
    uint32_t *cstart = reinterpret_cast<uint32_t *>(img + D.cstart);
    const int wpb = kGpThreads / 32;
    for (int w = blockIdx.x * wpb + (threadIdx.x >> 5); w < nsl; w += gridDim.x * wpb) {
        const uint32_t j = ccol[w * 32 + lane];
        uint32_t L = (j < (uint32_t)NL) ? cp[j + 1] - cp[j] : 0u;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) L = max(L, __shfl_xor_sync(0xffffffffu, L, o));
        if (lane == 0) cstart[w] = L;
    }
}

__global__ void __launch_bounds__(1024) gp_sell_scan_slices(GiantDesc *descs, uint8_t *gbuf)
{
    __shared__ uint32_t s_warp[34];
    GiantDesc &D = descs[blockIdx.y];
    uint8_t *img = gbuf + D.base;
    const int nsl = (D.N + 31) / 32;
    uint32_t *cstart = reinterpret_cast<uint32_t *>(img + D.cstart);
    const uint32_t slots = block_excl_scan<1024>(cstart, nsl, s_warp);
    if (threadIdx.x == 0) { cstart[nsl] = slots; D.nSliceC = nsl; D.slotsC = (int32_t)slots; }
}

__global__ void __launch_bounds__(kGpThreads) gp_sell_fill(GiantDesc *descs, uint8_t *gbuf)
{
    const GiantDesc &D = descs[blockIdx.y];
    uint8_t *img = gbuf + D.base;
    const int NL = D.N, nsl = (NL + 31) / 32, M = D.M;
    const uint32_t *cp = reinterpret_cast<const uint32_t *>(img + D.cp);
    const uint32_t *ridx = reinterpret_cast<const uint32_t *>(img + D.ridx);
    const double *rval = reinterpret_cast<const double *>(img + D.rval);
    const uint32_t *ccol = reinterpret_cast<const uint32_t *>(img + D.ccol);
    const uint32_t *cstart = reinterpret_cast<const uint32_t *>(img + D.cstart);
    double *sval = reinterpret_cast<double *>(img + D.sval);
    uint32_t *sidx = reinterpret_cast<uint32_t *>(img + D.sidx);
    for (int p = blockIdx.x * kGpThreads + threadIdx.x; p < nsl * 32; p += gridDim.x * kGpThreads) {
        const uint32_t j = ccol[p];
        const int ln = p & 31, w = p >> 5;
        const uint32_t s0 = cstart[w], Ls = cstart[w + 1] - s0;
        uint32_t k = 0;
        if (j < (uint32_t)NL) {
            for (uint32_t q = cp[j]; q < cp[j + 1]; ++q, ++k) {
                sval[(size_t)(s0 + k) * 32 + ln] = rval[q];
                sidx[(size_t)(s0 + k) * 32 + ln] = ridx[q];
            }
        }
        for (; k < Ls; ++k) {     // padding: value 0 times theta[M] = 0
            sval[(size_t)(s0 + k) * 32 + ln] = 0.0;
            sidx[(size_t)(s0 + k) * 32 + ln] = ((uint32_t)M << 1);
        }
    }
}

}  // namespace bambu
