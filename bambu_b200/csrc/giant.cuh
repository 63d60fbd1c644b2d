// giant.cuh -- the streamed tier: gene groups too large for one SM's shared memory
// (giant loci: thousands of transcripts, 10^5 read classes, millions of entries).
//
//   pack_giant_kernel   one CTA per giant group: raw arena slice -> global-memory image
//                       (live columns only, rows split into their even / odd streams,
//                       a column-major copy with rows ascending)
//   em_giant_kernel     cooperative launch, one CTA per SM; all CTAs iterate on ONE group
//                       at a time with two grid-wide barriers per EM update:
//                         E-phase  one thread per live column   c_j, ratio_j      (CSC)
//                         M-phase  one warp per half-row        ordered partial sums (CSR)
//                         update   every CTA forms theta' = (s_even + s_odd) / rs and the
//                                  stopping test redundantly -- no flag exchange needed
//
// Arithmetic and summation order are those of em_sell.cuh / src/em.cpp: each row sum is
// two sequential chains (even / odd dense positions) added at the end, so results and
// iteration counts are identical to the reference; a half-row's chain is evaluated by
// one warp (32 products at a time, added in order).
#pragma once
#include <cooperative_groups.h>

#include "common.cuh"
#include "em_sell.cuh"
#include "pack_kernel.cuh"

namespace bambu {

namespace cg = cooperative_groups;

// byte offsets (from the group's image base) of a giant image; sized on the host for the
// worst case (all columns live, no zero entries).  hp = half-row pointers: stream 2i holds
// row i's entries in even columns, stream 2i+1 those in odd columns.
struct GiantDesc {
    int32_t g;            // group id (absolute)
    int32_t M, N;
    int32_t NL, nnzL;     // written by the pack kernel
    int32_t pad;
    int64_t base;         // image base inside the giant buffer
    // sections (relative to base)
    int64_t rs, Yl, KY, cm, hp, cidx, cval, eq, cp, ridx, rval;
    // scratch for packing / iterating
    int64_t colmap, colcnt, cursor, hcnt, ratio, S, cntS;
    int64_t total;
};

__host__ inline void giant_layout(GiantDesc &d, int64_t M, int64_t N, int64_t e)
{
    auto al = [](int64_t x) { return (x + 15) / 16 * 16; };
    int64_t o = 0;
    d.rs = o; o = al(o + 8 * M);
    d.Yl = o; o = al(o + 8 * N);
    d.KY = o; o = al(o + 8 * N);
    d.cm = o; o = al(o + N + 1);
    d.hp = o; o = al(o + 4 * (2 * M + 1));
    d.cidx = o; o = al(o + 4 * e);
    d.cval = o; o = al(o + 8 * e);
    d.eq = o; o = al(o + e);
    d.cp = o; o = al(o + 4 * (N + 1));
    d.ridx = o; o = al(o + 4 * e);
    d.rval = o; o = al(o + 8 * e);
    d.colmap = o; o = al(o + 4 * N);
    d.colcnt = o; o = al(o + 4 * (N + 1));
    d.cursor = o; o = al(o + 4 * N);
    d.hcnt = o; o = al(o + 4 * (2 * M + 1));
    d.ratio = o; o = al(o + 8 * (N + 1));
    d.S = o; o = al(o + 8 * 6 * M);          // 2 streams x up to 3 sums
    d.cntS = o; o = al(o + 4 * 2 * M);
    d.total = o;
}

// ---- pack ------------------------------------------------------------------------------
template <int T>
__global__ void __launch_bounds__(T)
pack_giant_kernel(ArenaDev A, GiantDesc *__restrict__ descs, int n_giant, uint8_t *__restrict__ gbuf,
                  int64_t *__restrict__ nnz_nonzero, int *__restrict__ err_flag)
{
    __shared__ uint32_t s_warp[34];
    __shared__ unsigned long long s_nz;
    if ((int)blockIdx.x >= n_giant) return;
    GiantDesc &D = descs[blockIdx.x];
    const int tid = threadIdx.x;
    const int g = D.g, M = D.M, N = D.N;
    const int64_t row0 = A.grp_row_ptr[g], col0 = A.grp_col_ptr[g];
    const int64_t *rnp = A.row_nnz_ptr + row0;
    uint8_t *img = gbuf + D.base;
    double *rs = reinterpret_cast<double *>(img + D.rs);
    double *Yl = reinterpret_cast<double *>(img + D.Yl);
    double *KY = reinterpret_cast<double *>(img + D.KY);
    uint8_t *cm = img + D.cm;
    uint32_t *hp = reinterpret_cast<uint32_t *>(img + D.hp);
    uint32_t *cidx = reinterpret_cast<uint32_t *>(img + D.cidx);
    double *cval = reinterpret_cast<double *>(img + D.cval);
    uint8_t *eq = img + D.eq;
    uint32_t *cp = reinterpret_cast<uint32_t *>(img + D.cp);
    uint32_t *ridx = reinterpret_cast<uint32_t *>(img + D.ridx);
    double *rval = reinterpret_cast<double *>(img + D.rval);
    uint32_t *colmap = reinterpret_cast<uint32_t *>(img + D.colmap);
    uint32_t *colcnt = reinterpret_cast<uint32_t *>(img + D.colcnt);
    uint32_t *cursor = reinterpret_cast<uint32_t *>(img + D.cursor);
    uint32_t *hcnt = reinterpret_cast<uint32_t *>(img + D.hcnt);

    if (tid == 0) s_nz = 0;
    // A: live columns
    for (int j = tid; j < N; j += T) colmap[j] = (A.Y[col0 + j] != 0.0) ? 1u : 0u;
    __syncthreads();
    const uint32_t NL = block_excl_scan<T>(colmap, N, s_warp);
    for (int j = tid; j < N; j += T) {
        const bool live = (A.Y[col0 + j] != 0.0);
        const uint32_t lj = colmap[j];
        colmap[j] = live ? ((lj << 1) | 1u) : 0u;
        if (live) {
            const double y = A.Y[col0 + j];
            Yl[lj] = y;
            KY[lj] = __dmul_rn(A.K[col0 + j], y);
            cm[lj] = A.col_flags[col0 + j] & 1;
        }
    }
    for (int j = tid; j <= N; j += T) colcnt[j] = 0;
    __syncthreads();
    if (tid == 0) cm[NL] = 0;
    // B: rs over all entries, live entries per half-row
    unsigned long long my_nz = 0;
    for (int i = tid; i < M; i += T) {
        double a1 = 0.0, a2 = 0.0;
        uint32_t ne = 0, no = 0;
        int prev = -1;
        for (int64_t k = rnp[i]; k < rnp[i + 1]; ++k) {
            const int j = A.col_idx[k];
            if (j < 0 || j >= N) { atomicExch(err_flag, kErrColRange); continue; }
            if (j <= prev) atomicExch(err_flag, kErrColOrder);
            prev = j;
            const double a = A.aval[k];
            if (j & 1) a2 = __dadd_rn(a2, a); else a1 = __dadd_rn(a1, a);
            const bool nzv = (a != 0.0);
            my_nz += nzv;
            if (nzv && (colmap[j] & 1u)) { if (j & 1) ++no; else ++ne; }
        }
        rs[i] = __dadd_rn(a1, a2);
        hcnt[2 * i] = ne;
        hcnt[2 * i + 1] = no;
    }
    if (my_nz) atomicAdd(&s_nz, my_nz);
    __syncthreads();
    const uint32_t nnzL = block_excl_scan<T>(hcnt, 2 * M, s_warp);
    for (int t = tid; t < 2 * M; t += T) hp[t] = hcnt[t];
    if (tid == 0) hp[2 * M] = nnzL;
    __syncthreads();
    // C: compacted half-rows, column histogram
    for (int i = tid; i < M; i += T) {
        uint32_t pe = hcnt[2 * i], po = hcnt[2 * i + 1];
        for (int64_t k = rnp[i]; k < rnp[i + 1]; ++k) {
            const int j = A.col_idx[k];
            if (j < 0 || j >= N) continue;
            const double a = A.aval[k];
            const uint32_t c = colmap[j];
            if (a != 0.0 && (c & 1u)) {
                const uint32_t lj = c >> 1;
                const uint32_t pos = (j & 1) ? po++ : pe++;
                cidx[pos] = lj;
                cval[pos] = a;
                eq[pos] = A.nz_flags[k] & 1;
                atomicAdd(&colcnt[lj], 1u);
            }
        }
    }
    __syncthreads();
    block_excl_scan<T>(colcnt, (int)NL, s_warp);
    if (tid == 0) colcnt[NL] = nnzL;
    __syncthreads();
    for (int j = tid; j <= (int)NL; j += T) {
        cp[j] = colcnt[j];
        if (j < (int)NL) cursor[j] = colcnt[j];
    }
    __syncthreads();
    // D: scatter into columns
    for (int i = tid; i < M; i += T) {
        for (uint32_t p = hp[2 * i]; p < hp[2 * i + 2]; ++p) {
            const uint32_t q = atomicAdd(&cursor[cidx[p]], 1u);
            ridx[q] = ((uint32_t)i << 1) | (uint32_t)(i & 1);
            rval[q] = cval[p];
        }
    }
    __syncthreads();
    // E: rows ascending inside every column
    for (int j = tid; j < (int)NL; j += T) {
        const uint32_t lo = cp[j], hi = cp[j + 1];
        for (uint32_t p = lo + 1; p < hi; ++p) {
            const uint32_t kr = ridx[p];
            const double kv = rval[p];
            uint32_t q = p;
            while (q > lo && ridx[q - 1] > kr) {
                ridx[q] = ridx[q - 1];
                rval[q] = rval[q - 1];
                --q;
            }
            ridx[q] = kr;
            rval[q] = kv;
        }
    }
    if (tid == 0) {
        D.NL = (int32_t)NL;
        D.nnzL = (int32_t)nnzL;
        nnz_nonzero[g] = (int64_t)s_nz;
    }
}

// ---- EM ------------------------------------------------------------------------------------
// ordered sum of one stream [lo, hi): 32 products at a time, added in stream order.
// NSUM = 1: sum of (v*th)*x ; NSUM = 3: also the full-length and unique variants (epilogue).
template <bool HAZARD>
__device__ __forceinline__ double stream_sum(const uint32_t *__restrict__ cidx, const double *__restrict__ cval,
                                             uint32_t lo, uint32_t hi, int lane, double th,
                                             const double *ratio)
{
    double acc = 0.0;
    for (uint32_t base = lo; base < hi; base += 32) {
        const uint32_t k = base + lane;
        double t = 0.0;
        if (k < hi) {
            t = __dmul_rn(__dmul_rn(cval[k], th), __ldcg(ratio + cidx[k]));
            if (HAZARD && t != t) t = 0.0;   // replace(nan, 0)  (src/em.cpp:49)
        }
        const int n = min(32u, hi - base);
        if (n == 32) {
#pragma unroll
            for (int q = 0; q < 32; ++q) acc = __dadd_rn(acc, __shfl_sync(0xffffffffu, t, q));
        } else {
            for (int q = 0; q < n; ++q) acc = __dadd_rn(acc, __shfl_sync(0xffffffffu, t, q));
        }
    }
    return acc;
}

template <int T>
__global__ void __launch_bounds__(T)
em_giant_kernel(const GiantDesc *__restrict__ descs, int n_giant, const uint8_t *gbuf_c, uint8_t *gbuf,
                const int64_t *__restrict__ grp_row_ptr, int64_t total_rows, EmParams P, double *__restrict__ est,
                int32_t *__restrict__ iters, int32_t *__restrict__ status, unsigned int *gflags /* 8 words, zeroed */)
{
    extern __shared__ __align__(16) uint8_t smem[];
    cg::grid_group grid = cg::this_grid();
    double *theta = reinterpret_cast<double *>(smem);   // M, private copy per CTA
    const int tid = threadIdx.x, lane = tid & 31;
    const int64_t gtid = (int64_t)blockIdx.x * T + tid, gthreads = (int64_t)gridDim.x * T;
    const int64_t gwarp = gtid >> 5, gwarps = gthreads >> 5;
    (void)gbuf_c;

    for (int gi = 0; gi < n_giant; ++gi) {
        const GiantDesc D = descs[gi];
        const int M = D.M, NL = D.NL;
        uint8_t *img = gbuf + D.base;
        const double *rs = reinterpret_cast<const double *>(img + D.rs);
        const double *Yl = reinterpret_cast<const double *>(img + D.Yl);
        const double *KY = reinterpret_cast<const double *>(img + D.KY);
        const uint8_t *cm = img + D.cm;
        const uint32_t *hp = reinterpret_cast<const uint32_t *>(img + D.hp);
        const uint32_t *cidx = reinterpret_cast<const uint32_t *>(img + D.cidx);
        const double *cval = reinterpret_cast<const double *>(img + D.cval);
        const uint8_t *eq = img + D.eq;
        const uint32_t *cp = reinterpret_cast<const uint32_t *>(img + D.cp);
        const uint32_t *ridx = reinterpret_cast<const uint32_t *>(img + D.ridx);
        const double *rval = reinterpret_cast<const double *>(img + D.rval);
        double *ratio = reinterpret_cast<double *>(img + D.ratio);
        double *S = reinterpret_cast<double *>(img + D.S);
        uint32_t *cntS = reinterpret_cast<uint32_t *>(img + D.cntS);
        unsigned int *hz = gflags;        // [0..1] hazard flags by update parity, [2] Q
        const int64_t row0 = grp_row_ptr[D.g];

        for (int i = tid; i < M; i += T) theta[i] = 1.0;   // src/em.cpp:20-21
        if (gtid == 0) { hz[0] = hz[1] = 0; hz[2] = 0; }
        int nP = 0;                                          // non-finite thetas (same value in every CTA)
        int iter = 0;
        bool over = (1.0 > P.conv);
        grid.sync();
        while (over && iter < P.maxiter - 1) {
            ++iter;
            const int par = iter & 1;
            // E-phase: c_j = sum_i a_ij*theta_i, ratio_j = Y_j / c_j   (src/em.cpp:46-48)
            for (int64_t j = gtid; j < NL; j += gthreads) {
                double a1 = 0.0, a2 = 0.0;
                int cnt = 0;
                for (uint32_t k = cp[j]; k < cp[j + 1]; ++k) {
                    const uint32_t ri = ridx[k];
                    const double th = theta[ri >> 1];
                    const double t = __dmul_rn(rval[k], th);
                    if (nP) cnt += !is_finite(th);
                    if (ri & 1u) a2 = __dadd_rn(a2, t); else a1 = __dadd_rn(a1, t);
                }
                if (cnt < nP) a1 = qnan();   // dense 0*NaN at this column's structural zeros (Appendix C.14)
                const double r = __ddiv_rn(Yl[j], __dadd_rn(a1, a2));
                ratio[j] = r;
                if (!is_finite(r)) atomicOr(&hz[par], 1u);
            }
            grid.sync();
            // M-phase: one warp per half-row, ordered chain  (src/em.cpp:48-50)
            const bool hazard = (__ldcg(&hz[par]) != 0) || (nP != 0);
            if (gtid == 0) hz[par ^ 1] = 0;
            for (int64_t task = gwarp; task < 2 * (int64_t)M; task += gwarps) {
                const double th = theta[task >> 1];
                const uint32_t lo = hp[task], hi = hp[task + 1];
                const double s = hazard ? stream_sum<true>(cidx, cval, lo, hi, lane, th, ratio)
                                        : stream_sum<false>(cidx, cval, lo, hi, lane, th, ratio);
                if (lane == 0) S[task] = s;
            }
            grid.sync();
            // update: every CTA forms the whole theta' and the stopping test (src/em.cpp:50-63)
            int flag = (P.conv < 0.0) ? 1 : 0, nf = 0;
            for (int ib = 0; ib < M; ib += T) {      // uniform trip count: the band test votes per warp
                const int i = ib + tid;
                bool inband = false;
                double after = 0.0, diff = 0.0;
                if (i < M) {
                    const double th = theta[i];
                    after = __ddiv_rn(__dadd_rn(__ldcg(S + 2 * i), __ldcg(S + 2 * i + 1)), rs[i]);
                    theta[i] = after;
                    nf += !is_finite(after);
                    if (after > P.minvalue) {
                        diff = fabs(__dadd_rn(after, -th));
                        if (diff > __dmul_rn(after, P.conv_hi)) flag = 1;
                        else inband = (diff >= __dmul_rn(after, P.conv_lo));
                    }
                }
                if (__any_sync(0xffffffffu, inband)) {
                    if (inband && __ddiv_rn(diff, after) > P.conv) flag = 1;
                }
            }
            // block-wide: any flag, total non-finite count (identical in every CTA)
            over = __syncthreads_or(flag) != 0;
            {
                __shared__ int s_nf;
                if (tid == 0) s_nf = 0;
                __syncthreads();
                if (nf) atomicAdd(&s_nf, nf);
                __syncthreads();
                nP = s_nf;
                __syncthreads();
            }
        }

        // ---- epilogue: src/em.cpp:97-102 ----------------------------------------------------
        for (int64_t j = gtid; j < NL; j += gthreads) {
            double a1 = 0.0, a2 = 0.0;
            int cnt = 0;
            for (uint32_t k = cp[j]; k < cp[j + 1]; ++k) {
                const uint32_t ri = ridx[k];
                const double th = theta[ri >> 1];
                const double t = __dmul_rn(rval[k], th);
                cnt += !is_finite(th);
                if (ri & 1u) a2 = __dadd_rn(a2, t); else a1 = __dadd_rn(a1, t);
            }
            double c = __dadd_rn(a1, a2);
            if (cnt < nP) c = qnan();
            double b = __ddiv_rn(KY[j], c);
            if (b != b) b = 0.0;
            ratio[j] = b;
            if (fabs(b) > 1.7976931348623157e308) atomicAdd(&hz[2], 1u);
        }
        grid.sync();
        for (int64_t task = gwarp; task < 2 * (int64_t)M; task += gwarps) {
            const double th = theta[task >> 1];
            const uint32_t lo = hp[task], hi = hp[task + 1];
            double st = 0.0, sf = 0.0, su = 0.0;
            uint32_t cnt = 0;
            for (uint32_t base = lo; base < hi; base += 32) {
                const uint32_t k = base + lane;
                double t = 0.0, f = 0.0, u = 0.0;
                uint32_t isinf_b = 0;
                if (k < hi) {
                    const uint32_t col = cidx[k];
                    const double b = __ldcg(ratio + col);
                    const double a = cval[k];
                    const double af = eq[k] ? a : __dmul_rn(a, 0.0);
                    const double au = cm[col] ? __dmul_rn(a, 0.0) : a;
                    t = __dmul_rn(__dmul_rn(a, th), b);
                    f = __dmul_rn(__dmul_rn(af, th), b);
                    u = __dmul_rn(__dmul_rn(au, th), b);
                    isinf_b = (fabs(b) > 1.7976931348623157e308);
                }
                cnt += __popc(__ballot_sync(0xffffffffu, isinf_b));
                const int n = min(32u, hi - base);
                for (int q = 0; q < n; ++q) {
                    st = __dadd_rn(st, __shfl_sync(0xffffffffu, t, q));
                    sf = __dadd_rn(sf, __shfl_sync(0xffffffffu, f, q));
                    su = __dadd_rn(su, __shfl_sync(0xffffffffu, u, q));
                }
            }
            if (lane == 0) {
                S[3 * task] = st; S[3 * task + 1] = sf; S[3 * task + 2] = su;
                cntS[task] = cnt;
            }
        }
        grid.sync();
        const unsigned int Q = __ldcg(&hz[2]);
        int bad = 0;
        for (int64_t i = gtid; i < M; i += gthreads) {
            const double th = theta[i];
            double e0 = __dadd_rn(__ldcg(S + 6 * i), __ldcg(S + 6 * i + 3));
            double e1 = __dadd_rn(__ldcg(S + 6 * i + 1), __ldcg(S + 6 * i + 4));
            double e2 = __dadd_rn(__ldcg(S + 6 * i + 2), __ldcg(S + 6 * i + 5));
            const unsigned int cnt = __ldcg(cntS + 2 * i) + __ldcg(cntS + 2 * i + 1);
            if (!is_finite(th) || cnt < Q) e0 = e1 = e2 = qnan();
            est[row0 + i] = e0;
            est[total_rows + row0 + i] = e1;
            est[2 * total_rows + row0 + i] = e2;
            bad |= !(is_finite(e0) && is_finite(e1) && is_finite(e2));
        }
        if (bad) atomicOr(&hz[3], 1u);
        grid.sync();
        if (gtid == 0) {
            iters[D.g] = iter;
            status[D.g] = ((iter >= P.maxiter - 1 && over) ? 1 : 0) | (__ldcg(&hz[3]) ? 2 : 0);
            hz[3] = 0;
        }
        grid.sync();
    }
}

}  // namespace bambu
