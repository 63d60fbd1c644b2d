// giant.cuh -- the streamed tier: gene groups too large for one SM's shared memory
// (giant loci: thousands of transcripts, 10^5 read classes, millions of entries).
//
//   pack_giant_kernel   one CTA per group: v1 arena slice -> global-memory image (live columns only, rows split into
//                       their even / odd streams, a column-major copy with rows ascending, the column side in SELL
//                       order).  The wire form v2 is packed by giant_pack.cuh with the whole GPU.
//   em_giant_kernel     cooperative launch, one CTA per SM, split into teams of CTAs; a team iterates on one group at
//                       a time with two team-wide barriers per EM update:
//                         E-phase  one lane per live column on the SELL copy      c_j, ratio_j
//                         M-phase  one warp per block of rows: ordered sums of every parity stream, theta', stopping test
//                         then every CTA of the team refreshes its private copy of theta (shared memory)
//
// Arithmetic and summation order are those of em_sell.cuh / src/em.cpp: each row sum is
// two sequential chains (even / odd dense positions) added at the end, so results and
// iteration counts are identical to the reference.
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
    // column side in SELL order (slot-major inside slices of 32 length-sorted columns): E-phase reads are coalesced
    int64_t ccol, cstart, sval, sidx;
    int32_t nSliceC, slotsC;
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
    {
        const int64_t nsl = (N + 31) / 32;
        const int64_t slots_ub = e / 32 + M + nsl + 2;      // sum of slice maxima of length-sorted columns
        d.ccol = o; o = al(o + 4 * 32 * nsl);
        d.cstart = o; o = al(o + 4 * (nsl + 1));
        d.sval = o; o = al(o + 8 * 32 * slots_ub);
        d.sidx = o; o = al(o + 4 * 32 * slots_ub);
    }
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
    uint32_t NL, nnzL;
    {
    const int64_t row0 = A.grp_row_ptr[g], col0 = A.grp_col_ptr[g];
    const int64_t *rnp = A.row_nnz_ptr + row0;
    // A: live columns
    for (int j = tid; j < N; j += T) colmap[j] = (A.Y[col0 + j] != 0.0) ? 1u : 0u;
    __syncthreads();
    NL = block_excl_scan<T>(colmap, N, s_warp);
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
            if (j < 0 || j >= N) { atomicOr(err_flag, kErrColRange); continue; }
            if (j <= prev) atomicOr(err_flag, kErrColOrder);
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
    nnzL = block_excl_scan<T>(hcnt, 2 * M, s_warp);
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
    for (int j = tid; j < (int)NL; j += T) sort_column(ridx, rval, cp[j], cp[j + 1]);
    __syncthreads();
    // F: column side in SELL order.  colmap/colcnt/cursor are free now: reuse them as
    //    len[NL] (colmap), histogram (shared), position -> column (ccol)
    {
        __shared__ uint32_t hist[1024];
        uint32_t *ccol = reinterpret_cast<uint32_t *>(img + D.ccol);
        uint32_t *cstart = reinterpret_cast<uint32_t *>(img + D.cstart);
        double *sval = reinterpret_cast<double *>(img + D.sval);
        uint32_t *sidx = reinterpret_cast<uint32_t *>(img + D.sidx);
        const int nsl = ((int)NL + 31) / 32;
        for (int b = tid; b < 1024; b += T) hist[b] = 0;
        __syncthreads();
        for (int j = tid; j < (int)NL; j += T) atomicAdd(&hist[1023 - min(cp[j + 1] - cp[j], 1023u)], 1u);
        __syncthreads();
        block_excl_scan<T>(hist, 1024, s_warp);
        for (int j = tid; j < (int)NL; j += T) {
            const uint32_t pos = atomicAdd(&hist[1023 - min(cp[j + 1] - cp[j], 1023u)], 1u);
            ccol[pos] = (uint32_t)j;
        }
        for (int p = (int)NL + tid; p < nsl * 32; p += T) ccol[p] = NL;
        __syncthreads();
        const int lane = tid & 31, warp = tid >> 5;
        for (int w = warp; w < nsl; w += T / 32) {
            const uint32_t j = ccol[w * 32 + lane];
            uint32_t L = (j < NL) ? cp[j + 1] - cp[j] : 0u;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) L = max(L, __shfl_xor_sync(0xffffffffu, L, o));
            if (lane == 0) cstart[w] = L;
        }
        __syncthreads();
        const uint32_t slots = block_excl_scan<T>(cstart, nsl, s_warp);
        if (tid == 0) { cstart[nsl] = slots; D.nSliceC = nsl; D.slotsC = (int32_t)slots; }
        __syncthreads();
        for (int p = tid; p < nsl * 32; p += T) {
            const uint32_t j = ccol[p];
            const int ln = p & 31, w = p >> 5;
            const uint32_t s0 = cstart[w], Ls = cstart[w + 1] - s0;
            uint32_t k = 0;
            if (j < NL) {
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
    if (tid == 0) {
        D.NL = (int32_t)NL;
        D.nnzL = (int32_t)nnzL;
        nnz_nonzero[g] = (int64_t)s_nz;
    }
}

// ---- grid-wide barrier (cooperative launch => every CTA is resident) -------------------------
// bar[0] = arrival counter, bar[1] = generation.  One atomic per CTA, the last arrival opens the
// next generation; the others spin on it.  Cheaper than cg::grid_group::sync() (no L1 flush).
__device__ __forceinline__ void grid_barrier(unsigned int *bar, unsigned int n_ctas)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        volatile unsigned int *gen = bar + 1;
        const unsigned int my_gen = *gen;
        __threadfence();                                   // this CTA's writes before the arrival
        if (atomicAdd(bar, 1u) == n_ctas - 1) {
            bar[0] = 0;
            __threadfence();
            atomicAdd(bar + 1, 1u);                        // release
        } else {
            while (*gen == my_gen) { __nanosleep(32); }
        }
        __threadfence();                                   // acquire
    }
    __syncthreads();
}

// ---- EM ------------------------------------------------------------------------------------
// ---- M-phase of a block of `rb` rows by one warp ------------------------------------------------------------------
// An ordered sum is a chain of dependent FP64 adds (8.4 cycles each): one row per warp keeps 2 of 32 lanes adding.  Here
// a warp owns rb rows = 2*rb parity streams (segments): for every batch of 32 entries the whole warp forms the products
// of each segment in turn (coalesced value / index loads, one ratio gather per lane) and parks them in shared memory;
// then lane s adds the 32 products of segment s in index order -- 2*rb chains advance per instruction instead of 2.
// Same terms, same order per stream as row_streams / em_sell.cuh: identical doubles.
constexpr int kStageStride = 33;                 // doubles per segment in the staging area (33: conflict-free for the adding lanes)
__host__ __device__ inline uint32_t giant_stage_bytes(int rb) { return (uint32_t)(2 * rb * kStageStride * 8); }

template <bool HAZARD, int NSEG>
__device__ __forceinline__ double row_block_streams(const uint32_t *__restrict__ hp, const uint32_t *__restrict__ cidx,
                                                    const double *__restrict__ cval, int r0, int nr, int lane,
                                                    const double *theta, const double *ratio, double *stage)
{
    const int nseg = 2 * nr;          // <= NSEG
    // the lane's own segment in the add stage
    uint32_t my_lo = 0, my_n = 0;
    if (lane < nseg) { my_lo = hp[2 * r0 + lane]; my_n = hp[2 * r0 + lane + 1] - my_lo; }
    uint32_t nmax = my_n;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nmax = max(nmax, __shfl_xor_sync(0xffffffffu, nmax, o));
    // per segment (warp-uniform): bounds and the row's theta; lane `seg` holds the bounds
    uint32_t lo[NSEG], hi[NSEG];
    double th[NSEG];
#pragma unroll
    for (int seg = 0; seg < NSEG; ++seg) {
        lo[seg] = __shfl_sync(0xffffffffu, my_lo, seg);
        hi[seg] = lo[seg] + __shfl_sync(0xffffffffu, my_n, seg);
        th[seg] = (seg < nseg) ? theta[r0 + (seg >> 1)] : 0.0;
    }
    // software pipeline: the values / column indices of batch b + 1 are in flight while batch b gathers its ratios,
    // multiplies and adds (the SM issues in order: a load must be issued well before its first use)
    double v[NSEG];
    uint32_t c[NSEG];
#pragma unroll
    for (int seg = 0; seg < NSEG; ++seg) {
        const uint32_t k = lo[seg] + lane;
        const bool in = k < hi[seg];
        v[seg] = in ? cval[k] : 0.0;
        c[seg] = in ? cidx[k] : 0xffffffffu;
    }
    double acc = 0.0;
    for (uint32_t base = 0; base < nmax; base += 32) {
        double vn[NSEG];
        uint32_t cn[NSEG];
#pragma unroll
        for (int seg = 0; seg < NSEG; ++seg) {            // batch b + 1
            const uint32_t k = lo[seg] + base + 32 + lane;
            const bool in = k < hi[seg];
            vn[seg] = in ? cval[k] : 0.0;
            cn[seg] = in ? cidx[k] : 0xffffffffu;
        }
        double r[NSEG];
#pragma unroll
        for (int seg = 0; seg < NSEG; ++seg) r[seg] = (c[seg] != 0xffffffffu) ? __ldcg(ratio + c[seg]) : 0.0;
#pragma unroll
        for (int seg = 0; seg < NSEG; ++seg) {
            double t = __dmul_rn(__dmul_rn(v[seg], th[seg]), r[seg]);
            if (HAZARD) { if (t != t) t = 0.0; }          // replace(nan, 0)  (src/em.cpp:49); absent entries are 0 * th * 0
            if (c[seg] == 0xffffffffu) t = 0.0;
            stage[seg * kStageStride + lane] = t;
        }
        __syncwarp();
        if (lane < nseg && base < my_n) {
            const double *src = stage + lane * kStageStride;
            const int n = (int)min(32u, my_n - base);
            if (n == 32) {
#pragma unroll
                for (int q = 0; q < 32; ++q) acc = __dadd_rn(acc, src[q]);
            } else {
                for (int q = 0; q < n; ++q) acc = __dadd_rn(acc, src[q]);
            }
        }
        __syncwarp();
#pragma unroll
        for (int seg = 0; seg < NSEG; ++seg) { v[seg] = vn[seg]; c[seg] = cn[seg]; }
    }
    return acc;      // lane 2r: even-column stream of row r0 + r, lane 2r + 1: odd-column stream
}

template <bool HAZARD>
__device__ __forceinline__ double row_block_dispatch(int rb, const uint32_t *__restrict__ hp, const uint32_t *__restrict__ cidx,
                                                     const double *__restrict__ cval, int r0, int nr, int lane,
                                                     const double *theta, const double *ratio, double *stage)
{
    if (rb == 1) return row_block_streams<HAZARD, 2>(hp, cidx, cval, r0, nr, lane, theta, ratio, stage);
    return row_block_streams<HAZARD, 4>(hp, cidx, cval, r0, nr, lane, theta, ratio, stage);      // rb == 2
}

template <int T>
__global__ void __launch_bounds__(T)
em_giant_kernel(const GiantDesc *__restrict__ descs, int n_giant, const uint8_t *gbuf_c, uint8_t *gbuf,
                const int64_t *__restrict__ grp_row_ptr, int64_t total_rows, EmParams P, double *__restrict__ est,
                int32_t *__restrict__ iters, int32_t *__restrict__ status, unsigned int *gflags /* zeroed */,
                int n_teams, unsigned int *ticket /* zeroed */, int rb /* rows per warp block: 1 or 2 */)
{
    extern __shared__ __align__(16) uint8_t smem[];
    // The grid is split into n_teams teams of CTAs; a team iterates on one locus at a time with its
    // own barrier, so the barrier / latency-bound stretches of one locus overlap with the work of the
    // others.  Team t owns CTAs {t, t + n_teams, ...}; it starts on locus t and then takes whatever locus the ticket
    // hands out next.
    const int team = blockIdx.x % n_teams, team_rank = blockIdx.x / n_teams;
    const unsigned int n_ctas = (gridDim.x - team + n_teams - 1) / n_teams;      // CTAs of this team
    unsigned int *tflags = gflags + 64 * team;      // per team: [0..7] flags, [32..33] barrier (own cache line)
    unsigned int *bar = tflags + 32;
    double *theta = reinterpret_cast<double *>(smem);   // M, private copy per CTA
    const int tid = threadIdx.x, lane = tid & 31;
    const int64_t gtid = (int64_t)team_rank * T + tid, gthreads = (int64_t)n_ctas * T;   // team-relative
    // warp tasks are dealt round-robin over the CTAs (task w -> CTA w % n, warp w / n): a phase with
    // fewer tasks than warps then keeps every SM lightly and evenly loaded
    const int64_t gwarps = gthreads >> 5;
    const int64_t gwarp = (int64_t)(tid >> 5) * n_ctas + team_rank;
    (void)gbuf_c;

    for (int gi = team; gi < n_giant;) {
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
        const uint32_t *ccol = reinterpret_cast<const uint32_t *>(img + D.ccol);
        const uint32_t *cstart = reinterpret_cast<const uint32_t *>(img + D.cstart);
        const double *sval = reinterpret_cast<const double *>(img + D.sval);
        const uint32_t *sidx = reinterpret_cast<const uint32_t *>(img + D.sidx);
        const int nSliceC = D.nSliceC;
        double *ratio = reinterpret_cast<double *>(img + D.ratio);
        double *S = reinterpret_cast<double *>(img + D.S);          // [0, M): theta of the current update; then sums
        uint32_t *cntS = reinterpret_cast<uint32_t *>(img + D.cntS);
        // [0..1] hazard by update parity, [2] inf-column count, [3] non-finite estimate,
        // [4..5] "delta > conv" by update parity, [6..7] non-finite theta count by update parity
        unsigned int *hz = tflags;
        const int64_t row0 = grp_row_ptr[D.g];
        // per-warp staging of the products of 2 * rb segments behind theta[M+1]
        double *stage = reinterpret_cast<double *>(smem + align_up(8u * (uint32_t)(M + 1), 16) + giant_stage_bytes(rb) * (uint32_t)(tid >> 5));

        for (int i = tid; i < M; i += T) theta[i] = 1.0;   // src/em.cpp:20-21
        if (tid == 0) theta[M] = 0.0;                       // padding entries of the column SELL point here
        if (gtid == 0) { for (int q = 0; q < 8; ++q) if (q != 3) hz[q] = 0; }
        int nP = 0;                                          // non-finite thetas entering the update
        int iter = 0;
        bool over = (1.0 > P.conv);
        grid_barrier(bar, n_ctas);
        while (over && iter < P.maxiter - 1) {
            ++iter;
            const int par = iter & 1;
            // E-phase: c_j = sum_i a_ij*theta_i, ratio_j = Y_j / c_j   (src/em.cpp:46-48)
            // one lane per live column, a warp per slice of 32 length-sorted columns; slot-major storage
            // makes every load a coalesced 256 B / 128 B line
            for (int64_t w = gwarp; w < nSliceC; w += gwarps) {
                const uint32_t j = ccol[w * 32 + lane];
                const uint32_t s0 = cstart[w], s1 = cstart[w + 1];
                double a1 = 0.0, a2 = 0.0;
                int cnt = 0;
                for (uint32_t sb = s0; sb < s1; sb += 8) {            // 8 slots' loads in flight at a time
                    uint32_t ri[8];
                    double rv[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const bool ok = (sb + u < s1);
                        ri[u] = ok ? sidx[(size_t)(sb + u) * 32 + lane] : ((uint32_t)M << 1);
                        rv[u] = ok ? sval[(size_t)(sb + u) * 32 + lane] : 0.0;
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        if (sb + u < s1) {
                            const double th = theta[ri[u] >> 1];
                            const double t = __dmul_rn(rv[u], th);
                            if (nP) cnt += !is_finite(th);
                            if (ri[u] & 1u) a2 = __dadd_rn(a2, t); else a1 = __dadd_rn(a1, t);
                        }
                    }
                }
                if (j < (uint32_t)NL) {
                    if (cnt < nP) a1 = qnan();   // dense 0*NaN at this column's structural zeros (Appendix C.14)
                    const double r = __ddiv_rn(Yl[j], __dadd_rn(a1, a2));
                    ratio[j] = r;
                    if (!is_finite(r)) atomicOr(&hz[par], 1u);
                }
            }
            grid_barrier(bar, n_ctas);
            // M-phase: one warp per row, both parity streams in lockstep; the warp forms theta'_i and
            // the row's share of the stopping test  (src/em.cpp:48-63)
            const bool hazard = (__ldcg(&hz[par]) != 0) || (nP != 0);
            // every CTA is past the barrier, so nobody still reads the other parity's flags
            if (gtid == 0) { hz[par ^ 1] = 0; hz[4 + (par ^ 1)] = 0; hz[6 + (par ^ 1)] = 0; }
            const int64_t nBlocks = ((int64_t)M + rb - 1) / rb;
            for (int64_t blk = gwarp; blk < nBlocks; blk += gwarps) {
                const int r0 = (int)blk * rb, nr = min(rb, M - r0);
                const double acc = hazard ? row_block_dispatch<true>(rb, hp, cidx, cval, r0, nr, lane, theta, ratio, stage)
                                          : row_block_dispatch<false>(rb, hp, cidx, cval, r0, nr, lane, theta, ratio, stage);
                const double other = __shfl_down_sync(0xffffffffu, acc, 1);
                if (lane < 2 * nr && !(lane & 1)) {          // lane 2r forms theta' of row r0 + r and its share of the stopping test
                    const int i = r0 + (lane >> 1);
                    const double th = theta[i];
                    const double after = div_small_numerator(__dadd_rn(acc, other), rs[i]);     // == __ddiv_rn
                    int flag = (P.conv < 0.0) ? 1 : 0;
                    if (after > P.minvalue) {
                        const double diff = fabs(__dadd_rn(after, -th));
                        if (diff > __dmul_rn(after, P.conv_hi)) flag = 1;
                        else if (diff >= __dmul_rn(after, P.conv_lo) && __ddiv_rn(diff, after) > P.conv) flag = 1;
                    }
                    S[i] = after;
                    if (flag) atomicOr(&hz[4 + par], 1u);
                    if (!is_finite(after)) atomicAdd(&hz[6 + par], 1u);
                }
            }
            grid_barrier(bar, n_ctas);
            // every CTA refreshes its private theta; the flags are the same for all
            {
                const unsigned int f_over = __ldcg(&hz[4 + par]);
                for (int i0 = tid; i0 < M; i0 += 4 * T) {
                    double v[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) v[u] = (i0 + u * T < M) ? __ldcg(S + i0 + u * T) : 0.0;
#pragma unroll
                    for (int u = 0; u < 4; ++u) if (i0 + u * T < M) theta[i0 + u * T] = v[u];
                }
                over = f_over != 0;
            }
            nP = (int)__ldcg(&hz[6 + par]);
            __syncthreads();
        }
        grid_barrier(bar, n_ctas);   // S is reused for the epilogue sums: nobody may still be reading theta from it

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
        grid_barrier(bar, n_ctas);
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
        grid_barrier(bar, n_ctas);
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
        grid_barrier(bar, n_ctas);
        if (gtid == 0) {
            iters[D.g] = iter;
            status[D.g] = ((iter >= P.maxiter - 1 && over) ? 1 : 0) | (__ldcg(&hz[3]) ? 2 : 0);
            hz[3] = 0;
            tflags[40] = (unsigned int)n_teams + atomicAdd(ticket, 1u);      // this team's next locus
        }
        grid_barrier(bar, n_ctas);
        gi = (int)__ldcg(&tflags[40]);
    }
}

}  // namespace bambu
