// pack_kernel.cuh -- raw arena slice of one gene group -> its image (common.cuh).
//
// One CTA per group.  This is the device-side replacement of what run_parallel
// does in R before every emWithL1 call (three dcast()s into dense matrices,
// R/bambu-quantify_utilityFunctions.R:407-409,427-431): instead of densifying,
// it drops what cannot influence the result (columns with Y == 0, entries with
// aval == 0), keeps the position parities the summation order depends on, and
// lays the survivors out twice (by row for the M-step, by column for the E-step).
// It also evaluates rs_i = sum_j X_ij (src/em.cpp:33) over ALL entries and the
// closed form for single-transcript groups (R/...utilityFunctions.R:410-414).
#pragma once
#include "common.cuh"
#ifndef QD_FN
#define QD_FN __host__ __device__ __forceinline__
#endif
#include "x87_sum.h"

namespace bambu {

// Accumulator of the single-transcript closed form: R's sum() adds doubles into an x87 long double
// (SURVEY Appendix C.16).  Non-negative finite terms -- everything a real table produces -- go through the
// bit-exact integer emulation of that addition (x87_sum.h); a negative or non-finite term switches the sum
// to an error-free two-sum in doubles (exact up to the final rounding, NaN / Inf propagate like in R).
struct LongSum {
    qd::Ext x = {0, 0};
    double hi = 0.0, lo = 0.0;
    bool plain = false;
    __device__ __forceinline__ void add_dd(double v)
    {
        double s = __dadd_rn(hi, v);
        double bb = __dadd_rn(s, -hi);
        double err = __dadd_rn(__dadd_rn(hi, -__dadd_rn(s, -bb)), __dadd_rn(v, -bb));
        hi = s;
        if (is_finite(err)) lo = __dadd_rn(lo, err);
    }
    __device__ __forceinline__ void add(double v)
    {
        if (!plain && !(v >= 0.0 && is_finite(v))) {     // leave the exact path: replay what was summed so far
            bool oor = false;
            const double sofar = qd::ext_to_double(x, &oor);
            plain = true;
            hi = sofar; lo = 0.0;
        }
        if (plain) add_dd(v);
        else x = qd::ext_add(x, qd::ext_from_double(v));
    }
    __device__ __forceinline__ double value() const
    {
        if (plain) return is_finite(hi) ? __dadd_rn(hi, lo) : hi;
        bool oor = false;
        const double d = qd::ext_to_double(x, &oor);
        if (!oor) return d;
        // subnormal or overflowing total: the double sum of the (few) terms is the best available
        return (x.e > 0) ? __longlong_as_double(0x7ff0000000000000ll) : 0.0;
    }
};

// ---- D + E of both pack kernels: the column-major copy of a group's compacted rows -----------------
// rowcnt = exclusive row pointers (M+1), colcnt = exclusive column pointers (NL+1), cursor = scratch (NL).
template <int T>
__device__ __forceinline__ void pack_columns(int M, int NL, const uint32_t *rowcnt, const uint32_t *colcnt,
                                             uint32_t *cursor, const uint16_t *img_cidx, const double *img_cval,
                                             uint16_t *img_ridx, double *img_rval)
{
    const int tid = threadIdx.x;
    // D: scatter into columns (arbitrary order inside a column)
    for (int i = tid; i < M; i += T) {
        const uint32_t lo = rowcnt[i], hi = rowcnt[i + 1];
        for (uint32_t p = lo; p < hi; ++p) {
            const uint32_t lj = img_cidx[p] >> 1;
            const uint32_t q = atomicAdd(&cursor[lj], 1u);
            img_ridx[q] = (uint16_t)(((uint32_t)i << 1) | (uint32_t)(i & 1));
            img_rval[q] = img_cval[p];
        }
    }
    __syncthreads();
    // E: rows ascending inside every column (the order arma::sum adds them)
    for (int j = tid; j < NL; j += T) sort_column(img_ridx, img_rval, colcnt[j], colcnt[j + 1]);
}

template <int T>
__global__ void __launch_bounds__(T)
pack_resident_kernel(ArenaDev A, const int32_t *__restrict__ work, int n_work,
                     const int64_t *__restrict__ img_off, uint8_t *__restrict__ img_base,
                     PackInfo *__restrict__ info, int64_t *__restrict__ nnz_nonzero,
                     double *__restrict__ est, int32_t *__restrict__ status, int *__restrict__ err_flag,
                     int closed_form_single_tx)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    __shared__ uint32_t s_warp[34];
    __shared__ uint32_t s_nz;

    if ((int)blockIdx.x >= n_work) return;
    const int tid = threadIdx.x;
    const int g = work[blockIdx.x];
    const int64_t row0 = A.grp_row_ptr[g], col0 = A.grp_col_ptr[g];
    const int M = (int)(A.grp_row_ptr[g + 1] - row0);
    const int N = (int)(A.grp_col_ptr[g + 1] - col0);
    const int64_t *rnp = A.row_nnz_ptr + row0;

    if (tid == 0) s_nz = 0;

    // ---- single-transcript group: closed form, no EM ---------------------------
    if (M == 1 && closed_form_single_tx) {
        if (tid == 0) {
            LongSum s0, s1, s2;
            uint32_t nz = 0;
            for (int64_t k = rnp[0]; k < rnp[1]; ++k) {
                const int j = A.col_idx[k];
                if (j < 0 || j >= N) { atomicOr(err_flag, kErrColRange); break; }
                const double a = A.aval[k];
                nz += (a != 0.0);
                const double ky = __dmul_rn(A.K[col0 + j], A.Y[col0 + j]);
                const double af = (A.nz_flags[k] & 1) ? a : __dmul_rn(a, 0.0);
                const double au = (A.col_flags[col0 + j] & 1) ? __dmul_rn(a, 0.0) : a;
                s0.add(__dmul_rn(ky, a));
                s1.add(__dmul_rn(ky, af));
                s2.add(__dmul_rn(ky, au));
            }
            const double e0 = s0.value(), e1 = s1.value(), e2 = s2.value();
            est[row0] = e0;
            est[A.total_rows + row0] = e1;
            est[2 * A.total_rows + row0] = e2;
            if (!(is_finite(e0) && is_finite(e1) && is_finite(e2))) status[g] = 2;
            nnz_nonzero[g] = nz;
            info[g].NL = 0;
            info[g].nnzL = 0;
        }
        return;
    }

    uint32_t *colmap = reinterpret_cast<uint32_t *>(smem_raw);  // N    : live ? (live_idx<<1|1) : 0
    uint32_t *rowcnt = colmap + N;                              // M+1  : -> rp
    uint32_t *colcnt = rowcnt + (M + 1);                        // N+1  : -> cp
    uint32_t *cursor = colcnt + (N + 1);                        // N

    // ---- A: live columns --------------------------------------------------------
    for (int j = tid; j < N; j += T) colmap[j] = (A.Y[col0 + j] != 0.0) ? 1u : 0u;  // NaN counts as live
    __syncthreads();
    const uint32_t NL = block_excl_scan<T>(colmap, N, s_warp);
    // after the scan colmap[j] = live index if j is live; recover liveness from Y again
    for (int j = tid; j < N; j += T) {
        const bool live = (A.Y[col0 + j] != 0.0);
        colmap[j] = live ? ((colmap[j] << 1) | 1u) : 0u;
    }
    __syncthreads();

    // ---- B: rs_i over all entries, live entries per row -----------------------
    uint32_t my_nz = 0;
    double *img_rs;  // set once nnzL is known; rs is the first section so its offset is 0
    uint8_t *img = img_base + img_off[g];
    img_rs = reinterpret_cast<double *>(img);
    for (int i = tid; i < M; i += T) {
        double a1 = 0.0, a2 = 0.0;
        uint32_t cnt = 0;
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
            cnt += (nzv && (colmap[j] & 1u));
        }
        img_rs[i] = __dadd_rn(a1, a2);
        rowcnt[i] = cnt;
    }
    if (my_nz) atomicAdd(&s_nz, my_nz);
    for (int j = tid; j <= N; j += T) colcnt[j] = 0;
    __syncthreads();
    const uint32_t nnzL = block_excl_scan<T>(rowcnt, M, s_warp);
    if (tid == 0) rowcnt[M] = nnzL;
    __syncthreads();

    const ImgLayout L = img_layout((uint32_t)M, NL, nnzL);
    double *img_Y = reinterpret_cast<double *>(img + L.Yl);
    double *img_cval = reinterpret_cast<double *>(img + L.cval);
    double *img_rval = reinterpret_cast<double *>(img + L.rval);
    uint16_t *img_rp = reinterpret_cast<uint16_t *>(img + L.rp);
    uint16_t *img_cp = reinterpret_cast<uint16_t *>(img + L.cp);
    uint16_t *img_cidx = reinterpret_cast<uint16_t *>(img + L.cidx);
    uint16_t *img_ridx = reinterpret_cast<uint16_t *>(img + L.ridx);
    double *img_KY = reinterpret_cast<double *>(img + L.KY);
    uint8_t *img_eq = img + L.eq;
    uint8_t *img_cm = img + L.cm;

    // per live column: Y, K*Y (src/em.cpp:98 "K % Y"), multi_align
    for (int j = tid; j < N; j += T) {
        const uint32_t c = colmap[j];
        if (c & 1u) {
            const uint32_t lj = c >> 1;
            const double y = A.Y[col0 + j];
            img_Y[lj] = y;
            img_KY[lj] = __dmul_rn(A.K[col0 + j], y);
            img_cm[lj] = A.col_flags[col0 + j] & 1;
        }
    }

    // ---- C: compacted rows, column histogram -----------------------------------
    for (int i = tid; i < M; i += T) {
        uint32_t pos = rowcnt[i];
        for (int64_t k = rnp[i]; k < rnp[i + 1]; ++k) {
            const int j = A.col_idx[k];
            if (j < 0 || j >= N) continue;
            const double a = A.aval[k];
            const uint32_t c = colmap[j];
            if (a != 0.0 && (c & 1u)) {
                const uint32_t lj = c >> 1;
                img_cidx[pos] = (uint16_t)((lj << 1) | (uint32_t)(j & 1));
                img_cval[pos] = a;
                img_eq[pos] = A.nz_flags[k] & 1;
                atomicAdd(&colcnt[lj], 1u);
                ++pos;
            }
        }
        img_rp[i] = (uint16_t)rowcnt[i];
    }
    if (tid == 0) img_rp[M] = (uint16_t)nnzL;
    __syncthreads();
    block_excl_scan<T>(colcnt, (int)NL, s_warp);
    if (tid == 0) colcnt[NL] = nnzL;
    __syncthreads();
    for (int j = tid; j <= (int)NL; j += T) {
        img_cp[j] = (uint16_t)colcnt[j];
        if (j < (int)NL) cursor[j] = colcnt[j];
    }
    __syncthreads();

    pack_columns<T>(M, (int)NL, rowcnt, colcnt, cursor, img_cidx, img_cval, img_ridx, img_rval);
    if (tid == 0) {
        info[g].NL = (int32_t)NL;
        info[g].nnzL = (int32_t)nnzL;
        nnz_nonzero[g] = s_nz;
    }
}

// ---- the same from the wire form v2 (bambu_em_batched_v2) -----------------------------------------------
// The packer on the host already dropped the dead columns and the zero entries and computed rs, so what is
// left is: form Y = nobs / K and K*Y per live column, widen the bit-packed flags, copy the rows (checking what
// the EM relies on: indices in range and strictly increasing, one parity per column) and build the
// column-major copy.  Decoding happens here, in the kernel that builds the image -- there is no extra pass.
template <int T>
__global__ void __launch_bounds__(T)
pack_resident_wire_kernel(WireDev W, const int32_t *__restrict__ work, int n_work,
                          const int64_t *__restrict__ img_off, uint8_t *__restrict__ img_base,
                          PackInfo *__restrict__ info, int64_t *__restrict__ nnz_nonzero,
                          double *__restrict__ est, int32_t *__restrict__ status, int *__restrict__ err_flag,
                          int closed_form_single_tx)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    __shared__ uint32_t s_warp[34];

    if ((int)blockIdx.x >= n_work) return;
    const int tid = threadIdx.x;
    const int g = work[blockIdx.x];
    const int64_t row0 = W.grp_row_ptr[g], col0 = W.grp_col_ptr[g], ent0 = W.grp_ent_ptr[g];
    const int M = (int)(W.grp_row_ptr[g + 1] - row0);
    const int NL = (int)(W.grp_col_ptr[g + 1] - col0);
    const uint32_t nnzL0 = (uint32_t)(W.grp_ent_ptr[g + 1] - ent0);
    const uint32_t *rend = W.row_end + row0;
    const uint16_t *cidx16 = reinterpret_cast<const uint16_t *>(W.cidx + W.grp_idx_off[g]);
    const uint32_t *cidx32 = reinterpret_cast<const uint32_t *>(W.cidx + W.grp_idx_off[g]);
    const bool wide = wire_idx_width(NL) == 4;
    const double *av = W.aval + ent0;

    if (M == 1 && closed_form_single_tx) {      // R/bambu-quantify_utilityFunctions.R:410-414
        if (tid == 0) {
            LongSum s0, s1, s2;
            if (rend[0] != nnzL0) atomicOr(err_flag, kErrRowPtr);
            for (uint32_t k = 0; k < nnzL0; ++k) {
                const uint32_t c = wide ? cidx32[k] : (uint32_t)cidx16[k];
                const uint32_t lj = c >> 1;
                if (lj >= (uint32_t)NL) { atomicOr(err_flag, kErrColRange); break; }
                const double a = av[k];
                double y, kk;
                wire_col(W, col0 + lj, y, kk);
                const double ky = __dmul_rn(kk, y);
                const double af = wire_bit(W.eq_bits, ent0 + k) ? a : __dmul_rn(a, 0.0);
                const double au = wire_bit(W.multi_bits, col0 + lj) ? __dmul_rn(a, 0.0) : a;
                s0.add(__dmul_rn(ky, a));
                s1.add(__dmul_rn(ky, af));
                s2.add(__dmul_rn(ky, au));
            }
            const double e0 = s0.value(), e1 = s1.value(), e2 = s2.value();
            est[row0] = e0;
            est[W.total_rows + row0] = e1;
            est[2 * W.total_rows + row0] = e2;
            if (!(is_finite(e0) && is_finite(e1) && is_finite(e2))) status[g] = 2;
            nnz_nonzero[g] = nnzL0;
            info[g].NL = 0;
            info[g].nnzL = 0;
        }
        return;
    }

    uint32_t *parity = reinterpret_cast<uint32_t *>(smem_raw);  // NL   : bit0 / bit1 = parity 0 / 1 seen
    uint32_t *rowcnt = parity + NL;                             // M+1  : row pointers
    uint32_t *colcnt = rowcnt + (M + 1);                        // NL+1 : -> cp
    uint32_t *cursor = colcnt + (NL + 1);                       // NL
    __shared__ int s_bad;

    // ---- pass 1: validate what the EM relies on, column histogram ------------------------------------
    if (tid == 0) s_bad = 0;
    for (int j = tid; j < NL; j += T) { parity[j] = 0; colcnt[j] = 0; }
    if (tid == 0) colcnt[NL] = 0;
    __syncthreads();
    if (tid == 0 && M > 0 && rend[M - 1] != nnzL0) { atomicOr(err_flag, kErrRowPtr); s_bad = 1; }
    for (int i = tid; i < M; i += T) {
        const uint32_t lo = i ? rend[i - 1] : 0u, hi = rend[i];
        if (hi < lo || hi > nnzL0) { atomicOr(err_flag, kErrRowPtr); s_bad = 1; continue; }
        int prev = -1;
        for (uint32_t k = lo; k < hi; ++k) {
            const uint32_t c = wide ? cidx32[k] : (uint32_t)cidx16[k];
            const uint32_t lj = c >> 1;
            if (lj >= (uint32_t)NL) { atomicOr(err_flag, kErrColRange); s_bad = 1; break; }
            if ((int)lj <= prev) { atomicOr(err_flag, kErrColOrder); s_bad = 1; }
            prev = (int)lj;
            atomicAdd(&colcnt[lj], 1u);
            atomicOr(&parity[lj], 1u << (c & 1u));
        }
    }
    __syncthreads();
    for (int j = tid; j < NL; j += T) if (parity[j] == 3u) { atomicOr(err_flag, kErrParity); }
    // a malformed group is packed as an empty one: the run is reported as failed, memory stays consistent
    const bool bad = (s_bad != 0);
    const uint32_t nnzL = bad ? 0u : nnzL0;
    __syncthreads();
    if (bad) { for (int j = tid; j <= NL; j += T) colcnt[j] = 0; }

    uint8_t *img = img_base + img_off[g];
    const ImgLayout L = img_layout((uint32_t)M, (uint32_t)NL, nnzL);
    double *img_rs = reinterpret_cast<double *>(img + L.rs);
    double *img_Y = reinterpret_cast<double *>(img + L.Yl);
    double *img_cval = reinterpret_cast<double *>(img + L.cval);
    double *img_rval = reinterpret_cast<double *>(img + L.rval);
    uint16_t *img_rp = reinterpret_cast<uint16_t *>(img + L.rp);
    uint16_t *img_cp = reinterpret_cast<uint16_t *>(img + L.cp);
    uint16_t *img_cidx = reinterpret_cast<uint16_t *>(img + L.cidx);
    uint16_t *img_ridx = reinterpret_cast<uint16_t *>(img + L.ridx);
    double *img_KY = reinterpret_cast<double *>(img + L.KY);
    uint8_t *img_eq = img + L.eq;
    uint8_t *img_cm = img + L.cm;

    // ---- pass 2: the image.  Per live column: Y, K*Y (src/em.cpp:98 "K % Y"), multi_align ------------
    for (int j = tid; j < NL; j += T) {
        double y, kk;
        wire_col(W, col0 + j, y, kk);
        img_Y[j] = y;
        img_KY[j] = __dmul_rn(kk, y);
        img_cm[j] = (uint8_t)wire_bit(W.multi_bits, col0 + j);
    }
    for (int i = tid; i < M; i += T) {
        const uint32_t lo = bad ? 0u : (i ? rend[i - 1] : 0u), hi = bad ? 0u : rend[i];
        rowcnt[i] = lo;
        img_rs[i] = W.rs[row0 + i];
        img_rp[i] = (uint16_t)lo;
        for (uint32_t k = lo; k < hi; ++k) {
            img_cidx[k] = wide ? (uint16_t)cidx32[k] : cidx16[k];
            img_cval[k] = av[k];
            img_eq[k] = (uint8_t)wire_bit(W.eq_bits, ent0 + k);
        }
    }
    if (tid == 0) {
        rowcnt[M] = nnzL;
        img_rp[M] = (uint16_t)nnzL;
    }
    __syncthreads();
    block_excl_scan<T>(colcnt, NL, s_warp);
    if (tid == 0) colcnt[NL] = nnzL;
    __syncthreads();
    for (int j = tid; j <= NL; j += T) {
        img_cp[j] = (uint16_t)colcnt[j];
        if (j < NL) cursor[j] = colcnt[j];
    }
    __syncthreads();
    pack_columns<T>(M, NL, rowcnt, colcnt, cursor, img_cidx, img_cval, img_ridx, img_rval);
    if (tid == 0) {
        info[g].NL = (int32_t)NL;
        info[g].nnzL = (int32_t)nnzL;
        nnz_nonzero[g] = nnzL;
    }
}

}  // namespace bambu
