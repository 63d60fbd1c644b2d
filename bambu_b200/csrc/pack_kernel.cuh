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

namespace bambu {

// error-free accumulation (two-sum) used for the single-transcript closed form:
// R's sum() accumulates in long double there.
struct DD {
    double hi = 0.0, lo = 0.0;
    __device__ __forceinline__ void add(double x)
    {
        double s = __dadd_rn(hi, x);
        double bb = __dadd_rn(s, -hi);
        double err = __dadd_rn(__dadd_rn(hi, -__dadd_rn(s, -bb)), __dadd_rn(x, -bb));
        hi = s;
        if (is_finite(err)) lo = __dadd_rn(lo, err);
    }
    __device__ __forceinline__ double value() const { return is_finite(hi) ? __dadd_rn(hi, lo) : hi; }
};

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
            DD s0, s1, s2;
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

    // ---- D: scatter into columns (arbitrary order inside a column) ------------
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

    // ---- E: rows ascending inside every column (the order arma::sum adds them) --
    for (int j = tid; j < (int)NL; j += T) sort_column(img_ridx, img_rval, colcnt[j], colcnt[j + 1]);
    if (tid == 0) {
        info[g].NL = (int32_t)NL;
        info[g].nnzL = (int32_t)nnzL;
        nnz_nonzero[g] = s_nz;
    }
}

}  // namespace bambu
