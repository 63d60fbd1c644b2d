// em_resident.cuh -- the EM of one gene group, resident in shared memory.
//
// Restates, on the group's image (common.cuh), exactly the arithmetic of
//   em_theta   src/em.cpp:9-71     (fixed point on theta, group-level stopping rule)
//   emWithL1   src/em.cpp:77-110   (baseSum and the three weighted row sums)
// with every reduction adding its terms in increasing dense-index order into
// Armadillo's two interleaved accumulators (even position -> acc1, odd -> acc2,
// result acc1 + acc2), so results and iteration counts are those of the dense
// reference.  One thread owns one column in the E-phase and one row in the
// M-phase; a CTA owns one group at a time and pulls groups from its bucket's
// queue (largest first) until the queue is empty.
//
// FP64 throughout; products and sums use the non-contracting intrinsics
// (__dmul_rn/__dadd_rn): the reference rounds a_ij*theta_i before it is summed
// (src/em.cpp:46-47), so no FMA may be formed.
#pragma once
#include "common.cuh"

namespace bambu {

struct EmParams {
    int32_t maxiter;
    double minvalue, conv;
    double conv_lo, conv_hi;  // conv*(1 -+ 2^-40): division-free screen of |d|/theta > conv
};

template <int T>
__global__ void __launch_bounds__(T)
em_resident_kernel(const int32_t *__restrict__ work, int n_work, int *__restrict__ ticket,
                   const int64_t *__restrict__ img_off, const uint8_t *__restrict__ img_base,
                   const PackInfo *__restrict__ info, const int64_t *__restrict__ grp_row_ptr,
                   int64_t total_rows, EmParams P, double *__restrict__ est,
                   int32_t *__restrict__ iters, int32_t *__restrict__ status)
{
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ int s_next;
    __shared__ int s_nf[2];    // rows whose theta is non-finite, written by M-phase it into [it&1]
    __shared__ int s_rbad[2];  // a ratio Y/c of E-phase it is non-finite, [it&1]
    __shared__ int s_Q;        // columns whose baseSum is +-inf (epilogue)

    const int tid = threadIdx.x;
    for (;;) {
        if (tid == 0) s_next = atomicAdd(ticket, 1);
        __syncthreads();
        const int w = s_next;
        if (w >= n_work) break;
        const int g = work[w];
        const int64_t row0 = grp_row_ptr[g];
        const int M = (int)(grp_row_ptr[g + 1] - row0);
        const int NL = info[g].NL, nnzL = info[g].nnzL;
        const ImgLayout L = img_layout((uint32_t)M, (uint32_t)NL, (uint32_t)nnzL);
        const uint8_t *gimg = img_base + img_off[g];

        // ---- image -> shared memory ------------------------------------------
        {
            const uint4 *src = reinterpret_cast<const uint4 *>(gimg);
            uint4 *dst = reinterpret_cast<uint4 *>(smem);
            const int n16 = (int)(L.copy_bytes >> 4);
            for (int i = tid; i < n16; i += T) dst[i] = __ldg(src + i);
        }
        const double *rs = reinterpret_cast<const double *>(smem + L.rs);
        const double *Yl = reinterpret_cast<const double *>(smem + L.Yl);
        const double *cval = reinterpret_cast<const double *>(smem + L.cval);
        const double *rval = reinterpret_cast<const double *>(smem + L.rval);
        const uint16_t *rp = reinterpret_cast<const uint16_t *>(smem + L.rp);
        const uint16_t *cp = reinterpret_cast<const uint16_t *>(smem + L.cp);
        const uint16_t *cidx = reinterpret_cast<const uint16_t *>(smem + L.cidx);
        const uint16_t *ridx = reinterpret_cast<const uint16_t *>(smem + L.ridx);
        double *theta = reinterpret_cast<double *>(smem + L.copy_bytes);
        double *ratio = theta + M;
        for (int i = tid; i < M; i += T) theta[i] = 1.0;  // src/em.cpp:20-21
        if (tid == 0) { s_nf[0] = s_nf[1] = 0; s_rbad[0] = s_rbad[1] = 0; s_Q = 0; }
        __syncthreads();

        // ---- em_theta: while (delta > conv && iter < maxiter-1)  src/em.cpp:42 ----
        int iter = 0;
        bool over = (1.0 > P.conv);   // delta starts at 1 (src/em.cpp:38)
        while (over && iter < P.maxiter - 1) {
            ++iter;
            const int par = iter & 1;
            const int nP = s_nf[par ^ 1];          // non-finite thetas entering this update
            // E-phase: c_j = sum_i a_ij*theta_i (:46-47), ratio_j = Y_j / c_j (:48)
            if (tid == 0) s_nf[par] = 0;
            for (int j = tid; j < NL; j += T) {
                const int lo = cp[j], hi = cp[j + 1];
                double a1 = 0.0, a2 = 0.0;
                if (nP == 0) {
                    for (int k = lo; k < hi; ++k) {
                        const uint32_t ri = ridx[k];
                        const double t = __dmul_rn(rval[k], theta[ri >> 1]);
                        if (ri & 1u) a2 = __dadd_rn(a2, t); else a1 = __dadd_rn(a1, t);
                    }
                } else {
                    // dense semantics: a row with non-finite theta poisons every column it
                    // has a structural zero in (0*NaN), SURVEY Appendix C.14
                    int cnt = 0;
                    for (int k = lo; k < hi; ++k) {
                        const uint32_t ri = ridx[k];
                        const double th = theta[ri >> 1];
                        const double t = __dmul_rn(rval[k], th);
                        cnt += !is_finite(th);
                        if (ri & 1u) a2 = __dadd_rn(a2, t); else a1 = __dadd_rn(a1, t);
                    }
                    if (cnt < nP) a1 = __longlong_as_double(0x7ff8000000000000ll);
                }
                const double r = __ddiv_rn(Yl[j], __dadd_rn(a1, a2));
                ratio[j] = r;
                if (!is_finite(r)) s_rbad[par] = 1;
            }
            __syncthreads();
            // M-phase: theta_i = sum_j (a_ij*theta_i)*ratio_j / rs_i, NaN terms -> 0 (:48-50)
            const bool hazard = (s_rbad[par] != 0) || (nP != 0);
            if (tid == 0) s_rbad[par ^ 1] = 0;
            int flag = (P.conv < 0.0) ? 1 : 0;   // delta >= 0 > conv whenever conv is negative
            for (int i = tid; i < M; i += T) {
                const double th = theta[i];
                const int lo = rp[i], hi = rp[i + 1];
                double a1 = 0.0, a2 = 0.0;
                if (!hazard) {
                    for (int k = lo; k < hi; ++k) {
                        const uint32_t ci = cidx[k];
                        const double t = __dmul_rn(__dmul_rn(cval[k], th), ratio[ci >> 1]);
                        if (ci & 1u) a2 = __dadd_rn(a2, t); else a1 = __dadd_rn(a1, t);
                    }
                } else {
                    for (int k = lo; k < hi; ++k) {
                        const uint32_t ci = cidx[k];
                        double t = __dmul_rn(__dmul_rn(cval[k], th), ratio[ci >> 1]);
                        if (t != t) t = 0.0;
                        if (ci & 1u) a2 = __dadd_rn(a2, t); else a1 = __dadd_rn(a1, t);
                    }
                }
                const double after = __ddiv_rn(__dadd_rn(a1, a2), rs[i]);
                theta[i] = after;
                if (!is_finite(after)) atomicAdd(&s_nf[par], 1);
                // delta = max over {after > minvalue} of |after-before|/after  (:54-63);
                // only "delta > conv" is needed: screen without the division, divide
                // only inside the 2^-40 band around conv
                if (after > P.minvalue) {
                    const double diff = fabs(__dadd_rn(after, -th));
                    if (diff > __dmul_rn(after, P.conv_hi)) flag = 1;
                    else if (diff >= __dmul_rn(after, P.conv_lo)) {
                        if (__ddiv_rn(diff, after) > P.conv) flag = 1;
                    }
                }
            }
            over = __syncthreads_or(flag) != 0;
        }

        // ---- emWithL1 epilogue: src/em.cpp:97-102 ---------------------------------
        const int nP = (iter == 0) ? 0 : s_nf[iter & 1];
        const double *KY = reinterpret_cast<const double *>(gimg + L.KY);
        const uint8_t *eq = gimg + L.eq;
        const uint8_t *cm = gimg + L.cm;
        for (int j = tid; j < NL; j += T) {
            const int lo = cp[j], hi = cp[j + 1];
            double a1 = 0.0, a2 = 0.0;
            int cnt = 0;
            for (int k = lo; k < hi; ++k) {
                const uint32_t ri = ridx[k];
                const double th = theta[ri >> 1];
                const double t = __dmul_rn(rval[k], th);
                cnt += !is_finite(th);
                if (ri & 1u) a2 = __dadd_rn(a2, t); else a1 = __dadd_rn(a1, t);
            }
            double c = __dadd_rn(a1, a2);
            if (cnt < nP) c = __longlong_as_double(0x7ff8000000000000ll);
            double b = __ddiv_rn(__ldg(KY + j), c);   // (K % Y) / colsum  (:98)
            if (b != b) b = 0.0;                      // replace(nan, 0)   (:99)
            ratio[j] = b;
            if (fabs(b) > 1.7976931348623157e308) atomicAdd(&s_Q, 1);
        }
        __syncthreads();
        const int Q = s_Q;
        int bad = 0;
        for (int i = tid; i < M; i += T) {
            const double th = theta[i];
            const int lo = rp[i], hi = rp[i + 1];
            double t1 = 0, t2 = 0, f1 = 0, f2 = 0, u1 = 0, u2 = 0;
            int cnt = 0;
            for (int k = lo; k < hi; ++k) {
                const uint32_t ci = cidx[k];
                const uint32_t col = ci >> 1;
                const double b = ratio[col];
                const double a = cval[k];
                const double af = __ldg(eq + k) ? a : __dmul_rn(a, 0.0);       // fullAval   = aval*equal
                const double au = __ldg(cm + col) ? __dmul_rn(a, 0.0) : a;     // uniqueAval = aval*!multi
                const double t = __dmul_rn(__dmul_rn(a, th), b);
                const double f = __dmul_rn(__dmul_rn(af, th), b);
                const double u = __dmul_rn(__dmul_rn(au, th), b);
                cnt += (fabs(b) > 1.7976931348623157e308);
                if (ci & 1u) { t2 = __dadd_rn(t2, t); f2 = __dadd_rn(f2, f); u2 = __dadd_rn(u2, u); }
                else         { t1 = __dadd_rn(t1, t); f1 = __dadd_rn(f1, f); u1 = __dadd_rn(u1, u); }
            }
            double e0 = __dadd_rn(t1, t2), e1 = __dadd_rn(f1, f2), e2 = __dadd_rn(u1, u2);
            // dense 0*NaN / 0*Inf at the structural zeros of this row
            if (!is_finite(th) || cnt < Q) e0 = e1 = e2 = __longlong_as_double(0x7ff8000000000000ll);
            est[row0 + i] = e0;
            est[total_rows + row0 + i] = e1;
            est[2 * total_rows + row0 + i] = e2;
            bad |= !(is_finite(e0) && is_finite(e1) && is_finite(e2));
        }
        const int anybad = __syncthreads_or(bad);
        if (tid == 0) {
            iters[g] = iter;
            status[g] = ((iter >= P.maxiter - 1 && over) ? 1 : 0) | (anybad ? 2 : 0);
        }
        // the barrier above also protects smem against the next group's load
    }
}

}  // namespace bambu
