// em_sell.cuh -- the EM of one gene group on its SELL image, resident in shared memory.
//
// Restates exactly the arithmetic of
//   em_theta   src/em.cpp:9-71     (fixed point on theta, group-level stopping rule)
//   emWithL1   src/em.cpp:77-110   (baseSum and the three weighted row sums)
// Every reduction adds its terms in increasing dense-index order into Armadillo's two
// interleaved accumulators (even position -> acc1, odd -> acc2, result acc1 + acc2):
// the SELL image keeps the two streams apart, so a lane runs two independent add
// chains with no parity test, and a warp's loop trip count is uniform.  One lane owns
// one live column in the E-phase and one row in the M-phase; a warp owns slices of 32;
// a CTA owns one group at a time and pulls groups from its class's queue.
//
// Shared memory: [flag words 16 B][theta f64 M+1][ratio f64 NL+1][image].  Both vectors and the per-line constants
// (rs, Y) are stored BY POSITION (32*slice + lane), so a lane's own line needs no index lookup, and
// the image stores every gathered index as the byte offset of the element in that vector region, so
// a gather needs no address arithmetic.  theta[M] and ratio[NL] are always 0.0: padding entries
// (value 0.0) point there and add an exact +0.0.
//
// FP64 throughout; products and sums use the non-contracting intrinsics: the reference
// rounds a_ij*theta_i before it is summed (src/em.cpp:46-47), so no FMA may be formed.
#pragma once
#include "common.cuh"

#ifndef BAMBU_V_UNROLL
#define BAMBU_V_UNROLL 1
#endif

namespace bambu {

constexpr int kSlotUnroll = BAMBU_V_UNROLL;   // slots of a line in flight per lane

struct EmParams {
    int32_t maxiter;
    double minvalue, conv;
    double conv_lo, conv_hi;  // conv*(1 -+ 2^-40): division-free screen of |d|/theta > conv
};

__device__ __forceinline__ double qnan() { return __longlong_as_double(0x7ff8000000000000ll); }

// sum over one line of a SELL side: a1 = even stream, a2 = odd stream.
// MODE 0: plain; MODE 1: also count gathered elements that are not finite.
// ---- TMA 1-D bulk copy (cp.async.bulk, SASS UBLKCP) + mbarrier: the group image is pulled into
// shared memory by the copy engine of the SM while the threads initialise theta ------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t done = 0;
    while (!done) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    }
}

__device__ __forceinline__ double vec_at(const uint8_t *vec, uint32_t off)
{
    return *reinterpret_cast<const double *>(vec + off);
}

template <int MODE, int U = kSlotUnroll>
__device__ __forceinline__ void line_sums(const uint8_t *__restrict__ side, uint2 desc, int lane,
                                          const uint8_t *__restrict__ vec, double &a1, double &a2, int &cnt)
{
    const uint8_t *p = side + desc.x;        // desc = {byte offset of the slice's first slot, slots}
#pragma unroll U
    for (uint32_t s = 0; s < desc.y; ++s, p += kSlotBytes) {
        const double2 v = reinterpret_cast<const double2 *>(p)[lane];
        const uint32_t oe = reinterpret_cast<const uint16_t *>(p + 512)[lane];
        const uint32_t oo = reinterpret_cast<const uint16_t *>(p + 576)[lane];
        const double xe = vec_at(vec, oe), xo = vec_at(vec, oo);
        if (MODE == 1) cnt += (!is_finite(xe)) + (!is_finite(xo));
        a1 = __dadd_rn(a1, __dmul_rn(v.x, xe));
        a2 = __dadd_rn(a2, __dmul_rn(v.y, xo));
    }
}

#ifndef BAMBU_V_LAZY
#define BAMBU_V_LAZY 1
#endif

// T threads per CTA; U = slots of a line in flight per lane.  The throughput instances (U = 1: 64 / 128 threads for the
// small classes) keep as many groups resident per SM as registers and shared memory allow; the latency instance
// (256 threads, U = 4: one slice per warp, the loads of four slots in flight) gives ONE group a short update and is
// what a small arena -- a single sample -- runs on, where the longest-running groups bound the run time.
template <int T, int U = kSlotUnroll>
__global__ void __launch_bounds__(T)
em_sell_kernel(const int32_t *__restrict__ queue, const int *__restrict__ qcount, int *__restrict__ ticket,
               const GroupImg *__restrict__ ginfo, const uint8_t *__restrict__ pool,
               const int64_t *__restrict__ grp_row_ptr, int64_t total_rows, EmParams P,
               double *__restrict__ est, int32_t *__restrict__ iters, int32_t *__restrict__ status)
{
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ int s_next;
    // The per-update flag words sit in the first 16 bytes of the dynamic shared memory, whose base the kernel
    // keeps in a register anyway (as static __shared__ symbols their addresses were rebuilt every update):
    //   s_nf[2]:   rows whose theta is non-finite, written by M-phase `it` into [it&1]
    //   s_rbad[2]: a ratio Y/c of E-phase `it` is non-finite, [it&1]
    int *const s_nf = reinterpret_cast<int *>(smem);
    int *const s_rbad = s_nf + 2;
    __shared__ int s_Q;        // columns whose baseSum is +-inf (epilogue)
    __shared__ __align__(8) uint64_t s_mbar;   // completion of the image's bulk copy
    uint32_t load_parity = 0;

    constexpr int NW = T / 32;
    // large slices (classes of 24 KB and up) come in by TMA bulk copy; for the ~11 KB images of the
    // small classes a cooperative 16-byte load loop measured ~3 % faster (no mbarrier round trip)
    constexpr bool kTma = (T >= 128);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_work = *qcount;
    if (kTma) {
        if (tid == 0) mbar_init(&s_mbar, 1);
        __syncthreads();
    }
    for (;;) {
        if (tid == 0) s_next = atomicAdd(ticket, 1);
        __syncthreads();
        const int w = s_next;
        if (w >= n_work) break;
        const int g = queue[w];
        const GroupImg gi = ginfo[g];
        const int M = gi.M, NL = gi.NL, nSliceR = gi.nSliceR, nSliceC = gi.nSliceC;
        const SellLayout L = sell_layout(gi.M, gi.NL, gi.nSliceR, gi.nSliceC, gi.slotsR, gi.slotsC);
        const uint8_t *gimg = pool + ((size_t)gi.off16 << 4);
        const int64_t row0 = grp_row_ptr[g];

        const uint32_t vec_bytes = sell_vec_bytes((uint32_t)M, (uint32_t)NL);
        const uint8_t *vec = smem + kSellFlagBytes;      // theta[M+1] | ratio[NL+1]
        uint8_t *simg = smem + kSellFlagBytes + vec_bytes;

        // ---- image -> shared memory ------------------------------------------------
        if (kTma) {
          if (tid == 0) {
            // the previous group's generic-proxy accesses to this shared memory are complete (barrier
            // above); order them before the async-proxy writes of the bulk copy
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(&s_mbar, L.copy_bytes);
            for (uint32_t off = 0; off < L.copy_bytes; off += 65536u)
                tma_load_1d(simg + off, gimg + off, min(65536u, L.copy_bytes - off), &s_mbar);
          }
        } else {
            const uint4 *src = reinterpret_cast<const uint4 *>(gimg);
            uint4 *dst = reinterpret_cast<uint4 *>(simg);
            const int n16 = (int)(L.copy_bytes >> 4);
            for (int i = tid; i < n16; i += T) dst[i] = __ldg(src + i);
        }
        const double *rs = reinterpret_cast<const double *>(simg + L.rs);
        const double *Yl = reinterpret_cast<const double *>(simg + L.Yl);
        const uint8_t *sideR = simg + L.sideR;
        const uint8_t *sideC = simg + L.sideC;
        const uint2 *descR = reinterpret_cast<const uint2 *>(simg + L.descR);
        const uint2 *descC = reinterpret_cast<const uint2 *>(simg + L.descC);
        double *theta = reinterpret_cast<double *>(smem + kSellFlagBytes);  // M + 1, theta[M] = 0 (padding target)
        double *ratio = theta + (M + 1);                                    // NL + 1, ratio[NL] = 0
        for (int i = tid; i < M; i += T) theta[i] = 1.0;                    // src/em.cpp:20-21
        if (tid == 0) {
            theta[M] = 0.0;
            ratio[NL] = 0.0;
            s_nf[0] = s_nf[1] = 0; s_rbad[0] = s_rbad[1] = 0; s_Q = 0;
        }
        if (kTma) {
            mbar_wait(&s_mbar, load_parity);     // every thread observes the completed copy
            load_parity ^= 1;
        }
        __syncthreads();

        // ---- em_theta: while (delta > conv && iter < maxiter-1)  src/em.cpp:42 --------
        int iter = 0;
        bool over = (1.0 > P.conv);   // delta starts at 1 (src/em.cpp:38)
        while (over && iter < P.maxiter - 1) {
            ++iter;
            const int par = iter & 1;
            const int nP = s_nf[par ^ 1];          // non-finite thetas entering this update
            if (tid == 0) s_nf[par] = 0;
            // E-phase: c_j = sum_i a_ij*theta_i (:46-47), ratio_j = Y_j / c_j (:48)
            for (int sl = warp; sl < nSliceC; sl += NW) {
                const int j = sl * 32 + lane;          // the lane's column, by position
                double a1 = 0.0, a2 = 0.0;
                int cnt = 0;
                if (nP == 0) {
                    line_sums<0, U>(sideC, descC[sl], lane, vec, a1, a2, cnt);
                } else {
                    // dense semantics: a row with non-finite theta poisons every column it has a
                    // structural zero in (0*NaN), SURVEY Appendix C.14
                    line_sums<1>(sideC, descC[sl], lane, vec, a1, a2, cnt);
                    if (cnt < nP) a1 = qnan();
                }
                if (j < NL) {
                    const double r = __ddiv_rn(Yl[j], __dadd_rn(a1, a2));
                    ratio[j] = r;
                    if (!is_finite(r)) s_rbad[par] = 1;
                }
            }
            __syncthreads();
            // M-phase: theta_i = sum_j (a_ij*theta_i)*ratio_j / rs_i, NaN terms -> 0 (:48-50)
            const bool hazard = (s_rbad[par] != 0) || (nP != 0);
            if (tid == 0) s_rbad[par ^ 1] = 0;
            int flag = (P.conv < 0.0) ? 1 : 0;   // delta >= 0 > conv whenever conv is negative
            // slices are sorted by line length: deal them to the warps in snake order (w0 w1 | w1 w0 | ...) so that
            // the warp that got the longest slice gets the shortest of the next pass
            for (int base = 0; base < nSliceR; base += NW) {
                const int sl = base + (((base / NW) & 1) ? NW - 1 - warp : warp);
                if (sl >= nSliceR) continue;
                const int i = sl * 32 + lane;        // the lane's row, by position
                const double th = theta[min(i, M)];  // padding lanes read theta[M] = 0.0
                double a1 = 0.0, a2 = 0.0;
                const uint2 d = descR[sl];
                const uint8_t *p = sideR + d.x;
                const uint32_t ns = d.y;
                if (!hazard) {
#pragma unroll U
                    for (uint32_t s = 0; s < ns; ++s, p += kSlotBytes) {
                        const double2 v = reinterpret_cast<const double2 *>(p)[lane];
                        const uint32_t oe = reinterpret_cast<const uint16_t *>(p + 512)[lane];
                        const uint32_t oo = reinterpret_cast<const uint16_t *>(p + 576)[lane];
                        a1 = __dadd_rn(a1, __dmul_rn(__dmul_rn(v.x, th), vec_at(vec, oe)));
                        a2 = __dadd_rn(a2, __dmul_rn(__dmul_rn(v.y, th), vec_at(vec, oo)));
                    }
                } else {
                    for (uint32_t s = 0; s < ns; ++s, p += kSlotBytes) {
                        const double2 v = reinterpret_cast<const double2 *>(p)[lane];
                        const uint32_t oe = reinterpret_cast<const uint16_t *>(p + 512)[lane];
                        const uint32_t oo = reinterpret_cast<const uint16_t *>(p + 576)[lane];
                        double te = __dmul_rn(__dmul_rn(v.x, th), vec_at(vec, oe));
                        double to = __dmul_rn(__dmul_rn(v.y, th), vec_at(vec, oo));
                        if (te != te) te = 0.0;   // replace(nan, 0)  (:49)
                        if (to != to) to = 0.0;
                        a1 = __dadd_rn(a1, te);
                        a2 = __dadd_rn(a2, to);
                    }
                }
                // delta = max over {after > minvalue} of |after-before|/after  (:54-63); only
                // "delta > conv" is needed: screen without the division.  The exact division is
                // needed only inside the 2^-40 band around conv; it sits behind a warp-uniform
                // branch (vote) so that the compiler cannot evaluate it speculatively.
                bool inband = false;
                double after = 0.0, diff = 0.0;
                if (i < M) {
                    after = div_small_numerator(__dadd_rn(a1, a2), rs[i]);     // == __ddiv_rn, see common.cuh
                    theta[i] = after;
                    if (!is_finite(after)) atomicAdd(&s_nf[par], 1);
                    if (after > P.minvalue) {
                        diff = fabs(__dadd_rn(after, -th));
                        if (diff > __dmul_rn(after, P.conv_hi)) flag = 1;
                        else inband = (diff >= __dmul_rn(after, P.conv_lo));
                    }
                }
                if (BAMBU_V_LAZY ? __any_sync(0xffffffffu, inband) : inband) {
                    if (inband && __ddiv_rn(diff, after) > P.conv) flag = 1;
                }
            }
            over = __syncthreads_or(flag) != 0;
        }

        // ---- emWithL1 epilogue: src/em.cpp:97-102 -------------------------------------
        const int nP = (iter == 0) ? 0 : s_nf[iter & 1];
        const double *KY = reinterpret_cast<const double *>(gimg + L.KY);
        const uint8_t *cm = gimg + L.cm;
        const uint8_t *eq = gimg + L.eq;
        const uint32_t ratio_off = 8u * (uint32_t)(M + 1);
        const uint16_t *rowid = reinterpret_cast<const uint16_t *>(gimg + L.rowid);
        for (int sl = warp; sl < nSliceC; sl += NW) {
            const int j = sl * 32 + lane;
            double a1 = 0.0, a2 = 0.0;
            int cnt = 0;
            line_sums<1>(sideC, descC[sl], lane, vec, a1, a2, cnt);
            if (j < NL) {
                double c = __dadd_rn(a1, a2);
                if (cnt < nP) c = qnan();
                double b = __ddiv_rn(__ldg(KY + j), c);   // (K % Y) / colsum  (:98)
                if (b != b) b = 0.0;                      // replace(nan, 0)   (:99)
                ratio[j] = b;
                if (fabs(b) > 1.7976931348623157e308) atomicAdd(&s_Q, 1);
            }
        }
        __syncthreads();
        const int Q = s_Q;
        int bad = 0;
        for (int sl = warp; sl < nSliceR; sl += NW) {
            const int pos = sl * 32 + lane;
            const double th = theta[min(pos, M)];
            double t1 = 0, t2 = 0, f1 = 0, f2 = 0, u1 = 0, u2 = 0;
            int cnt = 0;
            const uint2 d = descR[sl];
            const uint32_t s0 = d.x / kSlotBytes, s1 = s0 + d.y;
            const uint8_t *p = sideR + d.x;
            for (uint32_t s = s0; s < s1; ++s, p += kSlotBytes) {
                const double2 v2 = reinterpret_cast<const double2 *>(p)[lane];
                const double ve = v2.x, vo = v2.y;
                const uint32_t oe = reinterpret_cast<const uint16_t *>(p + 512)[lane];
                const uint32_t oo = reinterpret_cast<const uint16_t *>(p + 576)[lane];
                const uint32_t ie = (oe - ratio_off) >> 3, io = (oo - ratio_off) >> 3;   // column position (cm is by position)
                const double be = vec_at(vec, oe), bo = vec_at(vec, oo);
                const double fe = __ldg(eq + (size_t)64 * s + lane) ? ve : __dmul_rn(ve, 0.0);        // fullAval = aval*equal
                const double fo = __ldg(eq + (size_t)64 * s + 32 + lane) ? vo : __dmul_rn(vo, 0.0);
                const double ue = __ldg(cm + ie) ? __dmul_rn(ve, 0.0) : ve;                            // uniqueAval = aval*!multi
                const double uo = __ldg(cm + io) ? __dmul_rn(vo, 0.0) : vo;
                cnt += (fabs(be) > 1.7976931348623157e308) + (fabs(bo) > 1.7976931348623157e308);
                t1 = __dadd_rn(t1, __dmul_rn(__dmul_rn(ve, th), be));
                t2 = __dadd_rn(t2, __dmul_rn(__dmul_rn(vo, th), bo));
                f1 = __dadd_rn(f1, __dmul_rn(__dmul_rn(fe, th), be));
                f2 = __dadd_rn(f2, __dmul_rn(__dmul_rn(fo, th), bo));
                u1 = __dadd_rn(u1, __dmul_rn(__dmul_rn(ue, th), be));
                u2 = __dadd_rn(u2, __dmul_rn(__dmul_rn(uo, th), bo));
            }
            if (pos < M) {
                const int i = __ldg(rowid + pos);       // the row this position holds
                double e0 = __dadd_rn(t1, t2), e1 = __dadd_rn(f1, f2), e2 = __dadd_rn(u1, u2);
                // dense 0*NaN / 0*Inf at the structural zeros of this row
                if (!is_finite(th) || cnt < Q) e0 = e1 = e2 = qnan();
                est[row0 + i] = e0;
                est[total_rows + row0 + i] = e1;
                est[2 * total_rows + row0 + i] = e2;
                bad |= !(is_finite(e0) && is_finite(e1) && is_finite(e2));
            }
        }
        const int anybad = __syncthreads_or(bad);
        if (tid == 0) {
            iters[g] = iter;
            status[g] = ((iter >= P.maxiter - 1 && over) ? 1 : 0) | (anybad ? 2 : 0);
        }
        // the barrier above also protects shared memory against the next group's load
    }
}

}  // namespace bambu
