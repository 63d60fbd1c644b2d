// sell_kernel.cuh -- stage-1 image (compact CSR + CSC, pack_kernel.cuh) -> SELL image
// (common.cuh) of one gene group, then classify the group by the shared memory its
// EM needs and queue it for the EM kernel of that class.
//
// One CTA per group.  Nothing here changes a value or the order in which a line's
// terms are added; it only splits each line by position parity (the two accumulators
// of arma::sum, src/em.cpp:47,50), pads, sorts lines by length and transposes.
#pragma once
#include "common.cuh"

namespace bambu {

struct SellSide {
    uint16_t *L;      // n          slots needed per line; after sell_measure: the line's position
    uint16_t *id;     // 32*nSlice  line at each position (n = padding)
    uint32_t *start;  // nSlice+1   first slot of each slice
    int n, nSlice;
    uint32_t slots;
};

// lengths, order, slices of one side.  ptr/idx: the side's compact arrays in the
// stage-1 image (idx bit 0 = position parity).
template <int T>
__device__ __forceinline__ void sell_measure(SellSide &S, const uint16_t *__restrict__ ptr,
                                             const uint16_t *__restrict__ idx, uint32_t *hist, uint32_t *s_warp)
{
    const int tid = threadIdx.x, n = S.n;
    for (int b = tid; b < 1024; b += T) hist[b] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += T) {
        int ne = 0, no = 0;
        for (int k = ptr[i]; k < ptr[i + 1]; ++k) {
            if (idx[k] & 1) ++no; else ++ne;
        }
        const int L = max(ne, no);
        S.L[i] = (uint16_t)L;
        atomicAdd(&hist[1023 - min(L, 1023)], 1u);   // reversed bins: ascending scan = descending length
    }
    __syncthreads();
    block_excl_scan<T>(hist, 1024, s_warp);
    for (int i = tid; i < n; i += T) {
        const uint32_t pos = atomicAdd(&hist[1023 - min((int)S.L[i], 1023)], 1u);
        S.id[pos] = (uint16_t)i;
    }
    for (int p = n + tid; p < S.nSlice * 32; p += T) S.id[p] = (uint16_t)n;
    __syncthreads();
    // slice length = longest line of the slice
    const int lane = tid & 31, warp = tid >> 5;
    for (int w = warp; w < S.nSlice; w += T / 32) {
        const int i = S.id[w * 32 + lane];
        int L = (i < n) ? (int)S.L[i] : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) L = max(L, __shfl_xor_sync(0xffffffffu, L, o));
        if (lane == 0) S.start[w] = (uint32_t)L;
    }
    __syncthreads();
    S.slots = block_excl_scan<T>(S.start, S.nSlice, s_warp);
    if (tid == 0) S.start[S.nSlice] = S.slots;
    for (int p = tid; p < n; p += T) S.L[S.id[p]] = (uint16_t)p;     // lengths are done with: line -> position
    __syncthreads();
}

// write one side's slots.  Slot = { (valE, valO) x 32 lanes | offE u16 x 32 | offO u16 x 32 };
// an index is stored as the BYTE OFFSET of the gathered element inside the EM kernel's vector
// region (theta at 0, ratio behind it): off = vec_off + 8 * index.  dummy = offset of the
// always-zero element of the vector this side gathers from.  pos_other: position of every line of the
// OTHER side (the gathered vector is stored by position).  eq_out/eq_in: per-entry "equal" flags (row side only).
template <int T>
__device__ __forceinline__ void sell_fill(const SellSide &S, const uint16_t *__restrict__ ptr,
                                          const uint16_t *__restrict__ idx, const double *__restrict__ val,
                                          uint8_t *__restrict__ side, uint32_t vec_off, uint16_t dummy,
                                          const uint16_t *__restrict__ pos_other,
                                          const uint8_t *__restrict__ eq_in, uint8_t *__restrict__ eq_out)
{
    const int tid = threadIdx.x;
    for (int p = tid; p < S.nSlice * 32; p += T) {
        const int i = S.id[p], lane = p & 31, w = p >> 5;
        const uint32_t s0 = S.start[w], Ls = S.start[w + 1] - s0;
        uint8_t *base = side + (size_t)kSlotBytes * s0;
        uint32_t ke = 0, ko = 0;
        if (i < S.n) {
            for (int k = ptr[i]; k < ptr[i + 1]; ++k) {
                const uint16_t ix = idx[k];
                const double v = val[k];
                const uint16_t off = (uint16_t)(vec_off + 8u * (uint32_t)pos_other[ix >> 1]);
                if (ix & 1) {
                    uint8_t *slot = base + (size_t)kSlotBytes * ko;
                    reinterpret_cast<double *>(slot)[2 * lane + 1] = v;
                    reinterpret_cast<uint16_t *>(slot + 576)[lane] = off;
                    if (eq_out) eq_out[(size_t)64 * (s0 + ko) + 32 + lane] = eq_in[k];
                    ++ko;
                } else {
                    uint8_t *slot = base + (size_t)kSlotBytes * ke;
                    reinterpret_cast<double *>(slot)[2 * lane] = v;
                    reinterpret_cast<uint16_t *>(slot + 512)[lane] = off;
                    if (eq_out) eq_out[(size_t)64 * (s0 + ke) + lane] = eq_in[k];
                    ++ke;
                }
            }
        }
        for (; ke < Ls; ++ke) {
            uint8_t *slot = base + (size_t)kSlotBytes * ke;
            reinterpret_cast<double *>(slot)[2 * lane] = 0.0;
            reinterpret_cast<uint16_t *>(slot + 512)[lane] = dummy;
            if (eq_out) eq_out[(size_t)64 * (s0 + ke) + lane] = 0;
        }
        for (; ko < Ls; ++ko) {
            uint8_t *slot = base + (size_t)kSlotBytes * ko;
            reinterpret_cast<double *>(slot)[2 * lane + 1] = 0.0;
            reinterpret_cast<uint16_t *>(slot + 576)[lane] = dummy;
            if (eq_out) eq_out[(size_t)64 * (s0 + ko) + 32 + lane] = 0;
        }
    }
}

struct ClassBounds {
    uint32_t max_bytes[kNumClasses];
};

template <int T>
__global__ void __launch_bounds__(T)
sell_kernel(const int32_t *__restrict__ work, int n_work, const int64_t *__restrict__ grp_row_ptr,
            const int64_t *__restrict__ img1_off, const uint8_t *__restrict__ img1_base,
            const PackInfo *__restrict__ info1, uint8_t *__restrict__ pool, unsigned long long pool_bytes,
            unsigned long long *__restrict__ pool_ctr, GroupImg *__restrict__ ginfo, ClassBounds cb,
            int32_t *__restrict__ queue, int64_t queue_stride, int *__restrict__ qcount, int *__restrict__ err_flag)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    __shared__ uint32_t s_warp[34];
    __shared__ unsigned long long s_off;

    if ((int)blockIdx.x >= n_work) return;
    const int tid = threadIdx.x;
    const int g = work[blockIdx.x];
    const int M = (int)(grp_row_ptr[g + 1] - grp_row_ptr[g]);
    const int NL = info1[g].NL, nnzL = info1[g].nnzL;
    const ImgLayout L1 = img_layout((uint32_t)M, (uint32_t)NL, (uint32_t)nnzL);
    const uint8_t *img1 = img1_base + img1_off[g];
    const uint16_t *rp = reinterpret_cast<const uint16_t *>(img1 + L1.rp);
    const uint16_t *cp = reinterpret_cast<const uint16_t *>(img1 + L1.cp);
    const uint16_t *cidx = reinterpret_cast<const uint16_t *>(img1 + L1.cidx);
    const uint16_t *ridx = reinterpret_cast<const uint16_t *>(img1 + L1.ridx);
    const double *cval = reinterpret_cast<const double *>(img1 + L1.cval);
    const double *rval = reinterpret_cast<const double *>(img1 + L1.rval);

    // scratch
    uint32_t *hist = reinterpret_cast<uint32_t *>(smem_raw);   // 1026
    SellSide R, C;
    R.n = M; R.nSlice = (M + 31) / 32;
    C.n = NL; C.nSlice = (NL + 31) / 32;
    R.start = hist + 1026;
    C.start = R.start + R.nSlice + 1;
    R.L = reinterpret_cast<uint16_t *>(C.start + C.nSlice + 1);
    C.L = R.L + M;
    R.id = C.L + NL;
    C.id = R.id + 32 * R.nSlice;

    sell_measure<T>(R, rp, cidx, hist, s_warp);
    sell_measure<T>(C, cp, ridx, hist, s_warp);

    const SellLayout L = sell_layout((uint32_t)M, (uint32_t)NL, (uint32_t)R.nSlice, (uint32_t)C.nSlice, R.slots, C.slots);
    const uint32_t smem_need = sell_smem_bytes(L, (uint32_t)M, (uint32_t)NL);
    if (tid == 0) {
        unsigned long long off = atomicAdd(pool_ctr, (unsigned long long)L.total);
        if (off + L.total > pool_bytes) {
            off = ~0ull;
            atomicOr(err_flag, kErrPool);
        }
        s_off = off;
    }
    __syncthreads();
    if (s_off == ~0ull) {
        if (tid == 0) { GroupImg z = {}; z.cls = -1; ginfo[g] = z; }
        return;
    }
    uint8_t *img = pool + s_off;

    // verbatim sections
    {
        const double *rs1 = reinterpret_cast<const double *>(img1 + L1.rs);
        const double *Y1 = reinterpret_cast<const double *>(img1 + L1.Yl);
        const double *KY1 = reinterpret_cast<const double *>(img1 + L1.KY);
        const uint8_t *cm1 = img1 + L1.cm;
        double *rs = reinterpret_cast<double *>(img + L.rs);
        double *Yl = reinterpret_cast<double *>(img + L.Yl);
        double *KY = reinterpret_cast<double *>(img + L.KY);
        uint8_t *cm = img + L.cm;
        // per-line constants by POSITION (R.id / C.id: line at each position)
        for (int p = tid; p < M; p += T) rs[p] = rs1[R.id[p]];
        for (int p = tid; p < NL; p += T) { const int j = C.id[p]; Yl[p] = Y1[j]; KY[p] = KY1[j]; cm[p] = cm1[j]; }
        if (tid == 0) cm[NL] = 0;
        uint16_t *rowid = reinterpret_cast<uint16_t *>(img + L.rowid);
        uint2 *descR = reinterpret_cast<uint2 *>(img + L.descR);
        uint2 *descC = reinterpret_cast<uint2 *>(img + L.descC);
        for (int p = tid; p < M; p += T) rowid[p] = R.id[p];
        for (int w = tid; w < R.nSlice; w += T) descR[w] = make_uint2(kSlotBytes * R.start[w], R.start[w + 1] - R.start[w]);
        for (int w = tid; w < C.nSlice; w += T) descC[w] = make_uint2(kSlotBytes * C.start[w], C.start[w + 1] - C.start[w]);
    }
    // vector region of the EM kernel: theta[M+1] at byte 0, ratio[NL+1] behind it.  The row side
    // gathers ratio[] (dummy element NL), the column side gathers theta[] (dummy element M).
    const uint32_t ratio_off = 8u * (uint32_t)(M + 1);
    sell_fill<T>(R, rp, cidx, cval, img + L.sideR, ratio_off, (uint16_t)(ratio_off + 8u * (uint32_t)NL), C.L, img1 + L1.eq,
                 img + L.eq);
    sell_fill<T>(C, cp, ridx, rval, img + L.sideC, 0u, (uint16_t)(8u * (uint32_t)M), R.L, nullptr, nullptr);

    if (tid == 0) {
        GroupImg gi;
        gi.off16 = (uint32_t)(s_off >> 4);
        gi.smem_bytes = smem_need;
        gi.M = (uint16_t)M; gi.NL = (uint16_t)NL;
        gi.nSliceR = (uint16_t)R.nSlice; gi.nSliceC = (uint16_t)C.nSlice;
        gi.slotsR = R.slots; gi.slotsC = C.slots;
        gi.nnzL = (uint32_t)nnzL;
        int c = 0;
        while (c < kNumClasses && smem_need > cb.max_bytes[c]) ++c;
        if (c == kNumClasses) {
            gi.cls = -1;
            gi.smem_bytes = 0;
            atomicOr(err_flag, kErrTooBig);
        } else {
            gi.cls = c;
            const int pos = atomicAdd(&qcount[c], 1);
            queue[(int64_t)c * queue_stride + pos] = g;
        }
        ginfo[g] = gi;
    }
}

}  // namespace bambu
