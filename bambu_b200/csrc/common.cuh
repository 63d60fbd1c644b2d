// common.cuh -- shared definitions of the device code (sm_100a only).
//
// Vocabulary: a *gene group* is the unit of one emWithL1 call in the reference
// (R/bambu-quantify_utilityFunctions.R:401-431); its *image* is the compacted,
// doubly-ordered copy of its compatibility slice that the EM kernels iterate on.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace bambu {

// ---- raw arena, device pointers (include/bambu_em.h) -------------------------
struct ArenaDev {
    const int64_t *grp_row_ptr, *grp_col_ptr, *row_nnz_ptr;
    const int32_t *col_idx;
    const double *aval;
    const uint8_t *nz_flags;
    const double *Y, *K;
    const uint8_t *col_flags;
    int64_t n_groups, total_rows;
};

// ---- wire form v2 ("live arena"), device pointers (include/bambu_em.h, bambu_em_batched_v2) -----------
// What the packer can drop without changing a result bit is dropped before the copy: columns with
// Y == 0, entries with aval == 0 (see the image comment below); rs_i -- which sums ALL entries of a
// row (src/em.cpp:33) -- therefore travels as a per-row value.  Entries carry the LIVE column index
// and the parity of the ORIGINAL column index (bit 0), which is all the summation order needs.
struct WireDev {
    const int64_t *grp_row_ptr, *grp_col_ptr, *grp_ent_ptr;   // rows / live columns / live entries per group
    const int64_t *grp_idx_off;     // byte offset of the group's index block inside cidx (library-computed)
    const uint32_t *row_end;        // per row: end of its entries, relative to the group's first entry
    const double *rs;               // per row
    const double *aval;             // per entry
    const uint8_t *cidx;            // per group u16[e] (NL <= 32767) or u32[e]: live_col << 1 | parity
    const uint32_t *eq_bits;        // bit k: entry k is `equal`
    const void *colY, *colK;        // col_mode 0: f64 Y, f64 K; 1: i32 nobs, i32 K (Y = nobs / K)
    const uint32_t *multi_bits;     // bit j: live column j is multi_align
    int32_t col_mode;
    int64_t n_groups, total_rows;
};

__host__ __device__ inline int wire_idx_width(int64_t NL) { return NL <= 32767 ? 2 : 4; }

__device__ __forceinline__ uint32_t wire_bit(const uint32_t *bits, int64_t k)
{
    return (bits[k >> 5] >> (uint32_t)(k & 31)) & 1u;
}

// Y_j and K_j of live column j (absolute index).  Mode 1 forms Y = nobs / K on the device exactly as
// R does on the host (R/bambu-quantify_utilityFunctions.R:221-228: n.obs = nobs / K, both integer-valued).
__device__ __forceinline__ void wire_col(const WireDev &W, int64_t j, double &y, double &k)
{
    if (W.col_mode == 0) {
        y = static_cast<const double *>(W.colY)[j];
        k = static_cast<const double *>(W.colK)[j];
    } else {
        k = (double)static_cast<const int32_t *>(W.colK)[j];
        y = __ddiv_rn((double)static_cast<const int32_t *>(W.colY)[j], k);
    }
}

// ---- what the pack kernel found in a group ------------------------------------
struct PackInfo {
    int32_t NL;     // live columns: Y != 0
    int32_t nnzL;   // entries with aval != 0 in live columns
};

// ---- image of a resident-tier group -------------------------------------------
// A group's EM only ever needs the columns with Y_j != 0 ("live"): a column with
// Y_j == 0 has ratio Y_j/c_j in {0, NaN}, so every term it adds to a row sum is
// +-0 or NaN->0 (src/em.cpp:48-49), and its baseSum K_j*Y_j/c_j is 0 or NaN->0
// (src/em.cpp:98-99); c_j itself is used nowhere else.  Entries with aval == 0
// are no-ops for the same reason.  rs_i (src/em.cpp:33) still sums ALL entries.
// The position parity that selects Armadillo's accumulator (even index ->
// acc1, odd -> acc2) is kept in bit 0 of every index.
//
//   [rs f64 M][Yl f64 NL][cval f64 nnzL][rval f64 nnzL]
//   [rp u16 M+1][cp u16 NL+1][cidx u16 nnzL][ridx u16 nnzL]      <- copied to shared memory
//   [KY f64 NL][eq u8 nnzL][cm u8 NL]                            <- read once, in the epilogue
//
//   cval/cidx: row-major (row i = [rp[i], rp[i+1])), cidx = live_col<<1 | (orig_col&1)
//   rval/ridx: column-major (col j = [cp[j], cp[j+1])), ridx = row<<1 | (row&1), rows ascending
struct ImgLayout {
    uint32_t rs, Yl, cval, rval, rp, cp, cidx, ridx, copy_bytes, KY, eq, cm, total;
};

__host__ __device__ inline uint32_t align_up(uint32_t x, uint32_t a) { return (x + a - 1) / a * a; }

__host__ __device__ inline ImgLayout img_layout(uint32_t M, uint32_t NL, uint32_t nnzL)
{
    ImgLayout L;
    L.rs = 0;
    L.Yl = L.rs + 8u * M;
    L.cval = L.Yl + 8u * NL;
    L.rval = L.cval + 8u * nnzL;
    L.rp = L.rval + 8u * nnzL;
    L.cp = L.rp + 2u * (M + 1);
    L.cidx = L.cp + 2u * (NL + 1);
    L.ridx = L.cidx + 2u * nnzL;
    L.copy_bytes = align_up(L.ridx + 2u * nnzL, 16);
    L.KY = L.copy_bytes;
    L.eq = L.KY + 8u * NL;
    L.cm = L.eq + nnzL;
    L.total = align_up(L.cm + NL, 16);
    return L;
}

// shared-memory bytes the resident EM kernel needs for a group (image + theta + ratio)
__host__ __device__ inline uint32_t em_smem_bytes(uint32_t M, uint32_t NL, uint32_t nnzL)
{
    return img_layout(M, NL, nnzL).copy_bytes + 8u * M + 8u * NL;
}

// shared-memory bytes the pack kernel needs for a group
__host__ __device__ inline uint32_t pack_smem_bytes(uint32_t M, uint32_t N)
{
    return 4u * (N + (M + 1) + (N + 1) + N);
}

// ---- SELL image: what the EM kernel iterates on -----------------------------------
// Every line (a row for the M-phase, a live column for the E-phase) is split into
// its even-position and odd-position entries -- the two accumulators of arma::sum --
// each kept in ascending index order.  Slot k of a line holds its k-th even and k-th
// odd entry, so a line needs L = max(#even, #odd) slots; missing entries are padded
// with (value 0.0, index of a dummy vector element that is always 0.0), which adds an
// exact +0.0.  Lines are ordered by L (descending) and cut into slices of 32 -- one
// warp -- padded to the slice's longest line; inside a slice storage is slot-major,
// lane-minor, so that a warp reads 32 consecutive values (no bank conflicts) and runs
// a loop whose trip count is uniform (no divergence, no parity branch).
//
// Lines live in POSITION order (position p = 32*slice + lane, lines sorted by length): the EM kernel's
// vectors theta / ratio and the per-line constants below are all indexed by position, so a lane finds its
// own line's data at a fixed place and only the gathers go through stored offsets.
//
//   [rs f64 M][Yl f64 NL]                                             (by position)
//   [row side : slotsR x {(valE, valO) f64x2 x32 | offE u16 x32 | offO u16 x32}]        640 B / slot
//               (an index is stored as the byte offset of the gathered element in the EM
//                kernel's vector region -- theta[M+1] then ratio[NL+1], by position -- so 8*(M+NL+2) < 65536)
//   [col side : slotsC x {same}]
//   [descR u32x2 nSliceR][descC u32x2 nSliceC]   {byte offset of the slice's first slot in its side, slots}
//                                                                   <- copied to shared memory
//   [KY f64 NL][cm u8 NL+1] (by position) [rowid u16 M: row at each position][eq u8 : slotsR x {E x32 | O x32}]
//                                                                   <- read once, in the epilogue
struct GroupImg {
    uint32_t off16;       // image offset inside the pool, in 16-byte units
    uint32_t smem_bytes;  // shared memory the EM kernel needs; 0 = not queued
    uint16_t M, NL;
    uint16_t nSliceR, nSliceC;
    uint32_t slotsR, slotsC;
    uint32_t nnzL;
    int32_t cls;
};
static_assert(sizeof(GroupImg) == 32, "GroupImg layout");

constexpr uint32_t kSlotBytes = 640;

struct SellLayout {
    uint32_t rs, Yl, sideR, sideC, descR, descC, copy_bytes, KY, cm, rowid, eq, total;
};

__host__ __device__ inline SellLayout sell_layout(uint32_t M, uint32_t NL, uint32_t nSliceR, uint32_t nSliceC,
                                                  uint32_t slotsR, uint32_t slotsC)
{
    SellLayout L;
    L.rs = 0;
    L.Yl = 8u * M;
    L.sideR = align_up(L.Yl + 8u * NL, 16);   // slots are read as double2
    L.sideC = L.sideR + kSlotBytes * slotsR;
    L.descR = L.sideC + kSlotBytes * slotsC;
    L.descC = L.descR + 8u * nSliceR;
    L.copy_bytes = align_up(L.descC + 8u * nSliceC, 16);
    L.KY = L.copy_bytes;
    L.cm = L.KY + 8u * NL;
    L.rowid = align_up(L.cm + NL + 1, 2);
    L.eq = align_up(L.rowid + 2u * M, 16);
    L.total = align_up(L.eq + 64u * slotsR, 16);
    return L;
}

// vector region of the EM kernel: theta[M+1] | ratio[NL+1]
__host__ __device__ inline uint32_t sell_vec_bytes(uint32_t M, uint32_t NL) { return align_up(8u * (M + NL + 2), 16); }

constexpr uint32_t kSellFlagBytes = 16;   // per-update flag words in front of the vectors

__host__ __device__ inline uint32_t sell_smem_bytes(const SellLayout &L, uint32_t M, uint32_t NL)
{
    return kSellFlagBytes + sell_vec_bytes(M, NL) + L.copy_bytes;
}

// shared-memory scratch of the SELL builder for a group (lengths, ids, slice starts, histogram)
__host__ __device__ inline uint32_t sell_scratch_bytes(uint32_t M, uint32_t N)
{
    return 4u * 1026u + 2u * (M + N) + 2u * (M + N + 64u) + 4u * ((M + 31) / 32 + (N + 31) / 32 + 4u) + 64u;
}

constexpr int kNumClasses = 6;
constexpr uint32_t kSmemMax = 227u * 1024u - 256u;  // per-CTA opt-in maximum minus static shared
constexpr int kMaxIndex = 32767;                    // u16 index carries a parity bit
constexpr int kMaxVecElems = 8190;                  // M + NL of a resident group: byte offsets must fit u16

// condition bits OR-ed into the device error word by the kernels
constexpr int kErrColRange = 1;   // col_idx outside [0, N)
constexpr int kErrColOrder = 2;   // col_idx not strictly increasing inside a row
constexpr int kErrPool = 4;       // image pool exhausted (the host grows it and reruns)
constexpr int kErrTooBig = 8;     // a group's SELL image exceeds shared memory (the host moves the group
                                  // to the streamed tier)
constexpr int kErrRowPtr = 16;    // wire v2: row_end decreasing or beyond the group's entries
constexpr int kErrParity = 32;    // wire v2: one live column carries both parities

// ---- block-wide exclusive scan, in place, n arbitrary ---------------------------
template <int T>
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t *a, int n, uint32_t *s_warp /* 34 words */)
{
    const int tid = threadIdx.x;
    const int chunk = (n + T - 1) / T;
    const int lo = min(n, tid * chunk), hi = min(n, lo + chunk);
    uint32_t sum = 0;
    for (int k = lo; k < hi; ++k) sum += a[k];
    uint32_t incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
        if ((tid & 31) >= o) incl += v;
    }
    if ((tid & 31) == 31) s_warp[tid >> 5] = incl;
    __syncthreads();
    if (tid < 32) {
        uint32_t w = (tid < T / 32) ? s_warp[tid] : 0u;
        uint32_t wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t v = __shfl_up_sync(0xffffffffu, wi, o);
            if (tid >= o) wi += v;
        }
        s_warp[tid] = wi - w;
        if (tid == 31) s_warp[32] = wi;
    }
    __syncthreads();
    uint32_t run = s_warp[tid >> 5] + (incl - sum);
    const uint32_t total = s_warp[32];
    for (int k = lo; k < hi; ++k) {
        uint32_t t = a[k];
        a[k] = run;
        run += t;
    }
    __syncthreads();
    return total;
}

// In-place Shell sort (Ciura gaps) of one column's (row index, value) pairs by row index: the scatter
// that builds a column leaves its entries in arbitrary order; arma::sum adds them by ascending row.
// Columns are short (tens of entries) but nothing bounds them: O(L^1.3) keeps a column shared by
// thousands of transcripts from stalling its thread for L^2 steps.
template <typename IdxT>
__device__ __forceinline__ void sort_column(IdxT *idx, double *val, uint32_t lo, uint32_t hi)
{
    const uint32_t n = hi - lo;
    if (n < 2) return;
    const uint32_t gaps[9] = {1750, 701, 301, 132, 57, 23, 10, 4, 1};
#pragma unroll 1
    for (int g = 0; g < 9; ++g) {
        const uint32_t gap = gaps[g];
        if (gap >= n) continue;
        for (uint32_t p = lo + gap; p < hi; ++p) {
            const IdxT kr = idx[p];
            const double kv = val[p];
            uint32_t q = p;
            while (q >= lo + gap && idx[q - gap] > kr) {
                idx[q] = idx[q - gap];
                val[q] = val[q - gap];
                q -= gap;
            }
            idx[q] = kr;
            val[q] = kv;
        }
    }
}

// exponent field not all ones; an integer test on the high word keeps it off the FP64 pipe
__device__ __forceinline__ bool is_finite(double x) { return (__double2hiint(x) & 0x7ff00000) != 0x7ff00000; }

// num / den, correctly rounded like __ddiv_rn, for the M-phase's theta' = sum / rs.
// The library division leaves its fast path when the numerator is below 2^-969 (or zero): thetas that have
// decayed that far are common late in long runs and each one stalls its warp for ~450 cycles.  Such a numerator
// is scaled by 2^960 first (exact) and the quotient scaled back (exact while it stays a normal number); zero is
// passed through; everything else -- subnormal quotients, den = 0, NaN -- still goes to the real division.
__device__ __forceinline__ double div_small_numerator(double num, double den)
{
    const bool tiny = (__double2hiint(num) & 0x7fffffff) < 0x03700000;      // |num| < 2^-968
    double top = num;
    if (tiny) top = (num == 0.0) ? 1.0 : __dmul_rn(num, 0x1p960);
    double q = __ddiv_rn(top, den);
    if (tiny) {
        if (num == 0.0) q = (den > 0.0) ? num : __ddiv_rn(num, den);
        else if (fabs(q) >= 0x1p-61 && fabs(q) <= 1.7976931348623157e308) q = __dmul_rn(q, 0x1p-960);
        else q = __ddiv_rn(num, den);
    }
    return q;
}

}  // namespace bambu
