// x87_sum.h -- `long double s; s += x;` on x86-64 (what R's sum() does) for non-negative finite doubles,
// in integer arithmetic, so that the GPU reproduces it bit for bit (tests/test_quant_dev.py::test_integer_x87_sum).
// Used by the device packer (R/bambu-quantify_utilityFunctions.R:280) and by the single-transcript closed form
// (R/bambu-quantify_utilityFunctions.R:410-414).  The includer defines QD_FN (host/device qualifiers).
#pragma once
#include <cstdint>
#include <cstring>

namespace qd {

// ---- x87 extended precision accumulation of non-negative doubles, in integers ---------------------
// value = m * 2^(e - 63); m has bit 63 set, or m == 0 for zero.  Round to nearest even at 64 bits after
// every addition, like `long double s; s += x;` on x86-64 (what R's sum() does).
struct Ext {
    uint64_t m;
    int32_t e;
};

QD_FN Ext ext_from_double(double d)       // d >= 0, finite
{
    uint64_t b;
    memcpy(&b, &d, 8);
    const int ex = (int)((b >> 52) & 0x7ff);
    const uint64_t fr = b & 0xfffffffffffffull;
    Ext r;
    if (ex == 0) {
        if (fr == 0) { r.m = 0; r.e = 0; return r; }
        int lz = 0;
        uint64_t t = fr;
        while (!(t >> 63)) { t <<= 1; ++lz; }
        r.m = t;
        r.e = -1011 - lz;
        return r;
    }
    r.m = (fr | 0x10000000000000ull) << 11;
    r.e = ex - 1023;
    return r;
}

QD_FN Ext ext_add(Ext a, Ext b)
{
    if (a.m == 0) return b;
    if (b.m == 0) return a;
    if (a.e < b.e) { Ext t = a; a = b; b = t; }
    const int diff = a.e - b.e;
    if (diff >= 65) return a;                            // b is below a quarter ulp of a
    typedef unsigned __int128 u128;
    const u128 A = (u128)a.m << 63, Bfull = (u128)b.m << 63;
    const u128 B = Bfull >> diff;
    const bool sticky = (B << diff) != Bfull;            // only possible for diff == 64
    const u128 S = A + B;
    uint64_t m, rem, half;
    int e = a.e;
    if ((uint64_t)(S >> 127)) { m = (uint64_t)(S >> 64); rem = (uint64_t)S; half = 1ull << 63; ++e; }
    else { m = (uint64_t)(S >> 63); rem = (uint64_t)S & 0x7fffffffffffffffull; half = 1ull << 62; }
    const bool up = rem > half || (rem == half && (sticky || (m & 1)));
    if (up) {
        ++m;
        if (m == 0) { m = 1ull << 63; ++e; }
    }
    Ext r;
    r.m = m;
    r.e = e;
    return r;
}

QD_FN double ext_to_double(Ext x, bool *out_of_range)   // (double) of a long double: round to nearest even at 53 bits
{
    if (x.m == 0) return 0.0;
    uint64_t q = x.m >> 11;
    const uint64_t rem = x.m & 0x7ff;
    int e = x.e;
    if (rem > 0x400 || (rem == 0x400 && (q & 1))) {
        ++q;
        if (q == (1ull << 53)) { q >>= 1; ++e; }
    }
    const int ex = e + 1023;
    if (ex <= 0 || ex >= 2047) { *out_of_range = true; return 0.0; }
    const uint64_t b = ((uint64_t)ex << 52) | (q & 0xfffffffffffffull);
    double d;
    memcpy(&d, &b, 8);
    return d;
}

}  // namespace qd
