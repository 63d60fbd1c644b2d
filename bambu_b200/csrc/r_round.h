// r_round.h -- R's round(x, digits) for doubles (R >= 4.0: the double closest to the correctly rounded decimal;
// exact ties of the binary value go to the even digit), in integer arithmetic so that host and device agree
// bit for bit.  bambu rounds counts / CPM / fullLengthCounts with it (R/bambu-quantify.R:30-36, sig.digit = 5).
// Same results as the host mirror bambu_b200/assays.py::r_round, which is pinned on the reference's stored assays.
// The includer defines QD_FN (host / device qualifiers).
#pragma once
#include <cstdint>
#include <cstring>

namespace qd {

QD_FN double r_round(double x, int digits)
{
    uint64_t b;
    memcpy(&b, &x, 8);
    const uint64_t sign = b & 0x8000000000000000ull;
    const int ex = (int)((b >> 52) & 0x7ff);
    if (ex == 0x7ff || digits < 0 || digits > 15) return x;            // NaN / Inf pass; digits outside [0, 15]: unchanged
    uint64_t m = b & 0xfffffffffffffull;
    int e;                                                             // |x| = m * 2^e
    if (ex == 0) { if (m == 0) return x; e = -1074; }
    else { m |= 0x10000000000000ull; e = ex - 1075; }
    if (e >= 0) return x;                                              // an integer already
    uint64_t p10 = 1;
    for (int k = 0; k < digits; ++k) p10 *= 10;
    typedef unsigned __int128 u128;
    const u128 P = (u128)m * p10;                                      // < 2^53 * 10^15 < 2^103
    const int s = -e;
    u128 n, rem, half;
    if (s >= 127) { n = 0; rem = (P != 0) ? 1 : 0; half = 2; }         // far below half a unit of the last digit
    else { n = P >> s; rem = P & (((u128)1 << s) - 1); half = (u128)1 << (s - 1); }
    if (rem > half || (rem == half && (n & 1))) ++n;
    if (n >= ((u128)1 << 53)) return x;                                // more digits asked for than the double holds
    const double q = (double)(uint64_t)n / (double)p10;                // correctly rounded: the double closest to n / 10^digits
    uint64_t qb;
    memcpy(&qb, &q, 8);
    qb |= sign;
    double r;
    memcpy(&r, &qb, 8);
    return r;
}

}  // namespace qd
