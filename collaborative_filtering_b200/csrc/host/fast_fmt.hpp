// Fast, byte-exact replacements for the two text primitives that bound the CLI tools once the math runs
// on the GPU (README.md:29 of the reference): printing a double the way the reference's
// `ostream << double` does (== printf("%g"), 6 significant digits) and reading it back (== strtod).
//
// Both have a fast path that is provably identical to the libc result and fall back to libc otherwise:
//   format_g6   scales |v| to [1e5, 1e6) by a correctly rounded power of ten (one double multiply: absolute
//               error < 3e-10 there), rounds to an integer; the libc path is taken when the scaled value is
//               within 1e-6 of a rounding tie or just below a decade boundary, for non-finite values and
//               outside 1e-290 .. 1e290
//   parse_double  Clinger's fast path: <= 15 significant digits and a decimal exponent within +-22 give an
//               exactly representable integer and power of ten, so one IEEE multiply/divide is correctly rounded
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

namespace gsihost {

struct Pow10Table {
    double p[2 * 300 + 1];           // 10^-300 .. 10^300, each within 1 ulp (built in 80-bit arithmetic, then rounded)
    Pow10Table() {
        long double up = 1.0L, dn = 1.0L;
        p[300] = 1.0;
        for (int i = 1; i <= 300; ++i) { up *= 10.0L; dn /= 10.0L; p[300 + i] = (double)up; p[300 - i] = (double)dn; }
    }
    double operator()(int e) const { return p[300 + e]; }
};
inline const Pow10Table& pow10_table() { static const Pow10Table t; return t; }

// writes printf("%g", v) into buf (no terminator needed by the caller: returns the length; buf >= 32 bytes)
inline int format_g6(char* buf, double v) {
    if (v == 0.0) {
        if (signbit(v)) { buf[0] = '-'; buf[1] = '0'; return 2; }
        buf[0] = '0';
        return 1;
    }
    const double a = fabs(v);
    if (!(a >= 1e-290 && a <= 1e290)) return snprintf(buf, 32, "%g", v);        // also NaN / inf
    uint64_t bits;
    memcpy(&bits, &a, 8);
    const int e2 = (int)(bits >> 52) - 1022;                                       // a = f * 2^e2, 0.5 <= f < 1 (normal numbers)
    int e10 = (int)floor((e2 - 1) * 0.30102999566398120);
    const Pow10Table& P = pow10_table();
    double scaled = a * P(5 - e10);
    if (scaled >= 1e6) { ++e10; scaled = a * P(5 - e10); }
    else if (scaled < 1e5) { --e10; scaled = a * P(5 - e10); }
    // just below a decade boundary the value may round up into the next decade: libc decides.  (Just ABOVE a
    // boundary both decades print the same digits "1" with this exponent, so no guard is needed there.)
    if (!(scaled >= 1e5 && scaled <= 999999.499999)) return snprintf(buf, 32, "%g", v);
    const uint32_t fl = (uint32_t)scaled;
    const double frac = scaled - (double)fl;
    if (fabs(frac - 0.5) < 1e-6) return snprintf(buf, 32, "%g", v);                // (near) tie: let libc decide
    uint32_t digits = fl + (frac > 0.5 ? 1u : 0u);                                 // 100000 .. 999999
    char d[6];
    for (int i = 5; i >= 0; --i) { d[i] = (char)('0' + digits % 10); digits /= 10; }
    int nd = 6;
    while (nd > 1 && d[nd - 1] == '0') --nd;                                       // %g strips trailing zeros
    char* o = buf;
    if (v < 0) *o++ = '-';
    if (e10 < -4 || e10 >= 6) {                                                    // scientific
        *o++ = d[0];
        if (nd > 1) { *o++ = '.'; for (int i = 1; i < nd; ++i) *o++ = d[i]; }
        *o++ = 'e';
        int x = e10;
        if (x < 0) { *o++ = '-'; x = -x; } else *o++ = '+';
        if (x >= 100) { *o++ = (char)('0' + x / 100); x %= 100; }
        *o++ = (char)('0' + x / 10);
        *o++ = (char)('0' + x % 10);
    } else if (e10 >= 0) {                                                         // fixed, integer part of e10 + 1 digits
        for (int i = 0; i <= e10; ++i) *o++ = (i < nd) ? d[i] : '0';
        if (nd > e10 + 1) { *o++ = '.'; for (int i = e10 + 1; i < nd; ++i) *o++ = d[i]; }
    } else {                                                                       // 0.000ddd
        *o++ = '0'; *o++ = '.';
        for (int i = 0; i < -e10 - 1; ++i) *o++ = '0';
        for (int i = 0; i < nd; ++i) *o++ = d[i];
    }
    return (int)(o - buf);
}

// strtod(p, &end) for the tokens the tools read; returns false when no number starts at p
inline bool parse_double(const char* p, const char* limit, double& out, const char*& end) {
    static const double kPow[23] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15, 1e16,
                                    1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
    const char* s = p;
    bool neg = false;
    if (s < limit && (*s == '-' || *s == '+')) { neg = (*s == '-'); ++s; }
    uint64_t mant = 0;
    int nsig = 0, dec_exp = 0;
    bool any = false, ok = true;
    while (s < limit && *s >= '0' && *s <= '9') {
        any = true;
        if (mant || *s != '0') { if (nsig < 18) { mant = mant * 10 + (uint64_t)(*s - '0'); ++nsig; } else { ++dec_exp; ok = false; } }
        ++s;
    }
    if (s < limit && *s == '.') {
        ++s;
        while (s < limit && *s >= '0' && *s <= '9') {
            any = true;
            if (mant || *s != '0') { if (nsig < 18) { mant = mant * 10 + (uint64_t)(*s - '0'); ++nsig; --dec_exp; } else ok = false; }
            else --dec_exp;
            ++s;
        }
    }
    if (!any) {                                          // "nan", "inf", garbage: libc decides
        char* q;
        out = strtod(p, &q);
        end = q;
        return q != p;
    }
    if (s < limit && (*s == 'e' || *s == 'E')) {
        const char* t = s + 1;
        bool eneg = false;
        if (t < limit && (*t == '-' || *t == '+')) { eneg = (*t == '-'); ++t; }
        if (t < limit && *t >= '0' && *t <= '9') {
            int ex = 0;
            while (t < limit && *t >= '0' && *t <= '9') { if (ex < 10000) ex = ex * 10 + (*t - '0'); ++t; }
            dec_exp += eneg ? -ex : ex;
            s = t;
        }
    }
    if (ok && nsig <= 15 && dec_exp >= -22 && dec_exp <= 22) {
        double v = (double)mant;                         // exact: < 10^15 < 2^53
        v = dec_exp >= 0 ? v * kPow[dec_exp] : v / kPow[-dec_exp];
        out = neg ? -v : v;
        end = s;
        return true;
    }
    char* q;                                             // long mantissas / extreme exponents
    out = strtod(p, &q);
    end = q;
    return q != p;
}

}  // namespace gsihost
