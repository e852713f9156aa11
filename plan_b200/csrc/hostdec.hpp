// hostdec.hpp -- exact host-side finalisation arithmetic of libplangpu.
//
// Aggregates leave the GPU as exact 128-bit integers; the value the reference would hold
// in a govalues Decimal (19-digit coefficient, half-even rounding; go.mod:15
// github.com/govalues/decimal v0.1.28) is derived here:
//   sum(DECIMAL)  -> function_aggr.go:684-689 (Decimal.Add fold)
//   avg(DECIMAL)  -> function_aggr.go:886-895 (sum.Quo(count))
//   avg(INT32)    -> function_aggr.go:881-885 (float64 sum / float64 count)
// Independent of oracle/ by construction (the product never links test code).
#pragma once
#include <stdint.h>

#include "common.cuh"

#ifdef __CUDACC__
#define HD_FN __host__ __device__ inline
#else
#define HD_FN inline
#endif

namespace pg {

struct HDec {
    u64 coef = 0;
    int scale = 0;
    bool neg = false;
};

constexpr int HD_MAXPREC = 19;
#define HD_MAXCOEF 9999999999999999999ULL

HD_FN u128 hd_pow10(int n)
{
    u128 r = 1;
    for (int i = 0; i < n; i++) r *= 10;
    return r;
}

HD_FN int hd_digits(u128 x)
{
    int d = 0;
    while (x > 0) { x /= 10; d++; }
    return d;
}

// divide by 10^shift, round half to even
HD_FN u128 hd_shift_right_even(u128 x, int shift)
{
    if (shift <= 0) return x;
    if (shift > 38) return 0;
    u128 p = hd_pow10(shift), q = x / p, r = x % p, half = p / 2;
    if (r > half || (r == half && (q & 1))) q += 1;
    return q;
}

// Bring an arbitrary-precision coefficient into the 19-digit format.  false = the integer
// part alone needs more than 19 digits (the reference panics with a decimal overflow).
HD_FN bool hd_normalise(bool neg, u128 coef, int scale, HDec *out)
{
    for (int guard = 0; guard < 4; guard++) {
        int prec = hd_digits(coef);
        if (prec - scale > HD_MAXPREC) return false;
        if (scale < 0) { coef *= hd_pow10(-scale); scale = 0; continue; }
        if (scale >= prec && scale > HD_MAXPREC) {
            coef = hd_shift_right_even(coef, scale - HD_MAXPREC);
            scale = HD_MAXPREC;
        } else if (prec > HD_MAXPREC) {
            coef = hd_shift_right_even(coef, prec - HD_MAXPREC);
            scale -= prec - HD_MAXPREC;
        }
        if (coef <= (u128)HD_MAXCOEF) {
            out->coef = (u64)coef;
            out->scale = scale;
            out->neg = neg;
            return true;
        }
    }
    return false;
}

HD_FN bool hd_from_i128(i128 v, int scale, HDec *out)
{
    bool neg = v < 0;
    u128 mag = neg ? (u128)(-(v + 1)) + 1 : (u128)v;
    return hd_normalise(neg, mag, scale, out);
}

HD_FN HDec hd_trim(HDec d, int min_scale)
{
    while (d.scale > min_scale && d.coef % 10 == 0) { d.coef /= 10; d.scale--; }
    return d;
}

// a / b with govalues' Quo contract: exact when the quotient terminates within 19 digits,
// otherwise the 38-digit truncated quotient rounded half-even to 19 digits; trailing zeros
// trimmed down to max(0, a.scale - b.scale).
HD_FN bool hd_quo(const HDec &a, const HDec &b, HDec *out)
{
    if (b.coef == 0) return false;
    bool neg = a.neg != b.neg;
    int pref = a.scale - b.scale;
    if (pref < 0) pref = 0;
    if (a.coef == 0) { out->coef = 0; out->scale = pref; out->neg = false; return true; }
    HDec f;
    bool done = false;
    {
        u128 num = a.coef, den = b.coef;
        int scale = a.scale - b.scale;
        int up = HD_MAXPREC - hd_digits(num);
        if (up > 0) { num *= hd_pow10(up); scale += up; }
        bool fits = true;
        if (scale > HD_MAXPREC) {
            int s2 = scale - HD_MAXPREC;
            if (s2 > 19 || den * hd_pow10(s2) > (u128)HD_MAXCOEF) fits = false;
            else { den *= hd_pow10(s2); scale = HD_MAXPREC; }
        }
        if (fits && scale < 0) {
            int s2 = -scale;
            if (s2 > 19 || num * hd_pow10(s2) > (u128)HD_MAXCOEF) fits = false;
            else { num *= hd_pow10(s2); scale = 0; }
        }
        if (fits && num % den == 0) done = hd_normalise(neg, num / den, scale, &f);
    }
    if (!done) {
        int up = 2 * HD_MAXPREC - hd_digits(a.coef);
        u128 num = (u128)a.coef * hd_pow10(up);
        if (!hd_normalise(neg, num / (u128)b.coef, a.scale - b.scale + up, &f)) return false;
    }
    *out = hd_trim(f, pref);
    return true;
}

}  // namespace pg
