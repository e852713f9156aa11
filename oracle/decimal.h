/*
 * oracle/decimal.h -- TEST INFRASTRUCTURE (CPU). Not part of the product path.
 *
 * Restatement of the fixed-point decimal the reference computes with:
 * github.com/govalues/decimal v0.1.28 (go.mod:15 of the reference; third-party,
 * NOT vendored under /root/reference, so this follows the library's published
 * contract).  Reference call sites this stands in for:
 *   Add  pkg/compute/function_operator_binary.go:135, pkg/common/decimal.go:20
 *   Sub  pkg/compute/function_operator_binary.go:162
 *   Mul  pkg/compute/function_operator_binary.go:185, pkg/common/decimal.go:28
 *   Quo  pkg/compute/function_aggr.go:888 (avg = sum.Quo(count))
 *   Int64(scale)    pkg/chunk/vector.go:124, pkg/compute/sort_encoder.go:66
 *   NewFromInt64    pkg/chunk/value.go:41, pkg/chunk/vector.go:257
 *   Float64         pkg/compute/function_cast.go:350
 *
 * Model: value = (-1)^neg * coef * 10^-scale, coef <= 10^19-1, 0 <= scale <= 19.
 * Results that need more than 19 digits are rounded half-to-even by dropping
 * low digits (scale shrinks); overflow only when the integer part alone needs
 * more than 19 digits.  PARITY NOTE: only the SF1 golden results pin this
 * end to end; the >19-digit rounding regime is "parity unpinned".
 */
#ifndef ORACLE_DECIMAL_H
#define ORACLE_DECIMAL_H
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;

typedef struct {
    uint64_t coef;
    int8_t scale;
    uint8_t neg;
} dec_t;

#define DEC_MAXPREC 19
#define DEC_MAXSCALE 19
#define DEC_MAXCOEF 9999999999999999999ULL

static const uint64_t DEC_POW10[20] = {
    1ULL, 10ULL, 100ULL, 1000ULL, 10000ULL, 100000ULL, 1000000ULL, 10000000ULL, 100000000ULL,
    1000000000ULL, 10000000000ULL, 100000000000ULL, 1000000000000ULL, 10000000000000ULL,
    100000000000000ULL, 1000000000000000ULL, 10000000000000000ULL, 100000000000000000ULL,
    1000000000000000000ULL, 10000000000000000000ULL};

static inline u128 dec_pow10_128(int n)
{
    u128 r = 1;
    while (n-- > 0) r *= 10;
    return r;
}

static inline int dec_prec128(u128 x)
{
    int p = 0;
    while (x > 0) { x /= 10; p++; }
    return p;
}

/* x / 10^shift rounded half to even (fint.rshHalfEven / bint.rshHalfEven) */
static inline u128 dec_rsh_half_even128(u128 x, int shift)
{
    if (shift <= 0) return x;
    if (shift > 38) return 0;
    u128 y = dec_pow10_128(shift);
    u128 q = x / y, r = x % y, half = y / 2;
    if (r > half || (r == half && (q & 1))) q++;
    return q;
}

static inline dec_t dec_make(int neg, uint64_t coef, int scale)
{
    dec_t d; d.coef = coef; d.scale = (int8_t)scale; d.neg = (uint8_t)(neg ? 1 : 0);
    return d;
}

/* newFromBint: normalise an arbitrary-precision coefficient to <=19 digits.
 * returns 0 on success, -1 on decimal overflow. */
static inline int dec_from_u128(int neg, u128 coef, int scale, int min_scale, dec_t *out)
{
    for (;;) {
        int prec = dec_prec128(coef);
        if (prec - scale > DEC_MAXPREC - min_scale) return -1;
        if (scale < min_scale) {
            coef *= dec_pow10_128(min_scale - scale);
            scale = min_scale;
        } else if (scale >= prec && scale > DEC_MAXSCALE) {
            coef = dec_rsh_half_even128(coef, scale - DEC_MAXSCALE);
            scale = DEC_MAXSCALE;
        } else if (prec > scale && prec > DEC_MAXPREC) {
            int drop = prec - DEC_MAXPREC;
            coef = dec_rsh_half_even128(coef, drop);
            scale -= drop;
        }
        if (coef > (u128)DEC_MAXCOEF) continue; /* 99..9 rounded up to 20 digits */
        break;
    }
    out->coef = (uint64_t)coef; out->scale = (int8_t)scale; out->neg = (uint8_t)(neg ? 1 : 0);
    return 0;
}

static inline dec_t dec_from_i64(int64_t v, int scale)
{
    dec_t d;
    d.neg = v < 0;
    d.coef = v < 0 ? (uint64_t)(-(v + 1)) + 1 : (uint64_t)v;
    d.scale = (int8_t)scale;
    return d;
}

/* Add (AddExact(e,0)): scale = max(scales); exact when it fits in 19 digits,
 * else rounded half-even to 19 digits. */
static inline int dec_add(dec_t a, dec_t b, dec_t *out)
{
    int scale = a.scale > b.scale ? a.scale : b.scale;
    u128 ca = (u128)a.coef * dec_pow10_128(scale - a.scale);
    u128 cb = (u128)b.coef * dec_pow10_128(scale - b.scale);
    u128 c; int neg;
    if (a.neg == b.neg) { c = ca + cb; neg = a.neg; }
    else if (ca >= cb) { c = ca - cb; neg = a.neg; }
    else { c = cb - ca; neg = b.neg; }
    if (c == 0) neg = 0; /* govalues: -0 does not exist after arithmetic unless both neg */
    return dec_from_u128(neg, c, scale, 0, out);
}

static inline int dec_sub(dec_t a, dec_t b, dec_t *out)
{
    b.neg = !b.neg;
    return dec_add(a, b, out);
}

/* Mul (MulExact(e,0)): scale = sum of scales, rounded back into range. */
static inline int dec_mul(dec_t a, dec_t b, dec_t *out)
{
    u128 c = (u128)a.coef * (u128)b.coef;
    int neg = a.neg != b.neg;
    return dec_from_u128(neg, c, a.scale + b.scale, 0, out);
}

static inline int dec_cmp(dec_t a, dec_t b)
{
    if (a.coef == 0 && b.coef == 0) return 0;
    if (a.coef == 0) return b.neg ? 1 : -1;
    if (b.coef == 0) return a.neg ? -1 : 1;
    if (a.neg != b.neg) return a.neg ? -1 : 1;
    int scale = a.scale > b.scale ? a.scale : b.scale;
    u128 ca = (u128)a.coef * dec_pow10_128(scale - a.scale);
    u128 cb = (u128)b.coef * dec_pow10_128(scale - b.scale);
    int r = ca < cb ? -1 : (ca > cb ? 1 : 0);
    return a.neg ? -r : r;
}

/* Trim(scale): strip trailing zeros down to at least `scale`. */
static inline dec_t dec_trim(dec_t d, int scale)
{
    while (d.scale > scale && d.coef % 10 == 0) { d.coef /= 10; d.scale--; }
    return d;
}

/* Quo (QuoExact(e,0)): exact if the division terminates within 19 digits,
 * otherwise a 38-digit truncated quotient rounded half-even to 19 digits;
 * then trailing zeros trimmed down to max(0, sa-sb). returns -1 on error. */
static inline int dec_quo(dec_t a, dec_t b, dec_t *out)
{
    if (b.coef == 0) return -1;
    int neg = a.neg != b.neg;
    int pref = a.scale - b.scale; if (pref < 0) pref = 0;
    if (a.coef == 0) { *out = dec_make(0, 0, pref); return 0; }
    dec_t f; int ok = 0;
    {   /* quoFint */
        u128 dc = a.coef; u128 ec = b.coef;
        int scale = a.scale - b.scale;
        int shift = DEC_MAXPREC - dec_prec128(dc);
        if (shift > 0) { dc *= dec_pow10_128(shift); scale += shift; }
        int fits = 1;
        if (scale > DEC_MAXSCALE) {
            int s2 = scale - DEC_MAXSCALE;
            if (s2 > 19 || ec * dec_pow10_128(s2) > (u128)DEC_MAXCOEF) fits = 0;
            else { ec *= dec_pow10_128(s2); scale = DEC_MAXSCALE; }
        }
        if (fits && scale < 0) {
            int s2 = -scale;
            if (s2 > 19 || dc * dec_pow10_128(s2) > (u128)DEC_MAXCOEF) fits = 0;
            else { dc *= dec_pow10_128(s2); scale = 0; }
        }
        if (fits && dc % ec == 0) {
            if (dec_from_u128(neg, dc / ec, scale, 0, &f) == 0) ok = 1;
        }
    }
    if (!ok) {   /* quoBint */
        int scale = a.scale - b.scale;
        int shift = 2 * DEC_MAXPREC - dec_prec128(a.coef);
        u128 dc = (u128)a.coef * dec_pow10_128(shift);
        scale += shift;
        u128 q = dc / (u128)b.coef;
        /* newFromBint with possibly huge scale */
        if (dec_from_u128(neg, q, scale, 0, &f) != 0) return -1;
    }
    *out = dec_trim(f, pref);
    return 0;
}

/* Int64(scale): whole and (half-even rounded / zero padded) fraction. */
static inline int dec_int64(dec_t d, int scale, int64_t *whole, int64_t *frac)
{
    if (scale < 0 || scale > DEC_MAXSCALE) return 0;
    u128 x = d.coef; u128 y = DEC_POW10[d.scale];
    if (scale < d.scale) { x = dec_rsh_half_even128(x, d.scale - scale); y = DEC_POW10[scale]; }
    u128 q = x / y, r = x % y;
    if (scale > d.scale) {
        r *= dec_pow10_128(scale - d.scale);
        if (r > (u128)DEC_MAXCOEF) return 0;
    }
    if (d.neg) {
        if (q > ((u128)1 << 63) || r > ((u128)1 << 63)) return 0;
        *whole = (int64_t)(-(__int128)q); *frac = (int64_t)(-(__int128)r);
        return 1;
    }
    if (q > (u128)INT64_MAX || r > (u128)INT64_MAX) return 0;
    *whole = (int64_t)q; *frac = (int64_t)r;
    return 1;
}

/* NewFromInt64(whole, frac, scale): fraction's trailing zeros are stripped. */
static inline int dec_new_from_int64(int64_t whole, int64_t frac, int scale, dec_t *out)
{
    dec_t d = dec_from_i64(whole, 0);
    dec_t f = dec_from_i64(frac, scale);
    if (f.coef != 0) {
        if (d.coef != 0 && d.neg != f.neg) return -1;
        if (f.coef >= DEC_POW10[scale]) return -1; /* must be within (-1,1) */
        f = dec_trim(f, 0);
        return dec_add(d, f, out);
    }
    *out = d;
    return 0;
}

/* String(): sign, integer digits, '.', exactly `scale` fraction digits. */
static inline int dec_string(dec_t d, char *buf, size_t n)
{
    char digits[24]; int nd = 0; uint64_t c = d.coef;
    do { digits[nd++] = (char)('0' + c % 10); c /= 10; } while (c);
    while (nd <= d.scale) digits[nd++] = '0';
    size_t k = 0;
    if (d.neg && k < n) buf[k++] = '-';
    for (int i = nd - 1; i >= 0; i--) {
        if (i == d.scale - 1 && k < n) buf[k++] = '.';
        if (k < n) buf[k++] = digits[i];
    }
    if (k >= n) k = n - 1;
    buf[k] = 0;
    return (int)k;
}

/* Float64(): strconv.ParseFloat(d.String(), 64) -- correctly rounded. */
static inline double dec_float64(dec_t d)
{
    char buf[48];
    dec_string(d, buf, sizeof buf);
    return strtod(buf, NULL);
}

#endif
