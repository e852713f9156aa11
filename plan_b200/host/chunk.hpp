// chunk.hpp -- C++ host-side mirror of the reference's pkg/chunk + pkg/common types, ABOVE the C ABI.
// (The reference host language is Go; there is no Go toolchain in this image or on the GPU box, so
// the host side of the drop-in is written in C++ with the reference's names and semantics.)
//   LType / LTID_*    /root/reference/pkg/common/ltype.go
//   Vector, Chunk     /root/reference/pkg/chunk/vector.go:15-22, chunk.go:16-20
//   Value::String     /root/reference/pkg/chunk/value.go:26-70
//   Vector::GetValue  /root/reference/pkg/chunk/vector.go:76-186
#pragma once
#include <stdint.h>
#include <string.h>

#include <charconv>
#include <string>
#include <vector>

#include "../../include/plangpu.h"
#include "../../include/plangpu_desc.h"

namespace planhost {

typedef unsigned __int128 u128;
constexpr int DefaultVectorSize = 2048;   // pkg/util/util.go:124

struct LType {
    int Id = 0, Width = 0, Scale = 0;
    static LType Boolean() { return {PG_LT_BOOLEAN, 0, 0}; }
    static LType Integer() { return {PG_LT_INTEGER, 0, 0}; }
    static LType Bigint() { return {PG_LT_BIGINT, 0, 0}; }
    static LType Date() { return {PG_LT_DATE, 0, 0}; }
    static LType Decimal(int w, int s) { return {PG_LT_DECIMAL, w, s}; }
    static LType Float() { return {PG_LT_FLOAT, 0, 0}; }
    static LType Double() { return {PG_LT_DOUBLE, 0, 0}; }
    static LType Varchar() { return {PG_LT_VARCHAR, 0, 0}; }
    static LType Hugeint() { return {PG_LT_HUGEINT, 0, 0}; }
};

// ---- govalues formatting contract on 128-bit integers -------------------------------------
inline u128 pow10_128(int n) { u128 r = 1; while (n-- > 0) r *= 10; return r; }
inline u128 div_half_even(u128 x, u128 p) { u128 q = x / p, r = x % p; if (2 * r > p || (2 * r == p && (q & 1))) q++; return q; }

inline std::string u128_to_string(u128 v)
{
    if (v == 0) return "0";
    std::string s;
    while (v > 0) { s.insert(s.begin(), (char)('0' + (int)(v % 10))); v /= 10; }
    return s;
}

inline std::string decimal_string(u128 coef, int scale, bool neg)
{
    std::string d = u128_to_string(coef);
    while ((int)d.size() <= scale) d.insert(d.begin(), '0');
    if (scale > 0) d.insert(d.size() - (size_t)scale, ".");
    return (neg && coef != 0 ? "-" : "") + d;
}

// Decimal.Int64(scale) followed by NewFromInt64(...).String(): round half-even to the TYPE scale,
// then strip the fraction's trailing zeros (vector.go:121-137, value.go:37-46)
inline std::string decimal_value_string(uint64_t coef, int scale, bool neg, int type_scale)
{
    u128 x = coef;
    int s = scale;
    if (type_scale < scale) { x = div_half_even(x, pow10_128(scale - type_scale)); s = type_scale; }
    u128 p = pow10_128(s), whole = x / p, frac = x % p;
    while (s > 0 && frac % 10 == 0) { frac /= 10; s--; }
    return decimal_string(whole * pow10_128(s) + frac, s, neg);
}

inline std::string go_float_string(double x)   // fmt %v = strconv 'g', shortest: %e form when the decimal exponent is < -4 or >= 6
{
    if (x != x) return "NaN";
    if (x == __builtin_inf()) return "+Inf";
    if (x == -__builtin_inf()) return "-Inf";
    if (x == 0) return __builtin_signbit(x) ? "-0" : "0";
    char buf[64];
    auto r = std::to_chars(buf, buf + sizeof buf, x, std::chars_format::scientific);     // shortest round-trip digits, d.ddde[+-]XX
    std::string s(buf, r.ptr);
    const size_t e = s.find('e');
    const int exp = std::stoi(s.substr(e + 1));
    if (exp < -4 || exp >= 6) return s;                 // to_chars already pads the exponent to two digits like Go
    r = std::to_chars(buf, buf + sizeof buf, x, std::chars_format::fixed);               // shortest digits, no exponent
    return std::string(buf, r.ptr);
}

inline std::string date_string(int32_t days)
{
    int64_t z = (int64_t)days + 719468;
    int64_t era = (z >= 0 ? z : z - 146096) / 146097, doe = z - era * 146097;
    int64_t yoe = (doe - doe / 1460 + doe / 36524 - doe / 146096) / 365, y = yoe + era * 400;
    int64_t doy = doe - (365 * yoe + yoe / 4 - yoe / 100), mp = (5 * doy + 2) / 153;
    int d = (int)(doy - (153 * mp + 2) / 5 + 1), m = (int)(mp < 10 ? mp + 3 : mp - 9);
    if (m <= 2) y++;
    char buf[32];
    snprintf(buf, sizeof buf, "%04d-%02d-%02d", (int)y, m, d);
    return buf;
}

struct Vector {
    LType Typ;
    int NativeType = 0;                 // pg_type of Data
    std::vector<uint8_t> Data;          // native encoding
    std::vector<std::string> Dict;      // PG_T_DICT8
    std::vector<std::string> Strings;   // PG_T_VARCHAR (copied out of the result: pg_string rows die with it)

    std::string ValueString(int64_t i) const   // GetValue(i).String()
    {
        const uint8_t *p = Data.data();
        switch (NativeType) {
        case PG_T_INT32: return std::to_string(((const int32_t *)p)[i]);
        case PG_T_INT64: return std::to_string((long long)((const int64_t *)p)[i]);
        case PG_T_DATE32: return date_string(((const int32_t *)p)[i]);
        case PG_T_DECIMAL64: {
            int64_t v = ((const int64_t *)p)[i];
            return decimal_value_string(v < 0 ? (uint64_t)(-(v + 1)) + 1 : (uint64_t)v, Typ.Scale, v < 0, Typ.Scale);
        }
        case PG_T_CHAR1: return std::string(1, (char)p[i]);
        case PG_T_DICT8: return p[i] < Dict.size() ? Dict[p[i]] : std::string("?");
        case PG_T_VARCHAR: return Strings[(size_t)i];
        case PG_T_FLOAT64: return go_float_string(((const double *)p)[i]);
        case PG_T_HUGEINT: {
            const pg_hugeint &h = ((const pg_hugeint *)p)[i];
            __int128 v = (__int128)(((u128)(uint64_t)h.upper << 64) | h.lower);
            return v < 0 ? "-" + u128_to_string((u128)(-v)) : u128_to_string((u128)v);
        }
        case PG_T_DECIMAL128: {
            const pg_decimal &d = ((const pg_decimal *)p)[i];
            return decimal_value_string(d.coef, d.scale, d.neg != 0, Typ.Scale);
        }
        default: return "usp";
        }
    }
};

struct Chunk {
    std::vector<Vector> Data;
    int64_t Count = 0;
    int64_t Card() const { return Count; }
    // Chunk.SaveToFile (chunk.go:196-220): tab separated, newline terminated
    void SaveToFile(FILE *f) const
    {
        for (int64_t r = 0; r < Count; r++) {
            for (size_t c = 0; c < Data.size(); c++) {
                if (c) fputc('\t', f);
                fputs(Data[c].ValueString(r).c_str(), f);
            }
            fputc('\n', f);
        }
    }
};

}  // namespace planhost
