// rowvm.cuh -- per-row expression evaluation for the ROW-EMITTING operators (rows.cu).
//
// The aggregate pipelines lower their expressions to ranges and affine products at plan time.  Filter, Project and
// Join operators whose parent is NOT an aggregate must produce ordinary rows from arbitrary expressions, so their
// Expr trees (/root/reference/pkg/compute/expr.go:49-60) are compiled once into a small postfix program and
// evaluated per row on the device with the reference's value semantics:
//   * ExprExec.execute / executeSelect (expr_exec.go:85-530): a NULL operand makes the result NULL; a filter keeps a
//     row only when its predicate is TRUE; AND / OR are evaluated on values as three-valued logic;
//   * executeCase (expr_exec.go:144-246): WHENs in order, a THEN branch is evaluated only for the rows its WHEN
//     selected (a division by zero in a branch not taken does not fire), ELSE for the rest -> conditional jumps;
//   * DECIMAL +, -, *, / (function_operator_binary.go:134-210) are govalues Add / Sub / Mul / Quo: a value is
//     (coefficient < 10^19, scale, sign); Add/Sub align to the larger scale, Mul adds the scales, Quo keeps 19
//     significant digits; whatever needs more than 19 digits is rounded half-even ONCE, and an integer part of more
//     than 19 digits is an error.  Values carry their scale at run time (hostdec.hpp restates the library contract);
//   * INTEGER +, - wrap in 32 bits like Go's int32 (binInt32Int32AddOp, function_operator_binary.go:143-146);
//   * cast(DECIMAL AS FLOAT) = float32(Float64(d)) (function_cast.go:349-354), FLOAT arithmetic and comparisons in
//     float32 (binFloat32*Op); comparisons between DECIMALs by value (common/decimal.go:12-40).
#pragma once
#include "hostdec.hpp"
#include "join.cuh"

namespace pg {

enum { RVK_BOOL = 1, RVK_INT = 2, RVK_DEC = 3, RVK_F32 = 4, RVK_CODE = 5 };        // static kind of a value
enum {
    RV_COL = 1,    // a: column slot                        push column value (NULL when the row id is negative: LEFT / MARK padding)
    RV_CONST,      // imm: value, b: scale (DEC) / kind
    RV_NULL,
    RV_MARK,       // push the MARK join's boolean: build row >= 0 true, -1 false, -2 NULL (NULL probe key)
    RV_ADD, RV_SUB, RV_MUL, RV_DIV,   // a: kind of the operation (RVK_INT / RVK_DEC / RVK_F32), b: 32 = wrap the INT result to int32
    RV_CMP,        // a: PG_FN_EQ..GE, b: kind compared
    RV_AND, RV_OR, RV_NOT,
    RV_INSET,      // a: mask slot (256-bit code set); pops a CODE, pushes BOOL
    RV_JZ,         // imm: target pc; pops a BOOL, jumps unless it is TRUE
    RV_JMP,        // imm: target pc
    RV_YEAR,       // DATE (days since 1970-01-01) -> calendar year
    RV_TOF32,      // a: kind of the operand (RVK_DEC / RVK_INT)
    RV_TODEC,      // INT -> DEC at scale 0
    RV_PRE,        // a, b: range of RvCode::pre -- only as the FIRST instruction of a filter program (rv_true): conjuncts of the
                   // form `column <cmp> constant` / `coded column in {codes}` are tested inline before the interpreter runs
};
enum { RV_ERR_OVERFLOW = 1, RV_ERR_DIVZERO = 2, RV_ERR_FLOAT = 3 };

constexpr int RV_MAXSTACK = 8, RV_MAXCODE = 384, RV_MAXCOL = 32, RV_MAXMASK = 16, RV_MAXOUT = 32, RV_MAXPRE = 24;

struct RvIns { int op, a, b, pad; i64 imm; };
struct RvCol { TypedCol col; int side; int scale; };     // side 0: the probe / scanned row, 1: the build row
struct RvVal { i128 v; int scale; int null; };

struct RvPre { int col, mask; i64 lo, hi; };             // mask >= 0: code-set test on a byte-coded column; else lo <= value <= hi
struct RvCode {
    RvIns ins[RV_MAXCODE];
    RvCol cols[RV_MAXCOL];
    unsigned masks[RV_MAXMASK][8];
    RvPre pre[RV_MAXPRE];
};

__device__ __forceinline__ float rv_as_f32(const RvVal &x) { return __int_as_float((int)(i64)x.v); }
__device__ __forceinline__ i128 rv_pow10(int n) { i128 r = 1; for (int i = 0; i < n; i++) r *= 10; return r; }

// bring an exact (value, scale) into govalues' 19-digit format; false = overflow of the integer part
__device__ __forceinline__ bool rv_fit(i128 v, int scale, RvVal *out)
{
    const bool neg = v < 0;
    const u128 mag = neg ? (u128)(-(v + 1)) + 1 : (u128)v;
    if (mag <= (u128)HD_MAXCOEF && scale >= 0 && scale <= HD_MAXPREC) { out->v = v; out->scale = scale; return true; }
    HDec d;
    if (!hd_normalise(neg, mag, scale, &d)) return false;
    out->v = d.neg ? -(i128)d.coef : (i128)d.coef;
    out->scale = d.scale;
    return true;
}

// evaluates program [pc0, pc1); row0 / row1 = row ids on side 0 / 1
static __device__ __noinline__ RvVal rv_eval(const RvCode &c, int pc0, int pc1, i64 row0, i64 row1, int *err)
{
    RvVal st[RV_MAXSTACK];
    int sp = 0;
    for (int pc = pc0; pc < pc1; pc++) {
        const RvIns in = c.ins[pc];
        switch (in.op) {
        case RV_COL: {
            const RvCol &rc = c.cols[in.a];
            const i64 row = rc.side ? row1 : row0;
            RvVal x;
            x.scale = rc.scale;
            x.null = row < 0 || !typed_valid(rc.col, row);
            x.v = x.null ? 0 : (i128)load_typed(rc.col, row);
            st[sp++] = x;
            break;
        }
        case RV_CONST: { RvVal x; x.v = (i128)in.imm; x.scale = in.b; x.null = 0; st[sp++] = x; break; }
        case RV_NULL: { RvVal x; x.v = 0; x.scale = 0; x.null = 1; st[sp++] = x; break; }
        case RV_MARK: { RvVal x; x.v = row1 >= 0 ? 1 : 0; x.scale = 0; x.null = row1 == -2; st[sp++] = x; break; }
        case RV_ADD: case RV_SUB: case RV_MUL: case RV_DIV: {
            const RvVal b = st[--sp], a = st[--sp];
            RvVal r;
            r.v = 0; r.scale = 0;
            r.null = a.null | b.null;
            if (!r.null) {
                if (in.a == RVK_F32) {
                    const float x = rv_as_f32(a), y = rv_as_f32(b);
                    const float z = in.op == RV_ADD ? x + y : in.op == RV_SUB ? x - y : in.op == RV_MUL ? x * y : x / y;
                    r.v = (i128)(i64)(unsigned)__float_as_int(z);
                } else if (in.a == RVK_INT) {
                    if (in.op == RV_DIV) { *err = RV_ERR_FLOAT; }
                    else {
                        const i128 z = in.op == RV_ADD ? a.v + b.v : in.op == RV_SUB ? a.v - b.v : a.v * b.v;
                        if (in.b == 32) r.v = (i128)(int32_t)(i64)z;
                        else { if (z > (i128)INT64_MAX || z < (i128)INT64_MIN) *err = RV_ERR_OVERFLOW; r.v = z; }
                    }
                } else if (in.op == RV_MUL) {
                    if (!rv_fit(a.v * b.v, a.scale + b.scale, &r)) *err = RV_ERR_OVERFLOW;
                } else if (in.op == RV_DIV) {
                    if (b.v == 0) *err = RV_ERR_DIVZERO;
                    else {
                        HDec x, y, q;
                        x.neg = a.v < 0; x.coef = (u64)(a.v < 0 ? -a.v : a.v); x.scale = a.scale;
                        y.neg = b.v < 0; y.coef = (u64)(b.v < 0 ? -b.v : b.v); y.scale = b.scale;
                        if (!hd_quo(x, y, &q)) *err = RV_ERR_OVERFLOW;
                        else { r.v = q.neg ? -(i128)q.coef : (i128)q.coef; r.scale = q.scale; }
                    }
                } else {
                    const int s = a.scale > b.scale ? a.scale : b.scale;
                    const i128 x = a.v * rv_pow10(s - a.scale), y = b.v * rv_pow10(s - b.scale);
                    if (!rv_fit(in.op == RV_ADD ? x + y : x - y, s, &r)) *err = RV_ERR_OVERFLOW;
                }
            }
            st[sp++] = r;
            break;
        }
        case RV_CMP: {
            const RvVal b = st[--sp], a = st[--sp];
            RvVal r;
            r.scale = 0; r.v = 0;
            r.null = a.null | b.null;
            if (!r.null) {
                int c3;
                if (in.b == RVK_F32) {
                    const float x = rv_as_f32(a), y = rv_as_f32(b);
                    c3 = x < y ? -1 : x > y ? 1 : x == y ? 0 : 2;       // 2: unordered (NaN) -- every comparison false but <>
                } else {
                    i128 x = a.v, y = b.v;
                    if (in.b == RVK_DEC) {
                        const int s = a.scale > b.scale ? a.scale : b.scale;
                        x *= rv_pow10(s - a.scale);
                        y *= rv_pow10(s - b.scale);
                    }
                    c3 = x < y ? -1 : x > y ? 1 : 0;
                }
                bool t;
                switch (in.a) {
                case PG_FN_EQ: t = c3 == 0; break;
                case PG_FN_NE: t = c3 != 0; break;
                case PG_FN_LT: t = c3 == -1; break;
                case PG_FN_LE: t = c3 == -1 || c3 == 0; break;
                case PG_FN_GT: t = c3 == 1; break;
                default: t = c3 == 1 || c3 == 0; break;
                }
                r.v = t ? 1 : 0;
            }
            st[sp++] = r;
            break;
        }
        case RV_AND: case RV_OR: {
            const RvVal b = st[--sp], a = st[--sp];
            RvVal r;
            r.scale = 0;
            const bool at = !a.null && a.v != 0, af = !a.null && a.v == 0, bt = !b.null && b.v != 0, bf = !b.null && b.v == 0;
            if (in.op == RV_AND) { r.v = (at && bt) ? 1 : 0; r.null = !(af || bf) && !(at && bt); }
            else { r.v = (at || bt) ? 1 : 0; r.null = !(at || bt) && !(af && bf); }
            st[sp++] = r;
            break;
        }
        case RV_NOT: { RvVal &a = st[sp - 1]; if (!a.null) a.v = a.v ? 0 : 1; break; }
        case RV_INSET: {
            RvVal &a = st[sp - 1];
            if (!a.null) { const unsigned code = (unsigned)(i64)a.v & 255u; a.v = (c.masks[in.a][code >> 5] >> (code & 31)) & 1u; }
            a.scale = 0;
            break;
        }
        case RV_JZ: { const RvVal a = st[--sp]; if (a.null || a.v == 0) pc = (int)in.imm - 1; break; }
        case RV_JMP: pc = (int)in.imm - 1; break;
        case RV_YEAR: { RvVal &a = st[sp - 1]; if (!a.null) a.v = (i128)year_of_days((i64)a.v); break; }
        case RV_TOF32: {
            RvVal &a = st[sp - 1];
            if (!a.null) {
                const i128 m = a.v < 0 ? -a.v : a.v;
                if (m >= ((i128)1 << 53) || a.scale > 19) *err = RV_ERR_FLOAT;       // outside the range where the IEEE division equals the parse
                double d = (double)(i64)a.v;
                if (in.a == RVK_DEC) { double p = 1.0; for (int i = 0; i < a.scale; i++) p *= 10.0; d = d / p; }
                a.v = (i128)(i64)(unsigned)__float_as_int((float)d);
            }
            a.scale = 0;
            break;
        }
        case RV_TODEC: st[sp - 1].scale = 0; break;
        default: break;
        }
    }
    return st[sp - 1];
}

// The inline pre-tests of a filter program (RV_PRE: conjuncts of the form `column <cmp> constant` / `coded column in {codes}`):
// most rows of a selective filter never reach the interpreter.  A NULL operand (or the NULL padding of a LEFT / MARK join,
// row < 0) makes the conjunct NULL, which is not TRUE.  true when the program has no pre-tests.
__device__ __forceinline__ bool rv_has_pre(const RvCode &c, int pc0, int pc1) { return pc1 > pc0 && c.ins[pc0].op == RV_PRE; }
__device__ __forceinline__ bool rv_pre(const RvCode &c, int pc0, int pc1, i64 row0, i64 row1)
{
    if (!rv_has_pre(c, pc0, pc1)) return true;
    const int p1 = c.ins[pc0].b;
    for (int i = c.ins[pc0].a; i < p1; i++) {
        const RvPre q = c.pre[i];
        const RvCol &rc = c.cols[q.col];
        const i64 row = rc.side ? row1 : row0;
        if (row < 0 || !typed_valid(rc.col, row)) return false;
        const i64 v = load_typed(rc.col, row);
        if (q.mask >= 0) { if (!((c.masks[q.mask][(v >> 5) & 7] >> (v & 31)) & 1u)) return false; }
        else if (v < q.lo || v > q.hi) return false;
    }
    return true;
}
// the same over descriptors a block copied into shared memory (pre-tests [i0, i1), column table, masks): one dependent
// global load per test -- the data -- instead of three
__device__ __forceinline__ bool rv_pre_smem(const RvPre *pre, int i0, int i1, const RvCol *cols, const unsigned (*masks)[8], i64 row0, i64 row1)
{
    for (int i = i0; i < i1; i++) {
        const RvPre &q = pre[i];
        const RvCol &rc = cols[q.col];
        const i64 row = rc.side ? row1 : row0;
        if (row < 0 || !typed_valid(rc.col, row)) return false;
        const i64 v = load_typed(rc.col, row);
        if (q.mask >= 0) { if (!((masks[q.mask][(v >> 5) & 7] >> (v & 31)) & 1u)) return false; }
        else if (v < q.lo || v > q.hi) return false;
    }
    return true;
}
// the interpreted rest of a filter program (everything after its RV_PRE instruction)
__device__ __forceinline__ bool rv_post(const RvCode &c, int pc0, int pc1, i64 row0, i64 row1, int *err)
{
    if (rv_has_pre(c, pc0, pc1)) pc0++;
    if (pc1 <= pc0) return true;
    const RvVal v = rv_eval(c, pc0, pc1, row0, row1, err);
    return !v.null && v.v != 0;
}
__device__ __forceinline__ bool rv_true(const RvCode &c, int pc0, int pc1, i64 row0, i64 row1, int *err)
{
    return rv_pre(c, pc0, pc1, row0, row1) && rv_post(c, pc0, pc1, row0, row1, err);
}

}  // namespace pg
