// rowvm_compile.hpp -- host side of rowvm.cuh: Expr trees (plan_ir.hpp) -> postfix programs.
// Shared by the row-emitting pipelines (rows.cu) and the expression-driven scan aggregate (scanagg.cu, vm_scanagg_kernel).
#pragma once
#include <algorithm>
#include <functional>
#include <string>
#include <vector>

#include "pipeline.hpp"
#include "rowvm.cuh"

namespace pg {

struct Src { int side = 0, col = -1; bool mark = false; };
typedef std::function<bool(int, Src *)> Resolver;

struct RvCompiler {
    RvCode code{};
    int ncode = 0, ncols = 0, nmasks = 0;
    const pg_table *tables[2] = {nullptr, nullptr};      // side 0 (probe / scanned) and side 1 (build)
    std::string why;

    const pg_table *tab(int side) const { return tables[side]; }

    bool fail(const std::string &s) { why = s; return false; }
    // depth of the evaluation stack where the next instruction runs: the interpreter's stack is RV_MAXSTACK values in
    // local memory and is never bounds-checked on the device, so a program that needs more is refused here
    int cur_sp = 0;
    bool emit(int op, int a = 0, int b = 0, i64 imm = 0)
    {
        if (ncode >= RV_MAXCODE) return fail("expression program too long");
        switch (op) {
        case RV_COL: case RV_CONST: case RV_NULL: case RV_MARK: cur_sp += 1; break;
        case RV_ADD: case RV_SUB: case RV_MUL: case RV_DIV: case RV_CMP: case RV_AND: case RV_OR: case RV_JZ: cur_sp -= 1; break;
        default: break;      // NOT, INSET, YEAR, TOF32, TODEC, JMP, PRE keep the depth
        }
        if (cur_sp > RV_MAXSTACK) return fail("expression needs a deeper evaluation stack than the interpreter has");
        if (cur_sp < 0) return fail("internal: evaluation stack underflow while compiling");
        code.ins[ncode++] = RvIns{op, a, b, 0, imm};
        return true;
    }
    int col_slot(int side, int col)
    {
        const pg_table *t = tab(side);
        const Column &c = t->cols[(size_t)col];
        for (int i = 0; i < ncols; i++)
            if (code.cols[i].side == side && code.cols[i].col.p == c.d_data) return i;
        if (ncols >= RV_MAXCOL) return -1;
        RvCol rc;
        rc.col.p = c.d_data;
        rc.col.width = c.phys_width();
        rc.col.base = c.base;
        rc.col.valid = c.has_nulls ? c.d_valid : nullptr;
        rc.side = side;
        rc.scale = c.type == PG_T_DECIMAL64 ? c.scale : 0;
        code.cols[ncols] = rc;
        return ncols++;
    }
    static int kind_of_column(const Column &c)
    {
        switch (c.type) {
        case PG_T_INT32: case PG_T_INT64: case PG_T_DATE32: return RVK_INT;
        case PG_T_DECIMAL64: return RVK_DEC;
        case PG_T_CHAR1: case PG_T_DICT8: return RVK_CODE;
        default: return 0;
        }
    }
    // code-set of a string predicate on a byte-coded column: fn applied to every code's string, once, on the host
    bool code_mask(const Column &c, int fn, const std::vector<std::string> &lits, int *mask_slot)
    {
        if (nmasks >= RV_MAXMASK) return fail("too many string predicates");
        unsigned *m = code.masks[nmasks];
        for (int i = 0; i < 8; i++) m[i] = 0;
        const int ncodes = c.type == PG_T_CHAR1 ? 256 : (int)c.dict.size();
        for (int k = 0; k < ncodes; k++) {
            const std::string s = c.type == PG_T_CHAR1 ? std::string(1, (char)k) : c.dict[(size_t)k];
            bool t = false;
            switch (fn) {
            case PG_FN_EQ: case PG_FN_IN: for (auto &l : lits) t = t || s == l; break;
            case PG_FN_NE: t = s != lits[0]; break;
            case PG_FN_LIKE: t = wildcard_match(lits[0].data(), lits[0].size(), s.data(), s.size()); break;
            case PG_FN_NOT_LIKE: t = !wildcard_match(lits[0].data(), lits[0].size(), s.data(), s.size()); break;
            default: return fail("string comparison other than =, <>, IN, LIKE");
            }
            if (t) m[k >> 5] |= 1u << (k & 31);
        }
        *mask_slot = nmasks++;
        return true;
    }
    // unify two numeric kinds (the binder casts INTEGER operands of a DECIMAL operation; be lenient about it)
    bool numeric_pair(int *ka, int *kb)
    {
        if (*ka == *kb) return true;
        if (*ka == RVK_INT && *kb == RVK_DEC) { *ka = RVK_DEC; return true; }      // INT values are DEC at scale 0 as they stand
        if (*ka == RVK_DEC && *kb == RVK_INT) { *kb = RVK_DEC; return true; }
        return fail("operand types differ (missing cast)");
    }

    // emits code that pushes the value of `e`; *kind = its static kind, *base = the byte-coded column it is (for string predicates)
    bool compile(const Expr &e, const Resolver &rs, int *kind, const Column **base = nullptr, int depth = 0)
    {
        if (base) *base = nullptr;
        if (depth == 0) cur_sp = 0;                  // a program starts on an empty stack
        if (depth > 24) return fail("expression too deep");
        switch (e.kind) {
        case PG_TK_COL: {
            Src s;
            if (e.side != 0 || !rs(e.idx, &s)) return fail("column reference out of scope");
            if (s.mark) { *kind = RVK_BOOL; return emit(RV_MARK); }
            const Column &c = tab(s.side)->cols[(size_t)s.col];
            *kind = kind_of_column(c);
            if (!*kind) return fail("column " + c.name + ": VARCHAR columns can be carried to the output but not computed on");
            const int cs = col_slot(s.side, s.col);
            if (cs < 0) return fail("too many columns referenced");
            if (base) *base = &c;
            return emit(RV_COL, cs);
        }
        case PG_TK_CONST:
            switch (e.ltype) {
            case PG_LT_BOOLEAN: *kind = RVK_BOOL; return emit(RV_CONST, 0, 0, e.v0 != 0);
            case PG_LT_INTEGER: case PG_LT_BIGINT: case PG_LT_DATE: *kind = RVK_INT; return emit(RV_CONST, 0, 0, e.v0);
            case PG_LT_DECIMAL: *kind = RVK_DEC; return emit(RV_CONST, 0, e.scale, e.v0);
            case PG_LT_FLOAT: case PG_LT_DOUBLE: {      // float constants are float32(val) (chunk/vector.go:205-207)
                double d;
                memcpy(&d, &e.v0, 8);
                const float f = (float)d;
                unsigned bits;
                memcpy(&bits, &f, 4);
                *kind = RVK_F32;
                return emit(RV_CONST, 0, 0, (i64)bits);
            }
            default: return fail("constant type");
            }
        case PG_TK_FUNC: break;
        default: return fail("string literal outside a comparison with a dictionary / char column");
        }
        const int fn = e.fn;
        const size_t na = e.args.size();
        if (fn == PG_FN_ADD || fn == PG_FN_SUB || fn == PG_FN_MUL || fn == PG_FN_DIV) {
            if (na != 2) return fail("arithmetic arity");
            int ka, kb;
            if (!compile(e.args[0], rs, &ka, nullptr, depth + 1) || !compile(e.args[1], rs, &kb, nullptr, depth + 1)) return false;
            if (!numeric_pair(&ka, &kb)) return false;
            if (ka != RVK_INT && ka != RVK_DEC && ka != RVK_F32) return fail("arithmetic on a non-numeric value");
            if (ka == RVK_INT && fn == PG_FN_DIV) return fail("integer division");
            *kind = ka;
            const int op = fn == PG_FN_ADD ? RV_ADD : fn == PG_FN_SUB ? RV_SUB : fn == PG_FN_MUL ? RV_MUL : RV_DIV;
            return emit(op, ka, (ka == RVK_INT && e.ltype == PG_LT_INTEGER) ? 32 : 0);
        }
        if (is_cmp(fn) || fn == PG_FN_LIKE || fn == PG_FN_NOT_LIKE) {
            if (na != 2) return fail("comparison arity");
            *kind = RVK_BOOL;
            const Expr *l = &e.args[0], *r = &e.args[1];
            int f2 = fn;
            if (l->kind == PG_TK_STR) { std::swap(l, r); f2 = flip_cmp(fn); }
            if (r->kind == PG_TK_STR) {
                int k;
                const Column *c;
                if (!compile(*strip_value_preserving_casts(l), rs, &k, &c, depth + 1)) return false;
                if (k != RVK_CODE || !c) return fail("string comparison on a column that is not dictionary / char coded");
                int ms;
                if (!code_mask(*c, f2, {r->str}, &ms)) return false;
                return emit(RV_INSET, ms);
            }
            if (!is_cmp(fn)) return fail("LIKE needs a string pattern");
            int ka, kb;
            if (!compile(*l, rs, &ka, nullptr, depth + 1) || !compile(*r, rs, &kb, nullptr, depth + 1)) return false;
            if (!numeric_pair(&ka, &kb)) return false;
            return emit(RV_CMP, f2, ka);
        }
        if (fn == PG_FN_IN) {
            if (na < 2) return fail("IN arity");
            *kind = RVK_BOOL;
            if (e.args[1].kind == PG_TK_STR) {
                std::vector<std::string> lits;
                for (size_t i = 1; i < na; i++) { if (e.args[i].kind != PG_TK_STR) return fail("IN list mixes types"); lits.push_back(e.args[i].str); }
                int k;
                const Column *c;
                if (!compile(*strip_value_preserving_casts(&e.args[0]), rs, &k, &c, depth + 1)) return false;
                if (k != RVK_CODE || !c) return fail("string IN on a column that is not dictionary / char coded");
                int ms;
                if (!code_mask(*c, PG_FN_IN, lits, &ms)) return false;
                return emit(RV_INSET, ms);
            }
            for (size_t i = 1; i < na; i++) {          // x = c1 OR x = c2 ... (inInt32Op, function_operator_boolean.go:393-504)
                int ka, kb;
                if (!compile(e.args[0], rs, &ka, nullptr, depth + 1) || !compile(e.args[i], rs, &kb, nullptr, depth + 1)) return false;
                if (!numeric_pair(&ka, &kb) || !emit(RV_CMP, PG_FN_EQ, ka)) return false;
                if (i > 1 && !emit(RV_OR)) return false;
            }
            return true;
        }
        if (fn == PG_FN_AND || fn == PG_FN_OR) {
            if (na < 2) return fail("AND / OR arity");
            *kind = RVK_BOOL;
            for (size_t i = 0; i < na; i++) {
                int k;
                if (!compile(e.args[i], rs, &k, nullptr, depth + 1)) return false;
                if (k != RVK_BOOL) return fail("AND / OR of a non-boolean");
                if (i > 0 && !emit(fn == PG_FN_AND ? RV_AND : RV_OR)) return false;
            }
            return true;
        }
        if (fn == PG_FN_NOT) {
            int k;
            if (na != 1 || !compile(e.args[0], rs, &k, nullptr, depth + 1) || k != RVK_BOOL) return why.empty() ? fail("NOT of a non-boolean") : false;
            *kind = RVK_BOOL;
            return emit(RV_NOT);
        }
        if (fn == PG_FN_EXTRACT) {
            if (na != 2 || e.args[0].kind != PG_TK_STR || e.args[0].str != "year") return fail("only EXTRACT(year ...) is off-loaded");
            int k;
            if (!compile(e.args[1], rs, &k, nullptr, depth + 1)) return false;
            if (k != RVK_INT || e.args[1].ltype != PG_LT_DATE) return fail("EXTRACT(year) of a non-date");
            *kind = RVK_INT;
            return emit(RV_YEAR);
        }
        if (fn == PG_FN_CAST) {
            if (na != 1) return fail("cast arity");
            int k;
            const Column *c;
            if (!compile(e.args[0], rs, &k, &c, depth + 1)) return false;
            if (base) *base = c;
            switch (e.ltype) {
            case PG_LT_DECIMAL:
                if (k == RVK_DEC) {
                    if (e.scale < e.args[0].scale) return fail("DECIMAL cast that drops fractional digits");
                    *kind = RVK_DEC;
                    return true;                 // value preserving (tryCastDecimalToDecimal, function_cast.go:380-404)
                }
                if (k == RVK_INT) { *kind = RVK_DEC; return emit(RV_TODEC); }      // tryCastInt32ToDecimal (:337-347)
                return fail("cast to DECIMAL from this type");
            case PG_LT_FLOAT: case PG_LT_DOUBLE:
                if (k == RVK_F32) { *kind = RVK_F32; return true; }
                if (k != RVK_DEC && k != RVK_INT) return fail("cast to FLOAT from this type");
                *kind = RVK_F32;
                return emit(RV_TOF32, k);
            case PG_LT_BIGINT: case PG_LT_HUGEINT: case PG_LT_INTEGER:
                if (k != RVK_INT || (e.ltype == PG_LT_INTEGER && e.args[0].ltype != PG_LT_INTEGER)) return fail("narrowing integer cast");
                *kind = RVK_INT;
                return true;
            case PG_LT_DATE: case PG_LT_VARCHAR: case PG_LT_BOOLEAN:
                if (e.ltype != e.args[0].ltype) return fail("cast between unrelated types");
                *kind = k;
                return true;
            default: return fail("cast target type");
            }
        }
        if (fn == PG_FN_CASE) {
            // children: [ELSE, WHEN1, THEN1, WHEN2, THEN2 ...] (executeCase, expr_exec.go:144-246)
            if (na < 3 || (na & 1) == 0) return fail("CASE arity");
            std::vector<int> to_end;
            int rk = 0;
            const int sp0 = cur_sp;                   // every path through the CASE leaves exactly one value above this depth
            auto branch = [&](const Expr &x) -> bool {
                int k;
                if (x.kind == PG_TK_CONST && x.ltype == 0) { k = rk; if (!emit(RV_NULL)) return false; }       // typeless NULL constant
                else if (!compile(x, rs, &k, nullptr, depth + 1)) return false;
                if (rk == 0) rk = k;
                else if (k != rk) {
                    if ((rk == RVK_DEC && k == RVK_INT) || (rk == RVK_INT && k == RVK_DEC)) rk = RVK_DEC;
                    else return fail("CASE branches of different types");
                }
                return true;
            };
            for (size_t i = 1; i + 1 < na; i += 2) {
                int k;
                if (!compile(e.args[i], rs, &k, nullptr, depth + 1)) return false;
                if (k != RVK_BOOL) return fail("CASE WHEN is not a boolean");
                const int jz = ncode;
                if (!emit(RV_JZ)) return false;
                if (!branch(e.args[i + 1])) return false;
                to_end.push_back(ncode);
                if (!emit(RV_JMP)) return false;
                code.ins[jz].imm = ncode;
                cur_sp = sp0;                         // the next WHEN starts where this one did (its THEN value is on another path)
            }
            if (!branch(e.args[0])) return false;
            for (int j : to_end) code.ins[j].imm = ncode;
            *kind = rk;
            return true;
        }
        return fail("function " + std::to_string(fn) + " is not off-loaded in row expressions");
    }

    // Upper bound of the run-time scale of a numeric expression (an aggregate accumulates at ONE scale and rescales
    // every value to it); -1 when there is none (a quotient keeps up to 19 significant digits).
    int scale_bound(const Expr &e, const Resolver &rs) const
    {
        switch (e.kind) {
        case PG_TK_COL: {
            Src s;
            if (e.side != 0 || !rs(e.idx, &s) || s.mark) return 0;
            const Column &c = tab(s.side)->cols[(size_t)s.col];
            return c.type == PG_T_DECIMAL64 ? c.scale : 0;
        }
        case PG_TK_CONST: return e.ltype == PG_LT_DECIMAL ? e.scale : 0;
        case PG_TK_FUNC: break;
        default: return 0;
        }
        auto of = [&](size_t i) { return scale_bound(e.args[i], rs); };
        switch (e.fn) {
        case PG_FN_ADD: case PG_FN_SUB: { if (e.args.size() != 2) return -1; const int a = of(0), b = of(1); return a < 0 || b < 0 ? -1 : std::max(a, b); }
        case PG_FN_MUL: { if (e.args.size() != 2) return -1; const int a = of(0), b = of(1); return a < 0 || b < 0 ? -1 : a + b; }
        case PG_FN_DIV: return -1;
        case PG_FN_CAST: return e.args.size() == 1 ? of(0) : -1;
        case PG_FN_CASE: {
            int m = e.args.empty() ? 0 : of(0);
            for (size_t i = 2; i < e.args.size() && m >= 0; i += 2) { const int b = of(i); m = b < 0 ? -1 : std::max(m, b); }
            return m;
        }
        default: return 0;
        }
    }

    // `column <cmp> constant` on an integer / date / decimal column, or a string predicate on a dictionary / char column:
    // lowered to an inline test (RvPre) instead of interpreter instructions.  false = not of that form (no error).
    int npre = 0;
    bool lower_pre(const Expr &f, const Resolver &rs, RvPre *out)
    {
        if (f.kind != PG_TK_FUNC) return false;
        auto column_of = [&](const Expr &x, Src *src) -> const Column * {
            const Expr *e = strip_value_preserving_casts(&x);
            if (e->kind != PG_TK_COL || e->side != 0 || !rs(e->idx, src) || src->mark) return nullptr;
            return &tab(src->side)->cols[(size_t)src->col];
        };
        const int fn = f.fn;
        if ((is_cmp(fn) || fn == PG_FN_LIKE || fn == PG_FN_NOT_LIKE || fn == PG_FN_IN) && f.args.size() >= 2) {
            // string predicates on byte-coded columns -> a code set
            bool all_str = true;
            for (size_t i = 1; i < f.args.size(); i++) all_str = all_str && f.args[i].kind == PG_TK_STR;
            if (all_str && (fn == PG_FN_IN || f.args.size() == 2) && (fn == PG_FN_EQ || fn == PG_FN_NE || fn == PG_FN_LIKE || fn == PG_FN_NOT_LIKE || fn == PG_FN_IN)) {
                Src src;
                const Column *c = column_of(f.args[0], &src);
                if (!c || !is_byte_family(c->type)) return false;
                std::vector<std::string> lits;
                for (size_t i = 1; i < f.args.size(); i++) lits.push_back(f.args[i].str);
                int ms;
                const int cs = col_slot(src.side, src.col);
                if (cs < 0 || !code_mask(*c, fn, lits, &ms)) { why.clear(); return false; }
                *out = RvPre{cs, ms, 0, 0};
                return true;
            }
        }
        if (!is_cmp(fn) || fn == PG_FN_NE || f.args.size() != 2) return false;
        const Expr *l = &f.args[0], *r = &f.args[1];
        int op = fn;
        Src src;
        const Column *c = column_of(*l, &src);
        if (!c) { c = column_of(*r, &src); std::swap(l, r); op = flip_cmp(fn); }
        if (!c || !is_int_family(c->type)) return false;
        const Expr *k = strip_value_preserving_casts(r);
        if (k->kind != PG_TK_CONST) return false;
        i128 v = k->v0;
        if (c->type == PG_T_DECIMAL64) {
            int ks;
            if (k->ltype == PG_LT_DECIMAL) ks = k->scale;
            else if (k->ltype == PG_LT_INTEGER || k->ltype == PG_LT_BIGINT) ks = 0;
            else return false;
            for (; ks < c->scale; ks++) v *= 10;
            for (; ks > c->scale; ks--) { if (v % 10 != 0) return false; v /= 10; }       // finer than the column: leave it to the interpreter
        } else if (!(k->ltype == PG_LT_INTEGER || k->ltype == PG_LT_BIGINT || k->ltype == PG_LT_DATE)) {
            return false;
        }
        if (v > (i128)INT64_MAX - 1 || v < (i128)INT64_MIN + 1) return false;
        const int cs = col_slot(src.side, src.col);
        if (cs < 0) return false;
        RvPre q{cs, -1, INT64_MIN, INT64_MAX};
        switch (op) {
        case PG_FN_EQ: q.lo = q.hi = (i64)v; break;
        case PG_FN_LT: q.hi = (i64)v - 1; break;
        case PG_FN_LE: q.hi = (i64)v; break;
        case PG_FN_GT: q.lo = (i64)v + 1; break;
        default: q.lo = (i64)v; break;      // GE
        }
        *out = q;
        return true;
    }

    // conjunction of filters -> one program [*p0, *p1).  A filter keeps a row only when EVERY conjunct is TRUE, so the
    // cheap conjuncts become inline tests (RV_PRE, lower_pre) and the interpreted rest leaves at the first conjunct
    // that is not TRUE (execSelectAnd narrows the selection the same way, expr_exec.go:444-480).
    bool compile_filters(const std::vector<const Expr *> &fs, const Resolver &rs, int *p0, int *p1)
    {
        *p0 = ncode;
        std::vector<const Expr *> rest;
        const int pre0 = npre;
        const bool no_pre = getenv("PG_VM_NO_PRE") && atoi(getenv("PG_VM_NO_PRE"));
        for (const Expr *f : fs) {
            RvPre q;
            if (!no_pre && npre < RV_MAXPRE && lower_pre(*f, rs, &q)) code.pre[npre++] = q;
            else rest.push_back(f);
        }
        if (npre > pre0 && !emit(RV_PRE, pre0, npre)) return false;
        std::vector<int> exits;
        for (size_t i = 0; i < rest.size(); i++) {
            int k;
            if (!compile(*rest[i], rs, &k)) return false;
            if (k != RVK_BOOL) return fail("filter is not a boolean expression");
            if (rest.size() > 1) { exits.push_back(ncode); if (!emit(RV_JZ)) return false; }
        }
        if (rest.size() > 1) {
            if (!emit(RV_CONST, 0, 0, 1)) return false;              // every conjunct was TRUE
            const int jmp = ncode;
            if (!emit(RV_JMP)) return false;
            for (int j : exits) code.ins[j].imm = ncode;
            cur_sp -= 1;                                             // (the FALSE path does not carry the TRUE path's value)
            if (!emit(RV_CONST, 0, 0, 0)) return false;              // some conjunct was FALSE or NULL
            code.ins[jmp].imm = ncode;
        }
        *p1 = ncode;
        cur_sp = 0;                                                  // a filter program leaves its one value to rv_true
        return true;
    }
};

}  // namespace pg
