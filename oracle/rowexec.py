"""TEST INFRASTRUCTURE (never imported by the product): a tree-walking CPU restatement of the reference's
ROW-EMITTING operators, used as the oracle for plan_b200/csrc/rows.cu.

It evaluates the same PhysicalOperator / Expr trees the GPU path is given (plan_b200.compute mirrors
/root/reference/pkg/compute/builder_physical_operator.go:49-66 and expr.go:49-60) over host numpy tables, row by row:

  Scan + pushed-down filters   executor_scan.go:225-241
  Filter                       executor_filter.go:12-118  (a row passes when every filter is TRUE; NULL is not TRUE)
  Project                      executor_project.go:24-82
  Join INNER / LEFT / SEMI / ANTI / MARK   executor_join.go:62-123, join_scan.go:67-299; NULL keys never match
                               (join_table.go:152-195); MARK: NULL mark for a NULL probe key (join_scan.go:132-165)
  CASE                         expr_exec.go:144-246 (THEN evaluated only where its WHEN is true)
  DECIMAL + - * /              function_operator_binary.go:134-210 -> govalues Add / Sub / Mul / Quo, through the C
                               restatement of the library in oracle/decimal.h (orc_dec_*)
  INTEGER + -                  int32 wrap-around (binInt32Int32AddOp, :143-146)
  cast(DECIMAL AS FLOAT)       float32(Float64(d)) (function_cast.go:349-354); FLOAT comparisons in float32

Pinning: the reference holds no golden vector for these operators in isolation (its Go tests assert no numeric results), so
this module is pinned indirectly -- its decimal arithmetic is oracle/decimal.h (checked against the reference's golden Q1 / Q6
files), CASE / OR / IN / LIKE inside aggregates over a join are pinned by the reference's golden q12.txt / q14.txt through
oracle.q12 / oracle.q14 and the GPU test that reproduces both files; MARK joins by q4.txt, multi-way join stacks by q5.txt / q7.txt /
q8.txt, the LEFT join (NULL padding, count(x) = NULL over it) by q13.txt (tests/test_oracle_golden.py runs those plans through this
module); a DECIMAL quotient by q8.txt.  DECIMAL division evaluated per row is PARITY UNPINNED against the reference.

Values: None (NULL) | bool | int | Dec(coef, scale, neg) | np.float32 | str.
"""
import ctypes as C
import datetime

import numpy as np

from . import oracle as O

# numbering shared with plan_b200.compute / plangpu_desc.h (restated here: the oracle does not import the product)
POT_Scan, POT_Filter, POT_Join, POT_Agg, POT_Project = 1, 2, 3, 4, 5
ET_Column, ET_Func, ET_Const = 0, 5, 7
JOIN_INNER, JOIN_SEMI, JOIN_ANTI, JOIN_MARK, JOIN_LEFT, JOIN_ANTI_MARK = 1, 2, 3, 4, 5, 6
LT_BOOLEAN, LT_INTEGER, LT_BIGINT, LT_DATE, LT_DECIMAL, LT_FLOAT, LT_DOUBLE, LT_VARCHAR, LT_HUGEINT = range(1, 10)
T_INT32, T_INT64, T_DATE32, T_DECIMAL64, T_CHAR1, T_DICT8, T_FLOAT64, T_HUGEINT, T_DECIMAL128, T_VARCHAR = range(1, 11)


class Dec:
    __slots__ = ("coef", "scale", "neg")

    def __init__(self, coef, scale, neg=False):
        self.coef, self.scale, self.neg = int(coef), int(scale), bool(neg) and coef != 0

    @staticmethod
    def from_int(v, scale=0):
        return Dec(abs(int(v)), scale, v < 0)

    def signed(self):
        return -self.coef if self.neg else self.coef

    def __repr__(self):
        return "Dec(%s%d e-%d)" % ("-" if self.neg else "", self.coef, self.scale)


class DecimalError(Exception):
    """the reference panics (decimal overflow / division by zero) and the query fails"""


def _dec_op(name, a, b):
    oc, os_, on = C.c_uint64(), C.c_int(), C.c_int()
    rc = getattr(O.lib(), name)(a.coef, a.scale, int(a.neg), b.coef, b.scale, int(b.neg), C.byref(oc), C.byref(os_), C.byref(on))
    if rc != 0:
        raise DecimalError(name)
    return Dec(oc.value, os_.value, bool(on.value))


def dec_add(a, b): return _dec_op("orc_dec_add", a, b)
def dec_sub(a, b): return _dec_op("orc_dec_add", a, Dec(b.coef, b.scale, not b.neg))
def dec_mul(a, b): return _dec_op("orc_dec_mul", a, b)


def dec_quo(a, b):
    if b.coef == 0:
        raise DecimalError("division by zero")
    return _dec_op("orc_dec_quo", a, b)


def dec_cmp(a, b):
    s = max(a.scale, b.scale)
    x, y = a.signed() * 10 ** (s - a.scale), b.signed() * 10 ** (s - b.scale)
    return (x > y) - (x < y)


def dec_to_f32(a):
    return np.float32(O.lib().orc_dec_float64(a.coef, a.scale, int(a.neg)))


def _num_pair(a, b):
    if isinstance(a, Dec) and isinstance(b, int) and not isinstance(b, bool):
        return a, Dec.from_int(b)
    if isinstance(b, Dec) and isinstance(a, int) and not isinstance(a, bool):
        return Dec.from_int(a), b
    return a, b


def _wrap32(v):
    return ((int(v) + (1 << 31)) & 0xFFFFFFFF) - (1 << 31)


def eval_expr(e, row):
    """row: list of values of the operator's input (child outputs)."""
    if e.Typ == ET_Column:
        return row[e.ColRef[1]]
    if e.Typ == ET_Const:
        v, t = e.ConstValue, e.DataTyp.Id
        if v is None:
            return None
        if t == LT_DECIMAL:
            return Dec.from_int(v, e.DataTyp.Scale)
        if t in (LT_FLOAT, LT_DOUBLE):
            return np.float32(v)
        if t == LT_BOOLEAN:
            return bool(v)
        if t == LT_VARCHAR:
            return v
        return int(v)
    fn, ch = e.FunImpl, e.Children
    if fn == "case":
        for i in range(1, len(ch) - 1, 2):
            if eval_expr(ch[i], row) is True:
                return _unify_case(eval_expr(ch[i + 1], row), e)
        return _unify_case(eval_expr(ch[0], row), e)
    if fn in ("and", "or"):
        vals = [eval_expr(c, row) for c in ch]
        if fn == "and":
            return False if any(v is False for v in vals) else (None if any(v is None for v in vals) else True)
        return True if any(v is True for v in vals) else (None if any(v is None for v in vals) else False)
    args = [eval_expr(c, row) for c in ch]
    if fn == "not":
        return None if args[0] is None else (not args[0])
    if fn == "cast":
        v, t = args[0], e.DataTyp.Id
        if v is None:
            return None
        if t == LT_DECIMAL:
            return v if isinstance(v, Dec) else Dec.from_int(v)
        if t in (LT_FLOAT, LT_DOUBLE):
            return v if isinstance(v, np.float32) else dec_to_f32(v) if isinstance(v, Dec) else np.float32(v)
        return v
    if fn == "extract":
        if args[1] is None:
            return None
        assert args[0] == "year"
        return (datetime.date(1970, 1, 1) + datetime.timedelta(days=int(args[1]))).year
    if fn == "in":
        if args[0] is None:
            return None
        return any(_compare("=", args[0], c) is True for c in args[1:])
    if fn in ("like", "not like"):
        if args[0] is None:
            return None
        m = O.wildcard_match(args[1].encode(), args[0].encode())
        return m if fn == "like" else not m
    if any(a is None for a in args):
        return None
    a, b = _num_pair(args[0], args[1])
    if fn in ("+", "-", "*", "/"):
        if isinstance(a, Dec):
            return {"+": dec_add, "-": dec_sub, "*": dec_mul, "/": dec_quo}[fn](a, b)
        if isinstance(a, np.float32) or isinstance(b, np.float32):
            a, b = np.float32(a), np.float32(b)
            with np.errstate(all="ignore"):
                return np.float32(a + b if fn == "+" else a - b if fn == "-" else a * b if fn == "*" else a / b)
        r = {"+": a + b, "-": a - b, "*": a * b}[fn]
        return _wrap32(r) if e.DataTyp.Id == LT_INTEGER else r
    return _compare(fn, a, b)


def _unify_case(v, e):
    if isinstance(v, int) and not isinstance(v, bool) and e.DataTyp.Id == LT_DECIMAL:
        return Dec.from_int(v)
    return v


def _compare(fn, a, b):
    a, b = _num_pair(a, b)
    if isinstance(a, Dec):
        c = dec_cmp(a, b)
    elif isinstance(a, np.float32) or isinstance(b, np.float32):
        a, b = np.float32(a), np.float32(b)
        if np.isnan(a) or np.isnan(b):
            return fn == "<>"
        c = int(a > b) - int(a < b)
    else:
        c = (a > b) - (a < b)
    return {"=": c == 0, "<>": c != 0, "<": c < 0, "<=": c <= 0, ">": c > 0, ">=": c >= 0}[fn]


def _all_true(filters, row):
    return all(eval_expr(f, row) is True for f in filters)


def table_rows(cols, schema, valid=None):
    """host table {name: numpy array} + schema [(name, pg_type, width, scale, dict)] -> list of value rows.
    valid: optional {name: bool array} (False = NULL)."""
    n = len(next(iter(cols.values()))) if cols else 0
    out = [[None] * len(schema) for _ in range(n)]
    for j, (name, typ, _w, scale, dic) in enumerate(schema):
        a = cols[name]
        ok = None if valid is None else valid.get(name)
        for i in range(n):
            if ok is not None and not ok[i]:
                continue
            x = a[i]
            if typ == T_DECIMAL64:
                out[i][j] = Dec.from_int(int(x), scale)
            elif typ == T_CHAR1:
                out[i][j] = chr(int(x))
            elif typ == T_DICT8:
                out[i][j] = dic[int(x)]
            elif typ == T_VARCHAR:
                out[i][j] = x.decode() if isinstance(x, bytes) else str(x)
            else:
                out[i][j] = int(x)
    return out


def execute(op, tables):
    """tables: {name: list of value rows} (table_rows).  Returns the operator's output rows (lists of values)."""
    if op.Typ == POT_Scan:
        return [r for r in tables[op.Info.Table] if _all_true(op.Filters, r)]
    if op.Typ == POT_Filter:
        return [r for r in execute(op.Children[0], tables) if _all_true(op.Filters, r)]
    if op.Typ == POT_Project:
        return [[eval_expr(e, r) for e in op.Outputs] for r in execute(op.Children[0], tables)]
    if op.Typ == POT_Agg:
        # aggExecutor over its child's rows (executor_aggr.go:106-262): groups in first-seen order; sum(DECIMAL) is a
        # left fold of Decimal.Add in row order, sum(INTEGER) a HUGEINT, avg(DECIMAL) = sum.Quo(count), avg(INTEGER) =
        # float64 sum / float64 count, NULL arguments are ignored, an aggregate without any input is NULL
        # (function_aggr.go:620-1032)
        groups = {}
        for r in execute(op.Children[0], tables):
            key = tuple(_gkey(eval_expr(g, r)) for g in op.Info.GroupBys)
            st = groups.get(key)
            if st is None:
                st = groups[key] = {"key": [eval_expr(g, r) for g in op.Info.GroupBys], "acc": [None] * len(op.Info.Aggs), "n": [0] * len(op.Info.Aggs)}
            for i, a in enumerate(op.Info.Aggs):
                fn = a.FunImpl
                if fn == "count" and not a.Children:
                    st["n"][i] += 1
                    continue
                v = eval_expr(a.Children[0], r)
                if v is None:
                    continue
                st["n"][i] += 1
                if fn == "count":
                    continue
                cur = st["acc"][i]
                if cur is None:
                    st["acc"][i] = v
                elif fn in ("sum", "avg"):
                    st["acc"][i] = dec_add(cur, v) if isinstance(v, Dec) else cur + v
                elif fn == "min":
                    st["acc"][i] = v if _compare("<", v, cur) else cur
                elif fn == "max":
                    st["acc"][i] = v if _compare(">", v, cur) else cur
        out = []
        for st in groups.values():
            vals = []
            for i, a in enumerate(op.Info.Aggs):
                fn, acc, n = a.FunImpl, st["acc"][i], st["n"][i]
                if fn == "count" and not a.Children:
                    vals.append(n)                      # count(*): a group exists only through a row
                elif n == 0:                            # no non-NULL input: NULL, count(x) included (CountOp.Finalize, function_aggr.go:949-960)
                    vals.append(None)
                elif fn == "count":
                    vals.append(n)
                elif n == 0:
                    vals.append(None)
                elif fn == "avg":
                    vals.append(dec_quo(acc, Dec.from_int(n)) if isinstance(acc, Dec) else float(acc) / float(n))
                else:
                    vals.append(acc)
            if not all(eval_having(f, st["key"], vals) for f in op.Filters):
                continue
            out.append([st["key"][o.ColRef[1]] if o.ColRef[0] == 0 else vals[o.ColRef[1]] for o in op.Outputs])
        if not op.Info.GroupBys and not out and not op.Filters:
            pass        # an aggregate over no rows at all: the reference emits nothing for these plans (executor_aggr.go:222-262)
        return out
    if op.Typ == POT_Join:
        left, right = execute(op.Children[0], tables), execute(op.Children[1], tables)
        lk = [c.Children[0] for c in op.Info.OnConds]
        rk = [c.Children[1] for c in op.Info.OnConds]
        ht = {}
        for b in right:
            key = tuple(_key(eval_expr(k, b)) for k in rk)
            if any(x is None for x in key):
                continue                                    # NULL keys are dropped on build (join_table.go:152-195)
            ht.setdefault(key, []).append(b)
        jt = op.Info.JoinTyp
        out = []

        def emit(l, r, mark=None):
            row = []
            for o in op.Outputs:
                side, idx = o.ColRef
                row.append(l[idx] if side == 0 else mark if side == 2 else (None if r is None else r[idx]))
            out.append(row)
        for l in left:
            key = tuple(_key(eval_expr(k, l)) for k in lk)
            null_key = any(x is None for x in key)
            matches = [] if null_key else ht.get(key, [])
            if jt == JOIN_INNER:
                for r in matches:
                    emit(l, r)
            elif jt == JOIN_LEFT:
                for r in matches:
                    emit(l, r)
                if not matches:
                    emit(l, None)
            elif jt == JOIN_SEMI:
                if matches:
                    emit(l, None)
            elif jt == JOIN_ANTI:
                if not matches:
                    emit(l, None)
            elif jt in (JOIN_MARK, JOIN_ANTI_MARK):        # NOT EXISTS is planned as AntiMARK + `mark = false`; same scan (join_scan.go:54)
                emit(l, None, None if null_key else bool(matches))
            else:
                raise ValueError("join type %r" % jt)
        return out
    raise ValueError("operator %r" % op.Typ)


def _gkey(v):
    return ("d", v.signed() * 10 ** (19 - v.scale)) if isinstance(v, Dec) else v


class _HavingRow:
    """HAVING references group keys as (0, i) and aggregates as (1, i) (expr_exec.go:248-265)"""

    def __init__(self, key, vals):
        self.key, self.vals = key, vals


def eval_having(f, key, vals):
    def sub(e):
        if e.Typ == ET_Column:
            side, idx = e.ColRef
            return (key if side == 0 else vals)[idx]
        if e.Typ == ET_Const:
            return eval_expr(e, [])
        # rebuild the node over already evaluated children: reuse eval_expr through a positional row
        row = [sub(c) for c in e.Children]
        shadow = type(e)(e.Typ, e.DataTyp, Children=[type(e)(ET_Column, c.DataTyp, ColRef=(0, i)) for i, c in enumerate(e.Children)], FunImpl=e.FunImpl)
        return eval_expr(shadow, row)
    return sub(f) is True


def _key(v):
    return v.signed() * 10 ** (19 - v.scale) if isinstance(v, Dec) else v


def format_value(v, typ):
    """the text Value.String prints for the value at output type `typ` (chunk/value.go:37-66, chunk/vector.go:121-137)."""
    if v is None:
        return "NULL"
    if isinstance(v, bool):
        return "true" if v else "false"
    if isinstance(v, float) and not isinstance(v, np.float32):
        return O.fmt_double(v)
    if isinstance(v, Dec):
        buf = C.create_string_buffer(96)
        O.lib().orc_format_decimal(v.coef, v.scale, int(v.neg), typ.Scale, buf, 96)
        return buf.value.decode()
    if typ.Id == LT_DATE:
        return (datetime.date(1970, 1, 1) + datetime.timedelta(days=int(v))).isoformat()
    return str(v)


def format_rows(rows, types):
    return sorted("\t".join(format_value(v, t) for v, t in zip(r, types)) for r in rows)
