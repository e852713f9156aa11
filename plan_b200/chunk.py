"""Host-side mirror of the reference's pkg/chunk and pkg/common types (Python stand-in
for the Go side, which cannot be compiled in this image).

  LType / LTID_*      /root/reference/pkg/common/ltype.go
  Vector              /root/reference/pkg/chunk/vector.go:15-22   (FLAT / CONST formats)
  Chunk               /root/reference/pkg/chunk/chunk.go:16-20
  Value.String        /root/reference/pkg/chunk/value.go:26-70
  Vector.GetValue     /root/reference/pkg/chunk/vector.go:76-186
  Chunk.SaveToFile    /root/reference/pkg/chunk/chunk.go:196-220

Vectors hold device-native encodings (include/plangpu.h): DECIMAL inputs are unscaled
int64 at the column scale, DATE is int32 days since 1970-01-01, VARCHAR(1) a byte.
DECIMAL / HUGEINT results use the 16-byte pg_decimal / pg_hugeint records.
"""
import datetime

import numpy as np

from . import _lib as L

DEFAULT_VECTOR_SIZE = 2048     # pkg/util/util.go:124

LTID_BOOLEAN, LTID_INTEGER, LTID_BIGINT, LTID_DATE, LTID_DECIMAL, LTID_FLOAT, LTID_DOUBLE, LTID_VARCHAR, \
    LTID_HUGEINT = range(1, 10)    # numbering of plangpu_desc.h PG_LT_*

DECIMAL128 = np.dtype([("coef", "<u8"), ("scale", "<i4"), ("neg", "<u4")])
HUGEINT = np.dtype([("lower", "<u8"), ("upper", "<i8")])


class LType:
    def __init__(self, id, width=0, scale=0):
        self.Id, self.Width, self.Scale = id, width, scale

    def __repr__(self):
        names = {1: "BOOLEAN", 2: "INTEGER", 3: "BIGINT", 4: "DATE", 5: "DECIMAL", 6: "FLOAT", 7: "DOUBLE",
                 8: "VARCHAR", 9: "HUGEINT"}
        if self.Id == LTID_DECIMAL:
            return "DECIMAL(%d,%d)" % (self.Width, self.Scale)
        return names.get(self.Id, "?")

    def __eq__(self, o):
        return (self.Id, self.Width, self.Scale) == (o.Id, o.Width, o.Scale)


def BooleanType(): return LType(LTID_BOOLEAN)
def IntegerType(): return LType(LTID_INTEGER)
def BigintType(): return LType(LTID_BIGINT)
def DateType(): return LType(LTID_DATE)
def DecimalType(w, s): return LType(LTID_DECIMAL, w, s)
def FloatType(): return LType(LTID_FLOAT)
def DoubleType(): return LType(LTID_DOUBLE)
def VarcharType(): return LType(LTID_VARCHAR)
def HugeintType(): return LType(LTID_HUGEINT)


PF_FLAT, PF_CONST = 0, 1


class Vector:
    def __init__(self, typ, data=None, phy_format=PF_FLAT, mask=None, dictionary=None):
        self._Typ, self.Data, self._PhyFormat, self.Mask = typ, data, phy_format, mask
        self.Dict = dictionary      # DICT8 columns: code -> string

    def Typ(self): return self._Typ
    def PhyFormat(self): return self._PhyFormat

    def GetValue(self, idx):
        if self._PhyFormat == PF_CONST:
            idx = 0
        if self.Mask is not None and not (self.Mask[idx >> 3] >> (idx & 7)) & 1:
            return Value(self._Typ, is_null=True)
        x = self.Data[idx]
        t = self._Typ.Id
        if t == LTID_DECIMAL:
            if self.Data.dtype == DECIMAL128:
                coef, scale, neg = int(x["coef"]), int(x["scale"]), bool(x["neg"])
            else:
                v = int(x)
                coef, scale, neg = abs(v), self._Typ.Scale, v < 0
            r = decimal_int64(coef, scale, neg, self._Typ.Scale)
            if r is None:
                return Value(self._Typ, s=decimal_string(coef, scale, neg))
            return Value(self._Typ, i64=r[0], i64_1=r[1])
        if t == LTID_HUGEINT:
            return Value(self._Typ, i64=int(x["upper"]), i64_1=int(x["lower"]))
        if t == LTID_VARCHAR:
            if isinstance(x, (bytes, str)):          # full VARCHAR column: the string itself
                return Value(self._Typ, s=x.decode() if isinstance(x, bytes) else x)
            if self.Dict is not None:
                return Value(self._Typ, s=self.Dict[int(x)])
            return Value(self._Typ, s=chr(int(x)))
        if t in (LTID_DOUBLE, LTID_FLOAT):
            return Value(self._Typ, f64=float(x))
        return Value(self._Typ, i64=int(x))


class Chunk:
    def __init__(self):
        self.Data, self.Count = [], 0

    def Init(self, types, cap=DEFAULT_VECTOR_SIZE):
        self.Data = [Vector(t) for t in types]
        self.Count = 0

    def Card(self): return self.Count
    def SetCard(self, n): self.Count = n
    def ColumnCount(self): return len(self.Data)

    def SaveToFile(self, fh):
        for r in range(self.Count):
            fh.write("\t".join(v.GetValue(r).String() for v in self.Data) + "\n")


class Value:
    def __init__(self, typ, is_null=False, i64=0, i64_1=0, f64=0.0, s=""):
        self.Typ, self.IsNull, self.I64, self.I64_1, self.F64, self.Str = typ, is_null, i64, i64_1, f64, s

    def String(self):
        if self.IsNull:
            return "NULL"
        t = self.Typ.Id
        if t == LTID_BOOLEAN:
            return "true" if self.I64 else "false"      # Value.String of a bool (chunk/value.go)
        if t in (LTID_INTEGER, LTID_BIGINT):
            return "%d" % self.I64
        if t == LTID_VARCHAR:
            return self.Str
        if t == LTID_DECIMAL:
            if self.Str:
                return self.Str
            return decimal_string(*new_from_int64(self.I64, self.I64_1, self.Typ.Scale))
        if t == LTID_DATE:
            return (datetime.date(1970, 1, 1) + datetime.timedelta(days=self.I64)).isoformat()
        if t in (LTID_DOUBLE, LTID_FLOAT):
            return go_float_string(self.F64)
        if t == LTID_HUGEINT:
            return str((self.I64 << 64) + (self.I64_1 & 0xFFFFFFFFFFFFFFFF))
        raise ValueError("usp")


# ---- govalues formatting contract (Int64 / NewFromInt64 / String) on Python ints ------

def _round_half_even_div(x, p):
    q, r = divmod(x, p)
    if 2 * r > p or (2 * r == p and q & 1):
        q += 1
    return q


def decimal_int64(coef, scale, neg, want_scale):
    """Decimal.Int64(scale): (whole, frac) with frac half-even rounded / zero padded."""
    x, y = coef, 10 ** scale
    if want_scale < scale:
        x = _round_half_even_div(coef, 10 ** (scale - want_scale))
        y = 10 ** want_scale
    q, r = divmod(x, y)
    if want_scale > scale:
        r *= 10 ** (want_scale - scale)
    if q > (1 << 63) - 1 or r > (1 << 63) - 1:
        return None
    return (-q, -r) if neg else (q, r)


def new_from_int64(whole, frac, scale):
    """decimal.NewFromInt64: trailing zeros of the fraction are stripped."""
    neg = whole < 0 or frac < 0
    f, s = abs(frac), scale
    while s > 0 and f % 10 == 0:
        f //= 10
        s -= 1
    return (abs(whole) * 10 ** s + f, s, neg)


def decimal_string(coef, scale, neg):
    digits = str(coef).rjust(scale + 1, "0")
    out = digits[:-scale] + "." + digits[-scale:] if scale > 0 else digits
    return ("-" if neg and coef != 0 else "") + out


def go_float_string(x):
    """fmt.Sprintf("%v", float64) (chunk/value.go:55-58): strconv's shortest 'g' -- the shortest digits that round-trip, %e form
    (d.ddde+XX, two exponent digits at least) when the decimal exponent is < -4 or >= 6, plain digits otherwise."""
    import decimal
    import math
    x = float(x)
    if math.isnan(x):
        return "NaN"
    if math.isinf(x):
        return "+Inf" if x > 0 else "-Inf"
    if x == 0:
        return "-0" if math.copysign(1.0, x) < 0 else "0"
    sign, digs, e10 = decimal.Decimal(repr(x)).as_tuple()        # repr: the shortest round-trip digits
    digs = list(digs)
    while len(digs) > 1 and digs[-1] == 0:
        digs.pop()
        e10 += 1
    text = "".join(map(str, digs))
    exp = len(text) + e10 - 1                                      # value = d.ddd * 10^exp
    neg = "-" if sign else ""
    if exp < -4 or exp >= 6:
        return "%s%s%se%s%02d" % (neg, text[0], "." + text[1:] if len(text) > 1 else "", "-" if exp < 0 else "+", abs(exp))
    if exp < 0:
        return neg + "0." + "0" * (-exp - 1) + text
    if e10 >= 0:
        return neg + text + "0" * e10
    return neg + text[:exp + 1] + "." + text[exp + 1:]


# pg_string {int64_t len; const char *data} (include/plangpu.h; the layout of common.String, string.go:10-13)
PG_STRING = np.dtype([("len", np.int64), ("data", np.uint64)])


def native_dtype(pg_type):
    return {L.PG_T_INT32: np.int32, L.PG_T_INT64: np.int64, L.PG_T_DATE32: np.int32, L.PG_T_DECIMAL64: np.int64,
            L.PG_T_CHAR1: np.uint8, L.PG_T_DICT8: np.uint8, L.PG_T_FLOAT64: np.float64, L.PG_T_HUGEINT: HUGEINT,
            L.PG_T_DECIMAL128: DECIMAL128, L.PG_T_VARCHAR: PG_STRING, L.PG_T_BOOL: np.uint8}[pg_type]
