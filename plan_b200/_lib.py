"""ctypes bindings of include/plangpu.h (libplangpu.so: the drop-in library) and include/plangpu_tpch.h
(libplangpu_tpch.so: the in-box TPC-H generator used by tests and benchmarks only)."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libplangpu.so")
TPCH_LIB_PATH = os.path.join(HERE, "libplangpu_tpch.so")

PG_OK, PG_EINVAL, PG_ENOMEM, PG_ECUDA, PG_ENCCL, PG_EOVERFLOW, PG_EUNSUPPORTED, PG_ESTATE = range(8)
STATUS_NAMES = ["PG_OK", "PG_EINVAL", "PG_ENOMEM", "PG_ECUDA", "PG_ENCCL", "PG_EOVERFLOW", "PG_EUNSUPPORTED",
                "PG_ESTATE"]

PG_DIST_SHARDED, PG_DIST_REPLICATED = 0, 1
PG_T_INT32, PG_T_INT64, PG_T_DATE32, PG_T_DECIMAL64, PG_T_CHAR1, PG_T_DICT8, PG_T_FLOAT64, PG_T_HUGEINT, \
    PG_T_DECIMAL128, PG_T_VARCHAR, PG_T_BOOL = range(1, 12)


class PlanGpuError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__("%s: %s" % (STATUS_NAMES[status] if 0 <= status < 8 else status, msg))
        self.status = status


class ColDesc(C.Structure):
    _fields_ = [("name", C.c_char_p), ("type", C.c_int32), ("width", C.c_int32), ("scale", C.c_int32),
                ("dict_len", C.c_int32), ("dict", C.POINTER(C.c_char_p))]


class DevInfo(C.Structure):
    _fields_ = [("device", C.c_int32), ("sm_count", C.c_int32), ("cc_major", C.c_int32), ("cc_minor", C.c_int32),
                ("hbm_bytes", C.c_int64), ("l2_bytes", C.c_int64), ("max_smem_per_block", C.c_int32),
                ("world_size", C.c_int32), ("rank", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [("exec_ms", C.c_double), ("kernel_ms", C.c_double), ("main_kernel_ms", C.c_double),
                ("comm_ms", C.c_double), ("rows_scanned", C.c_int64), ("algorithmic_bytes", C.c_int64),
                ("main_kernel_bytes", C.c_int64), ("kernel_launches", C.c_int32), ("reserved", C.c_int32),
                ("aux", C.c_int64 * 8)]


class ColBuf(C.Structure):
    """pg_colbuf: a host column buffer at 1/2/4/8 bytes per value with a frame of reference."""
    _fields_ = [("data", C.c_void_p), ("width", C.c_int32), ("reserved", C.c_int32), ("base", C.c_int64), ("valid", C.c_void_p)]


class PgDecimal(C.Structure):
    _fields_ = [("coef", C.c_uint64), ("scale", C.c_int32), ("neg", C.c_uint32)]


class PgHugeint(C.Structure):
    _fields_ = [("lower", C.c_uint64), ("upper", C.c_int64)]


# every symbol include/plangpu.h + include/plangpu_tpch.h declare: (name, restype, argtypes)
_P = C.c_void_p
SIGNATURES = [
    ("pg_abi_version", C.c_int, []),
    ("pg_trim", C.c_int, []),
    ("pg_init", C.c_int, [C.c_int]),
    ("pg_shutdown", C.c_int, []),
    ("pg_last_error", C.c_char_p, []),
    ("pg_device_info", C.c_int, [C.POINTER(DevInfo)]),
    ("pg_comm_unique_id", C.c_int, [_P]),
    ("pg_comm_init", C.c_int, [C.c_int, C.c_int, _P]),
    ("pg_comm_destroy", C.c_int, []),
    ("pg_table_create", C.c_int, [C.c_char_p, C.c_int, C.POINTER(ColDesc), C.POINTER(_P)]),
    ("pg_table_reserve", C.c_int, [_P, C.c_int64]),
    ("pg_table_append", C.c_int, [_P, C.c_int64, C.POINTER(_P), C.POINTER(_P)]),
    ("pg_table_append_cols", C.c_int, [_P, C.c_int64, C.POINTER(ColBuf)]),
    ("pg_table_column_encoding", C.c_int, [_P, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int64)]),
    ("pg_table_read_column", C.c_int, [_P, C.c_int, C.c_int64, C.c_int64, _P]),
    ("pg_table_read_column_stored", C.c_int, [_P, C.c_int, C.c_int64, C.c_int64, _P]),
    ("pg_table_device_column", C.c_int, [_P, C.c_int, C.POINTER(_P)]),
    ("pg_table_set_rows", C.c_int, [_P, C.c_int64]),
    ("pg_table_seal", C.c_int, [_P, C.c_int64]),
    ("pg_table_set_distribution", C.c_int, [_P, C.c_int]),
    ("pg_table_rows", C.c_int, [_P, C.POINTER(C.c_int64)]),
    ("pg_table_free", None, [_P]),
    ("pg_plan_compile", C.c_int, [C.POINTER(C.c_int64), C.c_size_t, C.POINTER(_P)]),
    ("pg_plan_bind", C.c_int, [_P, C.c_int, _P]),
    ("pg_plan_prepare", C.c_int, [_P]),
    ("pg_plan_explain", C.c_char_p, [_P]),
    ("pg_plan_execute", C.c_int, [_P, C.POINTER(_P)]),
    ("pg_plan_free", None, [_P]),
    ("pg_result_num_columns", C.c_int, [_P, C.POINTER(C.c_int)]),
    ("pg_result_column_type", C.c_int, [_P, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    ("pg_result_rows", C.c_int, [_P, C.POINTER(C.c_int64)]),
    ("pg_result_next", C.c_int, [_P, C.c_int64, C.POINTER(C.c_int64), C.POINTER(_P), C.POINTER(_P)]),
    ("pg_result_rewind", C.c_int, [_P]),
    ("pg_result_column_dict", C.c_int, [_P, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.POINTER(C.c_char_p))]),
    ("pg_result_stats", C.c_int, [_P, C.POINTER(Stats)]),
    ("pg_result_free", None, [_P]),
    ("pg_host_merge_partials", C.c_int, [_P, C.c_int64, C.c_int, C.c_int, C.c_int, _P, _P]),
    ("pg_init_devices", C.c_int, [C.c_int, C.POINTER(C.c_int)]),
    ("pg_use_device", C.c_int, [C.c_int]),
    ("pg_num_devices", C.c_int, []),
]
# include/plangpu_tpch.h (libplangpu_tpch.so)
TPCH_SIGNATURES = [
    ("pg_tpch_num_orders", C.c_int64, [C.c_double]),
    ("pg_tpch_num_customers", C.c_int64, [C.c_double]),
    ("pg_tpch_orders_lineitem", C.c_int, [C.c_double, C.c_int64, C.c_int64, C.POINTER(_P), C.POINTER(_P)]),
    ("pg_tpch_customer", C.c_int, [C.c_double, C.c_int64, C.c_int64, C.POINTER(_P)]),
    ("pg_tpch_part", C.c_int, [C.c_double, C.POINTER(_P)]),
    ("pg_tpch_supplier", C.c_int, [C.c_double, C.POINTER(_P)]),
    ("pg_tpch_partsupp", C.c_int, [C.c_double, C.POINTER(_P)]),
    ("pg_tpch_partsupp_range", C.c_int, [C.c_double, C.c_int64, C.c_int64, C.POINTER(_P)]),
    ("pg_tpch_num_parts", C.c_int64, [C.c_double]),
    ("pg_tpch_nation", C.c_int, [C.POINTER(_P)]),
]

_LIB = None


class _Libs:
    """Attribute access over both libraries: the operator ABI first, then the generator."""

    def __init__(self, core, tpch):
        self.core, self.tpch = core, tpch

    def __getattr__(self, name):
        try:
            return getattr(self.core, name)
        except AttributeError:
            return getattr(self.tpch, name)


def _load(path, signatures):
    if not os.path.exists(path):
        raise ImportError("%s is missing: build it with `python -m plan_b200.build` or __graft_entry__.build(); "
                          "the GPU path has no CPU fallback" % path)
    L = C.CDLL(path, mode=C.RTLD_GLOBAL)
    for name, res, args in signatures:
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    return L


def lib():
    """Load libplangpu.so (+ the generator library).  Raises if they were not built -- there is no CPU fallback."""
    global _LIB
    if _LIB is None:
        core = _load(LIB_PATH, SIGNATURES)
        _LIB = _Libs(core, _load(TPCH_LIB_PATH, TPCH_SIGNATURES))
    return _LIB


def check(status):
    if status != PG_OK:
        raise PlanGpuError(status, lib().pg_last_error().decode(errors="replace"))
