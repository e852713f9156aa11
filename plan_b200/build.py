"""Build libplangpu.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SOURCES = ["table.cu", "plan.cu", "scanagg.cu", "join.cu", "rows.cu", "comm.cu"]      # libplangpu.so: the drop-in operator library
TPCH_SOURCES = ["tpchgen.cu"]                                           # libplangpu_tpch.so: in-box data generator (tests / bench only)
OUT = os.path.join(HERE, "libplangpu.so")
TPCH_OUT = os.path.join(HERE, "libplangpu_tpch.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function", "--expt-relaxed-constexpr"] + os.environ.get("PG_NVCC_EXTRA", "").split()


def needs_build():
    if not os.path.exists(OUT) or not os.path.exists(TPCH_OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps += [os.path.join(HERE, "..", "include", f) for f in os.listdir(os.path.join(HERE, "..", "include"))]
    return any(os.path.getmtime(d) > t for d in deps)


def host_stale():
    """the host shim binaries are older than their own sources (plan_b200/host/*.cc, *.hpp) or than the library they link"""
    outs = [os.path.join(HERE, "host", "planhost_run"), os.path.join(HERE, "host", "planhost_append")]
    if not all(os.path.exists(o) for o in outs):
        return True
    t = min(os.path.getmtime(o) for o in outs)
    hdir = os.path.join(HERE, "host")
    deps = [os.path.join(hdir, f) for f in os.listdir(hdir) if f.endswith((".cc", ".hpp"))] + [OUT, TPCH_OUT]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        if host_stale():
            build_host()
        return OUT
    objs = []
    procs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    for s in SOURCES + TPCH_SOURCES:
        o = os.path.join(HERE, "build", s.replace(".cu", ".o"))
        cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, s), "-o", o]
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(o)
    ok = True
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            ok = False
            sys.stderr.write("nvcc failed for %s:\n%s\n" % (s, out))
        elif verbose:
            sys.stderr.write(out)
    if not ok:
        raise RuntimeError("libplangpu build failed")
    core = [o for o, src in zip(objs, SOURCES + TPCH_SOURCES) if src in SOURCES]
    tpch = [o for o, src in zip(objs, SOURCES + TPCH_SOURCES) if src in TPCH_SOURCES]
    subprocess.check_call([NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", OUT] + core + ["-lcudart", "-ldl"])
    subprocess.check_call([NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", TPCH_OUT] + tpch +
                          ["-L" + HERE, "-lplangpu", "-lcudart", "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN"])
    build_host()
    return OUT


HOST_BIN = os.path.join(HERE, "host", "planhost_run")
APPEND_BIN = os.path.join(HERE, "host", "planhost_append")


def build_host():
    """The C++ host shim above the C ABI (plan_b200/host): driver binaries linked against the two libraries only.
    planhost_run executes a TPC-H query like `tester tpch1g --query_id N`; planhost_append measures the ingest path
    the Go shim would take (2048-row appends from pageable memory)."""
    cxx = os.environ.get("CXX", "g++")
    for src, out in (("shim_main.cc", HOST_BIN), ("append_bench.cc", APPEND_BIN)):
        subprocess.check_call([cxx, "-std=c++17", "-O2", "-Wall", "-o", out, os.path.join(HERE, "host", src),
                               "-L" + HERE, "-lplangpu_tpch", "-lplangpu", "-Wl,-rpath,$ORIGIN/..", "-Wl,-rpath," + HERE,
                               "-Wl,-rpath,/usr/local/cuda/lib64"])
    return HOST_BIN


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(OUT)
