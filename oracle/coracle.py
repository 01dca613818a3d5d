"""ctypes loader for oracle/liboracle.so (the C restatement).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def lib():
    global _lib
    if _lib is None:
        p = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(p):
            raise RuntimeError("oracle/liboracle.so missing: run `make -C oracle`")
        _lib = C.CDLL(p)
        _lib.orc_max_threads.restype = C.c_int32
    return _lib


_native_note = "oracle/liboracle.so (-O3 -march=x86-64-v3, built with the repo)"


def use_native():
    """Rebuild the oracle's C sources with -O3 -march=native for THIS host's CPU (into the temp dir, never into the repo) and time
    that build from now on; falls back to the shipped x86-64-v3 build.  Returns a one-line note saying which build is loaded."""
    global _lib, _native_note
    import hashlib
    import subprocess
    import tempfile
    srcs = [os.path.join(_HERE, f) for f in ("oracle.c", "regex_bt.c")]
    h = hashlib.sha1(b"".join(open(f, "rb").read() for f in srcs)).hexdigest()[:12]
    out = os.path.join(tempfile.gettempdir(), "gpud_liboracle_native_%s_%d.so" % (h, os.getuid()))
    try:
        if not os.path.exists(out):
            cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
            subprocess.run([cc, "-O3", "-march=native", "-fPIC", "-shared", "-o", out + ".tmp"] + srcs + ["-lm", "-lpthread"], check=True,
                           stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, cwd=_HERE)
            os.replace(out + ".tmp", out)
        L = C.CDLL(out)
        L.orc_max_threads.restype = C.c_int32
        _lib = L
        _native_note = "oracle sources rebuilt on this host with gcc -O3 -march=native"
    except Exception as ex:           # no compiler on the box: keep the shipped build
        _native_note += " (native rebuild failed: %s)" % type(ex).__name__
    return _native_note


def build_note() -> str:
    return _native_note


def host_cpus() -> dict:
    """what the CPU arm may use: logical CPUs, affinity mask, cgroup quota, and the thread count the oracle takes from them"""
    d = {"logical": os.cpu_count(), "affinity": len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else None, "cgroup_quota": None}
    try:
        q, p = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        d["cgroup_quota"] = None if q == "max" else float(q) / float(p)
    except Exception:
        try:
            q = float(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
            p = float(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            d["cgroup_quota"] = q / p if q > 0 else None
        except Exception:
            pass
    d["threads_used"] = max_threads()
    return d


def windows_fields(ring: np.ndarray, W: int, thr: np.ndarray, alpha: float = 0.0, q_num: int = 0, q_den: int = 0, threads: int = 0):
    """ring: [F][n] f64 field-major.  Returns dict of [F][nw] arrays."""
    L = lib()
    ring = np.ascontiguousarray(ring, dtype=np.float64)
    F, n = ring.shape
    nw = (n + W - 1) // W
    out = {k: np.empty((F, nw), dtype=np.float64) for k in ("min", "max", "mean", "ema", "p99")}
    out["n_over"] = np.empty((F, nw), dtype=np.uint32)
    thr = np.ascontiguousarray(thr, dtype=np.float64)
    vp = C.c_void_p
    L.orc_windows_fields(vp(ring.ctypes.data), C.c_int32(F), C.c_int64(n), C.c_int32(W), vp(thr.ctypes.data), C.c_double(alpha),
                         C.c_int32(q_num), C.c_int32(q_den), C.c_int32(threads), vp(out["min"].ctypes.data), vp(out["max"].ctypes.data),
                         vp(out["mean"].ctypes.data), vp(out["ema"].ctypes.data), vp(out["p99"].ctypes.data), vp(out["n_over"].ctypes.data))
    return out


def max_threads() -> int:
    return int(lib().orc_max_threads())


class OrcHit(C.Structure):
    _fields_ = [("line", C.c_int64), ("offset", C.c_int64), ("kind", C.c_int32), ("code", C.c_int32), ("extended", C.c_int32),
                ("sub_code", C.c_int32), ("severity_fatal", C.c_int32), ("link", C.c_int64), ("intrinfo", C.c_uint32),
                ("error_status", C.c_uint32), ("device", C.c_char * 64), ("unit", C.c_char * 64)]


def xid_match(line: bytes):
    L = lib()
    assert L.orc_regex_ok() == 1
    h = OrcHit()
    L.orc_xid_match.restype = C.c_int32
    return h if L.orc_xid_match(line, C.c_int32(len(line)), C.byref(h)) else None


def sxid_match(line: bytes):
    L = lib()
    h = OrcHit()
    L.orc_sxid_match.restype = C.c_int32
    return h if L.orc_sxid_match(line, C.c_int32(len(line)), C.byref(h)) else None


N_EXT = 22    # ORC_N_EXT: hit kinds 3 .. 24


def ext_match(line: bytes):
    """(bitmask, captures): bit i = pattern of hit kind 3 + i fires; captures[kind] = group 1 of the cpu patterns"""
    L = lib()
    L.orc_ext_match.restype = C.c_int32
    cap = (C.c_int32 * (2 * N_EXT))()
    m = L.orc_ext_match(line, C.c_int32(len(line)), cap)
    caps = {3 + i: line[cap[2 * i]:cap[2 * i + 1]] for i in range(N_EXT) if (m >> i) & 1 and cap[2 * i] >= 0}
    return m, caps


def ext_groups(kind: int, line: bytes):
    """capture groups 0..9 of pattern `kind` on `line` as bytes (None = unset), or None if it does not match"""
    L = lib()
    L.orc_ext_groups.restype = C.c_int32
    c = (C.c_int32 * 20)()
    if not L.orc_ext_groups(C.c_int32(kind), line, C.c_int32(len(line)), c):
        return None
    return [line[c[2 * i]:c[2 * i + 1]] if c[2 * i] >= 0 else None for i in range(10)]


def scan_lines(buf: bytes, threads: int = 0, cap: int = 1 << 20, ext: bool = False):
    """Returns (hits list of OrcHit, n_lines) for the reference's split-on-newline scan form."""
    L = lib()
    L.orc_scan_lines_ext.restype = C.c_int64
    hits = (OrcHit * cap)()
    nl = C.c_int64()
    n = L.orc_scan_lines_ext(buf, C.c_int64(len(buf)), hits, C.c_int64(cap), C.byref(nl), C.c_int32(threads), C.c_int32(1 if ext else 0))
    return [hits[i] for i in range(min(n, cap))], nl.value
