"""oracle.py -- TEST INFRASTRUCTURE.  ctypes access to the two CPU checkers:

  * run_ref(variant, cfg)   the REFERENCE SOURCES compiled in draw-tape mode
                            (oracle/_ref/libref_w.so, libref_b.so; build_ref.sh)
  * run_port(cfg)           the C restatement (oracle/rach_oracle.c)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.  The product (the package next to it) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


class RefConfig(C.Structure):
    _fields_ = [("nUE", C.c_int), ("distribution", C.c_int), ("nPreamble", C.c_int),
                ("backoffIndicator", C.c_int), ("nGrantUL", C.c_int),
                ("maxRarWindow", C.c_int), ("maxMsg2TxCount", C.c_int),
                ("accessTime", C.c_int), ("cellRadius", C.c_float), ("hBS", C.c_float),
                ("hUT", C.c_float), ("geometry", C.c_int), ("seed", C.c_ulonglong),
                ("rep", C.c_int), ("useTape", C.c_int), ("stopMs", C.c_int),
                ("echo", C.c_int)]


class RefResult(C.Structure):
    _fields_ = [("simTimeMs", C.c_int), ("nSuccess", C.c_int),
                ("preambleTxSum", C.c_longlong), ("delaySum", C.c_longlong),
                ("failCountSum", C.c_longlong), ("continueFailed", C.c_longlong),
                ("collisionPreambles", C.c_longlong), ("totalPreambleTxop", C.c_longlong),
                ("collisionScans", C.c_longlong), ("totalScans", C.c_longlong),
                ("draws", C.c_longlong), ("maxDrawsPerUeMs", C.c_int), ("lastMs", C.c_int),
                ("aborted", C.c_int), ("captured", C.c_int),
                ("failCountsPrinted", C.c_int), ("nAccessUE", C.c_int),
                ("averageDelay", C.c_double), ("averagePreambleTx", C.c_double),
                ("ratioSuccess", C.c_double), ("seconds", C.c_double),
                ("lateRestarts", C.c_longlong), ("lateAbsorbed", C.c_longlong),
                ("pairTests", C.c_longlong), ("minPairMargin", C.c_double)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


DEFAULTS = dict(nUE=10000, distribution=2, nPreamble=54, backoffIndicator=20, nGrantUL=12,
                maxRarWindow=6, maxMsg2TxCount=9, accessTime=5, cellRadius=400.0, hBS=10.0,
                hUT=1.8, geometry=1, seed=0, rep=0, useTape=1, stopMs=0, echo=0)

DUMP_FIELDS = ["timer", "active", "txTime", "firstTxTime", "secondTxTime", "nowBackoff",
               "preamble", "preambleChange", "rarWindow", "maxRarCounter",
               "preambleTxCounter", "msg2Flag", "connectionRequest", "msg4Flag",
               "failCount", "sector"]


def make_config(**kw):
    d = dict(DEFAULTS)
    d.update(kw)
    return RefConfig(**d)


def build(force=False):
    """Compile the restatement and, where /root/reference exists, the tape-mode reference."""
    port = os.path.join(HERE, "_build", "librach_oracle.so")
    srcs = [os.path.join(HERE, f) for f in ("rach_oracle.c", "rach_oracle_n.c", "rach_oracle_u0.c", "ref_api.h")]
    if force or not os.path.exists(port) or os.path.getmtime(port) < max(os.path.getmtime(f) for f in srcs):
        subprocess.check_call(["make", "-C", HERE, "_build/librach_oracle.so"],
                              stdout=subprocess.DEVNULL)
    have_ref = all(os.path.exists(os.path.join(HERE, "_ref", f))
                   for f in ("libref_w.so", "libref_b.so", "libref_n.so", "libref_n2.so", "libref_u0.so"))
    if os.path.isdir("/root/reference") and (force or not have_ref):
        subprocess.check_call([os.path.join(HERE, "build_ref.sh")], stdout=subprocess.DEVNULL)


_libs = {}


def _lib(path, fn):
    key = (path, fn)
    if key not in _libs:
        lib = C.CDLL(path)
        f = getattr(lib, fn)
        f.restype = C.c_int
        f.argtypes = [C.POINTER(RefConfig), C.POINTER(RefResult), C.c_void_p, C.c_void_p]
        _libs[key] = f
    return _libs[key]


def ref_available(variant="w"):
    return os.path.exists(os.path.join(HERE, "_ref", "libref_%s.so" % variant))


def _run(f, cfg, per_ue, geom):
    res = RefResult()
    n = cfg.nUE
    ue = np.zeros((n, 16), dtype=np.int32) if per_ue else None
    gm = np.zeros((n, 6), dtype=np.float32) if geom else None
    rc = f(C.byref(cfg), C.byref(res),
           ue.ctypes.data_as(C.c_void_p) if per_ue else None,
           gm.ctypes.data_as(C.c_void_p) if geom else None)
    if rc != 0:
        raise RuntimeError("oracle run failed rc=%d" % rc)
    return res, ue, gm


def run_ref(variant, cfg, per_ue=True, geom=False):
    """variant 'w' (RandomAccessWithNOMA.c) or 'b' (RandomAccessSimulatorBeta.c)."""
    path = os.path.join(HERE, "_ref", "libref_%s.so" % variant)
    if not os.path.exists(path):
        raise FileNotFoundError(path)
    return _run(_lib(path, "ref_run"), cfg, per_ue, geom)


N_DEFAULTS = dict(nGrantUL=2, maxRarWindow=5, maxMsg2TxCount=10, cellRadius=500.0, geometry=1)
N_DUMP_FIELDS = ["timer", "active", "txTime", "firstTxTime", "secondTxTime", "nowBackoff", "preamble", "sector",
                 "rarWindow", "msg1ReTx", "nTxPreamble", "msg2", "msg3Wait", "RA", "msg3Faile", "RaFailed"]


def make_config_n(**kw):
    """NOMA.c defaults (N:41-57): 2 grants per sector, maxRarWindow 5, maxMsg1ReTx 10, 500 m cell."""
    d = dict(N_DEFAULTS)
    d.update(kw)
    return make_config(**d)


def _run_n(f, cfg):
    res = RefResult()
    n = cfg.nUE
    ue = np.zeros((n, 16), dtype=np.int32)
    gain = np.zeros(n, dtype=np.float64)
    rc = f(C.byref(cfg), C.byref(res), ue.ctypes.data_as(C.c_void_p), gain.ctypes.data_as(C.c_void_p))
    if rc != 0:
        raise RuntimeError("oracle run failed rc=%d" % rc)
    return res, ue, gain


def run_ref_n(cfg):
    """NOMA.c itself in tape mode -> (result, per-UE ints [n,16] in N_DUMP_FIELDS order, channelGain [n]).
    cfg.geometry == 0: the build with NOMA.c's alternative non-sector collision function switched in."""
    path = os.path.join(HERE, "_ref", "libref_n.so" if cfg.geometry else "libref_n2.so")
    if not os.path.exists(path):
        raise FileNotFoundError(path)
    return _run_n(_lib(path, "ref_run"), cfg)


def run_port_n(cfg):
    build()
    return _run_n(_lib(os.path.join(HERE, "_build", "librach_oracle.so"), "oracle_run_n"), cfg)


U0_DEFAULTS = dict(nPreamble=64, distribution=1, geometry=0)
U0_DUMP_FIELDS = ["timer", "active", "txTime", "preamble", "preambleChange", "rarWindow", "maxRarCounter",
                  "preambleTxCounter", "msg2Flag", "connectionRequest", "msg4Flag", "raFailed", "nowBackoff", "-", "-", "-"]


def make_config_u0(**kw):
    """RandomAccessSimulator.c as shipped (U0:48-59): 64 preambles, BI 20, Uniform 60 s."""
    d = dict(U0_DEFAULTS)
    d.update(kw)
    return make_config(**d)


def run_ref_u0(cfg):
    path = os.path.join(HERE, "_ref", "libref_u0.so")
    if not os.path.exists(path):
        raise FileNotFoundError(path)
    r, ue, _ = _run(_lib(path, "ref_run"), cfg, True, False)
    return r, ue


def run_port_u0(cfg):
    build()
    r, ue, _ = _run(_lib(os.path.join(HERE, "_build", "librach_oracle.so"), "oracle_run_u0"), cfg, True, False)
    return r, ue


def run_port(cfg, per_ue=True, geom=False):
    build()
    path = os.path.join(HERE, "_build", "librach_oracle.so")
    return _run(_lib(path, "oracle_run"), cfg, per_ue, geom)
