"""CPU-side checks of librach_gpu: it loads, exports every symbol of include/rach_gpu.h, the host
helpers give the known answers of the reference's formulas, and without a GPU it refuses loudly."""
import ctypes as C
import importlib
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def pkg():
    p = importlib.import_module("5g-nr-randomaccess_b200")
    p.build_lib()
    return p


def test_exports_every_declared_symbol(pkg):
    hdr = open(os.path.join(ROOT, "include", "rach_gpu.h")).read()
    declared = set(re.findall(r"\b(ra_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"ra_sim", "ra_params", "ra_stats", "ra_options"}
    lib = pkg.load_lib()
    assert declared, "no declarations parsed"
    for sym in sorted(declared):
        assert hasattr(lib, sym), sym
    assert set(pkg.SYMBOLS) == declared


def test_struct_sizes_match_header(pkg, tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include "rach_gpu.h"\n#include <stdio.h>\nint main(){printf("%zu %zu %zu\\n",'
                   'sizeof(ra_params),sizeof(ra_stats),sizeof(ra_options));return 0;}\n')
    exe = tmp_path / "sz"
    import subprocess
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    a, b, c = map(int, subprocess.check_output([str(exe)]).split())
    assert (a, b, c) == (C.sizeof(pkg.RaParams), C.sizeof(pkg.RaStats), C.sizeof(pkg.RaOptions))


def test_defaults_are_the_reference_defaults(pkg):
    p = pkg.default_params()          # RandomAccessWithNOMA.c:69-88
    assert (p.nPreamble, p.backoffIndicator, p.nGrantUL, p.maxRarWindow, p.maxMsg2TxCount,
            p.accessTime, p.distribution) == (54, 20, 12, 6, 9, 5, 2)
    assert (p.cellRadius, p.hBS) == (400.0, 10.0) and abs(p.hUT - 1.8) < 1e-6


@pytest.mark.parametrize("n,peak,peak_ms,all_at,total", [
    (10000, 11, 3350, 6855, 11240), (50000, 53, 3745, 7750, 51579), (100000, 105, 3745, 7970, 102052)])
def test_beta_schedule_known_answers(pkg, n, peak, peak_ms, all_at, total):
    """SURVEY section 4 known answers of W:285-287 (float/double mix, ceil of a float quotient)."""
    p = pkg.default_params(nUE=n)
    arr, at = pkg.arrival_schedule(p)
    assert arr[0] == 0 and list(arr[5:30:5]) == [1, 1, 1, 1, 1]
    assert arr.sum() == n and at == all_at
    assert arr.max() == peak and int(np.argmax(arr)) == peak_ms
    assert (arr[np.arange(len(arr)) % 5 != 0] == 0).all()
    # unclamped sum of ceil(): recompute with a huge nUE cap is not possible; check the clamp point
    assert arr[:all_at].sum() < n <= arr[:all_at + 1].sum()


def test_uniform_schedule_known_answers(pkg):
    """W:246: nAccessUE = 1,2,3,4,5,5,6,7,8,9 for 10k..100k."""
    exp = [1, 2, 3, 4, 5, 5, 6, 7, 8, 9]
    for n, k in zip(range(10000, 100001, 10000), exp):
        p = pkg.default_params(nUE=n, distribution=1)
        arr, at = pkg.arrival_schedule(p)
        assert arr[0] == k and arr.max() == k and arr.sum() == n
        assert len(arr) == 60000


def test_no_gpu_means_loud_failure(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(pkg.RachError, match="no CUDA device"):
        pkg.RachSim([pkg.default_params(nUE=100)], reps=1)


def test_product_never_imports_the_oracle():
    pk = os.path.join(ROOT, "5g-nr-randomaccess_b200")
    for dp, _, files in os.walk(pk):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".c")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "oracle" not in txt.lower().replace("# oracle", ""), os.path.join(dp, f)


def test_beta_schedule_equals_the_c_expression_for_many_sizes(pkg, oracle):
    """ra_arrival_schedule (C++ host code of the product) against the reference's C expression
    (W:285-287, 844-847) for 400 population sizes and three subframe lengths.  Guards the C-vs-C++
    pow() overload trap: pow(float, float) in C++ is powf and moves ceil() for some nUE."""
    import ctypes as C
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_build", "librach_oracle.so"))
    lib.oracle_beta_arrivals.argtypes = [C.c_int, C.c_int, C.c_void_p]
    ref = np.zeros(10000, dtype=np.int32)
    sizes = list(range(1000, 100001, 1000)) + list(range(105000, 1000001, 5000)) + [1, 7, 54321, 300000, 999983]
    for a in (5, 7, 10):
        for n in sizes if a == 5 else sizes[::7]:
            at_ref = lib.oracle_beta_arrivals(n, a, ref.ctypes.data_as(C.c_void_p))
            arr, at = pkg.arrival_schedule(pkg.default_params(nUE=n, accessTime=a))
            assert at == at_ref and np.array_equal(arr, ref), (n, a)


def test_validation_without_a_device(pkg):
    """ra_params_validate = the checks of ra_sim_create.  Variant N: the rejection loops of activeUE
    (NOMA.c:167-172 r > 35 m, NOMA.c:185-189 gain >= 1e-7) never end for a radius at or below 35 m (e.g. a
    zero-initialised ra_params) and practically never beyond a few km -- refused instead of hanging the GPU."""
    ok = pkg.default_params(variant=2)
    assert pkg.validate_params(ok) == (0, "")
    for bad in (0.0, 35.0, -1.0, float("nan"), float("inf"), 5001.0):
        rc, msg = pkg.validate_params(pkg.default_params(variant=2, cellRadius=bad))
        assert rc == -1 and "cellRadius" in msg, bad
    assert pkg.validate_params(pkg.default_params(variant=2, cellRadius=35.5))[0] == 0
    # W never feeds the radius back into the dynamics (W:392-415 are side outputs): any value is accepted there
    assert pkg.validate_params(pkg.default_params(cellRadius=0.0))[0] == 0
    # the move calendar of a replication is indexed with 32 bits: ring (next power of two of BI + subframe + window) x nUE
    rc, msg = pkg.validate_params(pkg.default_params(nUE=1 << 24, backoffIndicator=4096))
    assert rc == -1 and "calendar records" in msg
    assert pkg.validate_params(pkg.default_params(nUE=1 << 24, backoffIndicator=100))[0] == 0      # 128 x 2^24 = 2^31
    rc, msg = pkg.validate_params(pkg.default_params(nPreamble=0))
    assert rc == -1 and "nPreamble" in msg


@pytest.mark.parametrize("a", [5, 64, 4096])
def test_beta_schedule_beyond_the_reference_horizon_is_monotone(pkg, a):
    """maxTimeMs above the reference's 10 s: beta_dist's (1-x)^3 is negative for x > 1 (W:844-847); nobody
    arrives there, the cumulative schedule never decreases and never exceeds nUE."""
    p = pkg.default_params(nUE=100000, accessTime=a)
    p.maxTimeMs = 30000
    arr, at = pkg.arrival_schedule(p)
    assert len(arr) == 30000 and (arr >= 0).all() and arr.sum() <= 100000
    assert arr[10000:].sum() == 0 or a == 5      # with the default subframe everybody has arrived by 7970 ms
    ref, _ = pkg.arrival_schedule(pkg.default_params(nUE=100000, accessTime=a))
    assert np.array_equal(arr[:10000], ref)
