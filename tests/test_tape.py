"""Known-answer tests of the Philox4x32-10 draw tape (include/rach_tape.h) and of the float
threshold of the Msg3 test (RandomAccessWithNOMA.c:670-671)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SRC = r'''
#include "rach_tape.h"
void kat(const unsigned* ctr, const unsigned* key, unsigned* out) {
    rach_u32x4 r = rach_philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1]);
    for (int i = 0; i < 4; ++i) out[i] = r.v[i];
}
int draw(unsigned long long seed, unsigned rep, unsigned ue, unsigned ms, unsigned k) {
    return rach_tape_rand31(seed, rep, ue, ms, k, RACH_TAPE_TAG_UE);
}
int msg3(int r) { return rach_msg3_success(r); }
int msg3_ref(int r) { float p = (float)r / (float)2147483647; if (p > 0.1) return 1; return 0; }
'''


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    d = tmp_path_factory.mktemp("tape")
    (d / "t.c").write_text(SRC)
    so = str(d / "t.so")
    subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-I", os.path.join(ROOT, "include"),
                           str(d / "t.c"), "-o", so])
    return C.CDLL(so)


# Random123 kat_vectors, philox4x32 10 rounds
KAT = [
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


@pytest.mark.parametrize("ctr,key,exp", KAT)
def test_philox_known_answers(lib, ctr, key, exp):
    c = (C.c_uint * 4)(*ctr)
    k = (C.c_uint * 2)(*key)
    o = (C.c_uint * 4)()
    lib.kat(c, k, o)
    assert tuple(o) == exp


def test_draw_range_and_keying(lib):
    lib.draw.argtypes = [C.c_ulonglong, C.c_uint, C.c_uint, C.c_uint, C.c_uint]
    a = [lib.draw(5, 1, 7, 100, k) for k in range(8)]
    assert all(0 <= v <= 2147483647 for v in a)
    assert len(set(a)) == 8
    # every key component matters
    base = lib.draw(5, 1, 7, 100, 0)
    assert lib.draw(6, 1, 7, 100, 0) != base
    assert lib.draw(5, 2, 7, 100, 0) != base
    assert lib.draw(5, 1, 8, 100, 0) != base
    assert lib.draw(5, 1, 7, 101, 0) != base
    assert lib.draw(5 + (1 << 32), 1, 7, 100, 0) != base


def test_msg3_threshold(lib):
    """p > 0.1 <=> r >= 214748361 (float rounding of r)."""
    for r in list(range(214748300, 214748400)) + [0, 1, 2147483647, 214748352, 214748368]:
        assert lib.msg3(r) == lib.msg3_ref(r) == (1 if r >= 214748361 else 0)
    rng = np.random.default_rng(0)
    for r in rng.integers(0, 2**31, 2000):
        assert lib.msg3(int(r)) == (1 if r >= 214748361 else 0)
