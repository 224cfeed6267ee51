"""One-off wide fuzz of the engine's phase functions on the HOST (tests/emu build of rach_core.cuh, test infrastructure)
against the oracle restatement: random parameters, sizes and emulated block sizes, every UE compared.
    python tests/fuzz/emu_fuzz.py SECONDS SEED"""
import ctypes as C, os, random, subprocess, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O
O.build()
so = subprocess.check_output([os.path.join(ROOT, "tests", "emu", "build_emu.sh")]).decode().strip()
f = O._lib(so, "emu_run")
thr = C.c_int.in_dll(C.CDLL(so), "emu_threads")
fixed = C.c_int.in_dll(C.CDLL(so), "emu_use_fixed")
KEYS = ["simTimeMs", "nSuccess", "preambleTxSum", "delaySum", "failCountSum", "continueFailed", "collisionPreambles",
        "totalPreambleTxop", "collisionScans", "totalScans"]
budget, seed = float(sys.argv[1]), int(sys.argv[2])
rnd = random.Random(seed)
t0 = time.time(); n = bad = 0
while time.time() - t0 < budget:
    kw = dict(nUE=rnd.choice([1, 2, 9, 60, 400, 2000, 6000, 12000, 25000]), distribution=rnd.choice([1, 2, 2, 2]),
              nPreamble=rnd.choice([1, 2, 3, 8, 32, 54, 54, 64, 128]), backoffIndicator=rnd.choice([1, 2, 5, 10, 20, 20, 40, 80]),
              nGrantUL=rnd.choice([1, 2, 4, 8, 12, 12, 16, 54]), maxRarWindow=rnd.choice([2, 3, 6, 6, 6, 9, 11, 40]),
              maxMsg2TxCount=rnd.choice([0, 1, 2, 9, 9, 19, 49]), accessTime=rnd.choice([1, 2, 3, 5, 5, 5, 6, 8, 10]),
              seed=rnd.getrandbits(64), rep=rnd.randrange(100000), geometry=rnd.choice([0, 1]),
              stopMs=rnd.choice([0, 0, 0, 777, 3001, 6000]))
    if rnd.random() < 0.5:                       # the reference's default family: the compile-time point view
        kw.update(nPreamble=54, backoffIndicator=20, maxRarWindow=6, accessTime=5)
    thr.value = rnd.choice([32, 64, 128, 128, 256, 512])
    fixed.value = rnd.choice([1, 1, 0])
    cfg = O.make_config(**kw)
    p, ue, _ = O.run_port(cfg)
    e, ue2, _ = O._run(f, cfg, True, False)
    diff = [k for k in KEYS if getattr(p, k) != getattr(e, k)]
    nd = int((ue != ue2).any(axis=1).sum())
    n += 1
    if diff or nd:
        bad += 1
        print("MISMATCH", kw, thr.value, fixed.value, diff, nd, flush=True)
print("emu fuzz: %d cases, %d bad, %.0f s" % (n, bad, time.time() - t0))
