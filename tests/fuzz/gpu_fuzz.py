"""One-off wide parity fuzz on the GPU box: random sizes and parameters, W/B dynamics, every UE compared
with the oracle restatement.  python tests/fuzz/gpu_fuzz.py SECONDS SEED  (test infrastructure: the oracle is the checker)"""
import importlib, os, random, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
pkg = importlib.import_module("5g-nr-randomaccess_b200")
from oracle import oracle as O
budget, seed = float(sys.argv[1]), int(sys.argv[2])
rnd = random.Random(seed)
KEYS = ["simTimeMs", "nSuccess", "preambleTxSum", "delaySum", "failCountSum", "continueFailed", "collisionPreambles",
        "totalPreambleTxop", "collisionScans", "totalScans"]
t0 = time.time(); n = bad = 0
while time.time() - t0 < budget:
    kw = dict(nUE=rnd.randrange(1000, 160000), distribution=rnd.choice([1, 2, 2, 2]),
              nPreamble=rnd.choice([8, 32, 54, 54, 64, 128]), backoffIndicator=rnd.choice([5, 10, 20, 20, 40, 80]),
              nGrantUL=rnd.choice([2, 4, 8, 12, 12, 16, 54]), maxRarWindow=rnd.choice([3, 6, 6, 6, 11]),
              maxMsg2TxCount=rnd.choice([2, 9, 9, 19, 49]), accessTime=rnd.choice([5, 5, 5, 6, 8, 10]),
              seed=rnd.getrandbits(64), geometry=rnd.choice([0, 1]))
    if rnd.random() < 0.5:                       # the reference's default family: the compile-time point view
        kw.update(nPreamble=54, backoffIndicator=20, maxRarWindow=6, accessTime=5)
    shape, fixed = rnd.choice(["", "", "small", "big", "huge"]), rnd.choice(["1", "1", "0"])
    os.environ.pop("RACH_BLOCK", None)
    if shape:
        os.environ["RACH_BLOCK"] = shape
    os.environ["RACH_FIXED"] = fixed
    rep = rnd.randrange(100000)
    res, ue_ref, _ = O.run_port(O.make_config(rep=rep, **kw))
    with pkg.RachSim([pkg.default_params(**kw)], reps=1, devices=[0], rep_offset=rep, dump_ues=True) as sim:
        sim.run()
        st, ue = sim.stats(0, 0), sim.dump_ues(0, 0)
    diff = [k for k in KEYS if getattr(st, k) != getattr(res, k)]
    nd = int((ue != ue_ref).any(axis=1).sum())
    n += 1
    if diff or nd:
        bad += 1
        print("MISMATCH", kw, rep, shape, fixed, diff, nd, flush=True)
print("gpu fuzz: %d cases, %d bad, %.0f s" % (n, bad, time.time() - t0))
