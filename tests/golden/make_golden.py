#!/usr/bin/env python
"""Generate tests/golden/*.json|npz from the REFERENCE SOURCES in draw-tape mode.

Run here (where /root/reference exists):   python tests/golden/make_golden.py
It builds oracle/_ref/libref_{w,b}.so from /root/reference (oracle/build_ref.sh), runs every
case below through the reference's own main loop and stores
  * golden_stats.json   per case: the configuration, every counter, the floats the
                        reference printed, and sha256 of the per-UE (nUE x 16 int32) dump
                        and of the geometry (nUE x 6 float32) side outputs;
  * golden_ues.npz      the full per-UE dumps of the small cases (<= 3000 UEs).
The fixtures are what the GPU box (no /root/reference there) checks the oracle restatement
and the CUDA engine against.
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

W = dict(geometry=1)
B = dict(geometry=0)
CASES = [
    # name, variant, overrides
    ("w_default_2000", "w", dict(W, nUE=2000)),
    ("w_default_3000_seed7", "w", dict(W, nUE=3000, seed=7, rep=3)),
    ("w_lone_ue", "w", dict(W, nUE=1)),
    ("w_two_ues", "w", dict(W, nUE=2, nPreamble=1)),
    ("w_tiny_p3_g2", "w", dict(W, nUE=300, nPreamble=3, nGrantUL=2, seed=11)),
    ("w_bi1_immediate", "w", dict(W, nUE=1500, backoffIndicator=1, seed=5)),
    ("w_mrc1_limit_always", "w", dict(W, nUE=1500, maxMsg2TxCount=0, seed=6)),
    ("w_rc1_window2", "w", dict(W, nUE=1500, maxRarWindow=2, seed=8)),
    ("w_subframe7", "w", dict(W, nUE=2500, accessTime=7, seed=9)),
    ("w_uniform_3000", "w", dict(W, nUE=3000, distribution=1, seed=2)),
    ("w_p64_g4_bi40", "w", dict(W, nUE=3000, nPreamble=64, nGrantUL=4, backoffIndicator=40, seed=4)),
    ("b_default_3000", "b", dict(B, nUE=3000, nGrantUL=54)),
    ("b_g12_3000", "b", dict(B, nUE=3000, nGrantUL=12, seed=21)),
    # larger: stats + hashes only
    ("w_default_10000", "w", dict(W, nUE=10000)),                      # BASELINE configs[0]
    ("w_default_20000", "w", dict(W, nUE=20000, seed=1)),
    ("w_default_30000_retx20", "w", dict(W, nUE=30000, maxMsg2TxCount=19, seed=2)),
    ("w_uniform_20000", "w", dict(W, nUE=20000, distribution=1, seed=3)),
    ("w_p64_g16_bi10_20000", "w", dict(W, nUE=20000, nPreamble=64, nGrantUL=16, backoffIndicator=10, seed=4)),
    ("b_default_10000", "b", dict(B, nUE=10000, nGrantUL=54)),         # B as shipped
    ("b_g12_20000", "b", dict(B, nUE=20000, nGrantUL=12, seed=5)),
    # README retx-50 table (README.md:110-113: -mrc 50 -> maxMsg2TxCount 49, W:134)
    ("w_retx50_20000", "w", dict(W, nUE=20000, maxMsg2TxCount=49, seed=6)),
    ("w_retx50_30000", "w", dict(W, nUE=30000, maxMsg2TxCount=49, seed=7)),
    ("b_retx50_g12_20000", "b", dict(B, nUE=20000, nGrantUL=12, maxMsg2TxCount=49, seed=8)),
    # the headline size, upper end of the reference's sweep (W:221); ~14 min of the reference's O(nUE) scans
    ("w_default_100000", "w", dict(W, nUE=100000)),
]

N_CASES = [
    ("n_default_3000", dict(nUE=3000)),
    ("n_default_10000", dict(nUE=10000, seed=1)),                      # NOMA.c as shipped: 98.9 %
    ("n_default_30000", dict(nUE=30000, seed=2)),                      # 73.4 %
    ("n_g1_p8_4000", dict(nUE=4000, nGrantUL=1, nPreamble=8, seed=3)),
    ("n_g4_bi40_sub10_8000", dict(nUE=8000, nGrantUL=4, backoffIndicator=40, accessTime=10, seed=4)),
    ("n_retx3_r100_6000", dict(nUE=6000, maxMsg2TxCount=3, cellRadius=100.0, seed=5)),
    ("n_nonsector_8000", dict(nUE=8000, geometry=0, seed=6)),           # NOMA.c:325-447 switched in (N:688)
    ("n_nonsector_g12_20000", dict(nUE=20000, geometry=0, nGrantUL=12, seed=7)),   # with the commented nGrantUL = 12 (N:687)
    ("n_default_50000", dict(nUE=50000, seed=8)),                      # BASELINE configs[3] size (N:648 sweep point)
    ("n_default_100000", dict(nUE=100000, seed=9)),                    # upper end of NOMA.c's sweep
]

U0_CASES = [
    ("u0_default_3000", dict(nUE=3000)),
    ("u0_default_30000", dict(nUE=30000, seed=1)),                   # RandomAccessSimulator.c as shipped (64 preambles)
    ("u0_p1_overload_40000", dict(nUE=40000, nPreamble=1, seed=2)),   # collisions, drops, phantom colliders
    ("u0_p1_bi40_3000", dict(nUE=3000, nPreamble=1, backoffIndicator=40, seed=3)),
]

STAT_KEYS = ["simTimeMs", "nSuccess", "preambleTxSum", "delaySum", "failCountSum",
             "continueFailed", "collisionPreambles", "totalPreambleTxop", "collisionScans",
             "totalScans", "draws", "maxDrawsPerUeMs", "nAccessUE", "averageDelay",
             "averagePreambleTx", "ratioSuccess"]


def main():
    """python make_golden.py                regenerate everything
       python make_golden.py --only a,b,c   run only those cases and merge them into the committed files"""
    here = os.path.dirname(os.path.abspath(__file__))
    only = None
    if "--only" in sys.argv:
        only = set(sys.argv[sys.argv.index("--only") + 1].split(","))
    O.build(force=only is None)
    stats, ues = {}, {}
    if only is not None:
        with open(os.path.join(here, "golden_stats.json")) as f:
            stats = json.load(f)
        ues = dict(np.load(os.path.join(here, "golden_ues.npz")))
    sel = lambda name: only is None or name in only
    for name, variant, kw in CASES:
        if not sel(name):
            continue
        cfg = O.make_config(**kw)
        res, ue, geom = O.run_ref(variant, cfg, per_ue=True, geom=(variant == "w"))
        d = res.as_dict()
        entry = {"variant": variant, "config": {k: kw.get(k, O.DEFAULTS[k]) for k in O.DEFAULTS},
                 "stats": {k: d[k] for k in STAT_KEYS},
                 "ue_sha256": hashlib.sha256(np.ascontiguousarray(ue).tobytes()).hexdigest()}
        if variant == "w":
            entry["geom_sha256"] = hashlib.sha256(np.ascontiguousarray(geom).tobytes()).hexdigest()
        stats[name] = entry
        if cfg.nUE <= 3000:
            ues[name] = ue.astype(np.int16) if np.abs(ue).max() < 32768 else ue
        print("%-28s %s" % (name, {k: d[k] for k in ("simTimeMs", "nSuccess", "preambleTxSum", "delaySum")}))
    for name, kw in N_CASES:
        if not sel(name):
            continue
        cfg = O.make_config_n(**kw)
        res, ue, gain = O.run_ref_n(cfg)
        full = dict(O.DEFAULTS); full.update(O.N_DEFAULTS); full.update(kw)
        stats[name] = {"variant": "n", "config": full,
                       "stats": {k: getattr(res, k) for k in ("nSuccess", "preambleTxSum", "delaySum", "draws", "maxDrawsPerUeMs")},
                       "ue_sha256": hashlib.sha256(np.ascontiguousarray(ue).tobytes()).hexdigest(),
                       "gain_sha256": hashlib.sha256(np.ascontiguousarray(gain).tobytes()).hexdigest(),
                       "dropped": int((ue[:, 15] > 0).sum())}
        if cfg.nUE <= 3000:
            ues[name] = ue.astype(np.int16) if np.abs(ue).max() < 32768 else ue
        print("%-28s %s" % (name, stats[name]["stats"]))
    for name, kw in U0_CASES:
        if not sel(name):
            continue
        cfg = O.make_config_u0(**kw)
        res, ue = O.run_ref_u0(cfg)
        full = dict(O.DEFAULTS); full.update(O.U0_DEFAULTS); full.update(kw)
        stats[name] = {"variant": "u0", "config": full,
                       "stats": {k: getattr(res, k) for k in ("simTimeMs", "nSuccess", "preambleTxSum", "delaySum",
                                                              "collisionPreambles", "totalPreambleTxop", "draws")},
                       "ue_sha256": hashlib.sha256(np.ascontiguousarray(ue).tobytes()).hexdigest(),
                       "dropped": int((ue[:, 11] == -1).sum())}
        if cfg.nUE <= 3000:
            ues[name] = ue.astype(np.int16) if np.abs(ue).max() < 32768 else ue
        print("%-28s %s dropped %d" % (name, stats[name]["stats"], stats[name]["dropped"]))
    with open(os.path.join(here, "golden_stats.json"), "w") as f:
        json.dump(stats, f, indent=1, sort_keys=True)
    np.savez_compressed(os.path.join(here, "golden_ues.npz"), **ues)


if __name__ == "__main__":
    main()
