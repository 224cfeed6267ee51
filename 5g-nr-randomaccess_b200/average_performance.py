"""The role of the reference's AveragePerformance.py (lines 1-24) on top of the GPU engine:
run the nUE sweep 10000..100000 for `seeds` replications and write results.csv in the same layout
(10 rows: nUE, success %, #success, mean preamble tx, mean delay ms, cumulative seconds), each value
the mean over the seeds of the per-replication figure rounded as the reference's result files
round them (%.2lf per file, RandomAccessSimulatorBeta.c:466-480; np.around(mean, 3) at the end).

    python -m importlib ... or:  python 5g-nr-randomaccess_b200/average_performance.py --seeds 100 --out results.csv

Column 6 is the cumulative wall time of the sweep in seconds (the reference's clock() column);
here: the kernel time of the single launch, apportioned to the points by their update counts.
"""
import argparse
import csv
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def sweep_table(seeds=100, max_retx=10, grants=12, n_preamble=54, backoff=20, device=0, nues=None, seed64=0):
    pkg = importlib.import_module("5g-nr-randomaccess_b200")
    nues = list(nues or range(10000, 110000, 10000))
    pts = [pkg.default_params(nUE=n, maxMsg2TxCount=max_retx - 1, nGrantUL=grants, nPreamble=n_preamble,
                              backoffIndicator=backoff, seed=seed64) for n in nues]
    with pkg.RachSim(pts, reps=seeds, devices=[device]) as sim:
        sim.run()
        st = sim.stats_all()
        kernel_s = sim.kernel_ms / 1e3
    rows = []
    share = st["updates"].sum(axis=1).astype(np.float64)
    cum = np.cumsum(share / share.sum() * kernel_s)
    for k, n in enumerate(nues):
        s = st[k]
        ns = s["nSuccess"].astype(np.float64)
        # per-file values as the reference writes them: "%.2lf" of float32 expressions (B:434-438)
        ratio = np.round((ns.astype(np.float32) / np.float32(n) * 100.0).astype(np.float64), 2)
        avg_tx = np.round((s["preambleTxSum"].astype(np.float32) / ns.astype(np.float32)).astype(np.float64), 2)
        avg_delay = np.round((s["delaySum"].astype(np.float32) / ns.astype(np.float32)).astype(np.float64), 2)
        rows.append([float(n), ratio.mean(), ns.mean(), avg_tx.mean(), avg_delay.mean(), cum[k]])
    return np.around(np.asarray(rows), 3), st


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", type=int, default=100)
    ap.add_argument("--retx", type=int, default=10)
    ap.add_argument("--grants", type=int, default=12)
    ap.add_argument("--out", default="results.csv")
    a = ap.parse_args()
    table, _ = sweep_table(a.seeds, a.retx, a.grants)
    with open(a.out, "w") as f:
        w = csv.writer(f)
        for r in table:
            w.writerow(r)
    print(open(a.out).read())


if __name__ == "__main__":
    main()
