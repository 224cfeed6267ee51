"""The role of the reference's AveragePerformance.py (lines 1-24) on top of the GPU engine (the reference's tool is
Python + numpy, so is this one).

    python 5g-nr-randomaccess_b200/average_performance.py --seeds 100 --out results.csv
        run the nUE sweep 10000..100000 (RandomAccessWithNOMA.c:221) for `seeds` replications and write results.csv in
        the reference's layout (10 rows: nUE, success %, #success, mean preamble tx, mean delay ms, cumulative seconds),
        each value the mean over the seeds of the per-replication figure rounded as the reference's result files round
        them ("%.2lf" per file, RandomAccessSimulatorBeta.c:466-480; np.around(mean, 3) at the end,
        AveragePerformance.py:21-24) -- plus <out stem>_ci.csv with the 95 % confidence half-width of every cell
    ... --readme-tables [--seeds 100]
        the three README tables (README.md:91-113: retransmission limit 10 / 20 / 50) as markdown, mean +- 95 % CI
    ... --from-files DIR [--preambles 54]
        the reference's own input path: read DIR/{seed}_{preambles}_{nUE}_Results.txt (6 numbers, one per line, as
        written by RandomAccessSimulatorBeta.c:460-482 and by `rach_sim --format b`) instead of running anything

Column 6 is the cumulative wall time of the sweep in seconds (the reference's clock() column, B:67,205-208,481);
here: the kernel time of the single launch, apportioned to the points by their update counts.
The 95 % CI is the normal-approximation half-width 1.96 * s / sqrt(n) over the seeds (sample standard deviation).
"""
import argparse
import csv
import glob
import importlib
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

NUES = list(range(10000, 110000, 10000))
COLS = ["nUE", "success_pct", "n_success", "mean_preamble_tx", "mean_delay_ms", "cumulative_s"]


def _ci95(x):
    x = np.asarray(x, dtype=np.float64)
    if x.size < 2:
        return 0.0
    return 1.96 * x.std(ddof=1) / np.sqrt(x.size)


def per_seed_values(st, n):
    """Per-replication values as the reference writes them to its result files: "%.2lf" of float32 expressions
    (B:434-438,466-480).  st: structured ra_stats array of one point [seeds]."""
    ns = st["nSuccess"].astype(np.float64)
    ns32 = ns.astype(np.float32)
    ratio = np.round((ns32 / np.float32(n) * 100.0).astype(np.float64), 2)
    with np.errstate(divide="ignore", invalid="ignore"):
        avg_tx = np.round((st["preambleTxSum"].astype(np.float32) / ns32).astype(np.float64), 2)
        avg_delay = np.round((st["delaySum"].astype(np.float32) / ns32).astype(np.float64), 2)
    return ratio, ns, avg_tx, avg_delay


def sweep_table(seeds=100, max_retx=10, grants=12, n_preamble=54, backoff=20, device=0, nues=None, seed64=0,
                with_ci=False):
    """-> (table [points, 6] rounded to 3 decimals, ra_stats [points, seeds]) and, with_ci, the 95 % half-widths."""
    pkg = importlib.import_module("5g-nr-randomaccess_b200")
    nues = list(nues or NUES)
    pts = [pkg.default_params(nUE=n, maxMsg2TxCount=max_retx - 1, nGrantUL=grants, nPreamble=n_preamble,
                              backoffIndicator=backoff, seed=seed64) for n in nues]
    with pkg.RachSim(pts, reps=seeds, devices=[device]) as sim:
        sim.run()
        st = sim.stats_all()
        kernel_s = sim.kernel_ms / 1e3
    rows, cis = [], []
    share = st["updates"].sum(axis=1).astype(np.float64)
    cum = np.cumsum(share / share.sum() * kernel_s)
    for k, n in enumerate(nues):
        ratio, ns, avg_tx, avg_delay = per_seed_values(st[k], n)
        rows.append([float(n), ratio.mean(), ns.mean(), avg_tx.mean(), avg_delay.mean(), cum[k]])
        cis.append([0.0, _ci95(ratio), _ci95(ns), _ci95(avg_tx), _ci95(avg_delay), 0.0])
    table = np.around(np.asarray(rows), 3)
    if with_ci:
        return table, st, np.around(np.asarray(cis), 3)
    return table, st


def table_from_files(directory, n_preamble=54):
    """AveragePerformance.py:9-19: sum the six lines of every {seed}_{P}_{nUE}_Results.txt, divide by the seed count."""
    pat = re.compile(r"^(\d+)_%d_(\d+)_Results\.txt$" % n_preamble)
    by_nue = {}
    for path in glob.glob(os.path.join(directory, "*_Results.txt")):
        m = pat.match(os.path.basename(path))
        if not m:
            continue
        vals = []
        with open(path) as f:
            for line in f.readlines()[:6]:          # B writes six numbers (B:460-482); W five, then labelled lines (W:762-793)
                try:
                    vals.append(float(line.strip()))
                except ValueError:
                    break
        if len(vals) < 5:
            continue
        vals += [0.0] * (6 - len(vals))
        by_nue.setdefault(int(m.group(2)), []).append(vals)
    if not by_nue:
        raise SystemExit("no {seed}_%d_{nUE}_Results.txt files in %s" % (n_preamble, directory))
    rows, cis = [], []
    for n in sorted(by_nue):
        a = np.asarray(by_nue[n], dtype=np.float64)
        rows.append(a.mean(axis=0))
        cis.append([_ci95(a[:, c]) if c not in (0, 5) else 0.0 for c in range(6)])
    return np.around(np.asarray(rows), 3), np.around(np.asarray(cis), 3), {n: len(v) for n, v in by_nue.items()}


def write_csv(path, table):
    with open(path, "w") as f:
        w = csv.writer(f)
        for r in table:
            w.writerow(r)


def readme_tables(seeds=100, device=0, out=sys.stdout):
    """README.md:91-113: one table per retransmission limit, columns = the ten population sizes."""
    labels = [("Success ratio", 1), ("Number of successful devices", 2), ("Number of preamble tx", 3), ("Access delay (ms)", 4)]
    res = {}
    for retx in (10, 20, 50):
        table, _, ci = sweep_table(seeds, max_retx=retx, device=device, with_ci=True)
        res[retx] = (table, ci)
        out.write("#### Retransmission limit: %d   (%d replications per cell, mean +- 95 %% CI)\n" % (retx, seeds))
        out.write("| Number of devices per cell | " + " | ".join("{:,}".format(n) for n in NUES) + " |\n")
        out.write("|---|" + "---|" * len(NUES) + "\n")
        for name, c in labels:
            out.write("| %s | " % name + " | ".join("%.3f +- %.3f" % (table[k, c], ci[k, c]) for k in range(len(NUES))) + " |\n")
        out.write("\n")
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", type=int, default=100)
    ap.add_argument("--retx", type=int, default=10)
    ap.add_argument("--grants", type=int, default=12)
    ap.add_argument("--preambles", type=int, default=54)
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--out", default="results.csv")
    ap.add_argument("--readme-tables", action="store_true")
    ap.add_argument("--from-files", metavar="DIR")
    a = ap.parse_args()
    if a.readme_tables:
        readme_tables(a.seeds, a.device)
        return
    if a.from_files:
        table, ci, counts = table_from_files(a.from_files, a.preambles)
        print("seeds per point: %s" % counts, file=sys.stderr)
    else:
        table, _, ci = sweep_table(a.seeds, a.retx, a.grants, n_preamble=a.preambles, device=a.device, with_ci=True)
    write_csv(a.out, table)
    stem, ext = os.path.splitext(a.out)
    write_csv(stem + "_ci" + ext, ci)
    print(open(a.out).read())
    print("95 %% CI half-widths (%s):" % (stem + "_ci" + ext))
    print(open(stem + "_ci" + ext).read())


if __name__ == "__main__":
    main()
