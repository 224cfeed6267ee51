"""Statistical agreement with what the reference publishes (README.md:91-113, results.csv:1-10):
100-seed means of the stock rand() binary.  Our replications use the Philox tape, so agreement is
statistical: each published mean must lie within a few standard errors (of our own spread over
256 replications, plus the published figure's own 100-seed error) and within 1 % relative."""
import importlib
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# README.md:94-97 / :102-105 / :110-113  (success %, preamble tx, delay ms) per retransmission limit
README = {
    10: {10000: (100, 2.59, 47.114), 20000: (89.292, 5.358, 89.65), 30000: (60.928, 5.523, 92.254),
         50000: (37.252, 5.651, 94.289), 100000: (18.989, 5.76, 96.001)},
    20: {10000: (100, 2.561, 45.366), 20000: (89.548, 9.6, 155.95), 30000: (60.912, 10.062, 162.953),
         50000: (37.284, 10.371, 168.099), 100000: (18.993, 10.65, 172.233)},
    50: {20000: (89.451, 22.279, 347.504), 30000: (60.95, 23.34, 363.314), 50000: (37.298, 24.342, 379.32),
         100000: (19.012, 25.22, 392.61)},
}


@pytest.mark.parametrize("retx", [10, 20, 50])
def test_readme_tables(retx):
    pkg = importlib.import_module("5g-nr-randomaccess_b200")
    nues = sorted(README[retx])
    reps = 256
    pts = [pkg.default_params(nUE=n, maxMsg2TxCount=retx - 1, seed=2024) for n in nues]
    with pkg.RachSim(pts, reps=reps, devices=[0]) as sim:
        sim.run()
        st = sim.stats_all()
    for k, n in enumerate(nues):
        ns = st[k]["nSuccess"].astype(np.float64)
        ratio = 100.0 * ns / n
        tx = st[k]["preambleTxSum"] / ns
        delay = st[k]["delaySum"] / ns
        for name, ours, pub in (("success %", ratio, README[retx][n][0]), ("preamble tx", tx, README[retx][n][1]),
                                ("delay ms", delay, README[retx][n][2])):
            mean, sd = ours.mean(), ours.std(ddof=1)
            # our standard error + the published mean's own (100 seeds, same spread) ; 4 sigma
            # (the retx-50 table is visibly noisier than 100 seeds would give -- 24.671 @60k vs 24.631 @70k -- so
            #  never tighter than 1 % of the published figure)
            tol = max(4.0 * sd * np.sqrt(1.0 / reps + 1.0 / 100.0) + 0.002 * abs(pub), 0.01 * abs(pub)) + 1e-9
            if n == 10000:
                # the README contradicts itself at 10 000 UEs, where the limit is never reached and the three
                # tables should agree: tx 2.59 / 2.561 / 3.541, delay 47.1 / 45.4 / 58.8 -> 5 % is all it pins
                assert abs(mean - pub) <= 0.05 * abs(pub), (retx, n, name, mean, pub)
                continue
            assert abs(mean - pub) <= tol, (retx, n, name, mean, pub, tol)
            assert abs(mean - pub) <= 0.012 * abs(pub) + 0.05, (retx, n, name, mean, pub)


def test_results_csv_layout(tmp_path):
    """results.csv regenerated in the reference's layout (AveragePerformance.py:21-24) agrees with the
    published file on columns 2-5 (column 6 is wall time on other hardware)."""
    ap = importlib.import_module("5g-nr-randomaccess_b200.average_performance")
    table, _ = ap.sweep_table(seeds=64, nues=[10000, 30000, 60000, 100000])
    pub = {10000: (100.0, 10000.0, 2.59, 47.114), 30000: (60.913, 18273.85, 5.523, 92.254),
           60000: (31.263, 18758.26, 5.686, 94.849), 100000: (18.99, 18989.29, 5.76, 96.001)}   # results.csv:1,3,6,10
    assert table.shape == (4, 6)
    for row in table:
        p = pub[int(row[0])]
        for ours, ref in zip(row[1:5], p):
            assert abs(ours - ref) <= 0.012 * abs(ref) + 0.05, (row, p)
    assert (np.diff(table[:, 5]) > 0).all()
