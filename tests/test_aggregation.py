"""average_performance.py (the role of the reference's AveragePerformance.py:1-24): results.csv layout from result
files in the reference's format, 95 % confidence half-widths, and -- on the GPU -- the same table from the host
CLI's own files and from a direct sweep."""
import importlib
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ap():
    return importlib.import_module("5g-nr-randomaccess_b200.average_performance")


def test_table_from_reference_format_files(tmp_path):
    """Files as RandomAccessSimulatorBeta.c:460-482 writes them: nUE, ratio %.2lf, nSuccess, tx %.2lf, delay %.2lf, latency."""
    vals = {}
    for n in (10000, 20000, 30000):
        for seed in range(7):
            v = [n, 90.0 + seed, n - 10 * seed, 5.30 + 0.01 * seed, 89.0 + seed, 1.5 * (seed + 1)]
            vals.setdefault(n, []).append(v)
            (tmp_path / ("%d_54_%d_Results.txt" % (seed, n))).write_text("%d\n%.2f\n%d\n%.2f\n%.2f\n%f" % tuple(v))
    (tmp_path / "0_54_UE10000_Logs.txt").write_text("not a result file")
    # a W-format result file (W:762-793: five numbers, then labelled lines) next to the B-format ones: cumulative time 0
    wdir = tmp_path / "w"
    wdir.mkdir()
    (wdir / "0_54_10000_Results.txt").write_text("10000\n100.00\n10000\n2.60\n47.10\nNumber of total preamble tx: 26000\n"
                                                 "Finally Falied: 12\nFinally Success: 0.998801\n")
    tw, _, cw = _ap().table_from_files(str(wdir), 54)
    assert cw == {10000: 1} and list(tw[0]) == [10000.0, 100.0, 10000.0, 2.6, 47.1, 0.0]
    (tmp_path / "3_64_10000_Results.txt").write_text("1\n2\n3\n4\n5\n6")        # another preamble count: ignored
    table, ci, counts = _ap().table_from_files(str(tmp_path), 54)
    assert counts == {10000: 7, 20000: 7, 30000: 7} and table.shape == (3, 6)
    for k, n in enumerate((10000, 20000, 30000)):
        a = np.asarray(vals[n])
        np.testing.assert_allclose(table[k], np.around(a.mean(axis=0), 3))          # AveragePerformance.py:21-24
        np.testing.assert_allclose(ci[k, 1], np.around(1.96 * a[:, 1].std(ddof=1) / np.sqrt(7), 3))
        assert ci[k, 0] == 0 and ci[k, 5] == 0
    out = tmp_path / "results.csv"
    rc = subprocess.run([sys.executable, os.path.join(ROOT, "5g-nr-randomaccess_b200", "average_performance.py"),
                         "--from-files", str(tmp_path), "--out", str(out)], capture_output=True, text=True)
    assert rc.returncode == 0, rc.stderr
    rows = [[float(x) for x in line.split(",")] for line in out.read_text().strip().splitlines()]
    np.testing.assert_allclose(np.asarray(rows), table)
    assert (tmp_path / "results_ci.csv").exists()


@pytest.mark.gpu
def test_cli_files_and_direct_sweep_agree(tmp_path):
    """`rach_sim --format b` writes {seed}_54_{nUE}_Results.txt; averaging those files gives the same columns 1-5 as
    the direct sweep (same tape key and replication ids), and every cell carries a confidence interval."""
    pkg = importlib.import_module("5g-nr-randomaccess_b200")
    exe = pkg.build_host()
    seeds, nues = 12, [10000, 20000, 40000]
    rc = subprocess.run([exe, "--format", "b", "-g", "12", "-t", str(seeds), "--nue", ",".join(map(str, nues)), "--no-logs",
                         "--outdir", str(tmp_path)], capture_output=True, text=True)
    assert rc.returncode == 0, rc.stderr
    ap = _ap()
    t_files, ci_files, counts = ap.table_from_files(str(tmp_path / "BasicBetaSimulationResults"), 54)
    assert counts == {n: seeds for n in nues}
    # the B format has no activation draws (geometry off): the direct sweep must use the same tape consumption
    pts = [pkg.default_params(nUE=n, geometry=0) for n in nues]
    with pkg.RachSim(pts, reps=seeds, devices=[0]) as sim:
        sim.run()
        st = sim.stats_all()
    for k, n in enumerate(nues):
        ratio, ns, tx, delay = ap.per_seed_values(st[k], n)
        np.testing.assert_allclose(t_files[k, 1:5], np.around([ratio.mean(), ns.mean(), tx.mean(), delay.mean()], 3), atol=2e-3)
    assert (ci_files[1:, 1] > 0).all() and (ci_files[:, [0, 5]] == 0).all()


@pytest.mark.gpu
def test_readme_tables_with_confidence_intervals():
    """README.md:91-113 regenerated (16 replications per cell here): three tables, every cell mean +- CI, and the
    published 100-seed means inside a generous band around ours."""
    import io
    buf = io.StringIO()
    res = _ap().readme_tables(seeds=16, out=buf)
    txt = buf.getvalue()
    assert txt.count("#### Retransmission limit") == 3 and txt.count(" +- ") == 3 * 4 * 10 + 3      # every cell, and once in each heading
    readme_100k = {10: (18.989, 5.76, 96.001), 20: (18.993, 10.65, 172.233), 50: (19.012, 25.22, 392.61)}
    for retx, (ratio, tx, delay) in readme_100k.items():
        table, ci = res[retx]
        assert abs(table[9, 1] - ratio) < 4 * ci[9, 1] + 0.3
        assert abs(table[9, 3] - tx) < 4 * ci[9, 3] + 0.02 * tx
        assert abs(table[9, 4] - delay) < 4 * ci[9, 4] + 0.02 * delay
