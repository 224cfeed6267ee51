"""the reference driven by its own libc rand() and the same reference driven by the Philox draw
tape are two samples of one distribution.  10 seeds each of RandomAccessWithNOMA.c at 10 500 UEs (loaded but not
saturated, so every statistic has visible spread); means must agree within 4 standard errors."""
import numpy as np
import pytest


def test_rand_and_tape_builds_agree_statistically(oracle):
    if not oracle.ref_available("w"):
        pytest.skip("oracle/_ref not built")
    n, runs = 10500, 10
    out = {0: [], 1: []}
    for tape in (0, 1):
        for seed in range(runs):
            cfg = oracle.make_config(nUE=n, seed=1000 + seed, rep=seed, useTape=tape)
            r, _, _ = oracle.run_ref("w", cfg, per_ue=False)
            out[tape].append((100.0 * r.nSuccess / n, r.preambleTxSum / r.nSuccess, r.delaySum / r.nSuccess,
                              r.collisionPreambles / 1e5))
    a, b = np.asarray(out[0]), np.asarray(out[1])
    for k, name in enumerate(("success %", "preamble tx", "delay ms", "collided preambles / 1e5")):
        se = np.sqrt(a[:, k].var(ddof=1) / runs + b[:, k].var(ddof=1) / runs) + 1e-9
        assert abs(a[:, k].mean() - b[:, k].mean()) <= 4.0 * se + 1e-6 * abs(a[:, k].mean()), \
            (name, a[:, k].mean(), b[:, k].mean(), se)
