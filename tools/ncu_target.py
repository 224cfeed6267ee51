"""Smallest program that launches the step kernel once at full occupancy (ncu target)."""
import argparse, importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("5g-nr-randomaccess_b200")
ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=592); ap.add_argument("--nue", type=int, default=100000)
ap.add_argument("--ctas-per-sm", type=int, default=0); ap.add_argument("--runs", type=int, default=1); ap.add_argument("--distribution", type=int, default=2)
a = ap.parse_args()
p = pkg.default_params(nUE=a.nue, distribution=a.distribution)
with pkg.RachSim([p], reps=a.reps, devices=[0], ctas_per_sm=a.ctas_per_sm) as sim:
    for _ in range(a.runs):
        sim.run()
    st = sim.stats_all()
    print("kernel_ms %.2f updates/s %.4e" % (sim.kernel_ms, float(st["updates"].sum()) / sim.kernel_ms * 1e3))
