"""Smallest program that launches one step kernel at a chosen workload (ncu target / quick timing).
   --variant w|n|u0   --reps R --nue N --distribution 1|2   --runs K (timing of the last run is printed)"""
import argparse, importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("5g-nr-randomaccess_b200")
ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=592); ap.add_argument("--nue", type=int, default=100000)
ap.add_argument("--ctas-per-sm", type=int, default=0); ap.add_argument("--runs", type=int, default=1); ap.add_argument("--distribution", type=int, default=2)
ap.add_argument("--variant", default="w", choices=["w", "n", "u0"])
ap.add_argument("--retx", type=int, default=0, help="max retransmissions (README tables: 10, 20, 50); 0 = variant default")
a = ap.parse_args()
kw = dict(nUE=a.nue)
if a.variant == "n":
    kw["variant"] = 2
elif a.variant == "u0":
    kw["variant"] = 1
else:
    kw["distribution"] = a.distribution
if a.retx:
    kw["maxMsg2TxCount"] = a.retx - 1 if a.variant == "w" else a.retx
p = pkg.default_params(**kw)
with pkg.RachSim([p], reps=a.reps, devices=[0], ctas_per_sm=a.ctas_per_sm) as sim:
    for _ in range(a.runs):
        sim.run()
    st = sim.stats_all()
    print("variant %s nue %d reps %d kernel_ms %.3f updates/s %.4e reps/s %.1f" % (
        a.variant, a.nue, a.reps, sim.kernel_ms, float(st["updates"].sum()) / sim.kernel_ms * 1e3, a.reps / sim.kernel_ms * 1e3))
