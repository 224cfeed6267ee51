"""Per-phase cycle breakdown of the step kernel (thread 0 of every block), for profiles/."""
import argparse, importlib, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("5g-nr-randomaccess_b200")
ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=592); ap.add_argument("--nue", type=int, default=100000)
ap.add_argument("--ctas-per-sm", type=int, default=0); ap.add_argument("--no-timers", action="store_true"); ap.add_argument("--distribution", type=int, default=2)
a = ap.parse_args()
p = pkg.default_params(nUE=a.nue, distribution=a.distribution)
with pkg.RachSim([p], reps=a.reps, devices=[0], ctas_per_sm=a.ctas_per_sm, phase_timers=not a.no_timers) as sim:
    sim.run(); sim.run()
    cyc = sim.phase_cycles().astype(float)
    names = ["0 setup", "1 events", "2 c3", "3 uncertain", "4 late", "5 scans", "6 grants", "7 apply", "8", "9"]
    tot = cyc.sum()
    st = sim.stats_all()
    print(json.dumps({"reps": a.reps, "kernel_ms": sim.kernel_ms, "updates_per_s": float(st["updates"].sum()) / sim.kernel_ms * 1e3,
                      "phase_share": {n: round(c / tot, 4) for n, c in zip(names, cyc) if c},
                      "cycles_per_rep_M": {n: round(c / a.reps / 1e6, 2) for n, c in zip(names, cyc) if c}}))
