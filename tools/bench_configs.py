"""Timing of the secondary BASELINE configs (not the headline bench): Uniform 100k x 256 (configs[1]),
variant N 50k x 1024 (configs[3])."""
import importlib, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("5g-nr-randomaccess_b200")
out = {}
for name, p, reps in (("uniform_100k_x256", pkg.default_params(nUE=100000, distribution=1), 256),
                      ("noma_N_50k_x1024", pkg.default_params(variant=2, nUE=50000), 1024),
                      ("w_geometry_50k_x1024", pkg.default_params(nUE=50000, cellRadius=400.0), 1024),
                      ("legacy_U0_100k_x1024", pkg.default_params(variant=1, nUE=100000), 1024)):
    with pkg.RachSim([p], reps=reps, devices=[0]) as sim:
        sim.run(); sim.run()
        st = sim.stats_all()
        out[name] = {"kernel_ms": sim.kernel_ms, "updates_per_s": float(st["updates"].sum()) / sim.kernel_ms * 1e3,
                     "reps_per_s": reps / sim.kernel_ms * 1e3, "success_pct": 100.0 * float(st["nSuccess"].sum()) / (reps * p.nUE),
                     "mean_tx": float(st["preambleTxSum"].sum()) / max(float(st["nSuccess"].sum()), 1),
                     "mean_delay_ms": float(st["delaySum"].sum()) / max(float(st["nSuccess"].sum()), 1),
                     "mean_simTime_ms": float(st["simTimeMs"].mean())}
print(json.dumps(out, indent=1))
