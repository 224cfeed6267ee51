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
# BASELINE configs[2]: README sweep, Beta, nUE 5k..100k x maxRetx {10,20,50}, 1024 replications per point (33 points)
nues = [5000] + list(range(10000, 100001, 10000))
pts = [pkg.default_params(nUE=n, maxMsg2TxCount=r - 1) for r in (10, 20, 50) for n in nues]
with pkg.RachSim(pts, reps=1024, devices=[0]) as sim:
    sim.run()
    st = sim.stats_all()
    out["readme_sweep_33pts_x1024"] = {"kernel_ms": sim.kernel_ms, "updates_per_s": float(st["updates"].sum()) / sim.kernel_ms * 1e3,
                                       "replications": int(st.size), "reps_per_s": st.size / sim.kernel_ms * 1e3}
# BASELINE configs[4]: grid P {54,64} x G {4,8,12,16} x BI {10,20,40} at 100k UEs (24 points); 256 replications per point here
pts = [pkg.default_params(nUE=100000, nPreamble=P, nGrantUL=G, backoffIndicator=B) for P in (54, 64) for G in (4, 8, 12, 16) for B in (10, 20, 40)]
with pkg.RachSim(pts, reps=256, devices=[0]) as sim:
    sim.run()
    st = sim.stats_all()
    out["grid_24pts_x256"] = {"kernel_ms": sim.kernel_ms, "updates_per_s": float(st["updates"].sum()) / sim.kernel_ms * 1e3,
                              "replications": int(st.size), "reps_per_s": st.size / sim.kernel_ms * 1e3,
                              "success_pct_by_point": [round(100.0 * float(st[k]["nSuccess"].mean()) / 100000, 3) for k in range(len(pts))]}
print(json.dumps(out, indent=1))
