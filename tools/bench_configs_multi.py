"""BASELINE configs[2] and configs[4] at their full sizes through the library's own multi-device path
(ra_sim_create(devices[]), one host thread): python tools/bench_configs_multi.py [ndev]"""
import importlib, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("5g-nr-randomaccess_b200")
ndev = int(sys.argv[1]) if len(sys.argv) > 1 else torch.cuda.device_count()
devs = list(range(ndev))
out = {"devices": ndev}
# configs[2]: README sweep, Beta, nUE 5k..100k x max retx {10, 20, 50}, 1024 replications per point (33 points)
nues = [5000] + list(range(10000, 100001, 10000))
pts = [pkg.default_params(nUE=n, maxMsg2TxCount=r - 1) for r in (10, 20, 50) for n in nues]
t = time.perf_counter()
with pkg.RachSim(pts, reps=1024, devices=devs) as sim:
    sim.run(); sim.run()
    st = sim.stats_all()
    out["readme_sweep_33pts_x1024"] = {"kernel_ms": sim.kernel_ms, "updates_per_s": float(st["updates"].sum()) / sim.kernel_ms * 1e3,
                                       "replications": int(st.size), "wall_s_incl_setup": None,
                                       "success_pct_100k": {r: round(100.0 * float(st[10 + 11 * k]["nSuccess"].mean()) / 100000, 3) for k, r in enumerate((10, 20, 50))},
                                       "mean_tx_100k": {r: round(float(st[10 + 11 * k]["preambleTxSum"].sum()) / float(st[10 + 11 * k]["nSuccess"].sum()), 3) for k, r in enumerate((10, 20, 50))}}
out["readme_sweep_33pts_x1024"]["wall_s_incl_setup"] = time.perf_counter() - t
# configs[4]: P {54,64} x G {4,8,12,16} x BI {10,20,40}, 100k UEs, Beta, 4096 replications per point (24 points)
pts = [pkg.default_params(nUE=100000, nPreamble=P, nGrantUL=G, backoffIndicator=B) for P in (54, 64) for G in (4, 8, 12, 16) for B in (10, 20, 40)]
t = time.perf_counter()
with pkg.RachSim(pts, reps=4096, devices=devs) as sim:
    sim.run()
    st = sim.stats_all()
    out["grid_24pts_x4096"] = {"kernel_ms": sim.kernel_ms, "updates_per_s": float(st["updates"].sum()) / sim.kernel_ms * 1e3,
                               "replications": int(st.size), "reps_per_s": st.size / sim.kernel_ms * 1e3,
                               "success_pct_by_point": [round(100.0 * float(st[k]["nSuccess"].mean()) / 100000, 3) for k in range(len(pts))]}
out["grid_24pts_x4096"]["wall_s_incl_setup"] = time.perf_counter() - t
print(json.dumps(out, indent=1))
