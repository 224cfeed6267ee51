"""Timing of the BASELINE configs[4] grid on one device (runtime point view): P {54,64} x G {4,8,12,16} x BI {10,20,40}
at 100k UEs, REPS replications per point.  python tools/grid_target.py [REPS]"""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("5g-nr-randomaccess_b200")
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 256
pts = [pkg.default_params(nUE=100000, nPreamble=P, nGrantUL=G, backoffIndicator=B) for P in (54, 64) for G in (4, 8, 12, 16) for B in (10, 20, 40)]
with pkg.RachSim(pts, reps=reps, devices=[0]) as sim:
    sim.run(); sim.run()
    st = sim.stats_all()
    print("grid 24 points x %d: kernel_ms %.1f updates/s %.4e" % (reps, sim.kernel_ms, float(st["updates"].sum()) / sim.kernel_ms * 1e3))
