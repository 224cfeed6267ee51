"""Small run of every variant for compute-sanitizer (memcheck)."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("5g-nr-randomaccess_b200")
for name, pts, reps, dump in (
        ("W", [pkg.default_params(nUE=6000, seed=1), pkg.default_params(nUE=2500, nPreamble=3, nGrantUL=2, seed=2)], 3, True),
        ("W-nodump", [pkg.default_params(nUE=8000, seed=3, maxTimeMs=4000)], 4, False),
        ("N", [pkg.default_params(variant=2, nUE=5000, seed=4)], 3, True),
        ("U0", [pkg.default_params(variant=1, nUE=9000, nPreamble=2, seed=5, maxTimeMs=20000)], 4, True)):
    with pkg.RachSim(pts, reps=reps, devices=[0], dump_ues=dump) as sim:
        sim.run()
        st = sim.stats_all()
        if dump:
            sim.dump_ues(0, reps - 1)
        print(name, "ok", int(st["nSuccess"].sum()))
