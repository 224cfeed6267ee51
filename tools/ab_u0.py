"""A/B timing of library builds on the legacy (U0) variant: python tools/ab_u0.py libA.so libB.so ... (AB_REPS=256,1024)"""
import os, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = """
import importlib, os, sys
sys.path.insert(0, %r)
pkg = importlib.import_module("5g-nr-randomaccess_b200")
p = pkg.default_params(variant=1, nUE=100000)
for reps in [int(x) for x in os.environ.get("AB_REPS", "256,1024").split(",")]:
    with pkg.RachSim([p], reps=reps, devices=[0]) as sim:
        sim.run()
        st = sim.stats_all()
        print("reps %%d kernel_ms %%.1f nSuccess %%d" %% (reps, sim.kernel_ms, int(st["nSuccess"].sum())), end="; ")
print()
""" % root
for lib in sys.argv[1:]:
    env = dict(os.environ)
    if lib != "default":
        env["RACH_GPU_LIB"] = os.path.abspath(lib)
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    print(lib, out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-300:])
