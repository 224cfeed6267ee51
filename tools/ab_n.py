"""A/B timing of library builds on the variant-N config: python tools/ab_n.py libA.so libB.so ..."""
import os, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = """
import importlib, sys
sys.path.insert(0, %r)
pkg = importlib.import_module("5g-nr-randomaccess_b200")
p = pkg.default_params(variant=2, nUE=50000)
with pkg.RachSim([p], reps=2048, devices=[0]) as sim:
    sim.run(); sim.run()
    st = sim.stats_all()
    print("kernel_ms %%.2f nSuccess %%d" %% (sim.kernel_ms, int(st["nSuccess"].sum())))
""" % root
for lib in sys.argv[1:]:
    env = dict(os.environ)
    if lib != "default":
        env["RACH_GPU_LIB"] = os.path.abspath(lib)
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    print(lib, out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-300:])
