"""A/B timing of library builds: python tools/ab.py libA.so libB.so ... (each in a fresh process)."""
import json, os, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for lib in sys.argv[1:]:
    env = dict(os.environ)
    if lib != "default":
        env["RACH_GPU_LIB"] = os.path.abspath(lib)
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "ncu_target.py"), "--reps", os.environ.get("AB_REPS", "1184"), "--runs", "2", "--distribution", os.environ.get("AB_DIST", "2")],
                         env=env, capture_output=True, text=True)
    print(lib, out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-300:])
