"""Moved to tests/fuzz/emu_fuzz.py (test infrastructure: it checks against the oracle).  This stub forwards."""
import os, runpy, sys
runpy.run_path(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "fuzz", "emu_fuzz.py"), run_name="__main__")
