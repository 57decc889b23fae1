"""Build the full library with extra nvcc flags into build/variants/<tag>.so (development aid for scripts/ab.py)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mpc_planner_b200 import build
tag, flags = sys.argv[1], sys.argv[2:]
out = os.path.join(ROOT, "build", "variants", tag + ".so")
print(build.build(force=True, output=out, extra_flags=flags))
