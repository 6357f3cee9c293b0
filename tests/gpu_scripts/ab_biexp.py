"""A/B of library variants on config 4 (tests/gpu_scripts/biexp_c4.py in a child process per library).
Usage: python tests/gpu_scripts/ab_biexp.py <budgets, e.g. 4,3:4,2> [lib ...]   -> gpurun_out/ab_biexp.json"""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
budgets = sys.argv[1].split(":")
libs = sys.argv[2:] or sorted(glob.glob(os.path.join(ROOT, "dosma_b200", "libdfit*.so")))
res = {}
for lib in libs:
    env = dict(os.environ, DOSMA_B200_LIB=os.path.abspath(lib))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests/gpu_scripts/biexp_c4.py"), "5"] + budgets, env=env,
                       capture_output=True, text=True, cwd=ROOT)
    name = os.path.basename(lib)
    try:
        d = json.load(open(os.path.join(ROOT, "gpurun_out", "biexp_c4.json")))
        res[name] = {k: {"ms": v["ms"], "checksum": v["checksum"]} for k, v in d.items()}
    except Exception:
        res[name] = {"error": (r.stderr or r.stdout)[-600:]}
    if r.returncode != 0:
        res[name] = {"error": (r.stderr or r.stdout)[-600:]}
    print(name, json.dumps(res[name]), flush=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "ab_biexp.json"), "w"), indent=1)
