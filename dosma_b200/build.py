"""In-tree build of libdfit.so (hand-written sm_100a CUDA + the C-ABI of include/dfit.h).

    python -m dosma_b200.build [--force] [--verbose]

Every translation unit is compiled with
    nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3
(object files in dosma_b200/csrc/_build/, compiled in parallel) and linked into
dosma_b200/libdfit.so.  nvcc cross-compiles without a GPU, so this runs in the CPU-only build
container; the resulting .so is git-ignored but travels to the GPU box with the tree.
"""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(CSRC, "_build")
LIB = os.path.join(PKG, "libdfit.so")
INCLUDE = os.path.join(os.path.dirname(PKG), "include")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-O2",
    "-Xptxas", "-v",
    "--expt-relaxed-constexpr",
]


def _nvcc():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(INCLUDE, "dfit.h"))
    hs.append(os.path.abspath(__file__))
    return hs


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, variant=None, extra_flags=()):
    """Build libdfit.so, or -- for kernel experiments -- a variant `libdfit_<variant>.so` compiled with
    `extra_flags` (e.g. ("-DDFIT_M2_MIN_CTAS=6",)); a variant is loaded by setting DOSMA_B200_LIB to its path."""
    obj_dir = OBJ if variant is None else OBJ + "_" + variant
    lib_path = LIB if variant is None else os.path.join(PKG, f"libdfit_{variant}.so")
    os.makedirs(obj_dir, exist_ok=True)
    hdrs = _headers()
    jobs = []
    objs = []
    for src in sources():
        obj = os.path.join(obj_dir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if force or _stale(obj, [src] + hdrs):
            jobs.append((src, obj))

    def compile_one(job):
        src, obj = job
        cmd = [_nvcc()] + NVCC_FLAGS + list(extra_flags) + ["-I", INCLUDE, "-c", src, "-o", obj]
        res = subprocess.run(cmd, capture_output=True, text=True)
        log = obj[:-2] + ".ptxas.log"
        with open(log, "w") as f:
            f.write(" ".join(cmd) + "\n" + res.stdout + res.stderr)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{res.stdout}\n{res.stderr}")
        if verbose:
            print(res.stderr)
        return obj

    if jobs:
        # longest translation units first (the fp32 mono- / bi-exponential instances carry the two-voxel, TMA and rounds
        # kernels): with fewer cores than units the wall time is set by what starts last
        weight = {"inst_mono_f32": 0, "inst_biexp_f32": 1, "inst_mono_f64": 2, "inst_biexp_f64": 3, "inst_linear_f32": 4}
        jobs.sort(key=lambda j: min([w for k, w in weight.items() if os.path.basename(j[0]).startswith(k)] or [9]))
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            list(ex.map(compile_one, jobs))
    if jobs or force or _stale(lib_path, objs):
        cmd = [_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib_path] + objs
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    return lib_path


if __name__ == "__main__":
    variant, flags = None, []
    for arg in sys.argv[1:]:
        if arg.startswith("--variant="):
            variant = arg.split("=", 1)[1]
        elif arg.startswith("-D"):
            flags.append(arg)
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, variant=variant, extra_flags=flags)
    print(path)
