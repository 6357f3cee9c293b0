#!/usr/bin/env python
"""SASS instruction-mix digest of the kernels in dosma_b200/libdfit.so (cuobjdump -sass), written to
profiles/sass_digest.md: per kernel the static instruction count and the mnemonics that show what it is made of --
packed FP32 (FFMA2 / FMUL2 / FADD2), TMA (UTMALDG), bulk copies (UBLKCP), mbarrier traffic (SYNCS), MUFU, FP64.

    python profiles/sass_digest.py [kernel-name-regex ...]
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "dosma_b200", "libdfit.so")
WANT = sys.argv[1:] or [r"fit_kernel_mono2_tma<dfit::MonoExp, 8, false, float>", r"fit_kernel_mono2_tma<dfit::MonoExp, 8, true, float>",
                        r"fit_kernel_mono2_tma<dfit::MonoExp, 8, false, short>", r"fit_kernel_mono2_tma<dfit::MonoExp, 7, false, float>",
                        r"fit_kernel_mono2_list<dfit::MonoExp, 7, false>", r"fit_kernel_mono2<dfit::MonoExp, 8>",
                        r"fit_kernel<dfit::MonoExp, float, 8, true, false>", r"fit_kernel<dfit::BiExp, float, 16, true, false>",
                        r"fit_kernel_lmq<dfit::BiExp, 16, true>", r"fit_kernel_lmq<dfit::BiExp, 16, false>", r"fit_kernel_lmq<dfit::MonoExp, 8, true>",
                        r"mask_compact_kernel<2, float, 8>", r"qdess_kernel<float>", r"qdess_kernel<double>"]
KEYS = ["FFMA2", "FMUL2", "FADD2", "FFMA", "FMUL", "FADD", "MUFU", "FSETP", "FSEL", "FMNMX", "UTMALDG", "UBLKCP", "SYNCS", "LDS", "LDG",
        "STG", "DFMA", "DMUL", "DADD", "VOTE", "SHFL", "ATOMG", "REDG", "BRA"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
    blocks = re.split(r"\n\s*Function : \S+\n", sass)[1:]
    rows = []
    for name, body in zip(names, blocks):
        if not any(re.search(re.escape(w) if "<" in w else w, name) for w in WANT):
            continue
        ops = collections.Counter()
        for m in re.finditer(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", body):
            ops[m.group(1)] += 1
        rows.append((name, sum(ops.values()), ops))
    out = ["# SASS instruction-mix digest (cuobjdump -sass dosma_b200/libdfit.so, sm_100a cubins)", "",
           "Static instruction counts per kernel (all paths, including the rarely executed LM fallback); the dynamic mix of the hot",
           "loop is in the ncu summaries next to this file.  `FFMA2/FMUL2/FADD2` = packed FP32 (two voxels per instruction),",
           "`UTMALDG` = TMA tile loads (`cp.async.bulk.tensor`), `SYNCS` = mbarrier arrive / try_wait, `MUFU` = rcp / ex2 / lg2.", "",
           "| kernel | total | " + " | ".join(KEYS) + " |", "|---|---:|" + "---:|" * len(KEYS)]
    for name, tot, ops in sorted(rows):
        short = re.sub(r"\(.*", "", name).replace("void dfit::", "").replace("dfit::", "")
        out.append(f"| `{short}` | {tot} | " + " | ".join(str(ops.get(k, 0)) for k in KEYS) + " |")
    lib_ops = collections.Counter(re.findall(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", sass))
    out += ["", "Whole library: " + ", ".join(f"{k} {lib_ops.get(k, 0)}" for k in ("UTMALDG", "UBLKCP", "SYNCS", "FFMA2", "FMUL2", "FADD2", "MUFU", "DFMA")) +
            f"; tcgen05 / TMEM (UTC*MMA, LDTM): {sum(v for k, v in lib_ops.items() if k.startswith('UTC') or k in ('LDTM', 'STTM'))} "
            "(none: the path is a per-voxel iteration, not a contraction)."]
    with open(os.path.join(ROOT, "profiles", "sass_digest.md"), "w") as f:
        f.write("\n".join(out) + "\n")
    print("\n".join(out))


if __name__ == "__main__":
    main()
