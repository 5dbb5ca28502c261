"""Compile the library once per set of extra nvcc flags, here in the build container, into build/variants/<label>/libstratusgi.so
(git-ignored, but it travels to the GPU box), so that an A/B on the box is a file copy instead of a 40-second nvcc run.

    python tools/prebuild_variants.py LABEL:"-DFLAG ..." [LABEL:"..."] ...
"""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from stratusgfx_b200.build import CSRC, CU_SOURCES, NVCC_FLAGS  # noqa: E402


def build(label, flags):
    out = os.path.join(ROOT, "build", "variants", label)
    os.makedirs(out, exist_ok=True)
    objs = []
    base = [f for f in NVCC_FLAGS if f not in ("-Xptxas", "-v")]
    for s in CU_SOURCES:
        o = os.path.join(out, s[:-3] + ".o")
        subprocess.check_call(["nvcc"] + base + flags.split() + ["-c", os.path.join(CSRC, s), "-o", o])
        objs.append(o)
    lib = os.path.join(out, "libstratusgi.so")
    subprocess.check_call(["nvcc", "-shared", "-o", lib] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"])
    for o in objs:
        os.remove(o)
    return label, lib


if __name__ == "__main__":
    jobs = [a.split(":", 1) for a in sys.argv[1:]]
    with ThreadPoolExecutor(4) as ex:
        for label, lib in ex.map(lambda j: build(*j), jobs):
            print(label, lib)
