#!/usr/bin/env python
"""Stage the UNMODIFIED reference under baseline/_ref (git-ignored, NOT gpurun-ignored) so that the GPU
box -- which has no /root/reference -- can (i) time the reference's own CPU step loop beside the CUDA
path (`bench.py --impl reference`, `cpu_baseline.kind = "reference"`) and (ii) run the three unchanged
driver scripts against the CUDA-backed `Environment` module (tests/test_gpu_drivers.py).

Only what those two uses read is copied: the two Simulation-* directories without bytecode, logs, the
orphan SAC checkpoint and the plotting data.  Nothing under baseline/_ref is ever committed or edited."""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SKIP_DIRS = {"__pycache__", ".idea", "Data2", "runs", "ris_sac_model", "plt"}


def stage(src="/root/reference", dst=os.path.join(ROOT, "baseline", "_ref"), force=False):
    if not os.path.isfile(os.path.join(src, "Simulation-MARL-BCD", "Environment.py")):
        return None
    marker = os.path.join(dst, ".staged_from")
    if os.path.isfile(marker) and not force:
        return dst
    for sub in ("Simulation-MARL-BCD", "Simulation-SARL"):
        for dirpath, dirnames, filenames in os.walk(os.path.join(src, sub)):
            dirnames[:] = [d for d in dirnames if d not in SKIP_DIRS]
            rel = os.path.relpath(dirpath, src)
            os.makedirs(os.path.join(dst, rel), exist_ok=True)
            for f in filenames:
                if not f.endswith((".pyc", ".log")):
                    shutil.copy2(os.path.join(dirpath, f), os.path.join(dst, rel, f))
    with open(marker, "w") as fh:
        fh.write(src + "\n")
    return dst


if __name__ == "__main__":
    out = stage(force="--force" in sys.argv)
    print(out or "reference tree not mounted: nothing staged")
