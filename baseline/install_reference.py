"""Install the UNMODIFIED reference (bjo5029/causal-vae) into baseline/_ref/ so it travels to the GPU box.

    python baseline/install_reference.py            # /root/reference -> baseline/_ref  (git-ignored, shipped by gpurun)

The reference is a tree of research scripts without setup.py / pyproject, so `pip install --target baseline/_ref`
has nothing to build (tried: "does not appear to be a Python project"); this script is the equivalent install:
every `*.py` file is copied byte for byte under the same relative path, and a MANIFEST.json with the sha256 of each
file is written so `baseline/ref_harness.py` can state which bits it ran.  Nothing under baseline/_ref is tracked by
git and nothing in the product package reads it: it is the *comparator* of `bench.py --impl reference`, of the
`cpu_baseline` / `eager_gpu_baseline` legs and of tests/test_reference_crosscheck_gpu.py.
"""
import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(ROOT, "_ref")


def install(src="/root/reference"):
    if not os.path.isdir(src):
        raise SystemExit(f"{src} not present (this runs in the build container only)")
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    manifest = {}
    for d, _, files in os.walk(src):
        if "/.git" in d:
            continue
        for f in files:
            if not f.endswith(".py"):
                continue
            p = os.path.join(d, f)
            rel = os.path.relpath(p, src)
            q = os.path.join(DST, rel)
            os.makedirs(os.path.dirname(q), exist_ok=True)
            shutil.copyfile(p, q)
            manifest[rel] = hashlib.sha256(open(p, "rb").read()).hexdigest()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as fh:
        json.dump({"source": src, "files": manifest}, fh, indent=1, sort_keys=True)
    return len(manifest)


if __name__ == "__main__":
    n = install(*sys.argv[1:2])
    print(f"installed {n} reference files into {DST}")
