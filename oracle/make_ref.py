#!/usr/bin/env python
"""oracle/make_ref.py — TEST INFRASTRUCTURE: stage the UNMODIFIED reference modules of the ComA path for the GPU box.

    python oracle/make_ref.py            # dev container only (needs /root/reference)

The reference (snuvclab/coma) is pure Python; `/root/reference` does not exist on the GPU box, and gpurun ships the repo
tree including git-ignored files. This recipe copies the five modules the ComA path imports
    utils/coma.py  utils/coma_occupancy.py  utils/misc.py  utils/transformations.py  utils/load_3d.py
byte for byte into `oracle/_ref/utils/` (git-ignored: reference sources never enter the history; NOT gpurun-ignored, so they
travel) and writes a MANIFEST with their sha256. `oracle/ref_loader.py` imports them with open3d / trimesh / easydict stubbed
(SURVEY.md Appendix C). Users: tests (CUDA-vs-CPU agreement of the reference itself, parity of the kernels against the
reference run with device="cuda") and `bench.py --impl reference` / the `reference_cuda` key. Never imported by coma_b200/.
"""
import hashlib
import json
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("COMA_REFERENCE", "/root/reference")
FILES = ["utils/coma.py", "utils/coma_occupancy.py", "utils/misc.py", "utils/transformations.py", "utils/load_3d.py"]


def make(force=False):
    dst_root = os.path.join(HERE, "_ref")
    manifest_pth = os.path.join(dst_root, "MANIFEST.json")
    if not os.path.isdir(REF):
        return os.path.exists(manifest_pth)          # GPU box: use what travelled
    os.makedirs(os.path.join(dst_root, "utils"), exist_ok=True)
    manifest = {"source": REF, "files": {}}
    try:
        manifest["commit"] = subprocess.run(["git", "-C", REF, "rev-parse", "HEAD"], capture_output=True, text=True).stdout.strip() or None
    except Exception:
        manifest["commit"] = None
    for f in FILES:
        src, dst = os.path.join(REF, f), os.path.join(dst_root, f)
        shutil.copyfile(src, dst)
        manifest["files"][f] = hashlib.sha256(open(dst, "rb").read()).hexdigest()
    # the reference's `utils` is a namespace package (no __init__.py); this repo's pickle-compat shim `utils/` is a regular
    # package and would win the import even from a later sys.path entry — an (empty, generated) __init__.py levels that
    open(os.path.join(dst_root, "utils", "__init__.py"), "w").close()
    with open(manifest_pth, "w") as w:
        json.dump(manifest, w, indent=1)
    return True


if __name__ == "__main__":
    ok = make()
    print("oracle/_ref staged" if ok else "no reference available", file=sys.stderr)
    sys.exit(0 if ok else 1)
