"""TEST INFRASTRUCTURE ONLY — recipe that stages the UNMODIFIED reference next to the oracle.

The reference (poetrywanderer/CF-NeRF) is pure Python: there is nothing to compile, "building" it means making its
source files importable where the GPU box can see them.  This script copies every `*.py` of the read-only reference
tree, byte for byte and with its directory layout, from where it lies (`/root/reference`, or $CFNERF_REFERENCE_SRC)
into `oracle/_ref/`.  That directory is git-ignored (no reference source ever enters the history) but NOT
gpurun-ignored, so it travels to the GPU box like the repo's own built `.so`.  `oracle/refload.py` imports the
reference from `/root/reference` when present, else from `oracle/_ref/`.

Users: `bench.py --impl reference` and `bench.py`'s `cpu_baseline` leg (kind "reference": the reference's own
`render()` / trainer body timed on the host cores), and the `-m gpu` tests that put `cfnerf_b200.install()` behind
the real `run_nerf_uncertainty_NF` module.  Never imported by the product package.

    python oracle/build_ref.py            # stage (no-op when the source tree is absent)
    python oracle/build_ref.py --check    # verify the staged files still match the source byte for byte
"""
from __future__ import annotations

import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("CFNERF_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")
MANIFEST = os.path.join(DST, "MANIFEST.sha256")


def _py_files(root):
    out = []
    for d, dirs, files in os.walk(root):
        dirs[:] = [x for x in dirs if x not in (".git", "__pycache__")]
        for f in files:
            if f.endswith(".py"):
                out.append(os.path.relpath(os.path.join(d, f), root))
    return sorted(out)


def _sha(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def stage(verbose: bool = True) -> str | None:
    """Copy the reference's Python sources into oracle/_ref/.  Returns the directory, or None when there is no
    source tree here (the GPU box: it uses the already staged copy)."""
    if not os.path.isfile(os.path.join(SRC, "run_nerf_uncertainty_NF.py")):
        return DST if os.path.isfile(os.path.join(DST, "run_nerf_uncertainty_NF.py")) else None
    os.makedirs(DST, exist_ok=True)
    lines = []
    for rel in _py_files(SRC):
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(SRC, rel), dst)
        lines.append(f"{_sha(dst)}  {rel}")
    with open(MANIFEST, "w") as f:
        f.write("\n".join(lines) + "\n")
    if verbose:
        print(f"staged {len(lines)} unmodified reference files from {SRC} into {DST}")
    return DST


def check() -> bool:
    """True when every staged file has the hash recorded at staging time (and equals the source when it is present)."""
    if not os.path.isfile(MANIFEST):
        return False
    for line in open(MANIFEST):
        h, rel = line.strip().split("  ", 1)
        if _sha(os.path.join(DST, rel)) != h:
            return False
        src = os.path.join(SRC, rel)
        if os.path.isfile(src) and _sha(src) != h:
            return False
    return True


if __name__ == "__main__":
    if "--check" in sys.argv:
        ok = check()
        print("oracle/_ref matches the reference" if ok else "oracle/_ref is missing or differs")
        sys.exit(0 if ok else 1)
    stage()
