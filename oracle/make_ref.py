"""TEST / BENCH INFRASTRUCTURE, NOT PRODUCT CODE -- recipe for ``oracle/_ref/``.

Places the reference's own source files for the hot path, UNMODIFIED, where they can travel to the
GPU box:

    python -m oracle.make_ref            # in the build container, /root/reference present

``oracle/_ref/`` is git-ignored (the reference's sources never enter this repository's history) but
not gpurun-ignored, so the snapshot that goes to the GPU box carries it like a built ``.so``.  The
files are the ones SURVEY section 8(a) cites: ``settings.py``, ``__init__.py``,
``envs/{__init__,aircraft,kinematics,rewards,game,environment}.py``.  ``oracle/ref_shim.py`` imports them
under gym / pygame stand-ins; ``bench.py --impl reference`` and ``cpu_baseline`` time them
(``kind: "reference"``) after a golden-CSV gate (``oracle/ref_runner.py``).
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("ACAS2D_REFERENCE_ROOT", "/root/reference")
REF_DST = os.path.join(HERE, "_ref")

FILES = ("gym_ACAS2D/__init__.py", "gym_ACAS2D/settings.py", "gym_ACAS2D/envs/__init__.py",
         "gym_ACAS2D/envs/aircraft.py", "gym_ACAS2D/envs/kinematics.py", "gym_ACAS2D/envs/rewards.py",
         "gym_ACAS2D/envs/game.py", "gym_ACAS2D/envs/environment.py")


def make(verbose: bool = True) -> bool:
    """Copy the files byte for byte; returns False (and leaves any earlier copy alone) when the reference
    tree is not present, which is the situation on the GPU box."""
    if not os.path.isdir(os.path.join(REF_SRC, "gym_ACAS2D", "envs")):
        return os.path.isdir(os.path.join(REF_DST, "gym_ACAS2D", "envs"))
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(REF_SRC, rel), os.path.join(REF_DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[rel] = hashlib.sha256(open(dst, "rb").read()).hexdigest()
    with open(os.path.join(REF_DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": REF_SRC, "sha256": manifest}, f, indent=1)
    if verbose:
        print(f"oracle/_ref: {len(FILES)} reference files copied unmodified from {REF_SRC}")
    return True


if __name__ == "__main__":
    raise SystemExit(0 if make() else 1)
