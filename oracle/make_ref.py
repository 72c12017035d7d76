"""ORACLE recipe (test infrastructure) — stage a travelling copy of the UNMODIFIED reference path under
``oracle/_ref/`` so that the GPU box (which has no /root/reference) can run the reference's own PyTorch code as
the parity checker and as the CPU baseline arm of ``bench.py``.

``oracle/_ref/`` is git-ignored (reference sources and the MANO-licensed pkl are never committed) but NOT
gpurun-ignored, so it ships with the snapshot like the built ``.so``.  Run by ``__graft_entry__.build()`` whenever
/root/reference exists; a no-op elsewhere.  Files are copied byte for byte — nothing is edited."""
from __future__ import annotations

import hashlib
import os
import shutil

REF_ROOT = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
DST_ROOT = os.path.join(HERE, "_ref")

# the reference files of SURVEY §8(a) plus what they import
FILES = [
    "network/sub_modules/MANOLayer.py",
    "network/sub_modules/forwardKinematicsLayer.py",
    "network/Resnet50MANO3DHandPose.py",
    "network/sub_modules/resnet50MANO.py",
    "utils/coordinate_trans.py",
    "utils/util.py",
    "utils/general.py",
    "utils/relative_trafo.py",
    "utils/canonical_trafo.py",
    "criterions/metrics.py",
    "criterions/loss.py",
    "config/config.py",
    "config/mano/models/MANO_RIGHT.pkl",
]


def _sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as fh:
        h.update(fh.read())
    return h.hexdigest()


def stage(verbose: bool = False) -> bool:
    """Copy the files when the reference checkout is present; returns True when oracle/_ref is usable."""
    if not os.path.isdir(REF_ROOT):
        return os.path.isfile(os.path.join(DST_ROOT, FILES[0]))
    for rel in FILES:
        src = os.path.join(REF_ROOT, rel)
        dst = os.path.join(DST_ROOT, rel)
        if not os.path.isfile(src):
            continue
        if os.path.isfile(dst) and _sha(src) == _sha(dst):
            continue
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        if verbose:
            print("staged", rel)
    return True


if __name__ == "__main__":
    print("oracle/_ref ready" if stage(verbose=True) else "reference not present; nothing staged")
