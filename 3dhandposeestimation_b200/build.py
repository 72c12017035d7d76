"""Build recipe for libmano_b200.so (nvcc, sm_100a only, in-tree so the .so travels with gpurun)."""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmano_b200.so")
SOURCES = ["api.cu", "mano_pose.cu", "mano_pose_lh.cu", "sgemm.cu", "blend_tc.cu", "vskin.cu", "skin.cu", "fk.cu", "joint_epilogue.cu", "hand_trafo.cu", "viewpoint.cu", "reduce.cu", "affine.cu", "head_loss.cu"]
HEADERS = ["common.cuh", "hand_math.cuh", "fk_math.cuh", "blend_tc.cuh", "vskin.cuh", "tc_ptx.cuh", "skin.cuh", "ptx.cuh", os.path.join("..", "..", "include", "mano_b200.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
    "--expt-relaxed-constexpr",
] + os.environ.get("MANO_B200_NVCC_EXTRA", "").split()          # e.g. -DVS_PROFILE (diagnostic builds; part of the stamp)


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError("nvcc not found; libmano_b200.so cannot be built")


def _stamp() -> str:
    h = hashlib.sha256()
    for name in SOURCES + HEADERS:
        with open(os.path.join(CSRC, name), "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every translation unit for sm_100a and link the shared library.
    Skipped when the library is newer than its sources (stamp file)."""
    stamp_file = LIB + ".stamp"
    stamp = _stamp()
    if not force and os.path.isfile(LIB) and os.path.isfile(stamp_file) and open(stamp_file).read() == stamp:
        return LIB
    nvcc = _nvcc()
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    procs = []
    objs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        objs.append(obj)
        cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(f"--- nvcc {src} ---\n{out}\n")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed building libmano_b200.so")
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *objs, "-lcudart"]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout)
    with open(stamp_file, "w") as fh:
        fh.write(stamp)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
