#!/usr/bin/env python
"""Forward-only A/B of the fused lane = vertex kernel (csrc/vskin.cu) against the two separate kernels, and the
kernel's experiment variants (MANO_B200_VSKIN_VARIANT bits, MANO_B200_VSKIN_CLUSTER).  Device-resident inputs,
CUDA-event timing on the launching stream, 3 warm-up + N timed launches per configuration.
Usage (GPU box): python profiles/tools/vskin_variants.py [hands] [variant,variant,...]"""
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import __graft_entry__ as ge  # noqa: E402

ge.build()
pkg = importlib.import_module("3dhandposeestimation_b200")
cabi = pkg._cabi
lib = pkg.load_library()
dev = torch.device("cuda", 0)
H = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
variants = [v for v in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["0"])]
model = dict(pkg.assets.synthetic_mano())
model["hands_components"] = np.eye(45)
layer = pkg.ManoLayer(dev, model=model, pose_num=45)
rs = np.random.RandomState(1)
rot = torch.from_numpy(((rs.rand(H, 3) - .5) * 2).astype(np.float32)).to(dev)
pose = torch.from_numpy(((rs.rand(H, 45) - .5)).astype(np.float32)).to(dev)
beta = torch.from_numpy(((rs.rand(H, 10) - .5)).astype(np.float32)).to(dev)
verts = torch.empty(H, 778, 3, device=dev)
joints = torch.empty(H, 21, 3, device=dev)
mode = layer._mode
ws_bytes = lib.mb_mano_workspace_bytes(H, mode)
ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
stream = cabi.stream_handle(dev)
blob = layer._blob.data_ptr()


def run(flags, n=10):
    def once():
        cabi.check(lib.mb_mano_forward(blob, 45, rot.data_ptr(), pose.data_ptr(), beta.data_ptr(), H, mode | flags,
                                       verts.data_ptr(), joints.data_ptr(), ws.data_ptr(), ws_bytes, stream), "fwd")
    for _ in range(3):
        once()
    torch.cuda.synchronize()
    lib.mb_profile_enable(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        once()
    e1.record()
    torch.cuda.synchronize()
    prof = cabi.profile_collect()
    lib.mb_profile_enable(0)
    return e0.elapsed_time(e1) / n, {k: round(v[0] / v[1], 4) for k, v in prof.items()}


out = {"hands": H}
ms, st = run(cabi.FWD_INFERENCE | cabi.FWD_UNFUSED)
ref = verts.clone()
out["separate"] = {"ms": ms, "stages": st}
print("separate kernels", round(ms, 4), st, flush=True)
for cl in os.environ.get("VS_CLUSTERS", "2").split(","):
    os.environ["MANO_B200_VSKIN_CLUSTER"] = cl
    for v in variants:
        os.environ["MANO_B200_VSKIN_VARIANT"] = v
        ms, st = run(cabi.FWD_INFERENCE | cabi.FWD_FUSED)
        err = float((verts - ref).abs().max()) if int(v, 0) & 0xff00 == 0 else None
        out[f"fused cluster {cl} variant {v}"] = {"ms": ms, "stages": st, "max_abs_diff_vs_separate": err}
        print(f"fused cluster {cl} variant {v}", round(ms, 4), st, "diff vs separate", err, flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"vskin_variants_{H}.json"), "w"), indent=1)
