#!/usr/bin/env python
"""Per-role wait-time breakdown of the fused lane = vertex forward kernel (csrc/vskin.cu built with -DVS_PROFILE):
CTA 0's roles accumulate clock64 cycles per kind of wait.  Build and run on the GPU box:
    MANO_B200_NVCC_EXTRA=-DVS_PROFILE python profiles/tools/vskin_roles.py [hands] [variant]
(rebuild without the variable afterwards: the profile build is slower)."""
import importlib
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
variant = int(sys.argv[2], 0) if len(sys.argv) > 2 else 0
model = dict(pkg.assets.synthetic_mano())
model["hands_components"] = np.eye(45)
layer = pkg.ManoLayer(dev, model=model, pose_num=45)
rs = np.random.RandomState(1)
rot = torch.from_numpy(((rs.rand(H, 3) - .5) * 2).astype(np.float32)).to(dev)
pose = torch.from_numpy(((rs.rand(H, 45) - .5)).astype(np.float32)).to(dev)
beta = torch.from_numpy(((rs.rand(H, 10) - .5)).astype(np.float32)).to(dev)
verts = torch.empty(H, 778, 3, device=dev)
joints = torch.empty(H, 21, 3, device=dev)
mode = layer._mode | cabi.FWD_INFERENCE
ws = torch.empty(lib.mb_mano_workspace_bytes(H, layer._mode), dtype=torch.uint8, device=dev)
dbg = torch.zeros(8192 + 16 * 8, device=dev)
stream = cabi.stream_handle(dev)
for _ in range(3):
    rc = lib.mb_mano_forward_debug(layer._blob.data_ptr(), 45, rot.data_ptr(), pose.data_ptr(), beta.data_ptr(), H, mode,
                                   verts.data_ptr(), joints.data_ptr(), ws.data_ptr(), ws.numel(), dbg.data_ptr(), variant, stream)
    assert rc == 0, rc
torch.cuda.synchronize()
p = dbg[8192:].cpu().numpy().reshape(16, 8)
names = {0: ("blend issuer", ["feat_full", "vp_empty", "a_full", "a_empty"]),
         1: ("transform issuer", ["w_full", "bones_full", "t_empty"]),
         2: ("basis producer", ["a_empty"]),
         3: ("slow-stream producer (polling)", []),
         4: ("epilogue set 0 q0", ["vp_full", "t_full", "T load+release", "fma+sts", "row stores", "vp load"]),
         8: ("epilogue set 1 q0", ["vp_full", "t_full", "T load+release", "fma+sts", "row stores", "vp load"])}
ntiles = (H + 63) // 64
per_cta = (ntiles + 147) // 148
print(f"hands {H} variant {variant:#x}: CTA 0 ran ~{per_cta} hand tiles = {per_cta * 7} units")
for r, (nm, waits) in names.items():
    tot = p[r, 7]
    if tot <= 0:
        print(f"{nm}: no profile data (built without -DVS_PROFILE?)")
        continue
    parts = ", ".join(f"{w} {p[r, i] / tot * 100:.0f}% ({p[r, i] / (per_cta * 7):.0f} clk/unit)" for i, w in enumerate(waits))
    print(f"{nm}: loop {tot / 1e3:.0f} kclk = {tot / (per_cta * 7):.0f} clk/unit; waits: {parts}; other {100 - sum(p[r, :len(waits)]) / tot * 100:.0f}%")
