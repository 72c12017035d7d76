"""GPU diagnostic (not a pytest test): compares the tcgen05 blend path with the fp32 FFMA path
on identical inputs and prints where they differ (per 160-column n-tile / per 8-row group), which
pinpoints descriptor or layout mistakes in one run."""
import importlib
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge  # noqa: E402

ge.build()
pkg = importlib.import_module("3dhandposeestimation_b200")
from oracle import mano_oracle as mo  # noqa: E402

dev = torch.device("cuda", 0)
model = pkg.assets.synthetic_mano()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 300
rs = np.random.RandomState(0)
rot = ((rs.rand(B, 3) - .5) * 2 * np.pi).astype(np.float32)
pose = ((rs.rand(B, 45) - .5) * np.pi).astype(np.float32)
beta = (rs.rand(B, 10) - .5).astype(np.float32)
t = [torch.from_numpy(a).to(dev) for a in (rot, pose, beta)]
ov, oj = mo.mano_forward(model, rot, pose, beta)
out = {}
for mode in ("fp32", "f16x3", "f16"):
    try:
        layer = pkg.ManoLayer(dev, model=model, pose_num=45, mode=mode)
        v, j = layer(*t)
        torch.cuda.synchronize()
        out[mode] = v.cpu().numpy()
        err = np.abs(out[mode] - ov)
        print(f"{mode:6s} verts max err vs fp64 oracle {err.max():.3e}  mean {err.mean():.3e}  joints {np.abs(j.cpu().numpy() - oj).max():.3e}")
    except Exception as exc:  # noqa: BLE001
        print(mode, "FAILED:", repr(exc)[:300])
        break
if "f16x3" in out:
    d = np.abs(out["f16x3"] - out["fp32"]).reshape(B, 2334)
    print("per n-tile max |f16x3 - fp32|:", " ".join(f"{d[:, i * 160:(i + 1) * 160].max():.1e}" for i in range(15)))
    print("per 8-row group (first 16):", " ".join(f"{d[i * 8:(i + 1) * 8].max():.1e}" for i in range(min(16, B // 8))))
    print("per column mod 32 (max):", " ".join(f"{d[:, c::32].max():.0e}" for c in range(32)))
