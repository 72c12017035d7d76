#!/usr/bin/env python
"""Where the time of a config-3 step through the nn.Module drop-ins goes (B = 65 536): the pieces timed one after the other
on the device (CUDA events around 50 calls after a warm-up) and the host-side cost of issuing them (perf_counter, no sync)."""
import importlib
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
pkg = importlib.import_module("3dhandposeestimation_b200")
dev = torch.device("cuda", 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
rs = np.random.RandomState(0)
a = [((rs.rand(B, 3) - .5) * 2 * np.pi), ((rs.rand(B, 23) - .5) * np.pi), rs.rand(B, 20) + .1]
K = np.tile(np.array([[282.9, 0, 160], [0, 282.9, 160], [0, 0, 1]]), (B, 1, 1))
t = [torch.from_numpy(x.astype(np.float32)).to(dev) for x in (*a, K, rs.rand(B, 1) * .05 + .02, rs.randn(B, 3) * .05 + np.array([0, 0, .6]))]
for x in t[:3]:
    x.requires_grad_()
gt = torch.from_numpy((rs.randn(B, 21, 3) * .05 + np.array([0, 0, .6])).astype(np.float32)).to(dev)
vis = torch.from_numpy((rs.rand(B, 21, 1) < .8).astype(np.float32)).to(dev)
fk, mp, l2 = pkg.ForwardKinematics(dev), pkg.MPJPE(), pkg.L2Loss()
gt_uv = torch.rand(B, 21, 2, device=dev) * 320


def clear():
    for x in t[:3]:
        x.grad = None


def s_fwd():
    fk(*t)


def s_fwd_nograd():
    with torch.no_grad():
        fk(*t)


def s_l2():
    xyz, uv, _ = fk(*t)
    l2(xyz, gt, vis).backward()
    clear()


def s_l2_both():
    xyz, uv, _ = fk(*t)
    (l2(xyz, gt, vis) + l2(uv, gt_uv, vis)).backward()
    clear()


def s_bench():
    xyz, uv, _ = fk(*t)
    (mp(xyz, gt, vis) + 1e-3 * uv.sum()).backward()
    clear()


def s_fkloss():
    if not hasattr(pkg, "ForwardKinematicsLoss"):
        return
    lx, lu, _, _ = crit(*t, gt, gt_uv, vis)
    (lx + 1e-3 * lu).backward()
    clear()


crit = pkg.ForwardKinematicsLoss(dev) if hasattr(pkg, "ForwardKinematicsLoss") else None
out = {}
for name, fn in (("fk_forward_no_grad", s_fwd_nograd), ("fk_forward", s_fwd), ("fk+L2xyz+backward", s_l2),
                 ("fk+L2xyz+L2uv+backward", s_l2_both), ("bench_step(mpjpe+uv.sum)", s_bench), ("fk_loss_module", s_fkloss)):
    if name == "fk_loss_module" and crit is None:
        continue
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 50
    e0.record()
    h0 = time.perf_counter()
    for _ in range(n):
        fn()
    h1 = time.perf_counter()
    e1.record()
    torch.cuda.synchronize()
    out[name] = {"device_ms": e0.elapsed_time(e1) / n, "host_issue_ms": (h1 - h0) * 1e3 / n}
print(json.dumps({"B": B, **out}, indent=1))
if "--profile" in sys.argv:
    import cProfile
    import pstats
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(200):
        (s_fkloss if crit is not None else s_l2_both)()
    pr.disable()
    torch.cuda.synchronize()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(25)
