"""Worker of tests/test_gpu_multi.py (launched by torch.distributed.run, one rank per GPU over NCCL): the batched fitting loop
with the batch sharded over the ranks against the same loop on the whole batch on ONE GPU (rank 0) — the per-iteration
all-reduce of the objective partials is the only difference, so the parameters must agree to fp32 rounding."""
import importlib
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    out_path = sys.argv[1]
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    pkg = importlib.import_module("3dhandposeestimation_b200")
    fitting = pkg.fitting
    model = pkg.assets.synthetic_mano()
    B, nc, iters = 8192 + 37, 45, 6                     # uneven shards, both above / below nothing special
    rs = np.random.RandomState(3)
    rot = ((rs.rand(B, 3) - .5) * 2).astype(np.float32)
    pose = ((rs.rand(B, nc) - .5) * .6).astype(np.float32)
    beta = ((rs.rand(B, 10) - .5) * .5).astype(np.float32)
    tgt = (rs.randn(B, 21, 3) * .03).astype(np.float32)
    vis = (rs.rand(B, 21, 1) < .8).astype(np.float32)
    layer = pkg.ManoLayer(dev, model=model, pose_num=nc)

    def run(lo, hi, group):
        fit = fitting.ManoFitter(layer, hi - lo, lr=1e-2, group=group)
        fit.rot.copy_(torch.from_numpy(rot[lo:hi]))
        fit.pose.copy_(torch.from_numpy(pose[lo:hi]))
        fit.beta.copy_(torch.from_numpy(beta[lo:hi]))
        t, v = torch.from_numpy(tgt[lo:hi]).to(dev), torch.from_numpy(vis[lo:hi]).to(dev)
        losses = []
        for _ in range(iters):
            losses.append(fit.step(t, v).clone())
        torch.cuda.synchronize(dev)
        return fit, [float(x) for x in losses]

    lo, hi = fitting.shard_range(B, rank, world)
    fit_s, loss_s = run(lo, hi, None)                   # sharded: the default group's all-reduce every iteration
    shard = torch.cat([fit_s.rot, fit_s.pose, fit_s.beta], dim=1).contiguous()
    sizes = [fitting.shard_range(B, r, world) for r in range(world)]
    gathered = [torch.empty(b - a, 3 + nc + 10, device=dev) for a, b in sizes]
    dist.all_gather(gathered, shard)
    res = None
    if rank == 0:
        fit_w, loss_w = run(0, B, "local")              # the whole batch on this GPU, no collective
        whole = torch.cat([fit_w.rot, fit_w.pose, fit_w.beta], dim=1)
        got = torch.cat(gathered, dim=0)
        res = {"world": world, "max_param_diff": float((got - whole).abs().max()), "loss_sharded": loss_s, "loss_whole": loss_w,
               "fused": bool(fit_s.fused)}
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        with open(out_path, "w") as fh:
            json.dump(res, fh)


if __name__ == "__main__":
    main()
