"""CPU, world_size 2 over gloo: the host-side logic of the sharded fitting loop — contiguous batch
shards, the one all-reduce of the objective partials, and the global loss / gradient scales —
against the un-sharded objective of the reference (L2Loss + MANO regulariser, via the oracle)."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, B, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    fitting = importlib.import_module("3dhandposeestimation_b200.fitting")
    rs = np.random.RandomState(0)                     # every rank draws the same global problem
    joints = rs.randn(B, 21, 3) * .05
    target = rs.randn(B, 21, 3) * .05
    vis = (rs.rand(B, 21, 1) < .7).astype(np.float64)
    theta = rs.randn(B, 45)
    beta = rs.randn(B, 10)
    lo, hi = fitting.shard_range(B, rank, world)
    d2 = ((joints[lo:hi] - target[lo:hi]) ** 2).sum(-1, keepdims=True) * vis[lo:hi]
    partials = torch.tensor([d2.sum(), vis[lo:hi].sum(), (theta[lo:hi] ** 2).sum(), (beta[lo:hi] ** 2).sum()],
                            dtype=torch.float64)
    fitting.reduce_partials(partials)
    loss, inv_t, inv_b = fitting.objective_from_partials(partials)
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), np.array([loss.item(), inv_t.item(), inv_b.item(), lo, hi]))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_covers_batch_exactly():
    fitting = importlib.import_module("3dhandposeestimation_b200.fitting")
    for n in (0, 1, 7, 64, 1 << 20):
        for world in (1, 2, 3, 8):
            spans = [fitting.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_objective_all_invisible_and_zero_norms():
    fitting = importlib.import_module("3dhandposeestimation_b200.fitting")
    loss, inv_t, inv_b = fitting.objective_from_partials(torch.zeros(4, dtype=torch.float64))
    assert loss.item() == 0.0 and inv_t.item() == 0.0 and inv_b.item() == 0.0      # loss.py:20-21 returns 0


@pytest.mark.timeout(120)
def test_two_rank_gloo_objective_equals_unsharded_reference(tmp_path):
    from oracle import fk_oracle as fo

    B, world = 37, 2                                   # odd batch: uneven shards
    port = _free_port()
    mp.spawn(_worker, args=(world, port, B, str(tmp_path)), nprocs=world, join=True)
    rs = np.random.RandomState(0)
    joints = rs.randn(B, 21, 3) * .05
    target = rs.randn(B, 21, 3) * .05
    vis = (rs.rand(B, 21, 1) < .7).astype(np.float64)
    theta = rs.randn(B, 45)
    beta = rs.randn(B, 10)
    want = fo.l2loss(joints, target, vis) + fo.regularizer(theta, beta)
    got = [np.load(tmp_path / f"rank{r}.npy") for r in range(world)]
    for g in got:
        assert g[0] == pytest.approx(want, rel=1e-12)
        assert g[1] == pytest.approx(1.0 / np.linalg.norm(theta), rel=1e-12)
        assert g[2] == pytest.approx(1.0 / np.linalg.norm(beta), rel=1e-12)
    assert (got[0][3], got[0][4], got[1][3], got[1][4]) == (0, 19, 19, 37)
