"""Batched MANO fitting loop (BASELINE config 5): Adam on (rot, pose, beta) of many hands against
target keypoints, batch-sharded over the GPUs of one box.

Objective = the reference's training terms for the MANO heads:
  ``L2Loss`` (criterions/loss.py:10-25)   mean over the VISIBLE joints of the whole batch of ||joint - target||^2
  ``compute_regularization_loss`` (:113-117)   (||theta||_F + 10 ||beta||_F) / 100 over the whole batch
Both are normalised by batch-global quantities (visible count, Frobenius norms), so with the batch
sharded over ranks every iteration needs exactly one collective: an all-reduce (sum) of the four
partials ``[sum ||d||^2, N_vis, sum theta^2, sum beta^2]`` — 32 bytes over NCCL/NVLink.  Everything
else is rank-local, and for MANO's kinematic tree it is ONE kernel per iteration (``mb_mano_fit_step``:
joints-only forward, objective gradient, joints-only backward, regulariser gradient, Adam).  That is possible
because the batch-global scalars the gradient needs — the visible count and the Frobenius norms of the current
parameters — do not depend on the forward pass: the kernel of iteration i also emits the norms of the parameters it
has just updated, and the all-reduce that follows it delivers them to iteration i+1 together with iteration i's
L2 sum.  Other trees run the separate kernels (joints-only forward, masked reduction, joints-only backward, fused
Adam).  All sm_100a, through the C ABI; no autograd tape, no host synchronisation — the reduced scalars stay on
the device.

The reference itself has no fitting loop (its only optimiser is Adam on network weights,
trainval.py:119); the hyper-parameters here (lr 1e-2, betas (0.9, 0.999), eps 1e-8) are this
build's choice and are mirrored by the numpy-Adam check in tests/.
"""
from __future__ import annotations

import torch

from . import _cabi

N_PARTIALS = 4      # sum ||d||^2, N_vis, sum theta^2, sum beta^2


def shard_range(n: int, rank: int, world: int):
    """Contiguous slice [lo, hi) of n hands owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def reduce_partials(partials: torch.Tensor, group=None) -> torch.Tensor:
    """All-reduce (sum) the float64[4] objective partials over the ranks (no-op without a process group)."""
    import torch.distributed as dist

    if isinstance(group, str):          # "local": this rank's objective only — no collective (measurement / single-rank use)
        return partials
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(partials, op=dist.ReduceOp.SUM, group=group)
    return partials


def objective_from_partials(p: torch.Tensor):
    """(loss, inv_theta_norm, inv_beta_norm) from the GLOBAL partials, all on p's device.
    loss = S/N (0 if N == 0) + (sqrt(T) + 10 sqrt(Bt)) / 100."""
    s, n, t, b = p[0], p[1], p[2], p[3]
    l2 = torch.where(n > 0, s / torch.clamp(n, min=1.0), torch.zeros_like(s))
    tn, bn = torch.sqrt(t), torch.sqrt(b)
    loss = l2 + (tn + 10.0 * bn) / 100.0
    inv_t = torch.where(tn > 0, 1.0 / torch.clamp(tn, min=1e-300), torch.zeros_like(tn))
    inv_b = torch.where(bn > 0, 1.0 / torch.clamp(bn, min=1e-300), torch.zeros_like(bn))
    return loss, inv_t, inv_b


class ManoFitter:
    """Fits `n_hands` MANO parameter sets (this rank's shard) to target 21-joint keypoints."""

    def __init__(self, layer, n_hands: int, lr=1e-2, betas=(0.9, 0.999), eps=1e-8, group=None, regularize=True, fused=None):
        self.layer = layer
        # one-kernel iterations need MANO's tree (five chains of three joints below the wrist)
        can_fuse = bool(layer._mode & _cabi.MODEL_CHAINS_5X3)
        if fused and not can_fuse:
            raise _cabi.ManoB200Error("fused fitting steps need a MANO-shaped kinematic tree (MB_MODEL_CHAINS_5X3)")
        self.fused = can_fuse if fused is None else bool(fused)
        self.dev = layer._require_device()
        self.B = int(n_hands)
        self.nc = layer.pose_num
        self.lr, self.b1, self.b2, self.eps = float(lr), float(betas[0]), float(betas[1]), float(eps)
        self.group = group
        self.regularize = regularize
        width = 3 + self.nc + 10
        B = self.B
        # one flat buffer per role so that the Adam update is a single launch
        self.params = torch.zeros(B * width, device=self.dev)
        self.grads = torch.zeros_like(self.params)
        self.exp_avg = torch.zeros_like(self.params)
        self.exp_avg_sq = torch.zeros_like(self.params)
        o1, o2 = 3 * B, (3 + self.nc) * B
        self.rot, self.pose, self.beta = (self.params[:o1].view(B, 3), self.params[o1:o2].view(B, self.nc),
                                          self.params[o2:].view(B, 10))
        self.g_rot, self.g_pose, self.g_beta = (self.grads[:o1].view(B, 3), self.grads[o1:o2].view(B, self.nc),
                                                self.grads[o2:].view(B, 10))
        self.joints = torch.empty(B, 21, 3, device=self.dev)
        self.g_joints = torch.empty(B, 21, 3, device=self.dev)
        self.accum = torch.zeros(2, dtype=torch.float64, device=self.dev)
        self.partials = torch.zeros(N_PARTIALS, dtype=torch.float64, device=self.dev)
        self.l2_out = torch.zeros((), device=self.dev)
        self.one = torch.ones((), device=self.dev)
        self.steps = 0
        self.loss = torch.zeros((), dtype=torch.float64, device=self.dev)
        # fused path: {N_vis, sum theta^2, sum beta^2} of the current parameters (global), this rank's kernel partials
        self.globals = torch.zeros(3, dtype=torch.float64, device=self.dev)
        self.kernel_partials = torch.zeros(3, dtype=torch.float64, device=self.dev)
        self._vis_key = None
        self._param_version = None

    def refresh(self, keypoint_vis: torch.Tensor) -> None:
        """(Re)compute the batch-global scalars of the fused path from scratch: visible count and parameter norms,
        all-reduced.  Called automatically on the first step, when another visibility mask is passed and when the
        parameters were written from outside (``fit.pose.copy_(...)``); with a sharded batch every rank must do so
        in the same iteration (it is a collective)."""
        g = self.globals
        g[0] = keypoint_vis.ne(0).sum(dtype=torch.float64)
        g[1] = self.pose.double().square().sum()
        g[2] = self.beta.double().square().sum()
        reduce_partials(g, self.group)
        self._param_version = self.params._version

    def _step_fused(self, tgt, vis, vis_key, stream):
        lib = _cabi.lib()
        # vis_key identifies the mask the CALLER passed (a bool / non-contiguous mask is converted to a fresh fp32
        # tensor on every call, so the converted tensor's address says nothing)
        if self._vis_key != vis_key or self._param_version != self.params._version:
            self.refresh(vis)
            self._vis_key = vis_key
        self.steps += 1
        _cabi.check(lib.mb_mano_fit_step(self.layer._blob.data_ptr(), self.nc, self.params.data_ptr(), self.exp_avg.data_ptr(),
                                         self.exp_avg_sq.data_ptr(), tgt.data_ptr(), vis.data_ptr(), self.B, self.layer._mode,
                                         self.globals.data_ptr(), self.kernel_partials.data_ptr(), self.lr, self.b1, self.b2,
                                         self.eps, self.steps, int(self.regularize), stream), "mb_mano_fit_step")
        # the one collective of the iteration: {L2 sum of this iteration, norms of the updated parameters} (24 bytes) ...
        reduce_partials(self.kernel_partials, self.group)
        # ... and one single-thread launch: the loss of the iteration just done into self.loss, [S, N, sum theta^2,
        # sum beta^2] into self.partials, the updated norms into next iteration's globals — no torch op on the path
        _cabi.check(lib.mb_fit_finalize(self.globals.data_ptr(), self.kernel_partials.data_ptr(), int(self.regularize),
                                        self.partials.data_ptr(), self.loss.data_ptr(), stream), "mb_fit_finalize")
        self._param_version = self.params._version
        return self.loss

    @_cabi.on_tensor_device
    def step(self, target_joints: torch.Tensor, keypoint_vis: torch.Tensor) -> torch.Tensor:
        """One Adam iteration; returns the (global) loss as a 0-dim device tensor (no host sync)."""
        lib = _cabi.lib()
        B, nc, dev = self.B, self.nc, self.dev
        stream = _cabi.stream_handle(dev)
        blob = self.layer._blob.data_ptr()
        mode = self.layer._mode
        tgt = target_joints.contiguous()
        vis = keypoint_vis if (keypoint_vis.dtype == torch.float32 and keypoint_vis.is_contiguous()) else \
            keypoint_vis.to(torch.float32).contiguous()
        if self.fused:
            key = (keypoint_vis.data_ptr(), keypoint_vis._version, tuple(keypoint_vis.shape), keypoint_vis.dtype)
            return self._step_fused(tgt, vis, key, stream)
        # 1. joints-only forward (no 778-vertex contraction)
        _cabi.check(lib.mb_mano_forward(blob, nc, self.rot.data_ptr(), self.pose.data_ptr(), self.beta.data_ptr(), B, mode,
                                        None, self.joints.data_ptr(), None, 0, stream), "mb_mano_forward")
        # 2. rank-local partials of the objective
        _cabi.check(lib.mb_masked_joint_reduce(self.joints.data_ptr(), tgt.data_ptr(), vis.data_ptr(), _cabi.VIS_F32,
                                               B * 21, 3, _cabi.REDUCE_L2, self.accum.data_ptr(), self.l2_out.data_ptr(),
                                               stream), "mb_masked_joint_reduce")
        self.partials[0:2] = self.accum
        if self.regularize:
            self.partials[2] = self.pose.double().square().sum()
            self.partials[3] = self.beta.double().square().sum()
        else:
            self.partials[2:] = 0
        # 3. the one collective of the iteration
        reduce_partials(self.partials, self.group)
        self.loss, inv_t, inv_b = objective_from_partials(self.partials)
        # 4. gradient of the global L2 term w.r.t. this rank's joints, with the GLOBAL visible count
        self.accum.copy_(self.partials[0:2])
        _cabi.check(lib.mb_masked_l2_backward(self.joints.data_ptr(), tgt.data_ptr(), vis.data_ptr(), _cabi.VIS_F32, B * 21, 3,
                                              self.accum.data_ptr(), self.one.data_ptr(), self.g_joints.data_ptr(), stream),
                    "mb_masked_l2_backward")
        # 5. joints-only backward
        _cabi.check(lib.mb_mano_backward(blob, nc, self.rot.data_ptr(), self.pose.data_ptr(), self.beta.data_ptr(), None,
                                         self.g_joints.data_ptr(), B, mode, 0, self.g_rot.data_ptr(), self.g_pose.data_ptr(),
                                         self.g_beta.data_ptr(), None, 0, stream), "mb_mano_backward")
        if self.regularize:
            self.g_pose.add_(self.pose * (inv_t / 100.0).float())
            self.g_beta.add_(self.beta * (inv_b / 10.0).float())
        # 6. fused Adam on the flat parameter buffer
        self.steps += 1
        _cabi.check(lib.mb_adam_step(self.params.data_ptr(), self.grads.data_ptr(), self.exp_avg.data_ptr(),
                                     self.exp_avg_sq.data_ptr(), self.params.numel(), self.lr, self.b1, self.b2, self.eps,
                                     self.steps, stream), "mb_adam_step")
        return self.loss
