#!/usr/bin/env python
"""bench.py — MANO hands/sec fwd+bwd on N B200s (one process per GPU), the headline metric of
BASELINE.json, with the roofline of the dominant kernel, an end-to-end (host buffers) number
and the CPU baseline — the reference's own PyTorch ManoLayer (oracle/_ref, staged by
oracle/make_ref.py) forward + autograd backward on the box's host cores, with the numpy port beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--hands H] [--mode fp32|f16x3|f16]
    python bench.py --impl reference ...     # the reference's own CPU path (PyTorch; numpy port when it is not staged)
    torchrun --nproc-per-node N bench.py --gpus N ...

Workload (config.workload): BASELINE configs[3] scaled to fwd+bwd — full 45-D axis-angle MANO
(pose_num=45 with an identity PCA basis and zero pose mean, i.e. "no PCA"), H = 2^20 synthetic
hands PER GPU (weak scaling: every rank skins its own slice, no data-path collective), forward
(verts[H,778,3] + joints[H,21,3]) followed by the full backward (g_verts and g_joints dense).
configs[1] (B=4096) is an 80 MB working set that lives in L2 and takes ~12 us at the HBM
roofline, so it cannot be timed under the L2/clock rules; it is a parity-test case
(tests/test_gpu_parity.py) and can be timed with --hands 4096 --rotate 16.
One "step" = one fwd+bwd pass over the H hands.  Inputs are larger than L2 (>= 20 GB per step).
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "MANO hands/sec fwd+bwd"
UNIT = "hands/s"
# algorithmic bytes per hand (SURVEY 8d / BASELINE.md 4), fp32, nc = 45
BYTES_FWD = 4 * (3 + 45 + 10) + 4 * (2334 + 63)                 # 9820
BYTES_BWD = 4 * (2334 + 63) + 4 * (3 + 45 + 10) * 2             # 10052
BYTES_LBS = 19692                                               # stand-alone LBS forward: v_posed 9336 + bones 768 in, verts 9336 + tips 252 out
# stand-alone LBS backward: g_verts 9336 + tip grads 60 + v_posed_t 9408 + bones 768 in,
# dv_posed tiles (bf16 hi+mid, 4 B per coordinate) 9408 + per-bone sums 768 out
BYTES_LBS_BWD = 9336 + 60 + 9408 + 768 + 9408 + 768             # 29748
FLOP_BLEND = 2 * 145 * 2334                                     # 676860 per hand per contraction
FLOP_SKIN_T = 2 * 778 * 16 * 12                                 # 298752 per hand: T_v = sum_k w_vk A_k as a dense product (vskin.cu)
# the fused lane = vertex forward kernel (vskin.cu): feature tiles 640 + bone operand 1152 in, verts 9336 + tips 60 out;
# a training forward also leaves the rest-pose scratch of the skinning backward (9408)
BYTES_FUSED_FWD = 640 + 768 + 9336 + 60                         # 10804: feature tiles + fp32 bone transforms in, verts + tips out
BYTES_FUSED_FWD_TRAIN = BYTES_FUSED_FWD + 9408                  # 20212
STAGE_KERNEL = {"pose_fwd": "pose_forward_lh_kernel", "blend_fwd": "blend_tc_forward_mres_kernel", "lbs_fwd": "skin_forward_kernel",
                "fused_fwd": "vskin_forward_kernel",
                "lbs_bwd": "skin_backward_kernel", "blend_bwd": "blend_tc_backward_kernel", "pose_bwd": "pose_backward_lh_kernel"}


def ncu_traffic_per_hand():
    """DRAM bytes per hand per kernel from the committed `ncu --set full` captures (profiles/r2 overrides profiles/r1)."""
    out = {}
    for rnd in ("r1", "r2"):
        path = os.path.join(ROOT, "profiles", rnd, "traffic_final.json")
        if os.path.isfile(path):
            d = json.load(open(path))["kernels"]
            out.update({k: ((v["dram_read_bytes"] + v["dram_write_bytes"]) / v["hands"], rnd,
                            {x: round(y, 1) for x, y in v.items() if x.endswith("_pct")}) for k, v in d.items()})
    return out


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        d = json.load(open(path))
        return {"hbm_gbs": d.get("hbm_gbs", 6650.0), "bf16_tflops": d.get("bf16_tflops", 1590.0),
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", 1400.0), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons during the timed region (NVML)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.nv = None

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop_evt.wait(0.02)

    def stop(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=2)
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def no_pca_model(assets):
    """Synthetic MANO-shaped model with hands_components = I and hands_mean = 0, so the 45
    coefficients ARE the articulated axis-angles (the oracle of BASELINE config 4)."""
    import numpy as np

    m = assets.synthetic_mano()
    m["hands_components"] = np.eye(45)
    m["hands_mean"] = np.zeros(45)
    return m


def synth_inputs(H, seed):
    import numpy as np

    rs = np.random.RandomState(seed)
    rot = ((rs.rand(H, 3) - .5) * 2 * np.pi).astype(np.float32)
    pose = ((rs.rand(H, 45) - .5) * (np.pi / 2)).astype(np.float32)     # U(-pi/2, pi/2) * 0.5
    beta = (rs.rand(H, 10) - .5).astype(np.float32)
    return rot, pose, beta


def reference_layer(assets, model):
    """The UNMODIFIED reference's ManoLayer (network/sub_modules/MANOLayer.py) on CPU, constructed on `model` written as a
    pickle its own constructor opens — from /root/reference or its travelling copy oracle/_ref.  None when neither is there."""
    try:
        from oracle import ref_import

        if not ref_import.available():
            return None
        import tempfile

        import torch

        ref = ref_import.load()
        torch.set_num_threads(os.cpu_count() or 1)
        with tempfile.TemporaryDirectory() as td:
            pkl = os.path.join(td, "model.pkl")
            assets.write_reference_style_pkl(model, pkl)
            return ref.ManoLayer("cpu", pkl, pose_num=45)
    except Exception as exc:                                   # the baseline must never take the GPU arm down
        sys.stderr.write(f"reference ManoLayer unavailable: {exc!r}\n")
        return None


def reference_fwd_bwd(layer, rot, pose, beta, gv, gj):
    import torch

    t = [torch.from_numpy(a).requires_grad_() for a in (rot, pose, beta)]
    v, j = layer(*t)
    ((v * torch.from_numpy(gv)).sum() + (j * torch.from_numpy(gj)).sum()).backward()
    return v, j, t


def parity_check(model, kept):
    """fp64 oracle on the fixed subsample of the GPU arm's last timed step (8 192 hands): forward on
    all of it, gradients on its first 512 hands.  The one place the product arm runs oracle/ — as the checker."""
    import numpy as np
    from oracle import mano_oracle

    n = kept["rot"].shape[0]
    ev, ej = [], []
    for s0 in range(0, n, 1024):
        sl = slice(s0, s0 + 1024)
        ov, oj = mano_oracle.mano_forward(model, kept["rot"][sl], kept["pose"][sl], kept["beta"][sl])
        ev.append(np.abs(kept["verts"][sl] - ov).max(axis=(1, 2)))
        ej.append(np.abs(kept["joints"][sl] - oj).max(axis=(1, 2)))
    ev, ej = np.concatenate(ev), np.concatenate(ej)
    ng = min(n, 512)
    og = mano_oracle.mano_backward(model, kept["rot"][:ng], kept["pose"][:ng], kept["beta"][:ng], kept["gv"][:ng], kept["gj"][:ng])
    return {"hands": int(n), "verts_max_abs_err_m": float(ev.max()), "verts_p999_abs_err_m": float(np.quantile(ev, 0.999)),
            "joints_max_abs_err_m": float(ej.max()), "joints_p999_abs_err_m": float(np.quantile(ej, 0.999)),
            "grad_hands": int(ng),
            "grad_rel_err": max(float(np.abs(kept[k][:ng] - w).max() / np.abs(w).max())
                                for k, w in zip(("g_rot", "g_pose", "g_beta"), og)),
            "arbiter": "fp64 numpy oracle (oracle/mano_oracle.py), per-hand max over vertices"}


def cpu_baseline(assets, model, budget_s=12.0, chunk=512):
    """The reference's own PyTorch CPU path (BASELINE.md 3 protocol: forward + autograd backward, <= 1 024-hand chunks,
    torch.set_num_threads(cpu_count)) on the workload's synthetic no-PCA model for about `budget_s` seconds, with the
    numpy port (oracle/mano_oracle.py, fp32) timed beside it.  kind = "reference" unless the reference is not staged."""
    import numpy as np
    from oracle import mano_oracle

    rot, pose, beta = synth_inputs(chunk, 4242)
    rs = np.random.RandomState(1)
    gv = rs.randn(chunk, 778, 3).astype(np.float32)
    gj = rs.randn(chunk, 21, 3).astype(np.float32)

    def timed(fn, budget):
        fn()                                                   # warm-up
        done, t0 = 0, time.perf_counter()
        while True:
            fn()
            done += chunk
            dt = time.perf_counter() - t0
            if dt >= budget:
                return done / dt, done, dt

    port, pdone, pdt = timed(lambda: mano_oracle.mano_backward(model, rot, pose, beta, gv, gj, dtype=np.float32), budget_s * 0.4)
    layer = reference_layer(assets, model)
    if layer is None:
        return {"value": port, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                "sample": f"{pdone} hands in {chunk}-hand batches, numpy fp32 oracle fwd+bwd, {pdt:.1f} s (reference not staged)"}
    val, done, dt = timed(lambda: reference_fwd_bwd(layer, rot, pose, beta, gv, gj), budget_s)
    # config 1 of BASELINE.json: the reference forward at batch 64
    import torch

    r64 = [torch.from_numpy(a[:64]) for a in (rot, pose, beta)]
    with torch.no_grad():
        layer(*r64)
        best = min(_timeit(lambda: layer(*r64)) for _ in range(5))
    return {"value": val, "unit": UNIT, "cores": os.cpu_count(), "kind": "reference",
            "sample": f"{done} hands in {chunk}-hand batches, the reference's PyTorch ManoLayer forward + autograd backward "
                      f"(fp32, {os.cpu_count()} threads), {dt:.1f} s",
            "config1_forward_b64": {"hands_per_s": 64 / best, "ms": best * 1e3},
            "port_value": port, "port_note": "numpy fp32 port of the same algorithm (BLAS matmuls instead of repeat + bmm)"}


def _timeit(fn):
    t0 = time.perf_counter()
    fn()
    return time.perf_counter() - t0


def run_reference_arm(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path — its PyTorch ManoLayer forward + autograd
    backward from the staged copy oracle/_ref (the numpy port only when that is absent) — on the host cores; each step is
    a bounded sample (args.ref_hands hands) of the workload."""
    if rank != 0:
        return
    import numpy as np
    from oracle import mano_oracle

    assets = importlib.import_module("3dhandposeestimation_b200.assets")
    model = no_pca_model(assets)
    H = args.ref_hands
    rot, pose, beta = synth_inputs(H, 4242)
    rs = np.random.RandomState(1)
    gv = rs.randn(H, 778, 3).astype(np.float32)
    gj = rs.randn(H, 21, 3).astype(np.float32)
    layer = reference_layer(assets, model)
    if layer is not None:
        kind, what = "reference", "the reference's PyTorch ManoLayer forward + autograd backward (fp32)"
        fn = lambda: reference_fwd_bwd(layer, rot, pose, beta, gv, gj)
    else:
        kind, what = "port", "numpy fp32 oracle fwd+bwd (reference not staged)"
        fn = lambda: mano_oracle.mano_backward(model, rot, pose, beta, gv, gj, dtype=np.float32)
    for _ in range(args.warmup):
        fn()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        fn()
    dt = time.perf_counter() - t0
    value = H * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args), "sample_hands_per_step": H, "device": "cpu"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": os.cpu_count(), "kind": kind,
                         "sample": f"{H} hands per step x {args.steps} steps, {what}"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def run_secondary(args, pkg, layer, dev, rank, world, dist):
    """The other measured BASELINE configs (not the contract line): config 5 fitting loop, config 3 FK."""
    import numpy as np
    import torch

    H = args.hands
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(fn, n):
        for _ in range(max(args.warmup, 3)):
            fn()
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / n

    if args.workload == "fit":
        # config 5: Adam on (rot, pose, beta) of H hands per GPU against target keypoints; one NCCL all-reduce of
        # 4 doubles per iteration when world > 1
        rot, pose, beta = synth_inputs(H, 777 + rank)
        with torch.no_grad():
            _, tgt = layer.rot_pose_beta_to_mesh(*[torch.from_numpy(a).to(dev) for a in (rot, pose, beta)], joints_only=True)
            tgt = tgt + 1e-3 * torch.randn_like(tgt)
        vis = (torch.rand(H, 21, 1, device=dev) < 0.8).float()
        fitter = pkg.fitting.ManoFitter(layer, H)
        loss0 = float(fitter.step(tgt, vis))
        l2_0 = float(fitter.partials[0] / fitter.partials[1].clamp(min=1))
        ms = timed(lambda: fitter.step(tgt, vis), args.steps)
        loss1 = float(fitter.loss)
        l2_1 = float(fitter.partials[0] / fitter.partials[1].clamp(min=1))
        return {"metric": "MANO fitting-loop hand-iterations/sec (config 5)", "value": world * H / (ms * 1e-3), "unit": "hand-iterations/s",
                "n_gpus": world, "steps": args.steps, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
                "config": {"workload": f"mano_fit_adam_{H}_hands_per_gpu", "collective": "all-reduce of 4 doubles per iteration" if world > 1 else "none"},
                "loss_first": loss0, "loss_last": loss1, "l2_term_first_m2": l2_0, "l2_term_last_m2": l2_1,
                "note": "loss = masked L2 + the reference's batch-global Frobenius regulariser (loss.py:113-117), which grows with "
                        "sqrt(batch) and dominates at 2^20 hands; the L2 term is reported separately", "data": "synthetic"}
    if args.workload == "aux":
        # the joint epilogue, keypoint re-parameterisations, viewpoint epilogue and hand-mask loss,
        # each timed alone through the C ABI on H hands (every tensor is larger than L2) against the HBM roofline
        cabi = pkg._cabi
        lib = cabi.lib()
        peaks = load_peaks()
        st = torch.cuda.current_stream(dev).cuda_stream
        P = lambda t: t.data_ptr()
        g = torch.Generator(device=dev).manual_seed(5 + rank)
        R = lambda *shape: torch.randn(*shape, device=dev, generator=g)
        joints = R(H, 21, 3) * .05
        joints[:, 0] = 0
        xyz_in = (R(H, 21, 3) * .5 + torch.arange(21, device=dev)[:, None] * torch.tensor([.09, .06, .03], device=dev)).contiguous()
        scale = torch.rand(H, 1, device=dev, generator=g) * .05 + .02
        root = R(H, 3) * .05 + torch.tensor([0, 0, .6], device=dev)
        K = torch.tensor([[282.9, 0, 160], [0, 282.9, 160], [0, 0, 1.]], device=dev).repeat(H, 1, 1).contiguous()
        o63 = [torch.empty(H, 21, 3, device=dev) for _ in range(3)]
        o42 = torch.empty(H, 21, 2, device=dev)
        g63 = [R(H, 21, 3) for _ in range(2)]
        g42 = R(H, 21, 2) * 1e-3
        o9 = torch.empty(H, 3, 3, device=dev)
        g9 = R(H, 3, 3)
        u = [R(H) for _ in range(3)]
        gu = [torch.empty(H, device=dev) for _ in range(3)]
        gs, gr = torch.empty(H, 1, device=dev), torch.empty(H, 3, device=dev)
        Hm = min(H, 65536)                                       # hand-mask loss: [Hm][64][64] fp32 masks = 1 GB at 65 536
        mask = (torch.rand(Hm, 64, 64, device=dev, generator=g) < .4).float()
        uvp, uvg = torch.rand(Hm, 21, 2, device=dev, generator=g) * 70 - 3, torch.rand(Hm, 21, 2, device=dev, generator=g) * 70 - 3
        acc, out1 = torch.zeros(2, dtype=torch.float64, device=dev), torch.zeros((), device=dev)
        cases = [
            ("joint_epilogue_forward", 252 + 4 + 12 + 36 + 252 + 252 + 168, H,
             lambda: lib.mb_joint_epilogue_forward(P(joints), P(scale), P(root), P(K), H, 0, P(o63[0]), P(o63[1]), P(o42), st)),
            ("joint_epilogue_backward", 252 + 52 + 252 + 252 + 168 + 252 + 16, H,
             lambda: lib.mb_joint_epilogue_backward(P(joints), P(scale), P(root), P(K), P(g63[0]), P(g63[1]), P(g42), H, 0, P(o63[2]),
                                                    P(gs), P(gr), st)),
            ("bone_rel_trafo", 504, H, lambda: lib.mb_bone_rel_trafo(P(xyz_in), H, P(o63[0]), st)),
            ("bone_rel_trafo_inv", 504, H, lambda: lib.mb_bone_rel_trafo_inv(P(o63[0]), H, P(o63[1]), st)),
            ("canonical_trafo", 540, H, lambda: lib.mb_canonical_trafo(P(xyz_in), None, H, P(o63[2]), P(o9), st)),
            ("viewpoint_forward", 252 + 12 + 36 + 252, H,
             lambda: lib.mb_viewpoint_forward(P(xyz_in), P(u[0]), P(u[1]), P(u[2]), None, None, None, H, P(o9), P(o63[0]), None, None, st)),
            ("viewpoint_backward", 252 + 12 + 36 + 252 + 252 + 12, H,
             lambda: lib.mb_viewpoint_backward(P(xyz_in), P(u[0]), P(u[1]), P(u[2]), P(g9), P(g63[0]), H, P(o63[1]), P(gu[0]), P(gu[1]),
                                               P(gu[2]), st)),
            ("hand_mask_loss", 2 * 168 + 42 * 32, Hm,
             lambda: lib.mb_hand_mask_loss(P(uvp), P(uvg), P(mask), cabi.VIS_F32, Hm, 21, 64, 64, P(acc), P(out1), st)),
        ]
        res = {}
        for name, nbytes, n, fn in cases:
            def call(fn=fn, name=name):
                cabi.check(fn(), name)
            ms = timed(call, args.steps)
            gbs = nbytes * n / (ms * 1e-3) / 1e9
            res[name] = {"ms": ms, "items": n, "bytes_per_item": nbytes, "achieved": gbs, "unit": "GB/s", "peak": peaks["hbm_gbs"],
                         "frac": gbs / peaks["hbm_gbs"], "bound": "hbm"}
        return {"metric": "SURVEY 8(f) kernels, each alone (hands/s of the joint epilogue forward + backward)",
                "value": world * H / ((res["joint_epilogue_forward"]["ms"] + res["joint_epilogue_backward"]["ms"]) * 1e-3), "unit": "hands/s",
                "n_gpus": world, "steps": args.steps, "ms_per_step": res["joint_epilogue_forward"]["ms"] + res["joint_epilogue_backward"]["ms"],
                "higher_is_better": True, "scaling": "weak", "config": {"workload": f"aux_kernels_{H}_hands_per_gpu",
                "note": "hand_mask_loss: bytes = 2 x 21 uv pairs + 42 sampled 32-byte sectors per hand (gather)"},
                "roofline": res, "data": "synthetic"}
    # config 3: RHD 21-joint FK forward + backward + visible-joint MPJPE, rotating buffer sets (the working set of one
    # 65 536-sample call fits in L2)
    B = 65536 if args.hands == (1 << 20) else args.hands
    nsets = 16
    fk = pkg.ForwardKinematics(dev)
    mp = pkg.MPJPE()
    sets = []
    for i in range(nsets):
        rs = np.random.RandomState(31 + i + 100 * rank)
        a = [((rs.rand(B, 3) - .5) * 2 * np.pi), ((rs.rand(B, 23) - .5) * np.pi), rs.rand(B, 20) + .1]
        K = np.tile(np.array([[282.9, 0, 160], [0, 282.9, 160], [0, 0, 1]]), (B, 1, 1))
        sc = rs.rand(B, 1) * .05 + .02
        root = rs.randn(B, 3) * .05 + np.array([0, 0, .6])
        t = [torch.from_numpy(x.astype(np.float32)).to(dev) for x in (*a, K, sc, root)]
        for x in t[:3]:
            x.requires_grad_()
        gt = torch.from_numpy((rs.randn(B, 21, 3) * .05 + np.array([0, 0, .6])).astype(np.float32)).to(dev)
        vis = torch.from_numpy((rs.rand(B, 21, 1) < .8).astype(np.float32)).to(dev)
        sets.append((t, gt, vis))
    it = [0]

    def step():
        t, gt, vis = sets[it[0] % nsets]
        it[0] += 1
        xyz, uv, _ = fk(*t)
        loss = mp(xyz, gt, vis) + 1e-3 * uv.sum()
        loss.backward()
        for x in t[:3]:
            x.grad = None

    ms = timed(step, args.steps)

    # The same training step straight through the C ABI, all rotating sets captured in ONE CUDA graph (the call is
    # launch-bound through Python at this size): FK forward -> L2Loss partials -> dL/dxyz -> FK backward -> MPJPE.
    cabi = pkg._cabi
    lib = cabi.lib()
    P = lambda t: t.data_ptr()
    outs = [dict(xyz=torch.empty(B, 21, 3, device=dev), uv=torch.empty(B, 21, 2, device=dev), g_xyz=torch.empty(B, 21, 3, device=dev),
                 g=[torch.empty(B, n, device=dev) for n in (3, 23, 20)], acc=torch.zeros(2, dtype=torch.float64, device=dev),
                 acc2=torch.zeros(2, dtype=torch.float64, device=dev), l2=torch.zeros((), device=dev), mp=torch.zeros((), device=dev))
            for _ in range(nsets)]
    one = torch.ones((), device=dev)

    def raw_step(i, stream):
        t, gt, vis = sets[i]
        o = outs[i]
        ra, oa, bl, K, sc, root = t
        cabi.check(lib.mb_fk_forward(P(ra), P(oa), P(bl), P(K), P(sc), P(root), B, 0, P(o["xyz"]), P(o["uv"]), stream), "fk_forward")
        cabi.check(lib.mb_masked_joint_reduce(P(o["xyz"]), P(gt), P(vis), cabi.VIS_F32, B * 21, 3, cabi.REDUCE_L2, P(o["acc"]), P(o["l2"]),
                                              stream), "reduce")
        cabi.check(lib.mb_masked_l2_backward(P(o["xyz"]), P(gt), P(vis), cabi.VIS_F32, B * 21, 3, P(o["acc"]), P(one), P(o["g_xyz"]),
                                             stream), "l2_backward")
        cabi.check(lib.mb_fk_backward(P(ra), P(oa), P(bl), P(K), P(sc), P(root), P(o["g_xyz"]), None, B, 0, P(o["g"][0]), P(o["g"][1]),
                                      P(o["g"][2]), stream), "fk_backward")
        cabi.check(lib.mb_masked_joint_reduce(P(o["xyz"]), P(gt), P(vis), cabi.VIS_F32, B * 21, 3, cabi.REDUCE_MPJPE_MM, P(o["acc2"]),
                                              P(o["mp"]), stream), "mpjpe")

    side = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(side):
        for i in range(nsets):
            raw_step(i, side.cuda_stream)
    side.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        for i in range(nsets):
            raw_step(i, torch.cuda.current_stream(dev).cuda_stream)
    ms_graph = timed(graph.replay, max(3, args.steps // 4)) / nsets
    # per-kernel times (CUDA events around every launch, outside the graph) for the roofline of the two FK kernels
    lib.mb_profile_enable(1)
    cabi.profile_collect()
    for i in range(nsets):
        raw_step(i, torch.cuda.current_stream(dev).cuda_stream)
    prof = cabi.profile_collect()
    lib.mb_profile_enable(0)
    # The same training step with the two L2Loss terms INSIDE the FK kernels (mb_fk_loss_forward / _backward: one launch + one
    # 64-byte memset per direction) — through the C ABI as one graph over the rotating sets, and through the ForwardKinematicsLoss
    # module (one autograd node instead of three)
    gt_uvs = [torch.rand(B, 21, 2, device=dev) * 320 for _ in range(nsets)]
    louts = [dict(losses=torch.zeros(2, device=dev), ws=torch.zeros(8, dtype=torch.float64, device=dev)) for _ in range(nsets)]
    g_l = torch.tensor([1.0, 1e-3], device=dev)
    both = cabi.HEAD_XYZ | cabi.HEAD_UV

    def loss_step(i, stream):
        t, gt, vis = sets[i]
        o, lo = outs[i], louts[i]
        ra, oa, bl, K, sc, root = t
        cabi.check(lib.mb_fk_loss_forward(P(ra), P(oa), P(bl), P(K), P(sc), P(root), P(gt), P(gt_uvs[i]), P(vis), B, 0, both, P(o["xyz"]),
                                          P(o["uv"]), P(lo["losses"]), P(lo["ws"]), 64, stream), "fk_loss_forward")
        cabi.check(lib.mb_fk_loss_backward(P(ra), P(oa), P(bl), P(K), P(sc), P(root), P(gt), P(gt_uvs[i]), P(vis), B, 0, both, P(g_l),
                                           P(o["g"][0]), P(o["g"][1]), P(o["g"][2]), P(lo["ws"]), 64, stream), "fk_loss_backward")

    with torch.cuda.stream(side):
        for i in range(nsets):
            loss_step(i, side.cuda_stream)
    side.synchronize()
    lgraph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(lgraph, stream=side):
        for i in range(nsets):
            loss_step(i, torch.cuda.current_stream(dev).cuda_stream)
    ms_loss_graph = timed(lgraph.replay, max(3, args.steps // 4)) / nsets
    crit = pkg.ForwardKinematicsLoss(dev)

    def loss_module_step():
        t, gt, vis = sets[it[0] % nsets]
        gu = gt_uvs[it[0] % nsets]
        it[0] += 1
        lx, lu, _, _ = crit(*t, gt, gu, vis)
        (lx + 1e-3 * lu).backward()
        for x in t[:3]:
            x.grad = None

    ms_loss_module = timed(loss_module_step, args.steps)
    fk_loss = {"c_abi_graph_ms_per_step": ms_loss_graph, "c_abi_graph_samples_per_s": world * B / (ms_loss_graph * 1e-3),
               "module_api_ms_per_step": ms_loss_module, "module_api_samples_per_s": world * B / (ms_loss_module * 1e-3),
               "loss_xyz": float(louts[0]["losses"][0]), "loss_uv": float(louts[0]["losses"][1]),
               "note": "FK forward + L2Loss(xyz) + L2Loss(uv) and back: mb_fk_loss_forward / _backward (one kernel + one 64-byte "
                       "memset per direction); module = ForwardKinematicsLoss + (loss_xyz + 1e-3 loss_uv).backward()"}
    peaks = load_peaks()
    roof = {}
    for stage, kern, nbytes in (("fk_fwd", "fk_forward_kernel", 656), ("fk_bwd", "fk_backward_kernel", 840 + 252)):
        if stage in prof:
            kms = prof[stage][0] / prof[stage][1]
            gbs = nbytes * B / (kms * 1e-3) / 1e9
            roof[kern] = {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"],
                          "algorithmic_bytes_per_sample": nbytes, "avg_launch_ms": kms, "traffic": None}
    return {"metric": "RHD FK fwd+bwd+MPJPE samples/sec (config 3)", "value": world * B / (ms_graph * 1e-3), "unit": "samples/s",
            "n_gpus": world, "steps": args.steps, "ms_per_step": ms_graph, "higher_is_better": True, "scaling": "weak",
            "config": {"workload": f"fk_fwd+l2+bwd+mpjpe_{B}_samples_per_gpu", "l2": f"rotating {nsets} buffer sets",
                       "api": "C ABI, the rotating sets captured in one CUDA graph (7 kernels + 2 memsets per step)"},
            "module_api": {"value": world * B / (ms * 1e-3), "ms_per_step": ms,
                           "note": "ForwardKinematics + MPJPE nn.Modules with torch autograd: launch- and Python-bound at this size"},
            "fk_loss": fk_loss,
            "roofline": roof, "mpjpe_mm": float(outs[0]["mp"]), "l2": float(outs[0]["l2"]), "data": "synthetic"}


def bind_numa_local(gpu_index):
    """Pin this process (and therefore the pinned host buffers it first-touches) to the CPU cores of the GPU's NUMA node:
    with 8 ranks on one box, round 1's e2e arm scaled at 0.66 with every rank on the same cores.  Best effort."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        bdf = pynvml.nvmlDeviceGetPciInfo(h).busId
        bdf = (bdf.decode() if isinstance(bdf, bytes) else bdf).lower()
        if len(bdf.split(":")[0]) == 8:                       # 00000000:xx:yy.z -> 0000:xx:yy.z
            bdf = bdf[4:]
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0)
        use = (cpus & allowed) or allowed
        os.sched_setaffinity(0, use)
        # ... and ask for this process's new pages on that node even when the container's cpuset holds none of its cores
        # (set_mempolicy(MPOL_PREFERRED); x86-64 syscall 238, best effort)
        mem = None
        try:
            import ctypes

            mask = ctypes.c_ulong(1 << node)
            mem = ctypes.CDLL(None, use_errno=True).syscall(238, 1, ctypes.byref(mask), ctypes.c_ulong(8 * ctypes.sizeof(mask)))
        except Exception:
            pass
        return {"numa_node": node, "cpus": len(use), "node_cpus_in_cpuset": len(cpus & allowed), "set_mempolicy_rc": mem}
    except Exception:
        return None


def run_extras(args, pkg, layer45, dev, rank, world, dist, set0):
    """The other BASELINE configs, short, for the contract line: config 2 (the Resnet50MANO3DHandPose head workload:
    B = 4096 and the reference's own batch_size = 200, config.py:79; joints only as the heads consume it, and with verts),
    config 3 (FK fwd + L2 + bwd + MPJPE at 65 536 samples: C-ABI graph and the nn.Module API) and config 5 (fused Adam
    fitting iterations with the NCCL all-reduce inside the timed loop, and with it disabled)."""
    import numpy as np
    import torch

    cabi = pkg._cabi
    lib = cabi.lib()
    P = lambda t: t.data_ptr()
    st = torch.cuda.current_stream(dev).cuda_stream
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(fn, n, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / n

    out = {}
    # ---- config 2: head workload, pose_num = config.mano_pose_num = 10, value ranges of resnet50MANO.py:73-75 ----
    layer10 = pkg.ManoLayer(dev, model=pkg.assets.synthetic_mano(), pose_num=10, mode=args.mode)
    c2 = {}
    for B in (4096, 200):
        nsets = 16
        rs = np.random.RandomState(7 + B)
        sets = []
        for _ in range(nsets):
            sig = lambda *sh: torch.from_numpy(rs.rand(*sh).astype(np.float32)).to(dev)
            sets.append(dict(rot=(sig(B, 3) - .5) * 2 * np.pi, pose=(sig(B, 10) - .5) * 4, beta=(sig(B, 10) - .5) * .1,
                             joints=torch.empty(B, 21, 3, device=dev), verts=torch.empty(B, 778, 3, device=dev),
                             gj=torch.randn(B, 21, 3, device=dev), gv=torch.randn(B, 778, 3, device=dev),
                             g=[torch.empty(B, n, device=dev) for n in (3, 10, 10)]))
        ws = torch.empty(max(lib.mb_mano_workspace_bytes(B, layer10._mode), 16), dtype=torch.uint8, device=dev)
        it = [0]

        def head_step():
            s = sets[it[0] % nsets]
            it[0] += 1
            cabi.check(lib.mb_mano_forward(P(layer10._blob), 10, P(s["rot"]), P(s["pose"]), P(s["beta"]), B, layer10._mode, None,
                                           P(s["joints"]), None, 0, st), "fwd")
            cabi.check(lib.mb_mano_backward(P(layer10._blob), 10, P(s["rot"]), P(s["pose"]), P(s["beta"]), None, P(s["gj"]), B,
                                            layer10._mode, 0, P(s["g"][0]), P(s["g"][1]), P(s["g"][2]), None, 0, st), "bwd")

        def full_step():
            s = sets[it[0] % nsets]
            it[0] += 1
            cabi.check(lib.mb_mano_forward(P(layer10._blob), 10, P(s["rot"]), P(s["pose"]), P(s["beta"]), B, layer10._mode,
                                           P(s["verts"]), P(s["joints"]), P(ws), ws.numel(), st), "fwd")
            cabi.check(lib.mb_mano_backward(P(layer10._blob), 10, P(s["rot"]), P(s["pose"]), P(s["beta"]), P(s["gv"]), P(s["gj"]), B,
                                            layer10._mode, cabi.BWD_WORKSPACE_VALID, P(s["g"][0]), P(s["g"][1]), P(s["g"][2]),
                                            P(ws), ws.numel(), st), "bwd")

        ms_h = timed(head_step, 64)
        ms_f = timed(full_step, 64)
        # the heads' whole tail in one call per direction (mb_mano_head_loss_*: parameters -> joints -> match_mano_to_RHD ->
        # projection -> L2 xyz + L2 uv + regulariser, and back), per launch sequence and replayed as ONE captured CUDA graph
        hws = torch.empty(max(lib.mb_mano_head_loss_workspace_bytes(B), 16), dtype=torch.uint8, device=dev)
        hK = torch.tensor([[282.9, 0, 160], [0, 282.9, 160], [0, 0, 1.]], device=dev).repeat(B, 1, 1).contiguous()
        hL = torch.rand(B, device=dev) * .05 + .02
        hroot = torch.randn(B, 3, device=dev) * .05 + torch.tensor([0, 0, .6], device=dev)
        hgx, hgu = torch.randn(B, 21, 3, device=dev) * .05 + hroot[:, None, :], torch.rand(B, 21, 2, device=dev) * 320
        hvis = (torch.rand(B, 21, device=dev) < .8).float()
        hxyz, huv = torch.empty(B, 21, 3, device=dev), torch.empty(B, 21, 2, device=dev)
        hloss, hgl = torch.empty(3, device=dev), torch.ones(3, device=dev)
        hflags = cabi.HEAD_XYZ | cabi.HEAD_UV | cabi.HEAD_REG | cabi.HEAD_MATCH

        def head_loss_step(stream=st):
            s = sets[it[0] % nsets]
            it[0] += 1
            a = (P(layer10._blob), 10, P(s["rot"]), P(s["pose"]), P(s["beta"]), None, None, P(hL), P(hroot), P(hK), P(hgx), P(hgu),
                 P(hvis), B, layer10._mode, hflags, 0, 10.0)
            cabi.check(lib.mb_mano_head_loss_forward(*a, P(hxyz), P(huv), P(hloss), P(hws), hws.numel(), stream), "head fwd")
            cabi.check(lib.mb_mano_head_loss_backward(*a, P(hxyz), P(huv), P(hgl), P(s["g"][0]), P(s["g"][1]), P(s["g"][2]), None, None,
                                                      P(hws), hws.numel(), stream), "head bwd")

        ms_hl = timed(head_loss_step, 64)
        graph = torch.cuda.CUDAGraph()
        cap = torch.cuda.Stream(device=dev)
        cap.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(cap):
            it[0] = 0
            head_loss_step(cap.cuda_stream)                    # warm-up on the capture stream
            cap.synchronize()
            with torch.cuda.graph(graph, stream=cap):
                for _ in range(nsets):
                    head_loss_step(cap.cuda_stream)
        torch.cuda.current_stream(dev).wait_stream(cap)
        ms_hg = timed(graph.replay, 8) / nsets
        c2[f"B{B}"] = {"joints_only_fwd+bwd_ms": ms_h, "joints_only_hands_per_s": world * B / (ms_h * 1e-3),
                       "verts_fwd+bwd_ms": ms_f, "verts_hands_per_s": world * B / (ms_f * 1e-3),
                       "head_loss_fwd+bwd_ms": ms_hl, "head_loss_graph_fwd+bwd_ms": ms_hg,
                       "head_loss_hands_per_s": world * B / (ms_hg * 1e-3)}
        del sets, ws, graph
    c2["note"] = ("C ABI, 16 rotating buffer sets, pose_num 10, head value ranges; joints only = what the heads consume "
                  "(resnet50MANO.py:76,87); head_loss = the heads' whole tail, one call per direction (parameters -> joints -> "
                  "match_mano_to_RHD -> projection -> L2 xyz + L2 uv + regulariser and its backward), `graph` = 16 steps replayed as one "
                  "captured CUDA graph")
    out["config2_head"] = c2

    # ---- config 3: FK ----
    import copy

    a3 = copy.copy(args)
    a3.workload, a3.hands, a3.steps = "fk", 1 << 20, max(8, min(args.steps, 20))
    fk = run_secondary(a3, pkg, layer45, dev, rank, world, dist)
    # the same kernels on 2^20 - 1 samples (several tiles per persistent warp, a partial last tile): their HBM rooflines
    a3b = copy.copy(a3)
    a3b.hands, a3b.steps = (1 << 20) - 1, 4
    fkb = run_secondary(a3b, pkg, layer45, dev, rank, world, dist)
    out["config3_fk"] = {"samples": 65536, "c_abi_graph_ms_per_step": fk["ms_per_step"], "c_abi_graph_samples_per_s": fk["value"],
                         "at_1048575_samples": {"c_abi_graph_ms_per_step": fkb["ms_per_step"], "roofline": fkb["roofline"],
                                                "fk_loss_c_abi_graph_ms_per_step": fkb["fk_loss"]["c_abi_graph_ms_per_step"],
                                                "fk_loss_c_abi_graph_samples_per_s": fkb["fk_loss"]["c_abi_graph_samples_per_s"]},
                         "module_api_ms_per_step": fk["module_api"]["ms_per_step"], "module_api_samples_per_s": fk["module_api"]["value"],
                         "fk_loss": fk["fk_loss"],
                         "roofline": fk["roofline"], "mpjpe_mm": fk["mpjpe_mm"]}

    # ---- config 5: fitting iterations, with and without the collective ----
    H = args.hands
    with torch.no_grad():
        _, tgt = layer45.rot_pose_beta_to_mesh(set0["rot"], set0["pose"], set0["beta"], joints_only=True)
        tgt = tgt + 1e-3 * torch.randn_like(tgt)
    vis = (torch.rand(H, 21, 1, device=dev) < 0.8).float()
    res = {}
    for name, group in (("allreduce", None), ("local", "local")):
        fitter = pkg.fitting.ManoFitter(layer45, H, group=group)
        fitter.step(tgt, vis)
        n_it = max(10, min(args.steps, 30))
        ms = timed(lambda: fitter.step(tgt, vis), n_it)
        res[name] = {"ms_per_iteration": ms, "hand_iterations_per_s": world * H / (ms * 1e-3), "loss": float(fitter.loss)}
        del fitter
    res["collective"] = f"NCCL all-reduce of 3 doubles per iteration over {world} ranks" if world > 1 else "single rank: no collective is issued"
    res["hands_per_gpu"] = H
    out["config5_fit"] = res
    return out


def workload_name(args):
    return f"mano_full45_noPCA_fwd+bwd_{args.hands}_hands_per_gpu"


_JSON_OUT = None


def claim_stdout():
    """stdout carries exactly ONE line, the JSON record: everything else a library prints there (NCCL's version banner,
    torchrun's notices in a child) is sent to stderr by pointing fd 1 at fd 2 and keeping the real stdout for emit()."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--hands", type=int, default=1 << 20, help="hands per GPU per step")
    ap.add_argument("--mode", default=os.environ.get("MANO_B200_MODE", "f16x3"))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="mano", choices=["mano", "fit", "fk", "aux"],
                    help="mano (default, the contract line) | fit = BASELINE config 5 fitting loop | fk = config 3 FK fwd+bwd+MPJPE")
    ap.add_argument("--ref-hands", type=int, default=512, help="hands per step of the --impl reference arm")
    ap.add_argument("--rotate", type=int, default=1, help="number of distinct buffer sets cycled through (small --hands)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--unfused", action="store_true", help="A/B: the two separate forward kernels (MB_FWD_UNFUSED) instead of the fused one")
    ap.add_argument("--no-extras", action="store_true", help="skip the config 2 / 3 / 5 side measurements of the default line")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    import __graft_entry__ as ge

    if rank == 0:
        ge.build()
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # stdout carries exactly one JSON line
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist.barrier()
    pkg = importlib.import_module("3dhandposeestimation_b200")
    cabi = pkg._cabi
    lib = pkg.load_library()
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    H = args.hands
    model = no_pca_model(pkg.assets)
    layer = pkg.ManoLayer(dev, model=model, pose_num=45, mode=args.mode, fused_forward=False if args.unfused else None)
    mode = layer._mode | layer._fwd_flags
    stream = cabi.stream_handle(dev)

    if args.workload != "mano":
        line = run_secondary(args, pkg, layer, dev, rank, world, dist)
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        if rank == 0:
            emit(line)
        return

    nsets = max(1, args.rotate)
    sets = []
    for i in range(nsets):
        rot, pose, beta = synth_inputs(H, 1234 + rank * 131 + i)
        g = torch.Generator(device=dev).manual_seed(99 + rank * 7 + i)
        sets.append(dict(
            rot=torch.from_numpy(rot).to(dev), pose=torch.from_numpy(pose).to(dev), beta=torch.from_numpy(beta).to(dev),
            verts=torch.empty(H, 778, 3, device=dev), joints=torch.empty(H, 21, 3, device=dev),
            gv=torch.randn(H, 778, 3, device=dev, generator=g), gj=torch.randn(H, 21, 3, device=dev, generator=g),
            g_rot=torch.empty(H, 3, device=dev), g_pose=torch.empty(H, 45, device=dev), g_beta=torch.empty(H, 10, device=dev)))
    ws_bytes = lib.mb_mano_workspace_bytes(H, mode)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    blob = layer._blob.data_ptr()

    def step(i):
        s = sets[i % nsets]
        cabi.check(lib.mb_mano_forward(blob, 45, s["rot"].data_ptr(), s["pose"].data_ptr(), s["beta"].data_ptr(), H, mode,
                                       s["verts"].data_ptr(), s["joints"].data_ptr(), ws.data_ptr(), ws_bytes, stream), "fwd")
        cabi.check(lib.mb_mano_backward(blob, 45, s["rot"].data_ptr(), s["pose"].data_ptr(), s["beta"].data_ptr(),
                                        s["gv"].data_ptr(), s["gj"].data_ptr(), H, mode, cabi.BWD_WORKSPACE_VALID,
                                        s["g_rot"].data_ptr(), s["g_pose"].data_ptr(), s["g_beta"].data_ptr(),
                                        ws.data_ptr(), ws_bytes, stream), "bwd")

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident timing (value) ----------------------------------------------
    for i in range(max(args.warmup, 3)):
        step(i)
    sync_all()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = lib.mb_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step(i)
    e1.record()
    sync_all()
    clocks = sampler.stop()
    launches = lib.mb_launch_count() - launches0
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    ms_per_step = ms_total / args.steps
    value = world * H / (ms_per_step * 1e-3)

    # ---- roofline pass: per-stage CUDA events on the launch stream ---------------------
    lib.mb_profile_enable(1)
    for i in range(args.steps):
        step(i)
    torch.cuda.synchronize(dev)
    prof = cabi.profile_collect()
    lib.mb_profile_enable(0)
    peaks = load_peaks()
    stages = {k: {"ms": v[0] / v[1], "launches": v[1]} for k, v in prof.items()}
    tot_stage_ms = sum(s["ms"] for s in stages.values())
    for k, s in stages.items():
        s["share"] = s["ms"] / tot_stage_ms
    traffic = ncu_traffic_per_hand()

    def hbm_roofline(stage, bytes_per_hand, ms=None):
        ms = stages[stage]["ms"] if ms is None else ms
        gbs = bytes_per_hand * H / (ms * 1e-3) / 1e9
        kern = STAGE_KERNEL[stage]
        tr = traffic.get(kern)
        return {"kernel": kern, "bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": gbs / peaks["hbm_gbs"], "traffic": tr[0] * H if tr else None,
                "traffic_source": f"ncu dram__bytes_read+write per hand (profiles/{tr[1]}/traffic_final.json) x hands" if tr else None,
                "peak_source": peaks["source"], "algorithmic_bytes_per_hand": bytes_per_hand, "avg_launch_ms": ms,
                "share_of_step": stages[stage]["share"] if stage in stages else None,
                # utilisation of the units that can bound the kernel, from the committed ncu capture (percent of peak)
                "ncu_unit_utilisation_pct": tr[2] if tr else None}

    stage_bytes = {"lbs_bwd": BYTES_LBS_BWD, "lbs_fwd": BYTES_LBS, "fused_fwd": BYTES_FUSED_FWD_TRAIN}
    dominant = max(stages, key=lambda k: stages[k]["ms"])
    roof_stage = dominant if dominant in stage_bytes else max(stage_bytes.keys() & stages.keys(), key=lambda k: stages[k]["ms"])
    # `roofline` is the dominant kernel of the step; the north star's named "LBS GB/s" figure — the skinning forward,
    # now fused with the blend contraction — is reported next to it from the forward-only pass below
    roofline = dict(hbm_roofline(roof_stage, stage_bytes[roof_stage]), dominant_stage=dominant)
    step_gbs = (BYTES_FWD + BYTES_BWD) * H / (ms_per_step * 1e-3) / 1e9

    # ---- strong scaling (BASELINE config 4 as written: 2^20 hands sharded over the GPUs) beside the weak-scaling line -----
    strong = None
    if world > 1:
        Hs = H // world
        ws_s = lib.mb_mano_workspace_bytes(Hs, mode)

        def strong_step(i):
            s = sets[i % nsets]
            cabi.check(lib.mb_mano_forward(blob, 45, s["rot"].data_ptr(), s["pose"].data_ptr(), s["beta"].data_ptr(), Hs, mode,
                                           s["verts"].data_ptr(), s["joints"].data_ptr(), ws.data_ptr(), ws_s, stream), "fwd")
            cabi.check(lib.mb_mano_backward(blob, 45, s["rot"].data_ptr(), s["pose"].data_ptr(), s["beta"].data_ptr(),
                                            s["gv"].data_ptr(), s["gj"].data_ptr(), Hs, mode, cabi.BWD_WORKSPACE_VALID,
                                            s["g_rot"].data_ptr(), s["g_pose"].data_ptr(), s["g_beta"].data_ptr(),
                                            ws.data_ptr(), ws_s, stream), "bwd")

        for i in range(3):
            strong_step(i)
        sync_all()
        e0.record()
        for i in range(args.steps):
            strong_step(i)
        e1.record()
        sync_all()
        ms_s = max_over_ranks(e0.elapsed_time(e1)) / args.steps
        strong = {"hands_total": Hs * world, "hands_per_gpu": Hs, "ms_per_step": ms_s, "value": Hs * world / (ms_s * 1e-3), "unit": UNIT,
                  "note": "the same forward + backward with the batch of ONE GPU's weak-scaling step sharded over all ranks "
                          "(max over ranks, no collective); at this size a rank's slice no longer exceeds L2 by much"}

    # ---- forward only (north_star: >= 1e8 hands/s forward on 8 GPUs, LBS >= 70 % of the HBM roofline) -------------
    fwd_mode = mode | cabi.FWD_INFERENCE

    def fwd_step(i):
        s = sets[i % nsets]
        cabi.check(lib.mb_mano_forward(blob, 45, s["rot"].data_ptr(), s["pose"].data_ptr(), s["beta"].data_ptr(), H, fwd_mode,
                                       s["verts"].data_ptr(), s["joints"].data_ptr(), ws.data_ptr(), ws_bytes, stream), "fwd")

    for i in range(3):
        fwd_step(i)
    sync_all()
    e0.record()
    for i in range(args.steps):
        fwd_step(i)
    e1.record()
    sync_all()
    ms_fwd = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    lib.mb_profile_enable(1)
    cabi.profile_collect()
    for i in range(args.steps):
        fwd_step(i)
    torch.cuda.synchronize(dev)
    fprof = cabi.profile_collect()
    lib.mb_profile_enable(0)
    fstages = {k: v[0] / v[1] for k, v in fprof.items()}
    forward_only = {"value": world * H / (ms_fwd * 1e-3), "unit": "hands/s", "ms_per_step": ms_fwd, "steps": args.steps,
                    "stages_ms": fstages, "hbm_algorithmic_gbs": BYTES_FWD * H / (ms_fwd * 1e-3) / 1e9,
                    "hbm_frac": BYTES_FWD * H / (ms_fwd * 1e-3) / 1e9 / peaks["hbm_gbs"],
                    "algorithmic_bytes_per_hand": BYTES_FWD,
                    "note": "mb_mano_forward with MB_FWD_INFERENCE (no rest-pose scratch kept): verts[H,778,3] + joints[H,21,3]"}
    if not args.no_extras and (mode & 0xff) != cabi.MODE_FP32:
        # A/B: the same forward through the OTHER implementation (the library's default from 8 192 hands on is the fused blend +
        # skinning kernel with lane = vertex, vskin.cu; MB_FWD_UNFUSED runs the blend-contraction + lane = hand skinning kernels)
        other_is_unfused = "fused_fwd" in fstages
        um = (fwd_mode & ~(cabi.FWD_FUSED | cabi.FWD_UNFUSED)) | (cabi.FWD_UNFUSED if other_is_unfused else cabi.FWD_FUSED)

        def ufwd_step(i):
            s = sets[i % nsets]
            cabi.check(lib.mb_mano_forward(blob, 45, s["rot"].data_ptr(), s["pose"].data_ptr(), s["beta"].data_ptr(), H, um,
                                           s["verts"].data_ptr(), s["joints"].data_ptr(), ws.data_ptr(), ws_bytes, stream), "fwd")

        for i in range(2):
            ufwd_step(i)
        sync_all()
        lib.mb_profile_enable(1)
        cabi.profile_collect()
        e0.record()
        for i in range(5):
            ufwd_step(i)
        e1.record()
        sync_all()
        uprof = cabi.profile_collect()
        lib.mb_profile_enable(0)
        forward_only["unfused_ab" if other_is_unfused else "fused_ab"] = {
            "ms_per_step": e0.elapsed_time(e1) / 5, "stages_ms": {k: v[0] / v[1] for k, v in uprof.items()},
            "note": ("same launch with MB_FWD_UNFUSED: pose -> blend GEMM (writes v_posed_t) -> lane = hand skinning" if other_is_unfused else
                     "same launch with MB_FWD_FUSED: pose -> fused blend + skinning with lane = vertex (vskin.cu), no v_posed_t")}
    if "fused_fwd" in fstages:
        fms = fstages["fused_fwd"]
        lbs_fwd_roof = hbm_roofline("fused_fwd", BYTES_FUSED_FWD, ms=fms)
        lbs_fwd_roof["note"] = ("fused blend + skinning forward (vskin.cu + its bone-operand pre-pass), inference launch: feature tiles 640 + bone transforms 768 B in, "
                                "verts 9336 + fingertip joints 60 B out per hand.  The fused kernel is NOT HBM-bound (that is the point of fusing): its "
                                "bound is the SM's L1 / shared-memory data pipe — tensor-core operand reads + the result stores, together ~88 % busy in the "
                                "ncu capture (ncu_unit_utilisation_pct: l1_data_pipe_tensor_operand + l1_data_pipe_lsu); the HBM-bound figure of the unfused "
                                "skinning kernel is in forward_only.unfused_ab")
        tfl = (FLOP_BLEND + FLOP_SKIN_T) * H / (fms * 1e-3) / 1e12
        blend_roof = {"kernel": "vskin_forward_kernel", "bound": "tensor", "achieved": tfl, "peak": peaks["bf16_tflops_sustained"],
                      "unit": "TFLOP/s", "frac": tfl / peaks["bf16_tflops_sustained"],
                      "algorithmic_flop_per_hand": FLOP_BLEND + FLOP_SKIN_T, "avg_launch_ms": fms, "mode": args.mode,
                      "note": "algorithmic FLOP of the blend contraction (676 860) + the dense weight blend T = W A (298 752); executed: "
                              "x3 fp16 products for the blend, x4 for the transforms, on 896 padded vertex rows"}
    else:
        lbs_fwd_roof = hbm_roofline("lbs_fwd", BYTES_LBS, ms=fstages.get("lbs_fwd"))
        bms = fstages["blend_fwd"]
        tfl = FLOP_BLEND * H / (bms * 1e-3) / 1e12
        blend_roof = {"kernel": "blend_tc_forward_mres_kernel" if (H + 127) // 128 >= 64 else "blend_tc_forward_kernel", "bound": "tensor",
                      "achieved": tfl, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s", "frac": tfl / peaks["bf16_tflops_sustained"],
                      "algorithmic_flop_per_hand": FLOP_BLEND, "avg_launch_ms": bms, "mode": args.mode}

    # ---- parity: a fixed 8 192-hand subsample of the last timed step, checked in fp64 below ----
    parity = None
    cpu = None
    kept = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        step(args.steps - 1)                                   # the forward-only pass overwrote verts / joints: redo the last step
        torch.cuda.synchronize(dev)
        s = sets[(args.steps - 1) % nsets]
        nk = min(H, 8192)
        idx = np.unique(np.linspace(0, H - 1, nk).astype(np.int64))
        tidx = torch.from_numpy(idx).to(dev)
        kept = {k: s[k][tidx].cpu().numpy() for k in ("rot", "pose", "beta", "gv", "gj", "verts", "joints", "g_rot", "g_pose", "g_beta")}

    # ---- the other BASELINE configs, short, in the same line (every rank takes part: config 5 has the collective) ----
    extras = {}
    if not args.no_extras:
        try:
            extras = run_extras(args, pkg, layer, dev, rank, world, dist, sets[0])
        except Exception as exc:                               # never lose the contract line to an extra
            extras = {"error": repr(exc)}

    # ---- end to end through the public nn.Module API with host buffers ------------------
    e2e = None
    if not args.no_e2e:
        s = sets[0]
        # release the device-resident arm's big buffers before the module API allocates its own
        gv_keep, gj_keep = s["gv"], s["gj"]
        del ws
        for st in sets:
            st.clear()
        sets.clear()
        torch.cuda.empty_cache()
        numa = bind_numa_local(local_rank)                    # pinned buffers are first-touched on the GPU's NUMA node
        # ONE pinned buffer each way per step, laid out chunk by chunk as [rot | pose | beta] blocks so that a chunk is one
        # contiguous host range: one H2D and one D2H copy per chunk (round 1 issued three each, from three arrays)
        n_chunks = int(os.environ.get("MANO_B200_E2E_CHUNKS", "4")) if H >= 8 * 4096 else 1
        Hc = (H + n_chunks - 1) // n_chunks
        rot_h, pose_h, beta_h = synth_inputs(H, 555 + rank)
        h_in = torch.empty(H * 58, dtype=torch.float32).pin_memory()
        bounds = []
        off = 0
        for c in range(n_chunks):
            a, b = c * Hc, min(H, (c + 1) * Hc)
            n = b - a
            blk = h_in[off:off + n * 58]
            blk[:n * 3].copy_(torch.from_numpy(rot_h[a:b]).reshape(-1))
            blk[n * 3:n * 48].copy_(torch.from_numpy(pose_h[a:b]).reshape(-1))
            blk[n * 48:].copy_(torch.from_numpy(beta_h[a:b]).reshape(-1))
            bounds.append((a, b, off))
            off += n * 58
        h_out2 = [torch.empty(H * 58, dtype=torch.float32).pin_memory() for _ in range(2)]
        h_loss2 = [torch.empty(1, dtype=torch.float32).pin_memory() for _ in range(2)]
        # Three streams: host -> device copies, kernels, device -> host copies.  Every kernel of the step runs on ONE compute
        # stream, back to back as in the device-resident loop; the copies of the neighbouring chunks hide behind them (events
        # order a chunk's copy-in -> compute -> copy-out; device staging buffers alternate per step).  [round 2: two
        # symmetric streams, each doing copy + kernels + copy, let kernels of different chunks compete for the SMs — the
        # one-CTA-per-SM kernels serialise anyway — and measured 20.5 - 22.7 ms per step depending on how they interleaved]
        st_in, st_cmp, st_out = (torch.cuda.Stream(device=dev) for _ in range(3))
        side = [st_in, st_cmp, st_out]
        d_in2 = [[torch.empty((b - a) * 58, dtype=torch.float32, device=dev) for (a, b, _) in bounds] for _ in range(2)]
        free_ev = [[None] * n_chunks for _ in range(2)]
        step_no = [0]

        def e2e_step(copy_grads=True):
            par = step_no[0] & 1
            h_out, h_loss = h_out2[par], h_loss2[par]
            step_no[0] += 1
            for c, (a, b, o) in enumerate(bounds):
                n = b - a
                flat = d_in2[par][c]
                with torch.cuda.stream(st_in):
                    if free_ev[par][c] is not None:
                        st_in.wait_event(free_ev[par][c])                    # the step before last is done with this buffer
                    flat.copy_(h_in[o:o + n * 58], non_blocking=True)
                    ev_in = st_in.record_event()
                with torch.cuda.stream(st_cmp):
                    st_cmp.wait_event(ev_in)
                    d = [flat[:n * 3].view(n, 3).requires_grad_(), flat[n * 3:n * 48].view(n, 45).requires_grad_(),
                         flat[n * 48:].view(n, 10).requires_grad_()]
                    verts, joints = layer(*d)
                    torch.autograd.backward([verts, joints], [gv_keep[a:b], gj_keep[a:b]])
                    g = torch.cat([x.grad.reshape(-1) for x in d])           # 232 B per hand of gradients, one D2H copy
                    loss = joints[0, 0, :1] if c == 0 else None
                    ev_done = st_cmp.record_event()
                    free_ev[par][c] = ev_done
                with torch.cuda.stream(st_out):
                    st_out.wait_event(ev_done)
                    if copy_grads:
                        h_out[o:o + n * 58].copy_(g, non_blocking=True)
                    if loss is not None:
                        h_loss.copy_(loss, non_blocking=True)
                    g.record_stream(st_out)
                    joints.record_stream(st_out)
                for t in (gv_keep, gj_keep):
                    t.record_stream(st_cmp)
                del verts, joints, d, g, flat, loss

        def join_side():
            main = torch.cuda.current_stream(dev)
            for st in side:
                main.wait_stream(st)

        n_e2e = max(3, args.steps)                           # the same K steps as the device-resident loop
        sync_all()                                            # the side streams start after everything above
        for _ in range(2):
            e2e_step()
        join_side()
        sync_all()
        e0.record()
        for _ in range(n_e2e):
            e2e_step()
        join_side()                                           # the last step's copies are inside the timed region
        e1.record()
        sync_all()
        ms_e2e = max_over_ranks(e0.elapsed_time(e1)) / n_e2e

        # the contract's minimum for the device -> host leg is "the step's result (loss or metric)": the same steps with only
        # the 4-byte loss read back (the parameter gradients stay on the device, as they do inside the reference's heads)
        for _ in range(2):
            e2e_step(False)
        join_side()
        sync_all()
        e0.record()
        for _ in range(n_e2e):
            e2e_step(False)
        join_side()
        e1.record()
        sync_all()
        ms_e2e_loss = max_over_ranks(e0.elapsed_time(e1)) / n_e2e

        # the host side alone: the same pinned ranges over the same streams, no kernels — what the box's PCIe / host memory
        # path allows with `world` ranks copying at once (the ceiling of the e2e arm when it is below `value`)
        d_flat = torch.empty(H * 58, dtype=torch.float32, device=dev)

        def copy_step():
            for c, (a, b, o) in enumerate(bounds):
                n = b - a
                with torch.cuda.stream(st_in):
                    d_flat[o:o + n * 58].copy_(h_in[o:o + n * 58], non_blocking=True)
                    ev = st_in.record_event()
                with torch.cuda.stream(st_out):
                    st_out.wait_event(ev)
                    h_out2[0][o:o + n * 58].copy_(d_flat[o:o + n * 58], non_blocking=True)

        copy_step()
        join_side()
        sync_all()
        e0.record()
        for _ in range(5):
            copy_step()
        join_side()
        e1.record()
        sync_all()
        ms_copy = max_over_ranks(e0.elapsed_time(e1)) / 5
        numa_all = [numa]
        if world > 1:
            numa_all = [None] * world
            dist.all_gather_object(numa_all, numa)
        e2e = {"value": world * H / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": H * 58 * 4,
               "d2h_bytes_per_step": H * 58 * 4 + 4, "ms_per_step": ms_e2e, "steps": n_e2e,
               "loss_only_readback": {"value": world * H / (ms_e2e_loss * 1e-3), "ms_per_step": ms_e2e_loss, "d2h_bytes_per_step": 4,
                                      "note": "same steps, device -> host leg reduced to the loss scalar (the contract's minimum)"},
               "copy_only": {"ms_per_step": ms_copy, "gb_per_s_per_gpu_each_way": H * 58 * 4 / (ms_copy * 1e-3) / 1e9,
                             "note": "the same H2D + D2H copies with no kernels, all ranks at once (max over ranks): the host-side "
                                     "ceiling of this arm"},
               "numa": numa_all,
               "api": "ManoLayer.forward + autograd backward; ONE pinned host buffer in (rot | pose | beta per chunk) and one out "
                      f"(58 gradient floats per hand); {n_chunks} chunks, one H2D + one D2H copy per chunk on a copy-in and a copy-out "
                      "stream around ONE compute stream (events per chunk; device staging and pinned result buffers alternate per "
                      "step, so consecutive steps overlap); verts / joints and their upstream gradients "
                      "stay on the device (the step's result that crosses PCIe is the parameter gradient)"}

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        parity = parity_check(model, kept)
        cpu = cpu_baseline(pkg.assets, model)

    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32" if args.mode == "fp32" else ("f16x3->f32" if args.mode == "f16x3" else "f16->f32"),
        "data": "synthetic",
        "config": {"workload": workload_name(args), "hands_per_gpu": H, "pose_num": 45, "pca": "identity (no PCA)",
                   "blend_mode": args.mode, "parallelism": f"batch-sharded x{world}, no collective",
                   "l2": "inputs larger than L2 (>= 20 GB touched per step)" if H * 20000 > 4 * 126e6 else
                         f"rotating {nsets} buffer sets", "model": "seeded synthetic MANO-shaped model"},
        "clocks": clocks,
        "e2e": e2e,
        "gpu_launches": int(launches),
        "roofline": roofline,
        "roofline_lbs_forward": lbs_fwd_roof,
        "blend_gemm": blend_roof,
        "step_hbm": {"algorithmic_gbs": step_gbs, "frac_of_peak": step_gbs / peaks["hbm_gbs"],
                     "algorithmic_bytes_per_hand": BYTES_FWD + BYTES_BWD},
        "stages_ms": stages,
        "forward_only": forward_only,
        "strong_scaling": strong,
        "parity": parity,
        "cpu_baseline": cpu,
    }
    line.update(extras)
    emit(line)


if __name__ == "__main__":
    main()
