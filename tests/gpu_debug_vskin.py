"""Diagnostics (not a test): run the fused lane = vertex forward through mb_mano_forward_debug with both descriptor
variants and several split-product counts, print intermediate / output errors against the fp64 oracle."""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch

    import __graft_entry__ as ge
    from oracle import mano_oracle as mo
    from test_gpu_vskin import mano_inputs, run_debug

    ge.build()
    pkg = importlib.import_module("3dhandposeestimation_b200")
    dev = torch.device("cuda", 0)
    model = pkg.assets.synthetic_mano()
    B, nc = int(os.environ.get("VS_B", "100")), 45
    rot, pose, beta = mano_inputs(B, nc, seed=B)
    layer = pkg.ManoLayer(dev, model=model, pose_num=nc)
    ov, oj, cache = mo.mano_forward(model, rot, pose, beta, return_cache=True)
    Rp = np.einsum("bij,bkjl->bkil", cache["Rq"], cache["Rg"])
    tp = np.einsum("bij,bkj->bki", cache["Rq"], cache["tA"])
    A = np.concatenate([Rp, tp[..., None]], axis=-1)
    W = np.asarray(model["weights"], np.float64)
    n = min(B, 4)
    T_ref = np.einsum("vk,bkij->bvij", W[:128], A[:n]).reshape(n, 128, 12)
    for variant in (0, 3 << 4, 6 << 4):
        try:
            verts, joints, dbg = run_debug(pkg, layer, dev, rot, pose, beta, variant)
        except Exception as exc:                                     # a trap poisons the context: stop
            print(f"variant {variant}: FAILED {exc}")
            break
        eT = np.abs(dbg[:n, :, :12] - T_ref)
        evp = np.abs(dbg[:n, :, 12:15] - cache["v_posed"][:n, :128])
        ev = np.abs(verts - ov)
        print(f"variant {variant:#x}: err_T {eT.max():.3e} err_vp {evp.max():.3e} err_verts {ev.max():.3e} "
              f"err_joints {np.abs(joints - oj).max():.3e}  (verts err per vertex tile: "
              + " ".join(f"{ev[:, t * 128:(t + 1) * 128].max():.1e}" for t in range(7)) + ")")
        if eT.max() > 1e-5:
            print("  T[h0,v0] got ", np.array2string(dbg[0, 0, :12], precision=5))
            print("  T[h0,v0] want", np.array2string(T_ref[0, 0], precision=5))
            print("  T[h1,v5] got ", np.array2string(dbg[1, 5, :12], precision=5))
            print("  T[h1,v5] want", np.array2string(T_ref[1, 5], precision=5))
        if ev.max() > 1e-6:
            bad = np.argwhere(ev.max(axis=2) > 1e-6)
            print("  first bad (hand, vertex):", bad[:8].tolist(), " count", len(bad), "of", ev.shape[0] * 778)
    # error of the default path (separate kernels) and of the fused one over whole batches, per pose-kernel regime
    for Bt in (1000, 4096, 9001, 40001):
        r2, p2, b2 = mano_inputs(Bt, 45, seed=Bt + 45)
        want_v, want_j = [], []
        for s0 in range(0, Bt, 2048):
            a, b = mo.mano_forward(model, r2[s0:s0 + 2048], p2[s0:s0 + 2048], b2[s0:s0 + 2048])
            want_v.append(a); want_j.append(b)
        want_v, want_j = np.concatenate(want_v), np.concatenate(want_j)
        for name, lay in (("default", pkg.ManoLayer(dev, model=model, pose_num=45)),
                          ("fused", pkg.ManoLayer(dev, model=model, pose_num=45, fused_forward=True))):
            with torch.no_grad():
                v, j = lay(*[torch.from_numpy(x).to(dev) for x in (r2, p2, b2)])
            ev = np.abs(v.cpu().numpy() - want_v).max(axis=(1, 2))
            ej = np.abs(j.cpu().numpy() - want_j).max(axis=(1, 2))
            print(f"B={Bt:6d} {name:8s} verts max {ev.max():.3e} p99.9 {np.quantile(ev, .999):.3e} p50 {np.median(ev):.3e} | joints max {ej.max():.3e}")
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
