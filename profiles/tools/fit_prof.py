import importlib, sys, torch, numpy as np
sys.path.insert(0,'.')
import bench
pkg = importlib.import_module("3dhandposeestimation_b200")
dev = torch.device("cuda",0)
model = bench.no_pca_model(pkg.assets)
layer = pkg.ManoLayer(dev, model=model, pose_num=45)
H = 1<<20
rot,pose,beta = bench.synth_inputs(H, 777)
with torch.no_grad():
    _, tgt = layer.rot_pose_beta_to_mesh(*[torch.from_numpy(a).to(dev) for a in (rot,pose,beta)], joints_only=True)
vis = (torch.rand(H,21,1,device=dev) < 0.8).float()
f = pkg.fitting.ManoFitter(layer, H)
for _ in range(3): f.step(tgt, vis)
torch.cuda.synchronize()
lib = pkg.load_library()
lib.mb_profile_enable(1)
e0,e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): f.step(tgt, vis)
e1.record(); torch.cuda.synchronize()
print("ms/iter", e0.elapsed_time(e1)/10)
print({k:(round(v[0]/v[1],3),v[1]) for k,v in pkg._cabi.profile_collect().items()})
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3): f.step(tgt, vis)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=60))
