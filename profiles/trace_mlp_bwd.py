import ctypes as C, os, sys, json
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import bench
from remixfusion_b200 import abi, configs, synth
from remixfusion_b200.scene_rep import JointEncoding
from remixfusion_b200.global_volume import MapVolume
lib = abi.lib()
dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
hid = int(os.environ.get("RF_PROF_HIDDEN", 32)); hs = int(os.environ.get("RF_PROF_HASH", 16))
cfg = configs.replica(hidden=hid, hash_size=hs)
cam = cfg["cam"]; H, W = cam["H"], cam["W"]
K, poses, frames = bench.make_frames(cfg, 1, first=0, stride=50)
c2w, depth, rgb = frames[0]
bb = torch.from_numpy(np.array(cfg["mapping"]["bound"])).to(torch.float64)
torch.manual_seed(0)
model = JointEncoding(cfg, bb).to(dev)
with torch.no_grad():
    model.embed_res_fn.params.copy_((torch.rand_like(model.embed_res_fn.params) * 2 - 1) * 1e-2)
model.train()
params = [model.embed_res_fn.params] + list(model.decoder_res.fused_weights())
mvol = MapVolume(cfg, model, K); mvol.init_mapvolume()
d = torch.from_numpy(depth).to(dev); c = torch.from_numpy(rgb).to(dev)
mvol.integrate_kf({"rgb": c, "depth": d}, torch.from_numpy(c2w).float(), 1.0)
dirs = torch.from_numpy(synth.camera_dirs(K, H, W).reshape(-1, 3)).to(dev)
c2w_t = torch.from_numpy(c2w.astype(np.float32)).to(dev)
rays_d = torch.sum(dirs[..., None, :] * c2w_t[:3, :3], -1).contiguous()
rays_o = c2w_t[None, :3, -1].repeat(H * W, 1).contiguous()
tgt_d = d.reshape(-1, 1).contiguous(); tgt_c = c.reshape(-1, 3).contiguous()
def step():
    for p in params: p.grad = None
    configs.total_loss(cfg, model.mapping(rays_o, rays_d, tgt_c, tgt_d)).backward()
    torch.cuda.synchronize()
step(); step()
buf = (C.c_ulonglong * 32)()
lib.rf_debug_mlp_trace(buf, 1)
step()
lib.rf_debug_mlp_trace(buf, 1)
v = list(buf)
tiles = v[31]
names = {0:"vote+prefetch",1:"stage X",2:"barrier0",3:"tma wait",4:"issue P1",5:"wait P1",6:"epi1",7:"bar1",8:"issue P2",9:"wait P2",10:"epi2",11:"bar2",12:"issue P3",13:"wait P3",14:"epi3",15:"bar3",16:"issue P4",17:"wait P4",18:"epi4",19:"bar4",20:"issue P5",21:"wait P5",22:"epi5",24:"skipped iter",25:"final flush"}
tot = sum(v[:26])
print("tiles", tiles, "cycles/tile (group)", tot / max(tiles,1))
for i in range(26):
    if i in names: print(f"{names[i]:16s} {v[i]/max(tiles,1):9.1f}  {100*v[i]/tot:5.1f}%")
