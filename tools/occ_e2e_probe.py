import sys, time, numpy as np, torch
sys.path.insert(0, "/root/repo")
from coma_b200 import synth
from utils.coma_occupancy import ComA_Occupancy
import coma_b200.coma_occupancy as co
Hr=1310
occ = ComA_Occupancy(scale_tolerance=3.0, human_res=Hr, obj_res=4, normal_res=0, spatial_res=128, device="cuda:0", human_slice=(0, Hr))
host=[]; obj=None
for c0 in range(0, 1024, 256):
    ss = synth.make_samples(256, Hr, 4, seed=900 + c0 // 256)
    obj = (ss[0]["obj_verts"], ss[0]["obj_normals"]) if obj is None else obj
    for s in ss:
        s["obj_verts"], s["obj_normals"] = obj
        host.append(s)
print(type(host[0]["human_verts"]), host[0]["human_verts"].dtype, host[0]["human_verts"].flags.c_contiguous, host[0]["human_verts"].shape)
for rep in range(3):
    occ.spatial_occupancy_grids.zero_(); occ.debug_obj_vert = occ.debug_obj_normal = None
    torch.cuda.synchronize()
    t0=time.perf_counter()
    for s in host: occ.register_sample_to_cache(**s)
    t1=time.perf_counter()
    orig=occ._stage_chunk
    acc=[0.0]
    def timed(*a, **k):
        t=time.perf_counter(); r=orig(*a, **k); acc[0]+=time.perf_counter()-t; print("  stage_chunk ->", r); return r
    occ._stage_chunk=timed
    occ.aggregate_all_samples()
    torch.cuda.synchronize()
    t2=time.perf_counter()
    f=occ.return_aggregated_spatial_grids().cpu().numpy()
    t3=time.perf_counter()
    del occ._stage_chunk
    print(f"rep {rep}: register {t1-t0:.4f}  aggregate {t2-t1:.4f} (stage {acc[0]:.4f})  readout {t3-t2:.4f}")
