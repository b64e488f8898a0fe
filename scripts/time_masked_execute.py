import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/trapped-modes-ltg_b200")
import numpy as np, torch
from fcd_b200 import HeightMapPlan
from bench import make_frames_gpu, SEED
from fcd_b200 import synthetic as o
n=2048; dev=torch.device("cuda",0)
for F in (32, 64):
    plan=HeightMapPlan((n,n),F,dev)
    ref, frames = make_frames_gpu(n, F, SEED, dev)
    plan.bind(ref, square_size=o.board_square_size(n), height=1.0)
    m = torch.zeros((F,n,n),dtype=torch.bool,device=dev); m[:, 500:900, 600:1000]=True
    m8 = m.to(torch.uint8)
    for name, kw in (("nomask", {}), ("mask_bool", {"mask": m}), ("mask_u8", {"mask": m8}), ("mask_single", {"mask": m8[0]})):
        plan.execute(frames, **kw); torch.cuda.synchronize()
        plan.set_profiling(True)
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record(); plan.execute(frames, **kw); e1.record(); torch.cuda.synchronize()
        st=plan.stage_times(); plan.set_profiling(False)
        print(F, name, f"{e0.elapsed_time(e1)*1e3/F:.1f} us/frame", {k: round(v[0]*1e3/max(v[2],1),1) for k,v in st.items()})
    plan.close()
