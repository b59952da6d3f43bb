import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tray_b200 import ray, rand
from oracle import oracle as O
w,h,spp,d=1920,1080,64,50
ctx=ray.Context([0]); scene=ray.RichScene(rand.New(2)); flat=scene.flatten(); ctx.upload(flat)
tr=ray.New(w,h); tr.Camera=ray.RichSceneCamera(); tr.MaxDepth,tr.NumRaysPerPixel,tr.Seed=d,spp,2; tr.Context=ctx; tr._prepare(scene)
cam_c=tr.to_c(); params=tr._params(0,h)
host=np.zeros((h,w,4),dtype=np.uint8)
ctx.render(cam_c,params,None); ctx.upload(flat); st=ctx.render(cam_c,params,host)
print("host any", host.any(), host[540,960], st["paths"])
img,_=O.render_sampled_rows(O.rich_scene(2),O.camera_init(w,h,**O.RICH_CAMERA),O.make_params(w,h,spp=spp,max_depth=d,seed=2,stream_mode=1,fma_mode=0),135,0,16)
rows=list(range(0,h,135))
dd=np.abs(img[rows,:,:3].astype(int)-host[rows,:,:3].astype(int)).max(axis=2)
print("identical", (dd==0).mean(), "max", dd.max(), img[rows[4],960], host[rows[4],960])
