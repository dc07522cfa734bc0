import sys, torch
sys.path.insert(0,'gi-gs_b200'); sys.path.insert(0,'oracle'); sys.path.insert(0,'tests')
import diff_gaussian_rasterization as dgr, gigs_oracle as O, gpu_util as U
from gigs import scene
DEV='cuda:0'
P,W,H=4000,128,96
raw=scene.make_scene(P,seed=13); g=scene.activate(raw,DEV); cam=scene.orbit_camera(3,8,W,H).to(DEV)
fo=U.ours_forward(g,cam,torch.zeros(3,device=DEV))
fx,fy=W/(2*cam.tanfovx),H/(2*cam.tanfovy)
n,p=dgr.geometry_chain(W,H,fx,fy,cam.world_view_transform,fo['depth'],True)
occ=dgr._C.SSAO(W,H,fx,fy,0.8,0.01,0.05,0.0625,16,8,fo['normal_view'],p).cpu()
oo=O.ssao(W,H,fx,fy,0.8,0.01,0.05,0.0625,16,8,fo['normal_view'].cpu(),p.cpu())
d=(occ-oo).abs()[0]
print('ndiff>1e-4', int((d>1e-4).sum()), 'max', float(d.max()))
idx=torch.nonzero(d>1e-4)[:8]
nv=fo['normal_view'].cpu(); pc=p.cpu()
for y,x in idx.tolist():
    print(y,x,'occ',float(occ[0,y,x]),'ora',float(oo[0,y,x]),'n',nv[:,y,x].tolist(),'pos',pc[:,y,x].tolist())
