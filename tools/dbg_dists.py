import sys; sys.path.insert(0,'.')
import torch, ctypes
sys.path.insert(0,'tests')
from test_gpu_parity import _inputs,_posed_verts,_gpu_view,_gpu_raster
from dsf_b200.mano_layer import MANO_SMPL
from dsf_b200 import make_synthetic_mano
from oracle import mano_oracle as mo, raster_oracle as ro
m=make_synthetic_mano(0); layer=MANO_SMPL(m,'nyu'); c32=mo.ManoConstants(m)
inp=_inputs(2,seed=5); vw,_=_posed_verts(layer,inp)
view,xs,ys,_=_gpu_view(0,inp['center3d'],inp['cube'],128)
img,p2f,zbuf,bary,dists=_gpu_raster(layer,vw,view,xs,ys,fragments=True)
Lb=ro.lib(); b=0
vn=torch.empty(779,3); v=view[b].cpu()
Lb.orc_project_f32(ro._p(vw[b].cpu().contiguous()),779,ctypes.c_float(v[0]),ctypes.c_float(v[1]),ctypes.c_float(v[2]),ctypes.c_float(v[3]),ro._p(vn))
pf=torch.empty(128,128,dtype=torch.int32); zb=torch.empty(128,128); ds=torch.empty(128,128)
faces=c32.faces.int().contiguous()
Lb.orc_rasterize_f32(ro._p(vn),ro._p(faces,ctypes.c_int),faces.shape[0],ro._p(xs[b].cpu()),128,ro._p(ys[b].cpu()),128,1,ctypes.c_float(1e-8),0,ro._p(pf,ctypes.c_int),ro._p(zb),None,ro._p(ds))
d=dists[b].cpu()
ne=(d!=ds)
print('neq',ne.sum().item(),'of fg',(pf>=0).sum().item(),'maxabs',(d-ds).abs().max().item(), 'rel', ((d-ds).abs()/ds.abs().clamp(min=1e-30))[ne].max().item())
i=ne.nonzero()[:5]
for y,x in i.tolist(): print(y,x,d[y,x].item(),ds[y,x].item(),pf[y,x].item())
