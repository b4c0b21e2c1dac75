import sys, os, ctypes as C, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'lightmetrica-v2_b200'))
from lmb200py import capi, scenes
import torch
verts = scenes.soup(20000, seed=42, extent=10.0, edge=0.2)
lo, hi = scenes.bounds(verts); rays = scenes.random_rays(64, lo, hi, seed=7)
A = capi.Accel(0); A.build(verts)
d_rays = torch.from_numpy(rays).cuda(); out = torch.zeros((64, 4), dtype=torch.int32, device='cuda')
L = capi.lib(); L.lmb200_debug_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
print(L.lmb200_debug_trace(A.h, d_rays.data_ptr(), 64, out.data_ptr()))
o = out.cpu().numpy().view(np.uint32)
np.save(os.path.join(ROOT, 'gpurun_out', 'debug.npy'), o)
for i in range(16): print(i, hex(o[i,0]), o[i,1], o[i,2], o[i,3])
