#!/bin/bash
mkdir -p gpurun_out
cat > /tmp/pt_one.py <<'PY'
import sys, os, ctypes as C, torch
ROOT = os.environ.get("GRAFT_REPO_ROOT", "/root/repo")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'lightmetrica-v2_b200'))
from lmb200py import capi, scenedesc
sc = scenedesc.config2_scene(1000000, 1920, 1080)
S = capi.Scene(sc); W, H, spp = 1920, 1080, 32; N = W * H * spp
film = torch.zeros((H, W, 4), dtype=torch.float32, device='cuda'); L = capi.lib(); st = capi.RenderStats()
p = S.params(capi.MODE_PTDIRECT, N, seed=1)
capi.check(L.lmb200_render_dev(S.h_, C.byref(p), film.data_ptr(), torch.cuda.current_stream().cuda_stream, C.byref(st)))
print("ptdirect", N / st.seconds / 1e6, "Msamples/s", st.extend_rays, st.shadow_rays, st.iterations)
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/pt_launches.csv python /tmp/pt_one.py > gpurun_out/pt_under_ncu.log 2>&1
tail -2 gpurun_out/pt_under_ncu.log
