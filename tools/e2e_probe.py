import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vulkanexp_b200 import scene_format, synth
from vulkanexp_b200._lib import Context
from vulkanexp_b200.host_logic import OrientationGenerator
from vulkanexp_b200.pods import GridInfo, Light
flat = scene_format.flatten(synth.make_cfg2())
grid = GridInfo.make(flat["bounds_min"], flat["bounds_max"], (32,16,32), 256, hysteresis=0.9)
ctx = Context(0); ctx.scene_upload(flat); ctx.bvh_build(); ctx.probes_init(grid); ctx.probes_upload(state=np.ones(grid.probe_count, dtype=np.uint32))
gen = OrientationGenerator(); Rs=[gen.next() for _ in range(256)]
light = Light.default()
(ih,iw),(dh,dw)=grid.atlas_shapes()
outs=[]
for _ in range(2):
    pin=(torch.empty((ih,iw),dtype=torch.int32).pin_memory(), torch.empty((dh,dw),dtype=torch.int32).pin_memory(), torch.empty(grid.probe_count,dtype=torch.int32).pin_memory())
    outs.append((pin, tuple(t.numpy().view(np.uint32) for t in pin)))
idx=np.arange(grid.probe_count,dtype=np.uint32)
def run(mode, K=200):
    for i in range(5): ctx.probes_update(grid, light, Rs[i], None, sync=False)
    ctx.sync(); t0=time.perf_counter(); th=0
    for s in range(K):
        h0=time.perf_counter()
        ctx.probes_update(grid, light, Rs[5+s], idx if 'list' in mode else None, sync=False)
        if 'dl' in mode:
            o = outs[s&1][1]
            if 'st' in mode: o = (None, None, o[2])
            elif 'irr' in mode: o = (o[0], None, None)
            ctx.probes_download_async(o)
        th+=time.perf_counter()-h0
    if 'dl' in mode: ctx.probes_download_wait()
    ctx.sync(); ms=(time.perf_counter()-t0)*1e3/K
    print(mode, 'ms/step %.3f'%ms, 'host enqueue ms/step %.3f'%(th*1e3/K), 'device', ctx.probes_timings()['full'])
for m in ['null', 'null+dl', 'null+dl+st', 'null+dl+irr', 'null', 'null+dl']: run(m)
