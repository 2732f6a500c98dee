"""Kernel timeline of ONE CUDA-graph replay of the full step (trunk + hot path, two streams): name, stream, start (us from
the first kernel), duration -- from torch.profiler / CUPTI.  Profiler timings are not bench numbers; this is for seeing
which branch is the critical path.  Writes gpurun_out/timeline.json."""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as entry
entry.build()
import relnet_b200
from relnet_b200.pipeline import RelationHead, init_head_params, GraphedStep, Detector
from relnet_b200.trunk import make_trunk
import bench

dev = torch.device('cuda:0')
torch.backends.cudnn.benchmark = True
trunk = make_trunk(dev, torch.bfloat16)
head = RelationHead(init_head_params(0, dev), precision='f16')
image_h, im_info_h = bench.make_inputs(seed=0)
img = image_h.to(dev); im_info = im_info_h.to(dev)
step = GraphedStep(Detector(trunk, head, im_info), [img])
for _ in range(5):
    step(img)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        step(img)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.end > e.time_range.start]
ev.sort(key=lambda e: e.time_range.start)
# three identical replays were profiled: the last third of the kernel events is one replay
last = ev[len(ev) - len(ev) // 3:]
t0 = last[0].time_range.start
rows = [dict(name=e.name[:60], stream=int(getattr(e, 'device_resource_id', -1) if hasattr(e, 'device_resource_id') else -1), start_us=round(e.time_range.start - t0, 1),
             dur_us=round(e.time_range.end - e.time_range.start, 1)) for e in last]
os.makedirs('gpurun_out', exist_ok=True)
json.dump(rows, open('gpurun_out/timeline.json', 'w'))
print(len(rows), 'kernels, span', rows[-1]['start_us'] + rows[-1]['dur_us'], 'us')
