"""How many host threads make the CPU arm (torch-CPU trunk + numpy oracle) fastest on this box."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import relnet_b200
from relnet_b200.trunk import make_trunk
from bench import make_inputs
image, _ = make_inputs()
print('cpu_count', os.cpu_count(), 'affinity', len(os.sched_getaffinity(0)))
for nt in (8, 16, 32, 64, 128):
    if nt > (os.cpu_count() or 1):
        break
    torch.set_num_threads(nt)
    trunk = make_trunk('cpu', torch.float32)
    trunk(image)
    t0 = time.perf_counter(); trunk(image); trunk(image); dt = (time.perf_counter() - t0) / 2
    print('threads', nt, 'trunk s/image', round(dt, 3), flush=True)
