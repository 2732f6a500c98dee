"""Where a CTA of relation_attn_tile_kernel spends its time: per-CTA clock64() stamps (rn_debug_tile_trace) at N = M = 300,
d = 1024, H = 16 (144 CTAs).  Prints the median cycle count between consecutive points, warm L2 and after an L2 flush."""
import ctypes
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as entry
entry.build()
import relnet_b200
from relnet_b200 import ops, _lib
from oracle import relation_np as R

dev = torch.device('cuda:0')
c = R.make_relation_case(1, 300, 1024, 16)
t = [torch.from_numpy(c[k]).to(dev) for k in ('X', 'boxes', 'Wq', 'bq', 'Wk', 'bk', 'Wg', 'bg', 'Wout', 'bout')]
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
for _ in range(3):
    ops.relation(*t, group=16, residual_relu=True, precision='f16')
buf = torch.zeros(144 * 8, dtype=torch.int64, device=dev)
names = ['start->prologue done', 'prologue->Q/K/V landed (MMA thread)', 'prologue->geometry loads issued (thread 0)',
         'geometry issued->S visible', 'S visible->P written', 'P written->O visible', 'O visible->CTA end']
out = {}
for label, do_flush in (('warm', False), ('l2_flushed', True)):
    if do_flush:
        flush.fill_(1)
    _lib.lib().rn_debug_tile_trace(ctypes.c_void_p(buf.data_ptr()))
    ops.relation(*t, group=16, residual_relu=True, precision='f16', stage_mask=4)
    torch.cuda.synchronize()
    _lib.lib().rn_debug_tile_trace(ctypes.c_void_p(0))
    s = buf.view(144, 8).cpu().double()
    d = [s[:, 1] - s[:, 0], s[:, 2] - s[:, 1], s[:, 3] - s[:, 1], s[:, 4] - s[:, 3], s[:, 5] - s[:, 4], s[:, 6] - s[:, 5],
         s[:, 7] - s[:, 6]]
    out[label] = {n: int(x.median()) for n, x in zip(names, d)}
    out[label]['total start->end'] = int((s[:, 7] - s[:, 0]).median())
print(json.dumps(out, indent=1))
