"""Where a CTA of relation_fused_kernel spends its time.  Builds a PRIVATE copy of the library with -DRN_FUSED_TRACE
(build_trace/librelnet_b200_trace.so; the product library has no trace code), runs the module at N d H and prints, per
phase, the median / max over CTAs of the clock64() deltas of thread 0, plus the spread of CTA start / end times (ns).
    python tools/fused_trace.py 300 1024 16 [flush]"""
import ctypes
import os
import subprocess
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = os.path.join(ROOT, 'relation-networks-for-object-detection_b200')
env = dict(os.environ, RELNET_VARIANT='trace')
lib_path = subprocess.check_output([sys.executable, os.path.join(PKG, 'build.py')], env=env, text=True).strip().splitlines()[-1]
import numpy as np
import torch
import relnet_b200
from relnet_b200 import _lib, ops, synth
_lib.LIB_PATH = lib_path
lib = _lib.lib()
lib.rn_fused_trace_set.argtypes = [ctypes.c_void_p]

N, d, H = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda') if len(sys.argv) > 4 else None
c = synth.make_relation_case(N * 31 + d + H, N, d, H)
t = [torch.from_numpy(c[k]).cuda() for k in ('X', 'boxes', 'Wq', 'bq', 'Wk', 'bk', 'Wg', 'bg', 'Wout', 'bout')]
kw = dict(group=H, residual_relu=True, precision='f16')
for _ in range(3):
    ops.relation(*t, **kw)
buf = torch.zeros(148 * 16, dtype=torch.int64, device='cuda')
names = ['entry->prologue', 'phi block0', 'FC drain+readback', 'publish', 'wait team', 'S visible', 'g landed', 'softmax+P',
         "P.V'", 'l exchange+stage', 'partials+ticket', 'merge/final', 'exit sync']
runs = []
for it in range(5):
    if flush is not None:
        flush.fill_(it & 1)
    buf.zero_()
    torch.cuda.synchronize()
    assert lib.rn_fused_trace_set(ctypes.c_void_p(buf.data_ptr())) == 0
    ops.relation(*t, stage_mask=6, **kw)
    torch.cuda.synchronize()
    lib.rn_fused_trace_set(ctypes.c_void_p(0))
    a = buf.cpu().numpy().reshape(148, 16)
    a = a[a[:, 15] != 0]
    runs.append(a)
a = runs[-1]
print('N=%d d=%d H=%d  CTAs traced: %d  (last of %d runs%s)' % (N, d, H, a.shape[0], len(runs), ', L2 flushed' if flush is not None else ''))
print('CTA start spread %.2f us, end spread %.2f us, first start -> last end %.2f us' % (
    (a[:, 0].max() - a[:, 0].min()) / 1e3, (a[:, 15].max() - a[:, 15].min()) / 1e3, (a[:, 15].max() - a[:, 0].min()) / 1e3))
idx = [(1, 2), (2, 3), (3, 4), (4, 5), (5, 6), (6, 7), (7, 8), (8, 9), (9, 10), (10, 11), (11, 12), (12, 13), (13, 14)]
tot = a[:, 14] - a[:, 1]
print('%-20s %9s %9s' % ('phase (cycles)', 'median', 'max'))
for nm, (i0, i1) in zip(names, idx):
    dlt = a[:, i1] - a[:, i0]
    ok = (a[:, i1] != 0) & (a[:, i0] != 0)
    if ok.any():
        print('%-20s %9d %9d' % (nm, np.median(dlt[ok]), dlt[ok].max()))
print('%-20s %9d %9d' % ('total in-kernel', np.median(tot), tot.max()))
