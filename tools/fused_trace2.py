"""Per-stage stamps of the first phi round (RN_FUSED_TRACE2 variant): python tools/fused_trace2.py 300 1024 16"""
import ctypes, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = os.path.join(ROOT, 'relation-networks-for-object-detection_b200')
env = dict(os.environ, RELNET_VARIANT='trace2', RELNET_DEFINES='RN_FUSED_TRACE2')
lib_path = subprocess.check_output([sys.executable, os.path.join(PKG, 'build.py')], env=env, text=True).strip().splitlines()[-1]
os.environ['RELNET_LIB'] = lib_path
import numpy as np
import torch
import relnet_b200
from relnet_b200 import _lib, ops, synth
lib = _lib.lib()
lib.rn_fused_trace2_set.argtypes = [ctypes.c_void_p]
N, d, H = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
c = synth.make_relation_case(N * 31 + d + H, N, d, H)
t = [torch.from_numpy(c[k]).cuda() for k in ('X', 'boxes', 'Wq', 'bq', 'Wk', 'bk', 'Wg', 'bg', 'Wout', 'bout')]
kw = dict(group=H, residual_relu=True, precision='f16')
for _ in range(3):
    ops.relation(*t, **kw)
buf = torch.zeros(148 * 32, dtype=torch.int64, device='cuda')
torch.cuda.synchronize()
assert lib.rn_fused_trace2_set(ctypes.c_void_p(buf.data_ptr())) == 0
ops.relation(*t, stage_mask=6, **kw)
torch.cuda.synchronize()
a = buf.cpu().numpy().reshape(148, 2, 16)
a = a[a[:, 0, 0] != 0]
names = ['round start->key tab', 'st0 compute', 'st0 (pre-wait)', 'st0 a_free wait', 'st1 compute..', '', '', '', '', '', '', '', '', '']
for who, lbl in ((0, 'thread 0 (centre coordinate, j=0)'), (1, 'thread 256 (size coordinate, j=2)')):
    x = a[:, who, :]
    print(lbl)
    print('  key side ready            %6d' % np.median(x[:, 1] - x[:, 0]))
    for st in range(4):
        print('  stage %d: compute %6d   wait a_free %6d   (next stage starts +%d)' % (
            st, np.median(x[:, 3 + 3 * st] - x[:, 2 + 3 * st]), np.median(x[:, 4 + 3 * st] - x[:, 3 + 3 * st]),
            np.median((x[:, 5 + 3 * st] if st < 3 else x[:, 14]) - x[:, 4 + 3 * st])))
    print('  whole round (to last stores) %6d' % np.median(x[:, 14] - x[:, 0]))
