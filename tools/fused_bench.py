"""Fused relation kernel (relation_fused.cu) vs the round-1 decomposition, stage by stage.

For N = M in {300, 1000, 3000} (d = 1024, H = 16) and the other sweep shapes: the N x M part (stage mask 6 = geometry +
attention; the projection GEMM is common to both arms) timed as a CAPTURED GRAPH of that one stage (no host gaps, tensor-map
encode outside the timed region) with a 256 MB L2 flush between replays, CUDA events, median of 15.  One JSON line per point:
microseconds of both arms, achieved TFLOP/s = 4 N M d / t against the measured bf16 peak, module time.
"""
import json
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as entry
entry.build()
import relnet_b200
from relnet_b200 import ops, synth

dev = torch.device('cuda:0')
pk = os.path.join(ROOT, 'MEASURED_PEAKS.json')
peak = json.load(open(pk))['bf16_tflops'] if os.path.exists(pk) else 1590.0
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def graph_time(fn, reps=15):
    torch.cuda.synchronize()              # the capture stream must not overlap earlier work on the same workspace
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
        s.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            fn()
    torch.cuda.synchronize()
    ts = []
    for i in range(reps):
        flush.fill_(i & 1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


points = [(300, 1024, 16), (1000, 1024, 16), (3000, 1024, 16), (100, 256, 4), (300, 256, 16), (1000, 256, 16), (3000, 256, 16),
          (300, 256, 4), (1000, 256, 4), (3000, 256, 4)]
if len(sys.argv) > 1:
    points = [tuple(int(v) for v in a.split(',')) for a in sys.argv[1:]]
for N, d, H in points:
    c = synth.make_relation_case(N * 31 + d + H, N, d, H)
    t = [torch.from_numpy(c[k]).to(dev) for k in ('X', 'boxes', 'Wq', 'bq', 'Wk', 'bk', 'Wg', 'bg', 'Wout', 'bout')]
    rec = dict(N=N, d=d, H=H, dk=d // H, F_tc_gflop=round(4.0 * N * N * d / 1e9, 4))
    outs = {}
    for arm, on in (('fused', 1), ('unfused', 0)):
        if not on and d // H != 64:
            continue                      # the round-1 tile kernel is the d_k = 64 arm
        ops.relation_fused_enable(on)
        ws = torch.empty(ops.relation_workspace_bytes(N, N, d, d, d, H) + 4096, dtype=torch.uint8, device=dev)
        kw = dict(group=H, residual_relu=True, precision='f16', workspace=ws)
        outs[arm] = ops.relation(*t, **kw).clone()
        print('[%s N=%d] first call done' % (arm, N), file=sys.stderr, flush=True)
        rec[arm + '_nm_us'] = round(graph_time(lambda: ops.relation(*t, stage_mask=6, **kw)), 2)
        print('[%s N=%d] nm timed' % (arm, N), file=sys.stderr, flush=True)
        rec[arm + '_module_us'] = round(graph_time(lambda: ops.relation(*t, **kw)), 2)
        print('[%s N=%d] module timed' % (arm, N), file=sys.stderr, flush=True)
        ach = 4.0 * N * N * d / (rec[arm + '_nm_us'] * 1e-6) / 1e12
        rec[arm + '_tflops'] = round(ach, 2)
        rec[arm + '_frac_of_measured_bf16_peak'] = round(ach / peak, 4)
        del ws
    ops.relation_fused_enable(1)
    if len(outs) == 2:
        rec['fused_vs_unfused_rel'] = float((outs['fused'] - outs['unfused']).abs().max() / outs['unfused'].abs().max())
    print(json.dumps(rec), flush=True)
