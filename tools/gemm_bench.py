"""tcgen05 GEMM (gemm_tc.cu) at the head's shapes: M = 300 rows, K and N varied; microseconds per call (CUDA events, median
of 20), with and without an L2 flush between calls, and the split-K decision.  One JSON line per shape."""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as entry
entry.build()
import relnet_b200
from relnet_b200 import ops

dev = torch.device('cuda:0')
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timeit(fn, do_flush, reps=20):
    for _ in range(3):
        fn()
    evs = []
    for i in range(reps):
        if do_flush:
            flush.fill_(i & 1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    ts = sorted(x.elapsed_time(y) * 1e3 for x, y in evs)
    return round(ts[len(ts) // 2], 1)


for M, K, N in ((300, 1024, 64), (300, 1024, 256), (300, 1024, 1024), (300, 1024, 3072), (300, 128, 3072), (300, 12544, 1024),
                (100, 1024, 128), (2394, 1024, 512)):
    x = torch.randn(M, K, device=dev); W = torch.randn(N, K, device=dev) * 0.03; b = torch.zeros(N, device=dev)
    x16 = x.half()
    fn = lambda: ops.linear(x, W, b, precision='f16', x_f16=x16)
    print(json.dumps(dict(M=M, K=K, N=N, us_warm=timeit(fn, False), us_l2_flushed=timeit(fn, True),
                          gflop=round(2.0 * M * N * K / 1e9, 3))), flush=True)
