"""Data-parallel training step of the head: one image per rank, forward + backward through the C-ABI kernels
(relnet_b200.autograd), every gradient written straight into one flat GradientBucket, ONE NCCL SUM allreduce, SGD update.
Launch:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/train_step.py
(or plain `python tools/train_step.py` for one GPU).  Rank 0 prints a JSON line: step time (CUDA events, max over ranks),
the allreduce share of it, and a check that the reduced bucket equals the sum of the per-rank gradients (all_gather)."""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import __graft_entry__ as entry

rank = int(os.environ.get('RANK', 0)); world = int(os.environ.get('WORLD_SIZE', 1)); local = int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
if world > 1:
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
if rank == 0:
    entry.build()
if world > 1:
    dist.barrier()
import relnet_b200
from relnet_b200 import autograd as AG, ops, replicas
from oracle import relation_np as R, learn_nms_np as LN        # synthetic-case generators only (test infrastructure)

N, d, H, C, n = 300, 1024, 16, 80, 100
rel = [R.make_relation_case(100 + i, N, d, H) for i in (1, 2)]          # same weights on every rank
lp = LN.make_learn_nms_case(7, R=N, C=C, d=d)['P']
rng = np.random.RandomState(5)
shapes = {}
params = {}
for i, c in enumerate(rel, 1):
    for k in ('Wq', 'bq', 'Wk', 'bk', 'Wg', 'bg', 'Wout', 'bout'):
        params['rel%d_%s' % (i, k)] = torch.from_numpy(c[k]).to(dev)
params['fc2_weight'] = torch.from_numpy((rng.randn(d, d) / 32).astype(np.float32)).to(dev)
params['cls_score_weight'] = torch.from_numpy((rng.randn(C + 1, d) * 0.03).astype(np.float32)).to(dev)
params['bbox_pred_weight'] = torch.from_numpy((rng.randn(8, d) * 0.01).astype(np.float32)).to(dev)
for k, v in lp.items():
    params[k] = torch.from_numpy(v).to(dev)
bucket = replicas.GradientBucket({k: tuple(v.shape) for k, v in params.items()}, device=dev)
for k, v in params.items():
    v.requires_grad_(True)
    v.grad = bucket.views[k]                      # autograd accumulates in place into the bucket windows
img = LN.make_learn_nms_case(1000 + rank, R=N, C=C, d=d)                 # this rank's image: rois, head inputs
X0 = torch.from_numpy(img['feat']).to(dev)
rois = torch.from_numpy(img['rois']).to(dev); boxes = rois[:, 1:].contiguous()
im_info = torch.from_numpy(img['im_info']).to(dev)
cls_bias = torch.from_numpy(img['cls_score']).to(dev)
labels = torch.from_numpy((np.random.RandomState(rank).rand(N) < 0.25) * np.random.RandomState(rank + 9).randint(1, C + 1, N)).to(dev)
gt = torch.from_numpy(np.hstack([img['rois'][:8, 1:] + 3.0, np.arange(1, 9, dtype=np.float32)[:, None]]).astype(np.float32)).to(dev)
rel_keys = ('Wq', 'bq', 'Wk', 'bk', 'Wg', 'bg', 'Wout', 'bout')


def step(sync_only=False):
    bucket.zero_()
    A1 = AG.relation(X0, boxes, *[params['rel1_' + k] for k in rel_keys], group=H, residual_relu=True)
    X2 = A1 @ params['fc2_weight'].T
    A2 = AG.relation(X2, boxes, *[params['rel2_' + k] for k in rel_keys], group=H, residual_relu=True)
    cls_score = A2 @ params['cls_score_weight'].T + cls_bias
    bbox_pred = A2 @ params['bbox_pred_weight'].T
    multi, sbbox, sscore = AG.learn_nms(cls_score, bbox_pred, rois, im_info, A2, {k: params[k] for k in lp}, first_n=n,
                                        means=(0, 0, 0, 0), stds=(0.1, 0.1, 0.2, 0.2))
    target = ops.nms_multi_target(sbbox, gt, sscore, [0.5, 0.6, 0.7, 0.8, 0.9])
    pos, neg, d_multi = ops.nms_loss(multi.detach(), target)
    loss_cls = torch.nn.functional.cross_entropy(cls_score, labels.long())
    torch.autograd.backward([loss_cls, multi], [None, d_multi])
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    bucket.allreduce()
    e1.record()
    replicas.sgd_step(params_data, bucket, lr=0.0)          # lr 0: keep the weights fixed across the timed steps
    return float(loss_cls.detach()) + float(pos.sum() + neg.sum()), (e0, e1)


params_data = {k: v.data for k, v in params.items()}
# correctness of the exchange: reduced bucket == sum over ranks of the local buckets
bucket.zero_()
loss0, _ = step()
if world > 1:
    # redo the local part without the exchange to get this rank's own gradient
    reduced = bucket.flat.clone()
    saved = bucket.allreduce
    bucket.allreduce = lambda *a, **k: None
    step()
    bucket.allreduce = saved
    mine = bucket.flat.clone()
    allg = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allg, mine)
    want = torch.stack(allg).sum(0)
    err = float((reduced - want).abs().max() / want.abs().max())
else:
    err = 0.0
W, K = 3, 10
for _ in range(W):
    step()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
a.record()
ar = []
for _ in range(K):
    ar.append(step()[1])
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / K
ms_ar = sum(x.elapsed_time(y) for x, y in ar) / K
ms = replicas.max_over_ranks(ms, dev); ms_ar = replicas.max_over_ranks(ms_ar, dev)
if rank == 0:
    print(json.dumps(dict(tool='train_step', n_gpus=world, images_per_step=world, ms_per_step=round(ms, 3),
                          allreduce_ms=round(ms_ar, 3), bucket_mb=round(bucket.flat.numel() * 4 / 1e6, 2),
                          images_per_sec=round(world / (ms / 1e3), 1), allreduce_vs_sum_of_ranks_rel_err=err,
                          loss_rank0=round(loss0, 4),
                          note='head only (2x relation + fc2 + cls/bbox + learn-NMS fwd+bwd, fp32), eager launches')))
if world > 1:
    dist.destroy_process_group()
