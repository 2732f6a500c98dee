"""Data-parallel training step of the end-to-end graph (reference: train_end2end.py:57-177,
core/module.py:569-591): one image per rank per micro-batch, forward + backward, every trainable gradient lands in ONE flat
fp32 GradientBucket, ONE NCCL SUM allreduce per step (kvstore 'device' semantics: sum, rescale_grad = 1.0), SGD update.

What runs where:
  * trunk conv1 / res2 frozen (FIXED_PARAMS of the yaml), res3 .. res5 + RPN + conv_new_1: torch/cuDNN with autograd
    (library, out of scope as kernels -- bf16 autocast over fp32 master weights so the gradients are fp32 bucket views);
  * hot path: proposal, proposal_target, ROIPooling fwd/bwd, relation fwd/bwd (x2), learn-NMS fwd/bwd, nms_multi_target,
    learn-NMS loss -- the repo's own C-ABI kernels through relnet_b200.autograd;
  * fc_new_1/2, cls_score, bbox_pred and the two softmax / smooth-L1 losses: torch (library GEMMs under autocast).
RPN anchor labels come from the data loader in the reference (lib/rpn/rpn.py:80-243, CPU, out of scope): here they are a
seeded synthetic label map so that the RPN branch carries a gradient of the right shape.
"""
import torch
import torch.nn.functional as F
from . import autograd as AG
from . import ops, replicas
from .pipeline import HEAD_PARAM_SHAPES, NMS_NAMES, init_head_params

FROZEN_PREFIXES = ('conv1', 'res2')            # FIXED_PARAMS: conv1, bn*, res2 (BN is folded / frozen everywhere here)


class TrainStep(object):
    def __init__(self, trunk, device, num_gt=8, micro_batches=1, seed=0, lr=0.0, nongt_dim=300, first_n=100):
        self.dev, self.micro, self.lr = device, micro_batches, lr
        self.nongt_dim, self.first_n = nongt_dim, first_n
        self.trunk = trunk.float()                                   # fp32 master weights, bf16 autocast in the convs
        for n, q in self.trunk.named_parameters():
            q.requires_grad_(not n.startswith(FROZEN_PREFIXES))
        self.head = {k: v.clone().requires_grad_(True) for k, v in init_head_params(seed, device).items()}
        shapes = {'trunk.' + n: tuple(q.shape) for n, q in self.trunk.named_parameters() if q.requires_grad}
        shapes.update({k: tuple(v.shape) for k, v in self.head.items()})
        self.bucket = replicas.GradientBucket(shapes, device=device)
        self.params = {'trunk.' + n: q for n, q in self.trunk.named_parameters() if q.requires_grad}
        self.params.update(self.head)
        for n, q in self.params.items():
            # autograd accumulates in place into the bucket windows.  A channels-last conv weight gets a window with ITS strides
            # (same bytes, dense permutation): otherwise every accumulation converts the layout (one extra kernel per weight)
            v = self.bucket.views[n]
            if q.dim() == 4 and not q.is_contiguous() and q.is_contiguous(memory_format=torch.channels_last):
                v = v.view(-1).as_strided(q.shape, q.stride())
                self.bucket.views[n] = v
            q.grad = v
        self.state = {}
        g = torch.Generator().manual_seed(seed + 17)
        self.gt = torch.tensor([[100. + 90 * i, 60. + 50 * i, 260. + 90 * i, 300. + 50 * i, 1. + (i % 80)] for i in range(num_gt)],
                               device=device)
        self.rpn_label = None
        self.gen = g
        self._last = {}
        self.graph = None            # set by capture(): the whole accumulate phase (zero + fwd + bwd of every micro-batch) as ONE CUDA graph
        self.fwd_precision = 'tf32' if ops.device_info()['sm100'] else 'fp32'   # general kernels, GEMMs on the tcgen05 tf32 engine

    @property
    def last(self):
        """losses of the most recent step as Python numbers (one host sync, outside the step)"""
        return {k: (float(v) if isinstance(v, torch.Tensor) else v) for k, v in self._last.items()}

    # ------------------------------------------------------------------------------------------------ one image
    def forward_backward(self, image32, im_info):
        P, t, dev = self.head, self.trunk, self.dev
        from .trunk import plain_ops
        with plain_ops(), torch.autocast('cuda', dtype=torch.bfloat16):
            with torch.no_grad():                                    # frozen prefix (FIXED_PARAMS)
                x = image32.contiguous(memory_format=torch.channels_last)
                x = t.res2(F.max_pool2d(F.relu(t.conv1(x)), 3, 2, ceil_mode=True))
            c4 = t.res4(t.res3(x))
            r = F.relu(t.rpn_conv(c4))
            score, rpn_bbox = t.rpn_cls(r).float(), t.rpn_bbox(r).float()
            feat = F.relu(t.conv_new_1(t.res5(c4)))
        b, _, h, w = score.shape
        A = t.A
        logits = score.reshape(b, 2, A * h, w)
        if self.rpn_label is None:                                   # synthetic anchor labels: 1/8 labelled, 1/4 of them fg
            u = torch.rand((b, A * h, w), generator=self.gen).to(dev)
            self.rpn_label = torch.where(u < 0.03, 1, torch.where(u < 0.125, 0, -1)).long()
            self.rpn_bbox_t = (torch.randn(rpn_bbox.shape, generator=self.gen) * 0.1).to(dev)
        rpn_cls_loss = F.cross_entropy(logits, self.rpn_label, ignore_index=-1)
        fg = (self.rpn_label == 1).reshape(b, A, h, w).repeat_interleave(4, dim=1).float()
        rpn_bbox_loss = (F.smooth_l1_loss(rpn_bbox, self.rpn_bbox_t, reduction='none', beta=1.0 / 9) * fg).sum() / 256.0
        with torch.no_grad():                                        # CustomOps with zero input gradients (proposal.py:170-173)
            prob = F.softmax(logits, dim=1).reshape(b, 2 * A, h, w).contiguous()
            rois = ops.proposal(prob, rpn_bbox.detach().contiguous(), im_info)[0]
            rois, label, bbox_target, bbox_weight = ops.proposal_target(rois, self.gt)
            boxes = rois[:, 1:].contiguous()
        pooled = AG.roi_pool(feat.float().contiguous(), rois, (7, 7), 1.0 / 16)                   # SYM_REL_NMS:335
        with torch.autocast('cuda', dtype=torch.bfloat16):
            fc1 = F.linear(pooled.flatten(1), P['fc_new_1_weight'], P['fc_new_1_bias']).float()
        rel = lambda x, i: AG.relation(x, boxes, P['query_%d_weight' % i], P['query_%d_bias' % i], P['key_%d_weight' % i],
                                       P['key_%d_bias' % i], P['pair_pos_fc1_%d_weight' % i], P['pair_pos_fc1_%d_bias' % i],
                                       P['linear_out_%d_weight' % i], P['linear_out_%d_bias' % i], M=self.nongt_dim, group=16,
                                       residual_relu=True, precision=self.fwd_precision)
        a1 = rel(fc1, 1)                                                                          # :346-351
        with torch.autocast('cuda', dtype=torch.bfloat16):
            fc2 = F.linear(a1, P['fc_new_2_weight'], P['fc_new_2_bias']).float()
        a2 = rel(fc2, 2)                                                                          # :354-359
        cls_score = F.linear(a2, P['cls_score_weight'], P['cls_score_bias'])
        bbox_pred = F.linear(a2, P['bbox_pred_weight'], P['bbox_pred_bias'])
        cls_loss = F.cross_entropy(cls_score, label.long())
        bbox_loss = (F.smooth_l1_loss(bbox_pred, bbox_target, reduction='none', beta=1.0) * bbox_weight).sum() / rois.shape[0]
        multi, sbbox, sscore = AG.learn_nms(cls_score, bbox_pred, rois, im_info, a2, {k: P[k] for k in NMS_NAMES},
                                            first_n=self.first_n, means=(0, 0, 0, 0), stds=(0.1, 0.1, 0.2, 0.2),
                                            nongt_dim=self.nongt_dim, precision=self.fwd_precision)   # :424-501 (train graph)
        with torch.no_grad():
            target = ops.nms_multi_target(sbbox, self.gt, sscore, [0.5, 0.6, 0.7, 0.8, 0.9])      # :537-538
            pos, neg, d_multi = ops.nms_loss(multi.detach(), target)                              # :539-551
        torch.autograd.backward([rpn_cls_loss + rpn_bbox_loss + cls_loss + bbox_loss, multi], [None, d_multi])
        # loss values stay on the device (no host sync inside the step: it must be capturable); `last` reads them on demand
        self._last = dict(rpn_cls=rpn_cls_loss.detach(), cls=cls_loss.detach(), bbox=bbox_loss.detach(),
                          nms=(pos.sum() + neg.sum()).detach(), rois=int(rois.shape[0]))

    def accumulate(self, images32, im_info):
        """bucket <- sum over this rank's micro-batches of the gradients (zero, then fwd + bwd per image)"""
        self.bucket.zero_()
        for im in images32:
            self.forward_backward(im, im_info)

    def capture(self, images32, im_info, warmup=3):
        """Capture the accumulate phase as ONE CUDA graph (every launch of the step -- the library trunk's autograd, the C-ABI
        forward / backward pairs, the memsets -- is allocation-free after warm-up and host-sync-free); step() then replays it and
        runs the allreduce + SGD update behind it.  The eager step is CPU bound (≈20 ms of launches for ≈8 ms of GPU work)."""
        self._static_images = [im.clone() for im in images32]
        self._static_info = im_info.clone()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self.accumulate(self._static_images, self._static_info)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.accumulate(self._static_images, self._static_info)
        self.graph = g
        return self

    # ------------------------------------------------------------------------------------------------ one step
    def step(self, images32, im_info, exchange=True):
        """images32: list of `micro_batches` fp32 [1,3,H,W] images of this rank (gradients accumulate locally, like the
        reference's 2 images per GPU in the FPN config); then ONE allreduce, then the SGD update."""
        if self.graph is not None:
            for dst, src in zip(self._static_images, images32):
                if dst.data_ptr() != src.data_ptr():
                    dst.copy_(src, non_blocking=True)
            self.graph.replay()
        else:
            self.accumulate(images32, im_info)
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        ev[0].record()
        if exchange:
            self.bucket.allreduce()
        ev[1].record()
        replicas.sgd_step({n: q.data for n, q in self.params.items()}, self.bucket, lr=self.lr, state=self.state)
        return ev


class FPNTrainStep(TrainStep):
    """BASELINE.json configs[3] as a training step (train_rcnn.py -> get_symbol_rcnn, SYM_FPN_REL_NMS:979-1233): FPN trunk with
    autograd (library; conv1 / res2 frozen), rois GIVEN (the reference reads them from a proposal pickle through ROIIter and
    samples / labels them on the host, core/rcnn.py:153-223 -- here: `num_rois` seeded synthetic boxes labelled on the device by
    rn_proposal_target, which appends the gt boxes: N = num_rois + G rows, keys = the num_rois non-gt rows), ROIPooling per
    pyramid level (strides 4..32) through the C-ABI fwd/bwd pair, roi_pool_fc1/2 + two relation modules (C-ABI fwd/bwd),
    cls / bbox losses, learn-NMS with first_n = 150 (C-ABI fwd/bwd) and its loss.  `micro_batches` images per rank accumulate
    into the same bucket before the ONE allreduce (the reference's 2 images per GPU: batch 16 on 8 GPUs)."""

    STRIDES = (4, 8, 16, 32)

    def __init__(self, trunk, device, num_rois=1000, num_gt=8, micro_batches=2, seed=0, lr=0.0, first_n=150, canvas=(1024.0, 608.0)):
        super().__init__(trunk, device, num_gt=num_gt, micro_batches=micro_batches, seed=seed, lr=lr, nongt_dim=num_rois,
                         first_n=first_n)
        import numpy as np
        rng = np.random.default_rng(seed + 5)
        W, H = canvas
        sz = np.exp(rng.uniform(np.log(16.0), np.log(0.95 * min(W, H)), num_rois))
        ar = np.exp(rng.uniform(np.log(0.5), np.log(2.0), num_rois))
        w = np.minimum(sz * np.sqrt(ar), W - 2); h = np.minimum(sz / np.sqrt(ar), H - 2)
        x1 = rng.uniform(0, 1, num_rois) * (W - 1 - w); y1 = rng.uniform(0, 1, num_rois) * (H - 1 - h)
        self.rois_in = torch.tensor(np.stack([np.zeros(num_rois), x1, y1, x1 + w, y1 + h], 1), dtype=torch.float32, device=device)
        # pyramid dispatch of the (given, fixed) rois: done once on the host side, like the reference's loader (core/rcnn.py:153-223)
        from .pipeline import fpn_level
        with torch.no_grad():
            rois0 = ops.proposal_target(self.rois_in, self.gt)[0]
            lvl = fpn_level(rois0)
            self._idx = [torch.nonzero(lvl == l).flatten() for l in range(4)]
            self._inv = torch.argsort(torch.cat(self._idx))

    def forward_backward(self, image32, im_info):
        from .pipeline import fpn_level
        from .trunk import plain_ops
        P, t = self.head, self.trunk
        with plain_ops(), torch.autocast('cuda', dtype=torch.bfloat16):
            with torch.no_grad():                                    # frozen prefix (FIXED_PARAMS)
                x = image32.contiguous(memory_format=torch.channels_last)
                c2 = t.res2(F.max_pool2d(F.relu(t.conv1(x)), 3, 2, padding=1))
            c3 = t.res3(c2); c4 = t.res4(c3); c5 = t.res5(c4)
            p = t.lat[3](c5)
            feats = [None, None, None, t.smooth[3](p)]
            for l, c in ((2, c4), (1, c3), (0, c2)):
                p = t.lat[l](c) + F.interpolate(p, size=c.shape[-2:], mode='nearest')
                feats[l] = t.smooth[l](p)
        with torch.no_grad():
            rois, label, bbox_target, bbox_weight = ops.proposal_target(self.rois_in, self.gt)     # appends the gt rows
            boxes = rois[:, 1:].contiguous()
            idx, inv = self._idx, self._inv
        parts = [AG.roi_pool(feats[l].float().contiguous(), rois[idx[l]].contiguous(), (7, 7), 1.0 / self.STRIDES[l])
                 for l in range(4) if idx[l].numel()]                                             # SYM_FPN_REL_NMS:1108-1115
        pooled = torch.cat(parts, 0)[inv]
        with torch.autocast('cuda', dtype=torch.bfloat16):
            fc1 = F.linear(pooled.flatten(1), P['fc_new_1_weight'], P['fc_new_1_bias']).float()   # roi_pool_fc1 :1130
        rel = lambda x, i: AG.relation(x, boxes, P['query_%d_weight' % i], P['query_%d_bias' % i], P['key_%d_weight' % i],
                                       P['key_%d_bias' % i], P['pair_pos_fc1_%d_weight' % i], P['pair_pos_fc1_%d_bias' % i],
                                       P['linear_out_%d_weight' % i], P['linear_out_%d_bias' % i], M=self.nongt_dim, group=16,
                                       residual_relu=True, precision=self.fwd_precision)
        a1 = rel(fc1, 1)                                                                          # :1122-1133
        with torch.autocast('cuda', dtype=torch.bfloat16):
            fc2 = F.linear(a1, P['fc_new_2_weight'], P['fc_new_2_bias']).float()                  # roi_pool_fc2 :1138
        a2 = rel(fc2, 2)                                                                          # :1134-1141
        cls_score = F.linear(a2, P['cls_score_weight'], P['cls_score_bias'])
        bbox_pred = F.linear(a2, P['bbox_pred_weight'], P['bbox_pred_bias'])
        cls_loss = F.cross_entropy(cls_score, label.long())
        bbox_loss = (F.smooth_l1_loss(bbox_pred, bbox_target, reduction='none', beta=1.0) * bbox_weight).sum() / rois.shape[0]
        multi, sbbox, sscore = AG.learn_nms(cls_score, bbox_pred, rois, im_info, a2, {k: P[k] for k in NMS_NAMES},
                                            first_n=self.first_n, means=(0, 0, 0, 0), stds=(0.1, 0.1, 0.2, 0.2),
                                            nongt_dim=self.nongt_dim, precision=self.fwd_precision)
        with torch.no_grad():
            target = ops.nms_multi_target(sbbox, self.gt, sscore, [0.5, 0.6, 0.7, 0.8, 0.9])
            pos, neg, d_multi = ops.nms_loss(multi.detach(), target)
        torch.autograd.backward([cls_loss + bbox_loss, multi], [None, d_multi])
        self._last = dict(cls=cls_loss.detach(), bbox=bbox_loss.detach(), nms=(pos.sum() + neg.sum()).detach(),
                          rois=int(rois.shape[0]), per_level=[int(i.numel()) for i in idx])


def bus_gbs(nbytes, ms, world):
    """NCCL bus bandwidth of an allreduce: algorithmic bytes x 2 (n-1)/n / time"""
    if world <= 1 or ms <= 0:
        return 0.0
    return nbytes * 2.0 * (world - 1) / world / (ms * 1e-3) / 1e9
