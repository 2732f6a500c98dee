"""ResNet-101 v1 trunk + RPN + conv_new_1 as plain torch/cuDNN modules (random init, frozen-BN folded into the convs).

OUT OF SCOPE as kernels: the trunk only exists to feed the hot path realistic tensors and
to report images/sec on BASELINE.json's configs[1]; nothing here is hand-written.  Layer layout follows
relation_rcnn/symbols/resnet_v1_101_rcnn_base.py: conv1 7x7/2, ceil-mode 3x3/2 max pool, res2..res4 (3, 4, 23 blocks,
stride on the first 1x1), res5 with dilation 2 / stride 1 (:621-683), RPN 3x3+relu -> 1x1 cls (2A) / 1x1 bbox (4A)
(:685-693), softmax over {bg, fg} per anchor, conv_new_1 1x1 2048->256 + relu (SYM_REL:249-250).
"""
import os
import torch
import torch.nn as nn
import torch.nn.functional as F

# cuDNN's fused conv + bias (+ residual) + relu (cudnnConvolutionBiasActivationForward via ATen): one library launch per
# conv instead of conv / bias / relu / add -- 3 launches per bottleneck instead of ~11, and the elementwise passes over the
# activations disappear.  RELNET_TRUNK_FUSED=0 falls back to the plain module calls (same arithmetic, more launches).
FUSED = os.environ.get('RELNET_TRUNK_FUSED', '1') != '0'


_PLAIN = [False]


class plain_ops(object):
    """context: route every conv of the trunk through the plain (autograd- and autocast-aware) module calls"""

    def __enter__(self):
        self.prev = _PLAIN[0]
        _PLAIN[0] = True

    def __exit__(self, *a):
        _PLAIN[0] = self.prev


def _training(conv):
    """the fused library calls have no autograd formula (and ignore autocast): when a gradient is wanted (train.py) use
    the plain module ops"""
    return _PLAIN[0] or (torch.is_grad_enabled() and conv.weight.requires_grad)


def _conv_relu(conv, x):
    if FUSED and x.is_cuda and not _training(conv):
        return torch.cudnn_convolution_relu(x, conv.weight, conv.bias, conv.stride, conv.padding, conv.dilation, conv.groups)
    return F.relu(conv(x))


def _conv_add_relu(conv, x, z):
    if FUSED and x.is_cuda and not _training(conv):
        return torch.cudnn_convolution_add_relu(x, conv.weight, z, 1.0, conv.bias, conv.stride, conv.padding, conv.dilation,
                                                conv.groups)
    return F.relu(conv(x) + z)


class Bottleneck(nn.Module):
    def __init__(self, cin, mid, cout, stride=1, dilation=1, project=False):
        super().__init__()
        self.c1 = nn.Conv2d(cin, mid, 1, stride=stride)
        self.c2 = nn.Conv2d(mid, mid, 3, padding=dilation, dilation=dilation)
        self.c3 = nn.Conv2d(mid, cout, 1)
        self.proj = nn.Conv2d(cin, cout, 1, stride=stride) if project else None

    def forward(self, x):
        y = _conv_relu(self.c2, _conv_relu(self.c1, x))
        return _conv_add_relu(self.c3, y, self.proj(x) if self.proj is not None else x)


def _stage(cin, mid, cout, n, stride, dilation=1):
    blocks = [Bottleneck(cin, mid, cout, stride, dilation, project=True)]
    blocks += [Bottleneck(cout, mid, cout, 1, dilation) for _ in range(n - 1)]
    return nn.Sequential(*blocks)


class Trunk(nn.Module):
    def __init__(self, num_anchors=12):
        super().__init__()
        self.A = num_anchors
        self.conv1 = nn.Conv2d(3, 64, 7, stride=2, padding=3)
        self.res2 = _stage(64, 64, 256, 3, 1)
        self.res3 = _stage(256, 128, 512, 4, 2)
        self.res4 = _stage(512, 256, 1024, 23, 2)
        self.res5 = _stage(1024, 512, 2048, 3, 1, dilation=2)
        self.rpn_conv = nn.Conv2d(1024, 512, 3, padding=1)
        self.rpn_cls = nn.Conv2d(512, 2 * num_anchors, 1)
        self.rpn_bbox = nn.Conv2d(512, 4 * num_anchors, 1)
        self.conv_new_1 = nn.Conv2d(2048, 256, 1)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode='fan_in', nonlinearity='relu')
                nn.init.zeros_(m.bias)
        # frozen BN is folded into the convs; emulate its normalisation so 33 residual blocks stay O(1):
        self.conv1.weight.data.mul_(1.0 / 50.0)                 # input is mean-subtracted pixels, std ~50
        for m in self.modules():
            if isinstance(m, Bottleneck):
                m.c3.weight.data.mul_(0.25)

    @torch.no_grad()
    def calibrate(self, image):
        """Data-dependent rescale of the four head convs (deterministic given the seed): RPN scores std 1.5, deltas std
        0.15, pooled feature std 1 -- a realistic top-k / NMS / ROI workload instead of saturated garbage."""
        x = F.max_pool2d(F.relu(self.conv1(image)), 3, 2, ceil_mode=True)
        c4 = self.res4(self.res3(self.res2(x)))
        c5 = self.res5(c4)
        self.rpn_conv.weight.data.div_(F.relu(self.rpn_conv(c4)).float().std().clamp_min(1e-6))
        r = F.relu(self.rpn_conv(c4))
        self.rpn_cls.weight.data.mul_(1.5 / self.rpn_cls(r).float().std().clamp_min(1e-6))
        self.rpn_bbox.weight.data.mul_(0.15 / self.rpn_bbox(r).float().std().clamp_min(1e-6))
        self.conv_new_1.weight.data.div_(F.relu(self.conv_new_1(c5)).float().std().clamp_min(1e-6))

    @torch.no_grad()
    def prepare(self):
        """Inference-time rewrites (identical arithmetic): the projection shortcut's bias moves into c3's bias so the
        shortcut is a bias-free conv (one launch), and conv1's 7x7/2 weight is re-indexed into the 4x4/1 weight over the
        space-to-depth input that ops.image_s2d produces."""
        for m in self.modules():
            if isinstance(m, Bottleneck) and m.proj is not None and m.proj.bias is not None:
                m.c3.bias.data += m.proj.bias.data
                m.proj.bias = None
        w = self.conv1.weight.data.float()                               # [64,3,7,7]
        w8 = torch.zeros((w.shape[0], 3, 8, 8), dtype=torch.float32, device=w.device)
        w8[:, :, :7, :7] = w
        w12 = w8.view(-1, 3, 4, 2, 4, 2).permute(0, 1, 3, 5, 2, 4).reshape(-1, 12, 4, 4)   # [o, c*4 + r*2 + s, a, b]
        w16 = torch.zeros((w.shape[0], 16, 4, 4), dtype=torch.float32, device=w.device)
        w16[:, :12] = w12
        self.conv1_s2d_weight = w16.to(self.conv1.weight.dtype).contiguous(memory_format=torch.channels_last)
        return self

    @torch.no_grad()
    def stem(self, image):
        """conv1 + relu + pool1.  fp32 CUDA image + FUSED: hand-written space-to-depth / max-pool kernels around one cuDNN
        conv (ops.image_s2d, ops.maxpool3x3s2_nhwc); otherwise the plain module path."""
        if FUSED and image.is_cuda and image.dtype == torch.float32 and getattr(self, 'conv1_s2d_weight', None) is not None \
                and image.shape[0] == 1 and image.shape[2] % 2 == 0 and image.shape[3] % 2 == 0:
            from . import ops
            y = ops.image_s2d(image, pad=3)
            x = torch.cudnn_convolution_relu(y, self.conv1_s2d_weight, self.conv1.bias, (1, 1), (0, 0), (1, 1), 1)
            return ops.maxpool3x3s2_nhwc(x)
        if image.dtype != self.conv1.weight.dtype:
            image = image.to(self.conv1.weight.dtype)
            if image.is_cuda:
                image = image.contiguous(memory_format=torch.channels_last)
        return F.max_pool2d(_conv_relu(self.conv1, image), 3, 2, ceil_mode=True)

    @torch.no_grad()
    def c4(self, image):
        """conv1 .. res4: the stride-16 feature both the RPN and res5 read"""
        return self.res4(self.res3(self.res2(self.stem(image))))

    @torch.no_grad()
    def rpn(self, c4):
        """-> rpn_cls_prob [1,2A,h,w] fp32, rpn_bbox_pred [1,4A,h,w] fp32 (SYM_BASE:685-693 + softmax over {bg, fg})"""
        r = _conv_relu(self.rpn_conv, c4)
        if FUSED and r.is_cuda and r.dtype == torch.bfloat16 and r.shape[0] == 1 and self.A <= 16:
            from . import ops          # both 1x1 heads + softmax + fp32 NCHW outputs: one hand-written kernel
            return ops.rpn_head(r, self.rpn_cls.weight, self.rpn_cls.bias, self.rpn_bbox.weight, self.rpn_bbox.bias)
        score = self.rpn_cls(r).float()
        b, _, h, w = score.shape
        prob = F.softmax(score.reshape(b, 2, self.A * h, w), dim=1).reshape(b, 2 * self.A, h, w)
        return prob.contiguous(), self.rpn_bbox(r).float().contiguous()

    @torch.no_grad()
    def c5feat(self, c4, keep_dtype=False):
        """res5 (dilated) + conv_new_1 + relu -> [1,256,h,w], left channels-last (consumed as NHWC); fp32 unless keep_dtype
        (ops.roi_pool_fc pools straight from the bf16 map)"""
        y = _conv_relu(self.conv_new_1, self.res5(c4))
        return y if keep_dtype else y.float()

    # ---- Deformable Faster-RCNN form of res5 (BASELINE.json configs[2]; resnet_v1_101_rcnn_dcn_*.py:700-748) -------------
    @torch.no_grad()
    def enable_dcn(self, seed=3):
        """Adds the three `res5{a,b,c}_branch2b_offset` convs (3x3, dilate 2, 512 -> 72 = 2*9*4 deformable groups).  The
        reference zero-initialises them (:1525-1530: the deformable conv then equals the plain one); N(0, 0.01) here so the
        gather is non-trivial.  The 3x3 weights of res5 become the deformable conv weights
        (kept in fp32 for the one-time fp16 pack)."""
        g = torch.Generator().manual_seed(seed)
        dev, dt = self.conv1.weight.device, self.conv1.weight.dtype
        self.dcn_offset = nn.ModuleList([nn.Conv2d(512, 72, 3, padding=2, dilation=2) for _ in range(3)]).to(device=dev, dtype=dt)
        for m in self.dcn_offset:
            m.weight.data.copy_((torch.randn(m.weight.shape, generator=g) * 0.01).to(dev))
            m.bias.data.zero_()
            if dev.type == 'cuda':
                m.to(memory_format=torch.channels_last)
        self.dcn_weight = [b.c2.weight.detach().float().contiguous() for b in self.res5]
        self.dcn_bias = [b.c2.bias.detach().float().contiguous() for b in self.res5]
        return self

    @torch.no_grad()
    def c5feat_dcn(self, c4):
        """res5 with DeformableConvolution in the three 3x3 positions (offset conv: library; deformable conv: our
        channels-last sampler + tcgen05 GEMM, bias + relu fused) + conv_new_1 + relu -> [1,256,h,w] channels-last."""
        from . import ops
        x = c4
        for i, blk in enumerate(self.res5):
            y1 = _conv_relu(blk.c1, x)
            off = self.dcn_offset[i](y1).float().contiguous()                        # [1,72,h,w] NCHW fp32 (0.7 MB)
            y2 = ops.deform_conv_nhwc(y1, off, self.dcn_weight[i], self.dcn_bias[i], relu=True)     # fp16 channels_last
            y2 = y2.to(y1.dtype)
            x = _conv_add_relu(blk.c3, y2, blk.proj(x) if blk.proj is not None else x)
        return _conv_relu(self.conv_new_1, x)

    @torch.no_grad()
    def forward(self, image):
        """image [1,3,H,W] -> (rpn_cls_prob, rpn_bbox_pred, conv_new_1_relu)"""
        c4 = self.c4(image)
        prob, bbox = self.rpn(c4)
        return prob, bbox, self.c5feat(c4)


def make_trunk(device, dtype=torch.bfloat16, seed=0):
    torch.manual_seed(seed)
    t = Trunk().eval()
    g = torch.Generator().manual_seed(seed + 1)
    t = t.to(device=device)
    t.calibrate((torch.randn((1, 3, 600, 1000), generator=g) * 50.0).to(device))
    t = t.to(dtype=dtype)
    if device != 'cpu' and str(device) != 'cpu':
        t = t.to(memory_format=torch.channels_last)
        t.prepare()
    return t


class FPNTrunk(nn.Module):
    """ResNet-101 + FPN feature maps for BASELINE.json configs[3] (resnet_v1_101_rcnn_fpn_*.py get_resnet_v1_fpn_conv):
    C2..C5 at strides 4..32 (res5 NOT dilated here), 1x1 laterals + nearest 2x top-down + 3x3 smoothing -> four 256-channel
    maps fpn_ft4 / 8 / 16 / 32.  Library plumbing (torch/cuDNN), random init, only there to feed the head realistic maps."""

    def __init__(self):
        super().__init__()
        self.conv1 = nn.Conv2d(3, 64, 7, stride=2, padding=3)
        self.res2 = _stage(64, 64, 256, 3, 1)
        self.res3 = _stage(256, 128, 512, 4, 2)
        self.res4 = _stage(512, 256, 1024, 23, 2)
        self.res5 = _stage(1024, 512, 2048, 3, 2)
        self.lat = nn.ModuleList([nn.Conv2d(c, 256, 1) for c in (256, 512, 1024, 2048)])
        self.smooth = nn.ModuleList([nn.Conv2d(256, 256, 3, padding=1) for _ in range(4)])
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode='fan_in', nonlinearity='relu')
                nn.init.zeros_(m.bias)
        self.conv1.weight.data.mul_(1.0 / 50.0)
        for m in self.modules():
            if isinstance(m, Bottleneck):
                m.c3.weight.data.mul_(0.25)

    @torch.no_grad()
    def forward(self, image):
        """image [1,3,H,W] (H, W multiples of 32) -> [fpn_ft4, fpn_ft8, fpn_ft16, fpn_ft32]"""
        if image.dtype != self.conv1.weight.dtype:
            image = image.to(self.conv1.weight.dtype)
        if image.is_cuda:
            image = image.contiguous(memory_format=torch.channels_last)
        c2 = self.res2(F.max_pool2d(_conv_relu(self.conv1, image), 3, 2, padding=1))
        c3 = self.res3(c2); c4 = self.res4(c3); c5 = self.res5(c4)
        p = self.lat[3](c5)
        outs = [None, None, None, self.smooth[3](p)]
        for l, c in ((2, c4), (1, c3), (0, c2)):
            p = self.lat[l](c) + F.interpolate(p, size=c.shape[-2:], mode='nearest')
            outs[l] = self.smooth[l](p)
        return outs


def make_fpn_trunk(device, dtype=torch.bfloat16, seed=0):
    torch.manual_seed(seed)
    t = FPNTrunk().eval().to(device=device, dtype=dtype)
    if str(device) != 'cpu':
        t = t.to(memory_format=torch.channels_last)
    return t
