"""GPU: the hand-written stem helpers around the (library) trunk -- space-to-depth image conversion and the channels-last
bf16 max pool -- against plain torch, plus the re-indexed 4x4/1 stem weight against the original 7x7/2 convolution."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def ops(cuda_device):
    import __graft_entry__ as g
    g.build()
    import relnet_b200
    torch.cuda.set_device(cuda_device)
    return relnet_b200.ops


def test_image_s2d_layout_and_stem_equivalence(ops):
    torch.manual_seed(0)
    img = torch.randn(1, 3, 60, 100, device='cuda') * 50
    y = ops.image_s2d(img, pad=3)
    assert y.shape == (1, 16, 33, 53) and y.dtype == torch.bfloat16 and y.is_contiguous(memory_format=torch.channels_last)
    xp = F.pad(img.to(torch.bfloat16), (3, 3, 3, 3))
    want = torch.zeros_like(y)
    for c in range(3):
        for r in range(2):
            for s in range(2):
                want[0, c * 4 + r * 2 + s] = xp[0, c, r::2, s::2]
    assert torch.equal(y, want)                                   # pure data movement: bit-exact
    # 7x7 / stride 2 / pad 3 convolution == 4x4 / stride 1 convolution over the space-to-depth tensor
    from relnet_b200.trunk import Trunk
    t = Trunk().cuda().eval()
    t.prepare()
    ref = F.conv2d(img.to(torch.bfloat16).float(), t.conv1.weight.float(), None, stride=2, padding=3)
    got = F.conv2d(y.float(), t.conv1_s2d_weight.float(), None)
    assert got.shape == ref.shape
    assert (got - ref).abs().max().item() <= 1e-3 * ref.abs().max().item()      # same products, different sum order


@pytest.mark.parametrize('H,W', [(300, 500), (31, 17)])
def test_maxpool_nhwc_matches_torch_ceil_mode(ops, H, W):
    torch.manual_seed(1)
    x = torch.randn(1, 64, H, W, device='cuda').to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    got = ops.maxpool3x3s2_nhwc(x)
    want = F.max_pool2d(x, 3, 2, ceil_mode=True)
    assert got.shape == want.shape and torch.equal(got, want)
    with pytest.raises(Exception):
        ops.maxpool3x3s2_nhwc(x.float())


def test_trunk_fast_stem_equals_plain_path(ops):
    from relnet_b200 import trunk as TR
    t = TR.make_trunk('cuda')
    img = torch.randn(1, 3, 224, 320, device='cuda') * 50
    fast = t.stem(img)
    old = TR.FUSED
    try:
        TR.FUSED = False
        plain = t.stem(img)
    finally:
        TR.FUSED = old
    assert fast.shape == plain.shape
    assert (fast.float() - plain.float()).abs().max().item() <= 2e-2 * plain.float().abs().max().item()    # bf16 conv, different algo


def test_rpn_head_kernel_matches_module_path(ops):
    """rn_rpn_head_fwd (1x1 cls/bbox convs + {bg,fg} softmax + fp32 NCHW layout) vs the torch module path in float32"""
    from relnet_b200 import trunk as TR
    t = TR.make_trunk('cuda')
    torch.manual_seed(3)
    r = torch.randn(1, 512, 38, 63, device='cuda').clamp_min(0).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    prob, bbox = ops.rpn_head(r, t.rpn_cls.weight, t.rpn_cls.bias, t.rpn_bbox.weight, t.rpn_bbox.bias)
    rf = r.float()
    score = F.conv2d(rf, t.rpn_cls.weight.float(), t.rpn_cls.bias.float())
    want_prob = F.softmax(score.reshape(1, 2, 12 * 38, 63), dim=1).reshape(1, 24, 38, 63)
    want_bbox = F.conv2d(rf, t.rpn_bbox.weight.float(), t.rpn_bbox.bias.float())
    assert prob.shape == want_prob.shape and bbox.shape == want_bbox.shape and prob.dtype == torch.float32
    assert (prob - want_prob).abs().max().item() <= 1e-4                      # fp32 accumulate on both sides
    assert (bbox - want_bbox).abs().max().item() <= 1e-4 * max(1.0, want_bbox.abs().max().item())
    assert torch.allclose(prob[:, :12] + prob[:, 12:], torch.ones_like(prob[:, :12]), atol=1e-6)
    # the portable SIMT form (rn_rpn_head_fwd) and the tcgen05 form (rn_rpn_head_packed_fwd, the default on sm_100) agree
    ps, bs = ops.rpn_head(r, t.rpn_cls.weight, t.rpn_cls.bias, t.rpn_bbox.weight, t.rpn_bbox.bias, simt=True)
    assert (ps - prob).abs().max().item() <= 1e-5 and (bs - bbox).abs().max().item() <= 1e-4 * max(1.0, bbox.abs().max().item())
    # HW not a multiple of the 32-position tile
    r2 = r[:, :, :5, :7].contiguous(memory_format=torch.channels_last)
    p2, b2 = ops.rpn_head(r2, t.rpn_cls.weight, t.rpn_cls.bias, t.rpn_bbox.weight, t.rpn_bbox.bias)
    assert torch.allclose(p2, prob[:, :, :5, :7], atol=1e-6) and torch.allclose(b2, bbox[:, :, :5, :7], atol=1e-5)


def test_roi_pool_fc_from_bf16_equals_fp32_path(ops):
    """pooling straight from the trunk's bf16 channels-last map gives bitwise the pooled values of the fp32 kernel"""
    if not ops.device_info()['sm100']:
        pytest.skip('tcgen05 path only')
    rng = np.random.RandomState(4)
    feat = torch.from_numpy(rng.randn(1, 256, 38, 63).astype(np.float32)).cuda().clamp_min(0).to(torch.bfloat16) \
        .contiguous(memory_format=torch.channels_last)
    R = 300
    x1 = rng.uniform(0, 800, R); y1 = rng.uniform(0, 450, R)
    rois = torch.from_numpy(np.stack([np.zeros(R), x1, y1, np.minimum(x1 + rng.uniform(4, 600, R), 999),
                                      np.minimum(y1 + rng.uniform(4, 450, R), 599)], 1).astype(np.float32)).cuda()
    rois[0] = torch.tensor([0, 990.0, 590.0, 999.0, 599.0])             # a roi whose bins fall off the map
    W = torch.from_numpy((rng.randn(64, 256 * 49) * 0.01).astype(np.float32)).cuda(); b = torch.zeros(64, device='cuda')
    y_bf = ops.roi_pool_fc(feat, rois, W, b)
    y_32 = ops.roi_pool_fc(feat.float(), rois, W, b)
    assert torch.equal(y_bf, y_32)
