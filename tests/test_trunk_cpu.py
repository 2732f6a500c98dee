"""CPU: the inference-time rewrites of the (library) trunk are exact re-indexings -- the 7x7/2 stem conv as a 4x4/1 conv over
the space-to-depth input (the layout rn_image_s2d_bf16 writes), and the projection-shortcut bias folded into c3."""
import torch
import torch.nn.functional as F


def _s2d(img, pad=3):
    """reference layout of rn_image_s2d_bf16: channel = c*4 + (row parity)*2 + (col parity), 16 channels (12..15 zero)"""
    xp = F.pad(img, (pad, pad, pad, pad))
    B, C, H, W = xp.shape
    y = torch.zeros((B, 16, H // 2, W // 2), dtype=img.dtype)
    for c in range(3):
        for r in range(2):
            for s in range(2):
                y[:, c * 4 + r * 2 + s] = xp[:, c, r::2, s::2]
    return y


def test_stem_reindexing_is_exact_in_float64():
    import relnet_b200            # noqa: F401
    from relnet_b200.trunk import Trunk
    torch.manual_seed(0)
    t = Trunk().double().eval()
    t.prepare()
    img = torch.randn(1, 3, 40, 56, dtype=torch.float64) * 50
    ref = F.conv2d(img, t.conv1.weight, t.conv1.bias, stride=2, padding=3)
    got = F.conv2d(_s2d(img), t.conv1_s2d_weight, t.conv1.bias)
    assert got.shape == ref.shape
    assert (got - ref).abs().max().item() <= 1e-12 * ref.abs().max().item()


def test_projection_bias_fold_keeps_the_block_function():
    import relnet_b200            # noqa: F401
    from relnet_b200.trunk import Bottleneck, Trunk
    torch.manual_seed(1)
    blk = Bottleneck(16, 8, 32, stride=2, project=True).double().eval()
    for p in blk.parameters():
        torch.nn.init.normal_(p, std=0.3)
    x = torch.randn(2, 16, 9, 11, dtype=torch.float64)
    with torch.no_grad():
        ref = blk(x)
        holder = Trunk.__new__(Trunk)
        torch.nn.Module.__init__(holder)
        holder.blk = blk
        holder.conv1 = torch.nn.Conv2d(3, 4, 7, stride=2, padding=3).double()
        holder.prepare()
        assert blk.proj.bias is None
        got = blk(x)
    assert (got - ref).abs().max().item() <= 1e-12 * ref.abs().max().item()
