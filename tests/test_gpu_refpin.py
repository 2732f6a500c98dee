"""GPU: pin the C restatements (oracle/oracle_c.c) AND the product kernels of the two `operator_cxx` ops against the
REFERENCE's own CUDA kernels, compiled from /root/reference into oracle/_ref/libref_deform.so (oracle/Makefile,
oracle/ref_deform.cu + oracle/ref_stub/), at the sizes the Deformable Faster-RCNN config runs them
(res5 deformable conv: 512 ch, 38 x 63, 4 deformable groups; PS-ROI pooling: R = 300, 256 ch), plus an independent
implementation of ROIPooling (torchvision) since MXNet's own roi_pooling.cu is not in the reference tree.

Tolerances: sample counts bit-exact; values 1e-5 (the reference library is compiled with nvcc's default FMA contraction, the
oracle and the product without it: same operation order, last-bit differences); atomicAdd backward 1e-5 of the tensor max.
"""
import numpy as np
import pytest
import torch
from conftest import rel_err
from oracle import rois_np as RO
from oracle import relation_np as R

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def ops(cuda_device):
    import __graft_entry__ as g
    g.build()
    import relnet_b200
    torch.cuda.set_device(cuda_device)
    return relnet_b200.ops


@pytest.fixture(scope='module')
def refd(cuda_device):
    if not RO.ref_deform_available():
        pytest.skip('oracle/_ref/libref_deform.so not built (needs /root/reference at build time)')
    return RO


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _conv_case(seed, C=512, H=38, W=63):
    rng = np.random.default_rng(seed)
    im = rng.standard_normal((C, H, W)).astype(np.float32)
    off = (rng.standard_normal((4 * 18, H, W)) * 2.0).astype(np.float32)         # several pixels: samples leave the image
    off[:, :2] *= 8.0
    return im, off


def test_deform_im2col_reference_vs_oracle_vs_product(ops, refd):
    im, off = _conv_case(0)
    ref = refd.ref_deform_im2col(T(im), T(off)).cpu().numpy()
    orc = RO.deform_im2col(im, off)
    ours = ops.deform_im2col(T(im), T(off)).cpu().numpy()
    print('deformable im2col 512x38x63: oracle_c vs reference kernel %.2e, product vs reference kernel %.2e'
          % (rel_err(orc, ref), rel_err(ours, ref)))
    np.testing.assert_allclose(orc, ref, rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(ours, ref, rtol=1e-5, atol=1e-5)
    assert np.array_equal(ref == 0, orc == 0), 'out-of-image samples differ'


def test_deform_col2im_and_coord_reference_vs_oracle_vs_product(ops, refd):
    C, H, W, Co = 128, 38, 63, 64                                                # backward at full spatial size, fewer channels
    rng = np.random.default_rng(3)
    data = rng.standard_normal((1, C, H, W)).astype(np.float32)
    off = (rng.standard_normal((1, 4 * 18, H, W)) * 2.0).astype(np.float32)
    wgt = (rng.standard_normal((Co, C, 3, 3)) * 0.05).astype(np.float32)
    dout = rng.standard_normal((1, Co, H, W)).astype(np.float32)
    col = np.ascontiguousarray((wgt.reshape(Co, -1).T @ dout[0].reshape(Co, -1)).reshape(C * 9, H, W), np.float32)
    g_im = refd.ref_deform_col2im(T(col), T(off[0]), (C, H, W)).cpu().numpy()
    g_off = refd.ref_deform_col2im_coord(T(col), T(data[0]), T(off[0])).cpu().numpy()
    dd_o, doff_o, _ = RO.deform_conv_backward(dout, data, off, wgt)
    print('col2im: oracle_c vs reference %.2e | col2im_coord: oracle_c vs reference %.2e'
          % (rel_err(dd_o[0], g_im), rel_err(doff_o[0], g_off)))
    assert rel_err(dd_o[0], g_im) < 1e-5 and rel_err(doff_o[0], g_off) < 1e-5
    dd, doff, dw, _ = ops.deform_conv_backward(T(dout), T(data), T(off), T(wgt))
    assert rel_err(dd.cpu().numpy()[0], g_im) < 2e-5 and rel_err(doff.cpu().numpy()[0], g_off) < 2e-5


def _psroi_case(seed, R_=300, C=256, H=38, W=63):
    rng = np.random.default_rng(seed)
    data = rng.standard_normal((1, C, H, W)).astype(np.float32)
    boxes = R.make_boxes(rng, R_)
    rois = np.hstack([np.zeros((R_, 1), np.float32), boxes]).astype(np.float32)
    rois[:4, 1:] = [[0, 0, 0, 0], [990, 590, 999, 599], [-20, -20, 5, 5], [500, 300, 500.4, 300.4]]
    trans = rng.standard_normal((R_, 2, 7, 7)).astype(np.float32)
    return data, rois, trans


@pytest.mark.parametrize('with_trans', [False, True])
def test_deform_psroi_reference_vs_oracle_vs_product(ops, refd, with_trans):
    data, rois, trans = _psroi_case(1)
    tr = trans if with_trans else None
    kw = dict(output_dim=256, trans_std=0.1 if with_trans else 0.0)
    o_ref, c_ref = refd.ref_deform_psroi_pool(T(data), T(rois), T(tr) if with_trans else None, **kw)
    o_orc, c_orc = RO.deform_psroi_pool(data, rois, tr, **kw)
    o_our, c_our = ops.deform_psroi_pool(T(data), T(rois), T(tr) if with_trans else None, return_count=True, **kw)
    o_ref, c_ref = o_ref.cpu().numpy(), c_ref.cpu().numpy()
    print('PS-ROI pool R=300 C=256 trans=%s: oracle_c vs reference %.2e, product vs reference %.2e'
          % (with_trans, rel_err(o_orc, o_ref), rel_err(o_our.cpu().numpy(), o_ref)))
    np.testing.assert_array_equal(c_orc, c_ref)
    np.testing.assert_array_equal(c_our.cpu().numpy(), c_ref)
    np.testing.assert_allclose(o_orc, o_ref, rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(o_our.cpu().numpy(), o_ref, rtol=1e-5, atol=1e-5)
    # the channels-last forms (fp32, and bf16 = the trunk's layout): same table, vector taps
    d_cl = T(data).contiguous(memory_format=torch.channels_last)
    o_cl, c_cl = ops.deform_psroi_pool(d_cl, T(rois), T(tr) if with_trans else None, return_count=True, **kw)
    np.testing.assert_array_equal(c_cl.cpu().numpy(), c_ref)
    np.testing.assert_allclose(o_cl.cpu().numpy(), o_ref, rtol=1e-5, atol=1e-5)
    d_bf = d_cl.to(torch.bfloat16)
    o_bf = ops.deform_psroi_pool(d_bf, T(rois), T(tr) if with_trans else None, **kw)
    o_ref_bf, _ = refd.ref_deform_psroi_pool(d_bf.float().contiguous(), T(rois), T(tr) if with_trans else None, **kw)
    np.testing.assert_allclose(o_bf.cpu().numpy(), o_ref_bf.cpu().numpy(), rtol=1e-5, atol=1e-5)
    # backward of the same op (atomicAdd: summation order differs run to run)
    dout = np.random.default_rng(9).standard_normal(o_ref.shape).astype(np.float32)
    dd_ref, dt_ref = refd.ref_deform_psroi_pool_backward(T(dout), T(c_ref), T(data), T(rois), T(tr) if with_trans else None, **kw)
    dd_orc, dt_orc = RO.deform_psroi_pool_backward(dout, c_ref, data, rois, tr, **kw)
    dd_our, dt_our = ops.deform_psroi_pool_backward(T(dout), T(c_ref), T(data), T(rois), T(tr) if with_trans else None, **kw)
    assert rel_err(dd_orc, dd_ref.cpu().numpy()) < 1e-5 and rel_err(dd_our.cpu().numpy(), dd_ref.cpu().numpy()) < 1e-5
    if with_trans:
        assert rel_err(dt_orc, dt_ref.cpu().numpy()) < 1e-4 and rel_err(dt_our.cpu().numpy(), dt_ref.cpu().numpy()) < 1e-4


def test_roi_pool_vs_independent_implementation(ops):
    """MXNet's ROIPooling source is not in the reference tree; torchvision.ops.roi_pool is an independent implementation of
    the same Fast R-CNN definition (round(x * scale), max(end - start + 1, 1), floor/ceil bin edges, empty bin -> 0)."""
    tv = pytest.importorskip('torchvision.ops')
    rng = np.random.default_rng(4)
    data = rng.standard_normal((1, 256, 38, 63)).astype(np.float32)
    boxes = R.make_boxes(rng, 300)
    rois = np.hstack([np.zeros((300, 1), np.float32), boxes]).astype(np.float32)
    rois[:3, 1:] = [[0, 0, 0, 0], [990, 590, 999, 599], [500, 300, 500.4, 300.4]]
    ref = tv.roi_pool(T(data), T(rois), (7, 7), 1.0 / 16).cpu().numpy()
    orc, _ = RO.roi_pool(data, rois)
    ours = ops.roi_pool(T(data), T(rois)).cpu().numpy()
    assert np.array_equal(orc, ref), 'oracle_c ROIPooling differs from torchvision roi_pool'
    assert np.array_equal(ours, ref)


def test_deform_conv_channels_last_fast_path(ops):
    """rn_deform_conv_nhwc_fwd (bf16 channels_last in, fp16 column buffer, tcgen05 GEMM, bias + relu fused) at the res5 size
    against the float32 C oracle evaluated on the same bf16-rounded data: 2e-3 of the tensor max (fp16 operands)."""
    if not ops.device_info()['sm100']:
        pytest.skip('tcgen05 needs sm_100')
    rng = np.random.default_rng(8)
    C, H, W, Co = 512, 38, 63, 512
    data = rng.standard_normal((1, C, H, W)).astype(np.float32)
    off = (rng.standard_normal((1, 72, H, W)) * 2.0).astype(np.float32)
    wgt = (rng.standard_normal((Co, C, 3, 3)) * 0.02).astype(np.float32)
    bias = rng.standard_normal(Co).astype(np.float32)
    d_bf = T(data).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    ref = RO.deform_conv(d_bf.float().cpu().numpy(), off, wgt, bias)
    out = ops.deform_conv_nhwc(d_bf, T(off), T(wgt), T(bias), relu=False, out_dtype=torch.float32)
    assert out.is_contiguous(memory_format=torch.channels_last)
    e = rel_err(out.cpu().numpy(), ref)
    print('deformable conv 512->512 @38x63, channels_last bf16 in: rel err vs oracle %.2e' % e)
    assert e < 2e-3
    out16 = ops.deform_conv_nhwc(d_bf, T(off), T(wgt), T(bias), relu=True)
    assert rel_err(out16.float().cpu().numpy(), np.maximum(ref, 0)) < 3e-3
