"""ctypes front-end to oracle_c.c (TEST INFRASTRUCTURE): ROIPooling(max), DeformablePSROIPooling, deformable conv,
NMS sweep, and -- on the GPU box -- the reference's own compiled nms_kernel.cu (oracle/_ref/libref_gpu_nms.so)."""
import ctypes
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def build(quiet=True):
    subprocess.check_call(['make', '-C', HERE] + (['-s'] if quiet else []))


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(HERE, '_build', 'liboracle_c.so')
        if not os.path.exists(path):
            build()
        _lib = ctypes.CDLL(path)
    return _lib


def _p(a, t=ctypes.c_float):
    return a.ctypes.data_as(ctypes.POINTER(t))


def roi_pool(data, rois, pooled=(7, 7), spatial_scale=0.0625):
    data = np.ascontiguousarray(data, np.float32); rois = np.ascontiguousarray(rois, np.float32)
    B, C, H, W = data.shape; R = rois.shape[0]
    out = np.empty((R, C, pooled[0], pooled[1]), np.float32)
    arg = np.empty(out.shape, np.int32)
    lib().oracle_roi_pool(_p(data), _p(rois), R, C, H, W, pooled[0], pooled[1], ctypes.c_float(spatial_scale),
                          _p(out), _p(arg, ctypes.c_int))
    return out, arg


def deform_psroi_pool(data, rois, trans=None, spatial_scale=0.0625, output_dim=256, group_size=1, pooled_size=7,
                      part_size=0, sample_per_part=4, trans_std=0.0, no_trans=None):
    data = np.ascontiguousarray(data, np.float32); rois = np.ascontiguousarray(rois, np.float32)
    if no_trans is None:
        no_trans = trans is None
    part = part_size or pooled_size
    B, C, H, W = data.shape; R = rois.shape[0]
    ncls = 1 if no_trans else trans.shape[1] // 2
    t = np.zeros(1, np.float32) if no_trans else np.ascontiguousarray(trans, np.float32)
    out = np.empty((R, output_dim, pooled_size, pooled_size), np.float32)
    cnt = np.empty_like(out)
    lib().oracle_deform_psroi_pool(_p(data), _p(rois), _p(t), R, C, H, W, int(bool(no_trans)),
                                   ctypes.c_float(spatial_scale), output_dim, group_size, pooled_size, part,
                                   sample_per_part, ctypes.c_float(trans_std), ncls, _p(out), _p(cnt))
    return out, cnt


def deform_im2col(im, offset, kernel=(3, 3), pad=(2, 2), stride=(1, 1), dilate=(2, 2), num_deformable_group=4):
    im = np.ascontiguousarray(im, np.float32); offset = np.ascontiguousarray(offset, np.float32)
    C, H, W = im.shape
    kh, kw = kernel
    Ho = (H + 2 * pad[0] - (dilate[0] * (kh - 1) + 1)) // stride[0] + 1
    Wo = (W + 2 * pad[1] - (dilate[1] * (kw - 1) + 1)) // stride[1] + 1
    assert offset.shape == (num_deformable_group * 2 * kh * kw, Ho, Wo), offset.shape
    col = np.empty((C * kh * kw, Ho, Wo), np.float32)
    lib().oracle_deform_im2col(_p(im), _p(offset), C, H, W, kh, kw, pad[0], pad[1], stride[0], stride[1],
                               dilate[0], dilate[1], num_deformable_group, Ho, Wo, _p(col))
    return col


def deform_conv(data, offset, weight, bias=None, kernel=(3, 3), pad=(2, 2), stride=(1, 1), dilate=(2, 2),
                num_deformable_group=4, num_group=1):
    """DeformableConvolutionOp::Forward (deformable_convolution-inl.h:91-144): im2col then W[g].col[g] per group."""
    data = np.asarray(data, np.float32)
    outs = []
    for n in range(data.shape[0]):
        col = deform_im2col(data[n], offset[n], kernel, pad, stride, dilate, num_deformable_group)
        K, Ho, Wo = col.shape
        Co = weight.shape[0]
        w2 = np.asarray(weight, np.float32).reshape(num_group, Co // num_group, -1)
        c2 = col.reshape(num_group, K // num_group, Ho * Wo)
        o = np.matmul(w2, c2).reshape(Co, Ho, Wo)
        if bias is not None:
            o = o + np.asarray(bias, np.float32).reshape(-1, 1, 1)
        outs.append(o)
    return np.stack(outs).astype(np.float32)


def nms_sorted(boxes, thresh):
    boxes = np.ascontiguousarray(boxes, np.float32)
    keep = np.empty(boxes.shape[0], np.int32)
    n = lib().oracle_nms_sorted(_p(boxes), boxes.shape[0], boxes.shape[1], ctypes.c_float(thresh),
                                _p(keep, ctypes.c_int))
    return keep[:n].astype(np.int64)


# ---------------------------------------------------------------------------------------------------------
_ref = None


def ref_gpu_nms_available():
    return os.path.exists(os.path.join(HERE, '_ref', 'libref_gpu_nms.so'))


def ref_gpu_nms(sorted_dets, thresh, device_id=0):
    """Call the REFERENCE's own `_nms` (lib/nms/nms_kernel.cu:91-144, compiled into oracle/_ref) on host data.
    sorted_dets [n,5] float32 sorted by score.  Needs a GPU."""
    global _ref
    if _ref is None:
        _ref = ctypes.CDLL(os.path.join(HERE, '_ref', 'libref_gpu_nms.so'))
    fn = getattr(_ref, '_Z4_nmsPiS_PKfiifi')      # void _nms(int*, int*, const float*, int, int, float, int)
    d = np.ascontiguousarray(sorted_dets, np.float32)
    keep = np.zeros(d.shape[0], np.int32)
    num = ctypes.c_int(0)
    fn(_p(keep, ctypes.c_int), ctypes.byref(num), _p(d), d.shape[0], d.shape[1], ctypes.c_float(thresh), device_id)
    return keep[:num.value].astype(np.int64)


# ---------------------------------------------------------------------------------------------------------
# backward (training) restatements
def roi_pool_backward(dout, argmax, rois, data_shape):
    dout = np.ascontiguousarray(dout, np.float32); argmax = np.ascontiguousarray(argmax, np.int32)
    rois = np.ascontiguousarray(rois, np.float32)
    B, C, H, W = data_shape
    R, _, PH, PW = dout.shape
    dd = np.empty(data_shape, np.float32)
    lib().oracle_roi_pool_bwd(_p(dout), _p(argmax, ctypes.c_int), _p(rois), R, B, C, H, W, PH, PW, _p(dd))
    return dd


def deform_psroi_pool_backward(dout, top_count, data, rois, trans=None, spatial_scale=0.0625, output_dim=256,
                               group_size=1, pooled_size=7, part_size=0, sample_per_part=4, trans_std=0.0, no_trans=None):
    data = np.ascontiguousarray(data, np.float32); rois = np.ascontiguousarray(rois, np.float32)
    dout = np.ascontiguousarray(dout, np.float32); top_count = np.ascontiguousarray(top_count, np.float32)
    if no_trans is None:
        no_trans = trans is None
    part = part_size or pooled_size
    B, C, H, W = data.shape; R = rois.shape[0]
    ncls = 1 if no_trans else trans.shape[1] // 2
    t = np.zeros(1, np.float32) if no_trans else np.ascontiguousarray(trans, np.float32)
    dd = np.empty_like(data)
    dt = np.zeros(1, np.float32) if no_trans else np.empty_like(t)
    lib().oracle_deform_psroi_pool_bwd(_p(dout), _p(top_count), _p(data), _p(rois), _p(t), R, B, C, H, W,
                                       int(bool(no_trans)), ctypes.c_float(spatial_scale), output_dim, group_size,
                                       pooled_size, part, sample_per_part, ctypes.c_float(trans_std), ncls, _p(dd), _p(dt))
    return dd, (None if no_trans else dt)


def _conv_geom(shape, kernel, pad, stride, dilate):
    C, H, W = shape
    kh, kw = kernel
    Ho = (H + 2 * pad[0] - (dilate[0] * (kh - 1) + 1)) // stride[0] + 1
    Wo = (W + 2 * pad[1] - (dilate[1] * (kw - 1) + 1)) // stride[1] + 1
    return C, H, W, kh, kw, Ho, Wo


def deform_conv_backward(dout, data, offset, weight, kernel=(3, 3), pad=(2, 2), stride=(1, 1), dilate=(2, 2),
                         num_deformable_group=4, num_group=1, weight_grad_deformed=False, has_bias=False):
    """DeformableConvolutionOp::Backward (deformable_convolution-inl.h:145-233).  Returns ddata, doffset, dweight[, dbias].
    weight_grad_deformed=False is the REFERENCE: its dWeight uses the plain im2col of the data (:215), not the deformed
    sampling; True gives the mathematically consistent gradient (later MXNet releases)."""
    data = np.ascontiguousarray(data, np.float32); offset = np.ascontiguousarray(offset, np.float32)
    dout = np.ascontiguousarray(dout, np.float32); weight = np.asarray(weight, np.float32)
    B = data.shape[0]
    C, H, W, kh, kw, Ho, Wo = _conv_geom(data.shape[1:], kernel, pad, stride, dilate)
    Co = weight.shape[0]
    G = num_group
    w3 = weight.reshape(G, Co // G, -1)
    ddata = np.zeros_like(data); doff = np.empty_like(offset); dw = np.zeros_like(w3)
    geo = (C, H, W, kh, kw, pad[0], pad[1], stride[0], stride[1], dilate[0], dilate[1])
    for n in range(B):
        og = dout[n].reshape(G, Co // G, Ho * Wo)
        col = np.ascontiguousarray(np.matmul(w3.transpose(0, 2, 1), og).reshape(C * kh * kw, Ho, Wo), np.float32)
        lib().oracle_deform_col2im_coord(_p(col), _p(data[n]), _p(offset[n]), *geo, num_deformable_group, Ho, Wo, _p(doff[n]))
        lib().oracle_deform_col2im(_p(col), _p(offset[n]), *geo, num_deformable_group, Ho, Wo, _p(ddata[n]))
        if weight_grad_deformed:
            c2 = deform_im2col(data[n], offset[n], kernel, pad, stride, dilate, num_deformable_group)
        else:
            c2 = np.empty((C * kh * kw, Ho, Wo), np.float32)
            lib().oracle_im2col(_p(data[n]), *geo, Ho, Wo, _p(c2))
        dw += np.matmul(og, c2.reshape(G, -1, Ho * Wo).transpose(0, 2, 1))
    res = [ddata, doff, dw.reshape(weight.shape).astype(np.float32)]
    if has_bias:
        res.append(dout.sum(axis=(0, 2, 3)).astype(np.float32))
    return res


# ---------------------------------------------------------------------------------------------------------
# The REFERENCE's own deformable kernels (oracle/_ref/libref_deform.so, built by the Makefile from
# /root/reference/relation_rcnn/operator_cxx/{nn/deformable_im2col.cuh, deformable_psroi_pooling.cu} + ref_deform.cu).
# They take DEVICE pointers: arguments are torch CUDA tensors (float32, contiguous).  GPU box only.
_ref_deform = None


def ref_deform_available():
    return os.path.exists(os.path.join(HERE, '_ref', 'libref_deform.so'))


def _refd():
    global _ref_deform
    if _ref_deform is None:
        _ref_deform = ctypes.CDLL(os.path.join(HERE, '_ref', 'libref_deform.so'))
    return _ref_deform


def _dp(t):
    return ctypes.c_void_p(t.data_ptr())


def ref_deform_im2col(im, offset, kernel=(3, 3), pad=(2, 2), stride=(1, 1), dilate=(2, 2), num_deformable_group=4):
    """deformable_im2col (deformable_im2col.cuh:283-309) on im [C,H,W], offset [dg*2*kh*kw,Ho,Wo] -> col [C*kh*kw,Ho,Wo]."""
    import torch
    C, H, W = im.shape
    col = torch.empty((C * kernel[0] * kernel[1], offset.shape[1], offset.shape[2]), dtype=torch.float32, device=im.device)
    rc = _refd().ref_deformable_im2col(_dp(im), _dp(offset), C, H, W, kernel[0], kernel[1], pad[0], pad[1], stride[0],
                                       stride[1], dilate[0], dilate[1], num_deformable_group, _dp(col))
    assert rc == 0, rc
    return col


def ref_deform_col2im(col, offset, im_shape, kernel=(3, 3), pad=(2, 2), stride=(1, 1), dilate=(2, 2), num_deformable_group=4):
    """deformable_col2im (:381-411), req = kWriteTo semantics of the op (grad buffer zeroed by the caller: done here)."""
    import torch
    C, H, W = im_shape
    g = torch.zeros((C, H, W), dtype=torch.float32, device=col.device)
    rc = _refd().ref_deformable_col2im(_dp(col), _dp(offset), C, H, W, kernel[0], kernel[1], pad[0], pad[1], stride[0],
                                       stride[1], dilate[0], dilate[1], num_deformable_group, _dp(g))
    assert rc == 0, rc
    return g


def ref_deform_col2im_coord(col, im, offset, kernel=(3, 3), pad=(2, 2), stride=(1, 1), dilate=(2, 2), num_deformable_group=4):
    """deformable_col2im_coord (:490-519) -> grad_offset, shaped like offset."""
    import torch
    C, H, W = im.shape
    g = torch.zeros_like(offset)
    rc = _refd().ref_deformable_col2im_coord(_dp(col), _dp(im), _dp(offset), C, H, W, kernel[0], kernel[1], pad[0], pad[1],
                                             stride[0], stride[1], dilate[0], dilate[1], num_deformable_group, _dp(g))
    assert rc == 0, rc
    return g


def ref_deform_psroi_pool(data, rois, trans=None, spatial_scale=0.0625, output_dim=256, group_size=1, pooled_size=7,
                          part_size=0, sample_per_part=4, trans_std=0.0):
    """DeformablePSROIPoolingOp::Forward (-inl.h:64-95 + .cu:52-175): returns (out, top_count)."""
    import torch
    _, C, H, W = data.shape
    R = rois.shape[0]
    no_trans = trans is None
    t = torch.zeros(1, device=data.device) if no_trans else trans
    ncls = 1 if no_trans else trans.shape[1] // 2
    out = torch.empty((R, output_dim, pooled_size, pooled_size), dtype=torch.float32, device=data.device)
    cnt = torch.empty_like(out)
    rc = _refd().ref_deform_psroi_forward(_dp(data), _dp(rois), _dp(t), R, C, H, W, int(no_trans),
                                          ctypes.c_float(spatial_scale), output_dim, group_size, pooled_size, part_size,
                                          sample_per_part, ctypes.c_float(trans_std), ncls, _dp(out), _dp(cnt))
    assert rc == 0, rc
    return out, cnt


def ref_deform_psroi_pool_backward(dout, top_count, data, rois, trans=None, spatial_scale=0.0625, output_dim=256,
                                   group_size=1, pooled_size=7, part_size=0, sample_per_part=4, trans_std=0.0):
    """DeformablePSROIPoolingOp::Backward (-inl.h:97-151 + .cu:177-345): returns (ddata, dtrans or None)."""
    import torch
    _, C, H, W = data.shape
    R = rois.shape[0]
    no_trans = trans is None
    t = torch.zeros(1, device=data.device) if no_trans else trans
    ncls = 1 if no_trans else trans.shape[1] // 2
    dd = torch.empty_like(data)
    dt = torch.zeros(1, device=data.device) if no_trans else torch.empty_like(trans)
    rc = _refd().ref_deform_psroi_backward(_dp(dout), _dp(data), _dp(rois), _dp(t), _dp(top_count), R, C, H, W, int(no_trans),
                                           ctypes.c_float(spatial_scale), output_dim, group_size, pooled_size, part_size,
                                           sample_per_part, ctypes.c_float(trans_std), ncls, _dp(dd), _dp(dt))
    assert rc == 0, rc
    return dd, (None if no_trans else dt)
