"""Differentiable torch (float64) restatements of the two sampling operators -- TEST INFRASTRUCTURE.

They exist to check the BACKWARD restatements in oracle_c.c against autograd: the forward here must equal the C oracle's
forward (itself following deformable_im2col.cuh:77-113,216-262 and deformable_psroi_pooling.cu:29-138), and bilinear
sampling is differentiable almost everywhere, so autograd through it gives the gradients the reference's hand-written
backward kernels (deformable_im2col.cuh:116-207,315-458; deformable_psroi_pooling.cu:177-289) must reproduce away from
integer sample positions."""
import torch


def deform_conv(data, offset, weight, kernel=(3, 3), pad=(2, 2), stride=(1, 1), dilate=(2, 2), num_deformable_group=4):
    """data [B,C,H,W], offset [B,dg*2*kh*kw,Ho,Wo], weight [Co,C,kh,kw] (num_group = 1) -> [B,Co,Ho,Wo]"""
    B, C, H, W = data.shape
    kh, kw = kernel
    Ho = (H + 2 * pad[0] - (dilate[0] * (kh - 1) + 1)) // stride[0] + 1
    Wo = (W + 2 * pad[1] - (dilate[1] * (kw - 1) + 1)) // stride[1] + 1
    dg = num_deformable_group
    f = data.dtype
    h_in = (torch.arange(Ho, dtype=f) * stride[0] - pad[0]).view(1, 1, 1, Ho, 1)
    w_in = (torch.arange(Wo, dtype=f) * stride[1] - pad[1]).view(1, 1, 1, 1, Wo)
    ki = (torch.arange(kh, dtype=f) * dilate[0]).repeat_interleave(kw).view(1, 1, kh * kw, 1, 1)
    kj = (torch.arange(kw, dtype=f) * dilate[1]).repeat(kh).view(1, 1, kh * kw, 1, 1)
    off = offset.view(B, dg, kh * kw, 2, Ho, Wo)
    h = h_in + ki + off[:, :, :, 0]                       # [B,dg,K,Ho,Wo]
    w = w_in + kj + off[:, :, :, 1]
    valid = (h >= 0) & (w >= 0) & (h < H) & (w < W)
    hl = torch.floor(h.detach()); wl = torch.floor(w.detach())
    eh = hl >= H - 1; ew = wl >= W - 1
    hl = torch.where(eh, torch.full_like(hl, H - 1), hl); wl = torch.where(ew, torch.full_like(wl, W - 1), wl)
    hh = torch.where(eh, hl, hl + 1); wh = torch.where(ew, wl, wl + 1)
    lh = torch.where(eh, torch.zeros_like(h), h - hl); lw = torch.where(ew, torch.zeros_like(w), w - wl)
    cl = lambda t, m: t.clamp(0, m - 1).long()
    hl_, hh_, wl_, wh_ = cl(hl, H), cl(hh, H), cl(wl, W), cl(wh, W)
    cpg = C // dg
    d = data.view(B, dg, cpg, H * W)

    def tap(hi, wi):
        idx = (hi * W + wi).view(B, dg, 1, -1).expand(B, dg, cpg, -1)
        return torch.gather(d, 3, idx).view(B, dg, cpg, kh * kw, Ho, Wo)
    um = lambda t: t.unsqueeze(2)
    val = (um((1 - lh) * (1 - lw)) * tap(hl_, wl_) + um((1 - lh) * lw) * tap(hl_, wh_)
           + um(lh * (1 - lw)) * tap(hh_, wl_) + um(lh * lw) * tap(hh_, wh_)) * um(valid.to(f))
    col = val.reshape(B, C * kh * kw, Ho * Wo)
    out = torch.matmul(weight.reshape(weight.shape[0], -1), col)
    return out.view(B, -1, Ho, Wo)


def deform_psroi_pool(data, rois, trans=None, spatial_scale=0.0625, output_dim=256, group_size=1, pooled_size=7,
                      part_size=0, sample_per_part=4, trans_std=0.0):
    """data [B,C,H,W], rois [R,5], trans [R,2*ncls,part,part] or None -> out [R,output_dim,P,P], count"""
    f = data.dtype
    B, C, H, W = data.shape
    R = rois.shape[0]
    P = pooled_size; part = part_size or P; S = sample_per_part
    no_trans = trans is None
    ncls = 1 if no_trans else trans.shape[1] // 2
    cec = output_dim if no_trans else output_dim // ncls
    rnd = lambda x: torch.sign(x) * torch.floor(torch.abs(x) + 0.5)          # C round(): half away from zero
    rsw = rnd(rois[:, 1]) * spatial_scale - 0.5; rsh = rnd(rois[:, 2]) * spatial_scale - 0.5
    rew = (rnd(rois[:, 3]) + 1.) * spatial_scale - 0.5; reh = (rnd(rois[:, 4]) + 1.) * spatial_scale - 0.5
    roi_w = (rew - rsw).clamp(min=0.1); roi_h = (reh - rsh).clamp(min=0.1)
    bin_w = roi_w / P; bin_h = roi_h / P
    sub_w = bin_w / S; sub_h = bin_h / S
    pidx = torch.arange(P, dtype=f)
    part_i = torch.floor(pidx / P * part).long()                                  # [P]
    g_i = torch.floor(pidx * group_size / P).clamp(0, group_size - 1).long()
    ctop = torch.arange(output_dim)
    cls_id = ctop // cec
    if no_trans:
        tx = torch.zeros(R, output_dim, P, P, dtype=f); ty = tx
    else:
        t = trans.view(R, ncls, 2, part, part)[:, cls_id]                         # [R,O,2,part,part]
        t = t[:, :, :, part_i][:, :, :, :, part_i]                                # [R,O,2,P,P]
        tx = t[:, :, 0] * trans_std; ty = t[:, :, 1] * trans_std
    v = lambda x: x.view(R, 1, 1, 1)
    wstart = pidx.view(1, 1, 1, P) * v(bin_w) + v(rsw) + tx * v(roi_w)              # [R,O,P,P]
    hstart = pidx.view(1, 1, P, 1) * v(bin_h) + v(rsh) + ty * v(roi_h)
    s = torch.arange(S, dtype=f)
    w = wstart[..., None, None] + s.view(1, 1, 1, 1, 1, S) * sub_w.view(R, 1, 1, 1, 1, 1)       # [R,O,P,P,S(ih),S(iw)]
    h = hstart[..., None, None] + s.view(1, 1, 1, 1, S, 1) * sub_h.view(R, 1, 1, 1, 1, 1)
    w, h = torch.broadcast_tensors(w, h)
    ok = ~((w < -0.5) | (w > W - 0.5) | (h < -0.5) | (h > H - 0.5))
    w = w.clamp(0., W - 1.); h = h.clamp(0., H - 1.)
    x0 = torch.floor(w.detach()); x1 = torch.ceil(w.detach()); y0 = torch.floor(h.detach()); y1 = torch.ceil(h.detach())
    dx = w - x0; dy = h - y0
    c = ((ctop.view(1, -1, 1, 1) * group_size + g_i.view(1, 1, P, 1)) * group_size + g_i.view(1, 1, 1, P))   # [1,O,P,P]
    b = rois[:, 0].long().view(R, 1, 1, 1)
    base = ((b * C + c) * H * W)[..., None, None]
    flat = data.reshape(-1)
    tap = lambda yy, xx: flat[(base + yy.long() * W + xx.long())]
    val = (1 - dx) * (1 - dy) * tap(y0, x0) + (1 - dx) * dy * tap(y1, x0) + dx * (1 - dy) * tap(y0, x1) + dx * dy * tap(y1, x1)
    cnt = ok.to(f).sum(dim=(4, 5))
    out = (val * ok.to(f)).sum(dim=(4, 5)) / cnt.clamp(min=1.)
    return out, cnt
