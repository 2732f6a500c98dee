"""Host-side wrappers: torch CUDA tensors in, torch CUDA tensors out, every byte of arithmetic in librelnet_b200.so.

PyTorch is used for device memory, streams and (elsewhere) torch.distributed only.  No op here has a CPU or eager
fallback: non-CUDA tensors raise.
"""
import ctypes as C
c_i32 = C.c_int32
import torch
from . import _lib as L

PREC = {'fp32': 0, 'f16': 1, 'tf32': 2, 0: 0, 1: 1, 2: 2}
_ws = {}


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _workspace(nbytes, device):
    """Grow-only per-device scratch buffer handed to the C ABI (the library itself never allocates)."""
    key = (device.index, torch.cuda.current_stream().cuda_stream)
    buf = _ws.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(int(nbytes * 1.25) + 4096, dtype=torch.uint8, device=device)
        _ws[key] = buf
    return buf


def _f32(t, name):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise L.RelnetError('%s must be a CUDA tensor (relnet_b200 has no CPU path)' % name)
    if t.dtype != torch.float32 or not t.is_contiguous():
        t = t.contiguous().float()
    return t


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class _PackCache(object):
    """fp16 operand cache for RN_PREC_F16.

    Identity of an entry = (tag, data_ptr + shape of every source tensor); `tag` carries the descriptor fields that shape
    the packed layout (op, group, dv ...), so the same weights used under another descriptor get their own block.  Validity =
    the sources' in-place version counters: an optimizer step that bumps `_version` REPLACES the entry for that identity
    (no accumulation of dead packed buffers during training).  Updates that do not bump `_version` (`p.data.copy_()`,
    raw-pointer writes, checkpoint loaders writing through `.data`) must call `invalidate_packed()` afterwards."""

    MAX = 256

    def __init__(self):
        self.d = {}

    def get_tagged(self, tag, tensors, nbytes, pack_fn):
        return self.get(tensors, nbytes, pack_fn, tag)

    def get(self, tensors, nbytes, pack_fn, tag=None):
        ident = (tag,) + tuple((t.data_ptr(), tuple(t.shape)) for t in tensors)
        ver = tuple(t._version for t in tensors)
        hit = self.d.get(ident)
        if hit is not None and hit[0] == ver:
            return hit[1]
        if hit is None and len(self.d) >= self.MAX:
            self.d.pop(next(iter(self.d)))          # oldest identity out (dict keeps insertion order)
        buf = hit[1] if hit is not None and hit[1].numel() == nbytes else \
            torch.empty(nbytes, dtype=torch.uint8, device=tensors[0].device)
        pack_fn(buf)                                # same stream as every consumer: repacking in place is ordered
        self.d[ident] = (ver, buf, tensors)         # keep the sources alive so data_ptr stays unique
        return buf

    def invalidate(self, tensors=None):
        """forget every packed block (tensors=None) or the ones built from any of `tensors`"""
        if tensors is None:
            self.d.clear()
            return
        ptrs = set(t.data_ptr() for t in tensors)
        for k in [k for k in self.d if any(e[0] in ptrs for e in k[1:])]:
            del self.d[k]


_packs = _PackCache()


def invalidate_packed(tensors=None):
    """Call after weight updates that bypass torch's version counter (`.data` writes, raw-pointer kernels, checkpoint
    loads): drops the packed fp16 blocks of `tensors` (or all of them) so the next f16 call repacks."""
    _packs.invalidate(tensors)


def device_info():
    sm, ma, mi = C.c_int(), C.c_int(), C.c_int()
    ok = L.lib().rn_device_info(C.byref(sm), C.byref(ma), C.byref(mi))
    return dict(sm100=bool(ok), sm_count=sm.value, cc=(ma.value, mi.value))


def default_precision():
    return 'f16' if device_info()['sm100'] else 'fp32'


# ------------------------------------------------------------------------------------------------------------------
def relation(X, boxes, Wq, bq, Wk, bk, Wg, bg, Wout, bout, key_index=None, M=None, group=16, residual_relu=False,
             precision=None, wave_length=1000.0, return_softmax=False, stage_mask=7, workspace=None, out=None,
             x_f16=None, want_f16=False, allow_fp32_fallback=False):
    """Object-relation module (SYM_REL:30-151 + :267-268).  X [N,d] or [B,N,d]; boxes [N,4] or [B,N,4].
    RN_PREC_F16 only: x_f16 = the producer's fp16 copy of X (skips the cast launch); want_f16 -> returns (out, out_f16).
    A request the tcgen05 kernels do not cover (d_k or d_v > 64 per head, or the materialised softmax) RAISES under
    precision='f16' unless allow_fp32_fallback=True, in which case the library's general fp32 kernels run instead."""
    precision = precision or default_precision()
    X = _f32(X, 'X'); boxes = _f32(boxes, 'boxes')
    batched = X.dim() == 3
    B = X.shape[0] if batched else 1
    N, d = X.shape[-2], X.shape[-1]
    ws_ = [_f32(w, n) for w, n in ((Wq, 'Wq'), (bq, 'bq'), (Wk, 'Wk'), (bk, 'bk'), (Wg, 'Wg'), (bg, 'bg'),
                                   (Wout, 'Wout'), (bout, 'bout'))]
    Wq, bq, Wk, bk, Wg, bg, Wout, bout = ws_
    Wout2 = Wout.reshape(Wout.shape[0], -1)
    dq, dout = Wq.shape[0], Wout2.shape[0]
    if not (boxes.shape[-1] == 4 and boxes.shape[:-1] == X.shape[:-1] and tuple(Wq.shape) == (dq, d)
            and tuple(Wk.shape) == (dq, d) and Wout2.shape[1] == d and bq.numel() == dq and bk.numel() == dq
            and bout.numel() == dout and Wg.shape[0] == group and bg.numel() == group and dq % group == 0
            and dout % group == 0):
        raise L.RelnetError('relation: inconsistent shapes X%s boxes%s Wq%s Wk%s Wg%s Wout%s group=%d' % (
            tuple(X.shape), tuple(boxes.shape), tuple(Wq.shape), tuple(Wk.shape), tuple(Wg.shape), tuple(Wout.shape), group))
    assert Wq.shape[0] == Wk.shape[0], 'Matrix multiply requires same dimensions!'         # SYM_REL:119
    kidx = None
    if key_index is not None:
        kidx = key_index.to(device=X.device, dtype=torch.int32).contiguous()
        M = kidx.numel()
    M = int(M) if M is not None else N
    if precision == 'f16' and not relation_tc_supported(Wq.shape[0], Wout2.shape[0], group, return_softmax):
        if not allow_fp32_fallback:
            raise L.RelnetError('relation: precision=f16 does not cover dq/H=%d dout/H=%d return_softmax=%s (tcgen05 kernels: '
                                'd_k, d_v <= 64, no materialised softmax); pass precision="fp32" or allow_fp32_fallback=True'
                                % (Wq.shape[0] // group, Wout2.shape[0] // group, return_softmax))
        precision = 'fp32'      # the library's general fp32 kernels, asked for explicitly
    desc = L.RelationDesc(B, N, M, d, Wq.shape[0], Wout2.shape[0], group, Wg.shape[1], wave_length,
                          int(residual_relu), PREC[precision])
    if out is None:
        out = torch.empty((B, N, Wout2.shape[0]) if batched else (N, Wout2.shape[0]), dtype=torch.float32, device=X.device)
    sm = torch.empty((B, N, group, M) if batched else (N, group, M), dtype=torch.float32, device=X.device) \
        if return_softmax else None
    lib = L.lib()
    nbytes = lib.rn_relation_workspace_bytes(C.byref(desc))
    if workspace is not None:          # caller-owned scratch: lets stages of one module run on different streams
        if workspace.numel() < nbytes:
            raise L.RelnetError('relation: workspace of %d bytes < %d needed' % (workspace.numel(), nbytes))
        ws = workspace
    else:
        ws = _workspace(nbytes, X.device)
    if precision == 'f16':
        def pack(buf):
            L.check(lib.rn_relation_pack(C.byref(desc), _ptr(Wq), _ptr(bq), _ptr(Wk), _ptr(bk), _ptr(Wout2), _ptr(bout),
                                         _ptr(buf), _stream()), 'rn_relation_pack')
        packed = _packs.get((Wq, bq, Wk, bk, Wout2, bout), lib.rn_relation_packed_bytes(C.byref(desc)), pack,
                             tag=('relation', group, dq, dout))
        if x_f16 is not None or want_f16:
            if x_f16 is not None and not (x_f16.is_cuda and x_f16.dtype == torch.float16 and x_f16.is_contiguous()
                                          and x_f16.numel() == X.numel()):
                raise L.RelnetError('relation: x_f16 must be the contiguous fp16 CUDA copy of X')
            out16 = torch.empty(out.shape, dtype=torch.float16, device=X.device) if want_f16 else None
            L.check(lib.rn_relation_packed_fwd_f16io(C.byref(desc), _ptr(X), _ptr(x_f16), _ptr(boxes), _ptr(kidx),
                                                     _ptr(packed), _ptr(Wg), _ptr(bg), _ptr(out), _ptr(out16), _ptr(ws),
                                                     ws.numel(), stage_mask, _stream()), 'rn_relation_packed_fwd_f16io')
            return (out, out16) if want_f16 else out
        if stage_mask != 7:     # measurement hook: rerun a subset of the stages on the previous call's intermediates
            L.check(lib.rn_relation_packed_stages(C.byref(desc), _ptr(X), _ptr(boxes), _ptr(kidx), _ptr(packed), _ptr(Wg),
                                                  _ptr(bg), _ptr(out), _ptr(ws), ws.numel(), stage_mask, _stream()),
                    'rn_relation_packed_stages')
            return out
        L.check(lib.rn_relation_packed_fwd(C.byref(desc), _ptr(X), _ptr(boxes), _ptr(kidx), _ptr(packed), _ptr(Wg),
                                           _ptr(bg), _ptr(out), _ptr(ws), ws.numel(), _stream()), 'rn_relation_packed_fwd')
        return out
    L.check(lib.rn_relation_fwd(C.byref(desc), _ptr(X), _ptr(boxes), _ptr(kidx), _ptr(Wq), _ptr(bq), _ptr(Wk), _ptr(bk),
                                _ptr(Wg), _ptr(bg), _ptr(Wout2), _ptr(bout), _ptr(out), _ptr(sm), _ptr(ws), ws.numel(),
                                _stream()), 'rn_relation_fwd')
    return (out, sm) if return_softmax else out


def _grad_buffers(out, like):
    """gradient outputs: fresh tensors, or the caller's (e.g. windows of a replicas.GradientBucket) when `out` has them"""
    res = {}
    for n, t in like.items():
        o = out.get(n) if out else None
        if o is None:
            o = torch.empty_like(t)
        elif not (o.is_cuda and o.dtype == torch.float32 and o.is_contiguous() and o.numel() == t.numel()):
            raise L.RelnetError('gradient buffer %s must be a contiguous float32 CUDA tensor of %d elements' % (n, t.numel()))
        res[n] = o
    return res


def relation_backward(grad_out, X, boxes, Wq, bq, Wk, bk, Wg, bg, Wout, bout, key_index=None, M=None, group=16,
                      residual_relu=False, wave_length=1000.0, out=None, precision=None, forward_out=None):
    """Gradients of `relation` (forward intermediates are recomputed) -- rn_relation_bwd.  precision: 'f16' (default on
    sm_100) = every contraction on the library's tcgen05 tf32 GEMM (fp32 operands, fp32 accumulate); 'fp32' = cuBLAS fp32.
    forward_out: the output of the forward that was executed (any precision); with residual_relu its sign pattern is the relu
    mask (autograd semantics) instead of the recomputed forward's -- they differ only for units within rounding of zero.
    Returns a dict with the gradient of X and of every parameter, shaped like the argument it belongs to."""
    precision = precision or default_precision()
    X = _f32(X, 'X'); boxes = _f32(boxes, 'boxes'); grad_out = _f32(grad_out, 'grad_out')
    batched = X.dim() == 3
    B = X.shape[0] if batched else 1
    N, d = X.shape[-2], X.shape[-1]
    Wq, bq, Wk, bk, Wg, bg, Wout, bout = [_f32(w, n) for w, n in ((Wq, 'Wq'), (bq, 'bq'), (Wk, 'Wk'), (bk, 'bk'), (Wg, 'Wg'),
                                                                  (bg, 'bg'), (Wout, 'Wout'), (bout, 'bout'))]
    Wout2 = Wout.reshape(Wout.shape[0], -1)
    dq, dout = Wq.shape[0], Wout2.shape[0]
    if not (boxes.shape[-1] == 4 and boxes.shape[:-1] == X.shape[:-1] and tuple(Wq.shape) == (dq, d)
            and tuple(Wk.shape) == (dq, d) and Wout2.shape[1] == d and bq.numel() == dq and bk.numel() == dq
            and bout.numel() == dout and Wg.shape[0] == group and bg.numel() == group and dq % group == 0
            and dout % group == 0 and grad_out.numel() == B * N * dout):
        raise L.RelnetError('relation_backward: inconsistent shapes X%s boxes%s grad_out%s Wq%s Wk%s Wg%s Wout%s group=%d' % (
            tuple(X.shape), tuple(boxes.shape), tuple(grad_out.shape), tuple(Wq.shape), tuple(Wk.shape), tuple(Wg.shape),
            tuple(Wout.shape), group))
    kidx = None
    if key_index is not None:
        kidx = key_index.to(device=X.device, dtype=torch.int32).contiguous()
        M = kidx.numel()
    M = int(M) if M is not None else N
    desc = L.RelationDesc(B, N, M, d, dq, dout, group, Wg.shape[1], wave_length, int(residual_relu), PREC[precision])
    g = _grad_buffers(out, {'X': X, 'Wq': Wq, 'bq': bq, 'Wk': Wk, 'bk': bk, 'Wg': Wg, 'bg': bg, 'Wout': Wout, 'bout': bout})
    lib = L.lib()
    ws = _workspace(lib.rn_relation_bwd_workspace_bytes(C.byref(desc)), X.device)
    fo = None
    if forward_out is not None:
        fo = _f32(forward_out, 'forward_out')
        if fo.numel() != B * N * dout:
            raise L.RelnetError('relation_backward: forward_out has %d elements, expected %d' % (fo.numel(), B * N * dout))
    L.check(lib.rn_relation_bwd_masked(C.byref(desc), _ptr(X), _ptr(boxes), _ptr(kidx), _ptr(Wq), _ptr(bq), _ptr(Wk), _ptr(bk),
                                       _ptr(Wg), _ptr(bg), _ptr(Wout2), _ptr(bout), _ptr(fo), _ptr(grad_out), _ptr(g['X']),
                                       _ptr(g['Wq']), _ptr(g['bq']), _ptr(g['Wk']), _ptr(g['bk']), _ptr(g['Wg']), _ptr(g['bg']),
                                       _ptr(g['Wout']), _ptr(g['bout']), _ptr(ws), ws.numel(), _stream()), 'rn_relation_bwd_masked')
    return g


def relation_workspace_bytes(N, M, d, dq, dout, group, batch=1, E=64, precision='f16'):
    desc = L.RelationDesc(batch, N, M, d, dq, dout, group, E, 1000.0, 0, PREC[precision])
    return int(L.lib().rn_relation_workspace_bytes(C.byref(desc)))


def relation_tc_supported(dq, dout, group, return_softmax=False):
    """Shapes the tcgen05 relation kernels cover (relation_tc.cu:head_chunks): d_k <= 64 and d_v <= 64 per head (narrower
    heads are zero-padded to 64 columns at pack time), or d_k = d_v = 64 c run as `group * c` <= 16 virtual heads of 64 columns
    that share their head's geometry weight and softmax (fused kernel only)."""
    if return_softmax or dq % group or dout % group or dq < group or dout < group:
        return False
    dk, dv = dq // group, dout // group
    if dk <= 64 and dv <= 64:
        return True
    if dk != dv or dk % 64:
        return False
    team = group * (dk // 64)
    return team <= 16 and 128 % team == 0 and (128 // team) % 8 == 0 and bool(relation_fused_active())


def relation_fused_active():
    """Current setting of the A/B switch (0 = round-1 decomposition, 1 / 2 = fused)."""
    prev = int(L.lib().rn_relation_fused_enable(1))
    L.lib().rn_relation_fused_enable(prev)
    return prev


def relation_fused_enable(on):
    """A/B switch of the RN_PREC_F16 relation module: 1 = fused geometry + attention launch (default), 0 = round-1
    decomposition (geometry table -> tile attention -> combine), 2 = fused with phi's fp16 residual in the pair FC.
    Returns the previous setting."""
    return int(L.lib().rn_relation_fused_enable(int(on)))


def pos_embed(boxes, M=None, key_index=None, E=64, wave_length=1000.0, want_eps=True, want_emb=True):
    boxes = _f32(boxes, 'boxes')
    N = boxes.shape[0]
    kidx = key_index.to(device=boxes.device, dtype=torch.int32).contiguous() if key_index is not None else None
    M = kidx.numel() if kidx is not None else (int(M) if M is not None else N)
    eps = torch.empty((N, M, 4), dtype=torch.float32, device=boxes.device) if want_eps else None
    emb = torch.empty((N, M, E), dtype=torch.float32, device=boxes.device) if want_emb else None
    L.check(L.lib().rn_pos_embed_fwd(_ptr(boxes), _ptr(kidx), N, M, E, wave_length, _ptr(eps), _ptr(emb), _stream()),
            'rn_pos_embed_fwd')
    return eps, emb


def geometry_weight(boxes, Wg, bg, M=None, key_index=None, wave_length=1000.0):
    boxes = _f32(boxes, 'boxes'); Wg = _f32(Wg, 'Wg'); bg = _f32(bg, 'bg')
    batched = boxes.dim() == 3
    B = boxes.shape[0] if batched else 1
    N = boxes.shape[-2]
    kidx = key_index.to(device=boxes.device, dtype=torch.int32).contiguous() if key_index is not None else None
    M = kidx.numel() if kidx is not None else (int(M) if M is not None else N)
    H, E = Wg.shape
    g = torch.empty((B, H, N, M) if batched else (H, N, M), dtype=torch.float32, device=boxes.device)
    L.check(L.lib().rn_geometry_weight_fwd(_ptr(boxes), _ptr(kidx), B, N, M, H, E, wave_length, _ptr(Wg), _ptr(bg),
                                           _ptr(g), _stream()), 'rn_geometry_weight_fwd')
    return g


def linear(x, W, b=None, relu=False, precision=None, x_f16=None, want_f16=False):
    """y = act(x W^T + b).  RN_PREC_F16 only: x_f16 = fp16 copy of x (no cast launch); want_f16 -> returns (y, y_f16)."""
    precision = precision or default_precision()
    x = _f32(x, 'x'); W = _f32(W, 'W'); b = _f32(b, 'b') if b is not None else None
    x2 = x.reshape(x.shape[0], -1)
    rows, cin = x2.shape
    cout = W.shape[0]
    assert W.shape[1] == cin, (W.shape, cin)
    y = torch.empty((rows, cout), dtype=torch.float32, device=x.device)
    lib = L.lib()
    ws = _workspace(lib.rn_linear_workspace_bytes(rows, cin, cout, PREC[precision]), x.device)
    if precision == 'f16':
        def pack(buf):
            L.check(lib.rn_linear_pack(_ptr(W), cin, cout, _ptr(buf), _stream()), 'rn_linear_pack')
        packed = _packs.get((W,), lib.rn_linear_packed_bytes(cin, cout), pack, tag=('linear',))
        if (x_f16 is not None or want_f16) and cin % 8 == 0:
            if x_f16 is None:
                x_f16 = x2.to(torch.float16)
            elif not (x_f16.is_cuda and x_f16.dtype == torch.float16 and x_f16.is_contiguous() and x_f16.numel() == x2.numel()):
                raise L.RelnetError('linear: x_f16 must be the contiguous fp16 CUDA copy of x')
            y16 = torch.empty((rows, cout), dtype=torch.float16, device=x.device) if want_f16 else None
            L.check(lib.rn_linear_packed_f16in_fwd(_ptr(x_f16), _ptr(packed), _ptr(b), _ptr(y), _ptr(y16), rows, cin, cout,
                                                   int(relu), _ptr(ws), ws.numel(), _stream()), 'rn_linear_packed_f16in_fwd')
            return (y, y16) if want_f16 else y
        L.check(lib.rn_linear_packed_fwd(_ptr(x2), _ptr(packed), _ptr(b), _ptr(y), rows, cin, cout, int(relu), _ptr(ws),
                                         ws.numel(), _stream()), 'rn_linear_packed_fwd')
        return (y, y.to(torch.float16)) if want_f16 else y
    L.check(lib.rn_linear_fwd(_ptr(x2), _ptr(W), _ptr(b), _ptr(y), rows, cin, cout, int(relu), PREC[precision], _ptr(ws),
                              ws.numel(), _stream()), 'rn_linear_fwd')
    return y


def linear_multi(x_f16, layers):
    """Several FullyConnected layers over the same fp16 input as ONE tcgen05 GEMM (rn_linear_multi_packed_f16in_fwd).
    layers: list of (W [out_i, in], b [out_i] or None), at most 4.  Returns the list of contiguous fp32 [rows, out_i] outputs."""
    if not (x_f16.is_cuda and x_f16.dtype == torch.float16 and x_f16.is_contiguous() and x_f16.dim() == 2):
        raise L.RelnetError('linear_multi: x_f16 must be a contiguous 2-D fp16 CUDA tensor')
    rows, cin = x_f16.shape
    n = len(layers)
    Ws = [_f32(W, 'W') for W, _ in layers]
    bs = [(_f32(b, 'b') if b is not None else None) for _, b in layers]
    if not (1 <= n <= 4 and cin % 8 == 0 and all(W.dim() == 2 and W.shape[1] == cin for W in Ws)):
        raise L.RelnetError('linear_multi: 1..4 layers of [out_i, %d] weights (in %% 8 == 0)' % cin)
    outs = (c_i32 * n)(*[W.shape[0] for W in Ws])
    lib = L.lib()

    def pack(buf):
        wp = (C.c_void_p * n)(*[W.data_ptr() for W in Ws])
        bp = (C.c_void_p * n)(*[(b.data_ptr() if b is not None else None) for b in bs])
        L.check(lib.rn_linear_multi_pack(wp, bp, outs, n, cin, _ptr(buf), _stream()), 'rn_linear_multi_pack')
    keys = tuple(Ws) + tuple(b for b in bs if b is not None)
    packed = _packs.get(keys, lib.rn_linear_multi_packed_bytes(outs, n, cin), pack, tag=('multi', n))
    ys = [torch.empty((rows, W.shape[0]), dtype=torch.float32, device=x_f16.device) for W in Ws]
    yp = (C.c_void_p * n)(*[y.data_ptr() for y in ys])
    ws = _workspace(256, x_f16.device)
    L.check(lib.rn_linear_multi_packed_f16in_fwd(_ptr(x_f16), _ptr(packed), yp, outs, n, rows, cin, _ptr(ws), ws.numel(),
                                                 _stream()), 'rn_linear_multi_packed_f16in_fwd')
    return ys


# ------------------------------------------------------------------------------------------------------------------
def _learn_nms_desc(cls_score, bbox_pred, rois, feat, first_n, num_thresh, class_thresh, class_agnostic, means, stds,
                    nongt_dim, non_gt_index, merge_method, precision):
    R, NC = cls_score.shape
    desc = L.LearnNmsDesc()
    desc.R, desc.num_classes, desc.num_reg_classes = R, NC, bbox_pred.shape[1] // 4
    desc.feat_dim, desc.first_n, desc.num_thresh, desc.class_thresh = feat.shape[1], first_n, num_thresh, class_thresh
    desc.class_agnostic = int(class_agnostic)
    desc.has_means_stds = int(means is not None and stds is not None)
    if desc.has_means_stds:
        desc.means = (C.c_float * 4)(*[float(v) for v in means]); desc.stds = (C.c_float * 4)(*[float(v) for v in stds])
    kidx = None
    desc.nongt_dim = int(nongt_dim) if nongt_dim is not None else 0
    if nongt_dim is None and non_gt_index is not None:
        kidx = non_gt_index.to(device=rois.device, dtype=torch.int32).contiguous()
        desc.num_non_gt = kidx.numel()
    desc.merge_method, desc.precision = merge_method, PREC[precision]
    return desc, kidx


def learn_nms(cls_score, bbox_pred, rois, im_info, feat, weights, first_n=100, num_thresh=5, class_thresh=0.01,
              class_agnostic=True, means=None, stds=None, nongt_dim=None, non_gt_index=None, merge_method=-1,
              precision=None, feat_f16=None, emb=None):
    """learn_nms CustomOp forward (LNMS:238-401) + merge.  ``weights``: dict by checkpoint name (LNMS:429-441).
    feat_f16 (RN_PREC_F16 only): the producer's fp16 copy of feat, saves the cast launch; emb (RN_PREC_F16 only):
    roi_feat_embedding(feat) [R,128] when the caller already evaluated it (e.g. merged with cls_score / bbox_pred)."""
    precision = precision or default_precision()
    cls_score = _f32(cls_score, 'cls_score'); bbox_pred = _f32(bbox_pred, 'bbox_pred'); rois = _f32(rois, 'rois')
    im_info = _f32(im_info, 'im_info').reshape(-1); feat = _f32(feat, 'feat')
    R_, NC = cls_score.shape
    desc, kidx = _learn_nms_desc(cls_score, bbox_pred, rois, feat, first_n, num_thresh, class_thresh, class_agnostic, means,
                                 stds, nongt_dim, non_gt_index, merge_method, precision)
    keep = [_f32(weights[n], n) for n in L.LearnNmsWeights.NAMES]
    w = L.LearnNmsWeights(*[t.data_ptr() for t in keep])
    C_ = NC - 1
    dev = rois.device
    multi = torch.empty((first_n, C_, num_thresh), dtype=torch.float32, device=dev)
    sbbox = torch.empty((first_n, C_, 4), dtype=torch.float32, device=dev)
    sscore = torch.empty((first_n, C_), dtype=torch.float32, device=dev)
    final = torch.empty((first_n, C_), dtype=torch.float32, device=dev)
    lib = L.lib()
    ws = _workspace(lib.rn_learn_nms_workspace_bytes(C.byref(desc)), dev)
    pk_bytes = lib.rn_learn_nms_packed_bytes(C.byref(desc))
    if pk_bytes:        # RN_PREC_F16 class-agnostic: weight-only work cached per weight version (like rn_relation_pack)
        def pack(buf):
            L.check(lib.rn_learn_nms_pack(C.byref(desc), C.byref(w), _ptr(buf), _ptr(ws), ws.numel(), _stream()),
                    'rn_learn_nms_pack')
        packed = _packs.get(tuple(keep), pk_bytes, pack, tag=('learn_nms', first_n, feat.shape[1]))
        if feat_f16 is not None and not (feat_f16.is_cuda and feat_f16.dtype == torch.float16 and feat_f16.is_contiguous()
                                         and feat_f16.numel() == feat.numel() and feat.shape[1] % 8 == 0):
            raise L.RelnetError('learn_nms: feat_f16 must be the contiguous fp16 CUDA copy of feat (feat_dim % 8 == 0)')
        if emb is not None:
            emb = _f32(emb, 'emb')
            if tuple(emb.shape) != (R_, 128):
                raise L.RelnetError('learn_nms: emb must be [R, 128]')
        L.check(lib.rn_learn_nms_packed_fwd(C.byref(desc), _ptr(cls_score), _ptr(bbox_pred), _ptr(rois), _ptr(im_info),
                                            _ptr(feat), _ptr(feat_f16), _ptr(emb), C.byref(w), _ptr(packed), _ptr(kidx), _ptr(multi), _ptr(sbbox),
                                            _ptr(sscore), _ptr(final), _ptr(ws), ws.numel(), _stream()),
                'rn_learn_nms_packed_fwd')
        return multi, sbbox, sscore, final
    L.check(lib.rn_learn_nms_fwd(C.byref(desc), _ptr(cls_score), _ptr(bbox_pred), _ptr(rois), _ptr(im_info), _ptr(feat),
                                 C.byref(w), _ptr(kidx), _ptr(multi), _ptr(sbbox), _ptr(sscore), _ptr(final), _ptr(ws),
                                 ws.numel(), _stream()), 'rn_learn_nms_fwd')
    return multi, sbbox, sscore, final


def learn_nms_backward(grad_multi, cls_score, bbox_pred, rois, im_info, feat, weights, first_n=100, num_thresh=5,
                       class_thresh=0.0, class_agnostic=True, means=None, stds=None, nongt_dim=None, non_gt_index=None,
                       out=None, precision=None):
    """Gradients of nms_multi_score (train graph SYM_REL_NMS:424-501) -- rn_learn_nms_bwd.  precision as in
    relation_backward (contraction engine: tcgen05 tf32 GEMM / cuBLAS fp32).  Returns (dict of the 14 weight
    gradients by checkpoint name, d_cls_score [R,num_classes], d_feat [R,feat_dim])."""
    precision = precision or default_precision()
    cls_score = _f32(cls_score, 'cls_score'); bbox_pred = _f32(bbox_pred, 'bbox_pred'); rois = _f32(rois, 'rois')
    im_info = _f32(im_info, 'im_info').reshape(-1); feat = _f32(feat, 'feat'); grad_multi = _f32(grad_multi, 'grad_multi')
    NC = cls_score.shape[1]
    if tuple(grad_multi.shape) != (first_n, NC - 1, num_thresh):
        raise L.RelnetError('learn_nms_backward: grad_multi shape %s != %s' % (tuple(grad_multi.shape), (first_n, NC - 1, num_thresh)))
    desc, kidx = _learn_nms_desc(cls_score, bbox_pred, rois, feat, first_n, num_thresh, class_thresh, class_agnostic, means,
                                 stds, nongt_dim, non_gt_index, -1, precision)
    keep = [_f32(weights[n], n) for n in L.LearnNmsWeights.NAMES]
    w = L.LearnNmsWeights(*[t.data_ptr() for t in keep])
    grads = _grad_buffers(out, dict(zip(L.LearnNmsWeights.NAMES, keep)))
    g = L.LearnNmsWeights(*[grads[n].data_ptr() for n in L.LearnNmsWeights.NAMES])
    d_cls = torch.empty_like(cls_score); d_feat = torch.empty_like(feat)
    lib = L.lib()
    ws = _workspace(lib.rn_learn_nms_bwd_workspace_bytes(C.byref(desc)), rois.device)
    L.check(lib.rn_learn_nms_bwd(C.byref(desc), _ptr(cls_score), _ptr(bbox_pred), _ptr(rois), _ptr(im_info), _ptr(feat),
                                 C.byref(w), _ptr(kidx), _ptr(grad_multi), C.byref(g), _ptr(d_cls), _ptr(d_feat), _ptr(ws),
                                 ws.numel(), _stream()), 'rn_learn_nms_bwd')
    return grads, d_cls, d_feat


def nms_loss(nms_multi_score, nms_multi_target, loss_scale=1.0, pos_grad_scale=4.0, eps=1e-8):
    """learn-NMS loss terms and gradient (SYM_REL_NMS:539-551) -- rn_nms_loss.  Returns (pos_loss, neg_loss, d_multi)."""
    m = _f32(nms_multi_score, 'nms_multi_score'); t = _f32(nms_multi_target, 'nms_multi_target')
    if m.dim() != 3 or m.shape != t.shape:
        raise L.RelnetError('nms_loss: score %s / target %s must both be [n,C,T]' % (tuple(m.shape), tuple(t.shape)))
    pos = torch.empty_like(m); neg = torch.empty_like(m); d = torch.empty_like(m)
    L.check(L.lib().rn_nms_loss(_ptr(m), _ptr(t), m.shape[0], m.shape[1], m.shape[2], loss_scale, pos_grad_scale, eps,
                                _ptr(pos), _ptr(neg), _ptr(d), _stream()), 'rn_nms_loss')
    return pos, neg, d


def box_annotator_ohem(cls_score, bbox_pred, labels, bbox_targets, bbox_weights, num_classes, num_reg_classes, roi_per_img,
                       return_loss=False):
    """'BoxAnnotatorOHEM' CustomOp forward (box_annotator_ohem.py:26-53) -> labels_ohem [R], bbox_weights_ohem."""
    cls_score = _f32(cls_score, 'cls_score'); bbox_pred = _f32(bbox_pred, 'bbox_pred'); labels = _f32(labels, 'labels')
    bbox_targets = _f32(bbox_targets, 'bbox_targets'); bbox_weights = _f32(bbox_weights, 'bbox_weights')
    R = cls_score.shape[0]
    D = 4 * int(num_reg_classes)
    if not (cls_score.shape[1] == int(num_classes) and labels.numel() == R and tuple(bbox_pred.shape) == (R, D)
            and bbox_targets.shape == bbox_pred.shape and bbox_weights.shape == bbox_pred.shape):
        raise L.RelnetError('box_annotator_ohem: inconsistent shapes')
    lab = torch.empty_like(labels); w = torch.empty_like(bbox_weights)
    loss = torch.empty(R, dtype=torch.float32, device=cls_score.device)
    L.check(L.lib().rn_box_annotator_ohem(_ptr(cls_score), _ptr(bbox_pred), _ptr(labels), _ptr(bbox_targets),
                                          _ptr(bbox_weights), R, int(num_classes), int(num_reg_classes), int(roi_per_img),
                                          _ptr(lab), _ptr(w), _ptr(loss), _stream()), 'rn_box_annotator_ohem')
    return (lab, w, loss) if return_loss else (lab, w)


def nms_multi_target(bbox, gt_boxes, score, target_thresh):
    """nms_multi_target CustomOp forward: bbox [n,C,4], gt_boxes [G,5] or [1,G,5], score [n,C] -> [n,C,T] of 0/1."""
    bbox = _f32(bbox, 'bbox'); score = _f32(score, 'score'); gt = _f32(gt_boxes, 'gt_boxes')
    if gt.dim() == 3:
        assert gt.shape[0] == 1, 'only support batch_image=1, but receive %d' % gt.shape[0]
        gt = gt[0].contiguous()
    assert gt.shape[-1] == 5, 'code_size of gt should be 5, but receive %d' % gt.shape[-1]
    assert score.dim() == 2 and score.shape[1] == bbox.shape[1], 'number of fg classes should be same for boxes and scores'
    n, Cc = bbox.shape[0], bbox.shape[1]
    T = len(target_thresh)
    th = (C.c_double * T)(*[float(t) for t in target_thresh])
    out = torch.empty((n, Cc, T), dtype=torch.float32, device=bbox.device)
    L.check(L.lib().rn_nms_multi_target_fwd(_ptr(bbox), _ptr(gt), _ptr(score), n, Cc, gt.shape[0], th, T, _ptr(out),
                                            _stream()), 'rn_nms_multi_target_fwd')
    return out


# ------------------------------------------------------------------------------------------------------------------
def proposal(cls_prob, bbox_pred, im_info, feat_stride=16, scales=(4, 8, 16, 32), ratios=(0.5, 1, 2),
             pre_nms_top_n=6000, post_nms_top_n=300, thresh=0.7, min_size=0, return_num_kept=False):
    cls_prob = _f32(cls_prob, 'cls_prob'); bbox_pred = _f32(bbox_pred, 'bbox_pred')
    im_info = _f32(im_info, 'im_info').reshape(-1)
    if cls_prob.shape[0] != 1:
        raise ValueError('Sorry, multiple images each device is not implemented')      # proposal.py:54-56
    Hf, Wf = cls_prob.shape[2], cls_prob.shape[3]
    desc = L.ProposalDesc(Hf, Wf, feat_stride, len(scales), len(ratios), pre_nms_top_n, post_nms_top_n, thresh,
                          float(min_size))
    sc = (C.c_float * len(scales))(*[float(s) for s in scales])
    ra = (C.c_float * len(ratios))(*[float(r) for r in ratios])
    dev = cls_prob.device
    rois = torch.empty((post_nms_top_n, 5), dtype=torch.float32, device=dev)
    scores = torch.empty((post_nms_top_n, 1), dtype=torch.float32, device=dev)
    nk = torch.zeros(1, dtype=torch.int32, device=dev)
    lib = L.lib()
    ws = _workspace(lib.rn_proposal_workspace_bytes(C.byref(desc)), dev)
    L.check(lib.rn_proposal_fwd(C.byref(desc), sc, ra, _ptr(cls_prob), _ptr(bbox_pred), _ptr(im_info), _ptr(rois),
                                _ptr(scores), _ptr(nk), _ptr(ws), ws.numel(), _stream()), 'rn_proposal_fwd')
    return (rois, scores, nk) if return_num_kept else (rois, scores)


def nms(boxes_sorted, thresh, max_keep=None):
    """Device NMS over boxes already sorted by score; returns (keep int32 [max_keep], num int32[1]) on the device."""
    b = _f32(boxes_sorted, 'boxes')
    n, dim = b.shape
    max_keep = max_keep or max(n, 1)
    keep = torch.empty(max_keep, dtype=torch.int32, device=b.device)
    num = torch.zeros(1, dtype=torch.int32, device=b.device)
    lib = L.lib()
    ws = _workspace(lib.rn_nms_workspace_bytes(n), b.device)
    L.check(lib.rn_nms(_ptr(b), n, dim, thresh, max_keep, _ptr(keep), _ptr(num), _ptr(ws), ws.numel(), _stream()), 'rn_nms')
    return keep, num


def bbox_overlaps(boxes, query):
    if not boxes.is_cuda:
        raise L.RelnetError('bbox_overlaps: CUDA tensors required')
    b = boxes.contiguous().double(); q = query.contiguous().double()
    out = torch.zeros((b.shape[0], q.shape[0]), dtype=torch.float64, device=b.device)
    L.check(L.lib().rn_bbox_overlaps(_ptr(b), _ptr(q), b.shape[0], q.shape[0], _ptr(out), _stream()), 'rn_bbox_overlaps')
    return out


def proposal_target(rois, gt_boxes, num_reg_classes=2, class_agnostic=True, bg_thresh_hi=0.5, normalize=True,
                    means=(0.0, 0.0, 0.0, 0.0), stds=(0.1, 0.1, 0.2, 0.2), bbox_weights=(1.0, 1.0, 1.0, 1.0)):
    rois = _f32(rois, 'rois'); gt = _f32(gt_boxes, 'gt_boxes')
    N, G = rois.shape[0], gt.shape[0]
    desc = L.ProposalTargetDesc()
    desc.N, desc.G, desc.num_reg_classes, desc.class_agnostic = N, G, num_reg_classes, int(class_agnostic)
    desc.bg_thresh_hi, desc.normalize = bg_thresh_hi, int(normalize)
    desc.means = (C.c_double * 4)(*means); desc.stds = (C.c_double * 4)(*stds)
    desc.bbox_weights = (C.c_float * 4)(*bbox_weights)
    R = 2 if class_agnostic else num_reg_classes
    dev = rois.device
    ro = torch.empty((N + G, 5), dtype=torch.float32, device=dev)
    lab = torch.empty((N + G,), dtype=torch.float32, device=dev)
    bt = torch.empty((N + G, 4 * R), dtype=torch.float32, device=dev)
    bw = torch.empty((N + G, 4 * R), dtype=torch.float32, device=dev)
    L.check(L.lib().rn_proposal_target_fwd(C.byref(desc), _ptr(rois), _ptr(gt), _ptr(ro), _ptr(lab), _ptr(bt), _ptr(bw),
                                           _stream()), 'rn_proposal_target_fwd')
    return ro, lab, bt, bw


# ------------------------------------------------------------------------------------------------------------------
def roi_pool(data, rois, pooled_size=(7, 7), spatial_scale=0.0625, return_argmax=False):
    data = _f32(data, 'data'); rois = _f32(rois, 'rois')
    B, Cc, H, W = data.shape
    R = rois.shape[0]
    out = torch.empty((R, Cc, pooled_size[0], pooled_size[1]), dtype=torch.float32, device=data.device)
    arg = torch.empty(out.shape, dtype=torch.int32, device=data.device) if return_argmax else None
    L.check(L.lib().rn_roi_pool_fwd(_ptr(data), _ptr(rois), R, Cc, H, W, pooled_size[0], pooled_size[1], spatial_scale,
                                    _ptr(out), _ptr(arg), _stream()), 'rn_roi_pool_fwd')
    return (out, arg) if return_argmax else out


def roi_pool_backward(grad_out, argmax, rois, data_shape):
    """Gradient of `roi_pool` w.r.t. data (argmax from roi_pool(..., return_argmax=True)) -- rn_roi_pool_bwd."""
    grad_out = _f32(grad_out, 'grad_out'); rois = _f32(rois, 'rois')
    if argmax.dtype != torch.int32 or not argmax.is_cuda or tuple(argmax.shape) != tuple(grad_out.shape):
        raise L.RelnetError('roi_pool_backward: argmax must be the int32 CUDA tensor returned by roi_pool')
    B, Cc, H, W = data_shape
    R, _, PH, PW = grad_out.shape
    dd = torch.empty(tuple(data_shape), dtype=torch.float32, device=grad_out.device)
    L.check(L.lib().rn_roi_pool_bwd(_ptr(grad_out), _ptr(argmax.contiguous()), _ptr(rois), R, B, Cc, H, W, PH, PW, _ptr(dd),
                                    _stream()), 'rn_roi_pool_bwd')
    return dd


def roi_pool_nhwc_f16(data, rois, spatial_scale, pooled_size=(7, 7), out=None):
    """ROIPooling (max) from a channels-last map (fp32 or bf16, consumed without a copy) to fp16 [R, PH*PW*C] (bin-major,
    channel-minor: the K order of a weight packed by rn_linear_pack_chw_to_hwc)."""
    if not data.is_cuda:
        raise L.RelnetError('roi_pool_nhwc_f16: CUDA tensors required (no CPU path)')
    B, Cc, H, Wd = data.shape
    rois = _f32(rois, 'rois')
    R = rois.shape[0]
    cin = Cc * pooled_size[0] * pooled_size[1]
    lib = L.lib()
    if out is None:
        out = torch.empty((R, cin), dtype=torch.float16, device=data.device)
    elif not (out.is_cuda and out.dtype == torch.float16 and out.is_contiguous() and tuple(out.shape) == (R, cin)):
        raise L.RelnetError('roi_pool_nhwc_f16: out must be a contiguous fp16 [R, PH*PW*C] CUDA tensor')
    if R == 0:
        return out
    if data.dtype == torch.bfloat16 and Cc % 8 == 0:       # the trunk's native output: pooled straight from bf16
        nhwc = data.contiguous(memory_format=torch.channels_last)
        L.check(lib.rn_roi_pool_nhwc_bf16in_f16_fwd(_ptr(nhwc), _ptr(rois), R, Cc, H, Wd, pooled_size[0], pooled_size[1],
                                                    spatial_scale, _ptr(out), _stream()), 'rn_roi_pool_nhwc_bf16in_f16_fwd')
    else:
        nhwc = data.float().contiguous(memory_format=torch.channels_last)      # no-op for an fp32 channels-last map
        L.check(lib.rn_roi_pool_nhwc_f16_fwd(_ptr(nhwc), _ptr(rois), R, Cc, H, Wd, pooled_size[0], pooled_size[1],
                                             spatial_scale, _ptr(out), _stream()), 'rn_roi_pool_nhwc_f16_fwd')
    return out


def linear_pooled_hwc(pooled16, W, b, Cc, S, relu=False, want_f16=False):
    """FullyConnected on a bin-major pooled tensor (roi_pool_nhwc_f16) with the reference-layout weight [out, C*S]: the
    weight is packed once with its K axis permuted from (c, s) to (s, c)."""
    W = _f32(W, 'W'); b = _f32(b, 'b') if b is not None else None
    R, cin = pooled16.shape
    cout = W.shape[0]
    assert W.shape[1] == cin == Cc * S
    lib = L.lib()

    def pack(buf):
        L.check(lib.rn_linear_pack_chw_to_hwc(_ptr(W), cout, Cc, S, _ptr(buf), _stream()), 'rn_linear_pack_chw_to_hwc')
    packed = _packs.get_tagged('chw2hwc', (W,), lib.rn_linear_packed_bytes(cin, cout), pack)
    y = torch.empty((R, cout), dtype=torch.float32, device=pooled16.device)
    ws = _workspace(lib.rn_linear_workspace_bytes(R, cin, cout, 1), pooled16.device)
    y16 = torch.empty((R, cout), dtype=torch.float16, device=pooled16.device) if want_f16 else None
    L.check(lib.rn_linear_packed_f16in_fwd(_ptr(pooled16), _ptr(packed), _ptr(b), _ptr(y), _ptr(y16), R, cin, cout, int(relu),
                                           _ptr(ws), ws.numel(), _stream()), 'rn_linear_packed_f16in_fwd')
    return (y, y16) if want_f16 else y


def roi_pool_fc(data, rois, W, b, pooled_size=(7, 7), spatial_scale=0.0625, relu=False, want_f16=False):
    """ROIPooling + FullyConnected fused at the data-layout level (SYM_REL:252-262 with RN_PREC_F16): channels-last
    feature map -> fp16 pooled [R, PH*PW*C] -> tcgen05 GEMM against the K-permuted packed weight.  Returns fp32 [R, out].
    ``data`` is an NCHW-shaped tensor; a channels_last one is consumed without a copy."""
    pooled = roi_pool_nhwc_f16(data, rois, spatial_scale, pooled_size)
    return linear_pooled_hwc(pooled, W, b, data.shape[1], pooled_size[0] * pooled_size[1], relu=relu, want_f16=want_f16)


def deform_psroi_pool(data, rois, trans=None, spatial_scale=0.0625, output_dim=256, group_size=1, pooled_size=7,
                      part_size=0, sample_per_part=4, trans_std=0.0, no_trans=None, return_count=False):
    """DeformablePSROIPooling forward.  `data` NCHW fp32 (the reference layout) or a channels_last fp32 / bf16 map (the
    trunk's layout: the fast form, every bilinear tap a 16-byte channel-vector load -- rn_deform_psroi_pool_nhwc_fwd)."""
    if not isinstance(data, torch.Tensor) or not data.is_cuda:
        raise L.RelnetError('data must be a CUDA tensor (relnet_b200 has no CPU path)')
    nhwc = (data.dim() == 4 and data.dtype in (torch.float32, torch.bfloat16) and data.shape[1] > 1
            and data.is_contiguous(memory_format=torch.channels_last) and not data.is_contiguous())
    if not nhwc:
        data = _f32(data, 'data')
    rois = _f32(rois, 'rois')
    if no_trans is None:
        no_trans = trans is None
    t = _f32(trans, 'trans') if not no_trans else None
    B, Cc, H, W = data.shape
    R = rois.shape[0]
    desc = L.PsroiDesc(R, Cc, H, W, spatial_scale, output_dim, group_size, pooled_size, part_size or pooled_size,
                       sample_per_part, trans_std, int(bool(no_trans)), 1 if no_trans else t.shape[1] // 2)
    out = torch.empty((R, output_dim, pooled_size, pooled_size), dtype=torch.float32, device=data.device)
    cnt = torch.empty_like(out) if return_count else None
    if nhwc:
        L.check(L.lib().rn_deform_psroi_pool_nhwc_fwd(C.byref(desc), _ptr(data), int(data.dtype == torch.bfloat16), _ptr(rois),
                                                      _ptr(t), _ptr(out), _ptr(cnt), _stream()), 'rn_deform_psroi_pool_nhwc_fwd')
    else:
        L.check(L.lib().rn_deform_psroi_pool_fwd(C.byref(desc), _ptr(data), _ptr(rois), _ptr(t), _ptr(out), _ptr(cnt),
                                                 _stream()), 'rn_deform_psroi_pool_fwd')
    return (out, cnt) if return_count else out


def deform_psroi_pool_backward(grad_out, top_count, data, rois, trans=None, spatial_scale=0.0625, output_dim=256,
                               group_size=1, pooled_size=7, part_size=0, sample_per_part=4, trans_std=0.0, no_trans=None):
    """Gradients of `deform_psroi_pool` w.r.t. data and trans -- rn_deform_psroi_pool_bwd."""
    data = _f32(data, 'data'); rois = _f32(rois, 'rois'); grad_out = _f32(grad_out, 'grad_out')
    top_count = _f32(top_count, 'top_count')
    if no_trans is None:
        no_trans = trans is None
    t = _f32(trans, 'trans') if not no_trans else None
    B, Cc, H, W = data.shape
    R = rois.shape[0]
    if tuple(grad_out.shape) != (R, output_dim, pooled_size, pooled_size) or grad_out.shape != top_count.shape:
        raise L.RelnetError('deform_psroi_pool_backward: grad_out / top_count shape %s / %s' % (tuple(grad_out.shape), tuple(top_count.shape)))
    desc = L.PsroiDesc(R, Cc, H, W, spatial_scale, output_dim, group_size, pooled_size, part_size or pooled_size,
                       sample_per_part, trans_std, int(bool(no_trans)), 1 if no_trans else t.shape[1] // 2)
    dd = torch.empty_like(data)
    dt = torch.empty_like(t) if t is not None else None
    L.check(L.lib().rn_deform_psroi_pool_bwd(C.byref(desc), B, _ptr(grad_out), _ptr(top_count), _ptr(data), _ptr(rois),
                                             _ptr(t), _ptr(dd), _ptr(dt), _stream()), 'rn_deform_psroi_pool_bwd')
    return dd, dt


def _dc_desc(data, weight, kernel, pad, stride, dilate, num_group, num_deformable_group, precision):
    B, Cc, H, W = data.shape
    return L.DeformConvDesc(B, Cc, H, W, weight.shape[0], kernel[0], kernel[1], pad[0], pad[1], stride[0], stride[1],
                            dilate[0], dilate[1], num_group, num_deformable_group, PREC[precision])


def deform_conv(data, offset, weight, bias=None, kernel=(3, 3), pad=(2, 2), stride=(1, 1), dilate=(2, 2),
                num_group=1, num_deformable_group=4, precision='fp32'):
    data = _f32(data, 'data'); offset = _f32(offset, 'offset'); weight = _f32(weight, 'weight')
    bias = _f32(bias, 'bias') if bias is not None else None
    desc = _dc_desc(data, weight, kernel, pad, stride, dilate, num_group, num_deformable_group, precision)
    Ho, Wo = offset.shape[2], offset.shape[3]
    out = torch.empty((data.shape[0], weight.shape[0], Ho, Wo), dtype=torch.float32, device=data.device)
    lib = L.lib()
    ws = _workspace(lib.rn_deform_conv_workspace_bytes(C.byref(desc)), data.device)
    L.check(lib.rn_deform_conv_fwd(C.byref(desc), _ptr(data), _ptr(offset), _ptr(weight), _ptr(bias), _ptr(out), _ptr(ws),
                                   ws.numel(), _stream()), 'rn_deform_conv_fwd')
    return out


def deform_conv_backward(grad_out, data, offset, weight, kernel=(3, 3), pad=(2, 2), stride=(1, 1), dilate=(2, 2),
                         num_group=1, num_deformable_group=4, has_bias=False, weight_grad_deformed=False):
    """Gradients of `deform_conv` -- rn_deform_conv_bwd.  Returns (ddata, doffset, dweight, dbias or None).
    weight_grad_deformed=False reproduces the reference's dWeight (plain im2col, deformable_convolution-inl.h:215)."""
    data = _f32(data, 'data'); offset = _f32(offset, 'offset'); weight = _f32(weight, 'weight')
    grad_out = _f32(grad_out, 'grad_out')
    desc = _dc_desc(data, weight, kernel, pad, stride, dilate, num_group, num_deformable_group, 'fp32')
    if tuple(grad_out.shape) != (data.shape[0], weight.shape[0], offset.shape[2], offset.shape[3]):
        raise L.RelnetError('deform_conv_backward: grad_out shape %s' % (tuple(grad_out.shape),))
    dd = torch.empty_like(data); do = torch.empty_like(offset); dw = torch.empty_like(weight)
    db = torch.empty(weight.shape[0], dtype=torch.float32, device=data.device) if has_bias else None
    lib = L.lib()
    ws = _workspace(lib.rn_deform_conv_workspace_bytes(C.byref(desc)), data.device)
    L.check(lib.rn_deform_conv_bwd(C.byref(desc), _ptr(grad_out), _ptr(data), _ptr(offset), _ptr(weight),
                                   int(bool(weight_grad_deformed)), _ptr(dd), _ptr(do), _ptr(dw), _ptr(db), _ptr(ws),
                                   ws.numel(), _stream()), 'rn_deform_conv_bwd')
    return dd, do, dw, db


def deform_conv_nhwc(data, offset, weight, bias=None, relu=False, kernel=(3, 3), pad=(2, 2), stride=(1, 1), dilate=(2, 2),
                     num_deformable_group=4, out_dtype=torch.float16):
    """DeformableConvolution on a channels_last map (fp32 or bf16, one image): rn_deform_conv_nhwc_fwd -- the sampler
    writes an fp16 K-major column buffer from 16-byte channel vectors, the tcgen05 GEMM emits the channels_last output
    with bias (+relu) fused.  Returns [1, Co, Ho, Wo] in channels_last memory format (fp16 or fp32)."""
    if not (isinstance(data, torch.Tensor) and data.is_cuda and data.dim() == 4 and data.shape[0] == 1
            and data.dtype in (torch.float32, torch.bfloat16) and data.is_contiguous(memory_format=torch.channels_last)):
        raise L.RelnetError('deform_conv_nhwc: data must be a [1,C,H,W] channels_last fp32 / bf16 CUDA tensor')
    offset = _f32(offset, 'offset'); weight = _f32(weight, 'weight')
    b = _f32(bias, 'bias') if bias is not None else None
    desc = _dc_desc(data, weight, kernel, pad, stride, dilate, 1, num_deformable_group, 'f16')
    lib = L.lib()

    def pack(buf):
        L.check(lib.rn_deform_conv_pack(C.byref(desc), _ptr(weight), _ptr(buf), _stream()), 'rn_deform_conv_pack')
    packed = _packs.get_tagged('deform_conv', (weight,), lib.rn_deform_conv_packed_bytes(C.byref(desc)), pack)
    Ho, Wo = offset.shape[-2], offset.shape[-1]
    Co = weight.shape[0]
    out = torch.empty((1, Ho, Wo, Co), dtype=out_dtype, device=data.device)
    ws = _workspace(lib.rn_deform_conv_workspace_bytes(C.byref(desc)), data.device)
    o32 = out if out_dtype == torch.float32 else None
    o16 = out if out_dtype == torch.float16 else None
    if o32 is None and o16 is None:
        raise L.RelnetError('deform_conv_nhwc: out_dtype must be float16 or float32')
    L.check(lib.rn_deform_conv_nhwc_fwd(C.byref(desc), _ptr(data), int(data.dtype == torch.bfloat16), _ptr(offset), _ptr(packed),
                                        _ptr(b), int(relu), _ptr(o32), _ptr(o16), _ptr(ws), ws.numel(), _stream()),
            'rn_deform_conv_nhwc_fwd')
    return out.permute(0, 3, 1, 2)


def deform_im2col(im, offset, kernel=(3, 3), pad=(2, 2), stride=(1, 1), dilate=(2, 2), num_deformable_group=4):
    im = _f32(im, 'im'); offset = _f32(offset, 'offset')
    Cc, H, W = im.shape
    desc = L.DeformConvDesc(1, Cc, H, W, 1, kernel[0], kernel[1], pad[0], pad[1], stride[0], stride[1], dilate[0],
                            dilate[1], 1, num_deformable_group, 0)
    col = torch.empty((Cc * kernel[0] * kernel[1], offset.shape[1], offset.shape[2]), dtype=torch.float32, device=im.device)
    L.check(L.lib().rn_deform_im2col(C.byref(desc), _ptr(im), _ptr(offset), _ptr(col), _stream()), 'rn_deform_im2col')
    return col


def rpn_head(r, Wcls, bcls, Wbbox, bbbox, simt=False):
    """RPN head after rpn_conv: r = channels-last bf16 [1,Cin,h,w]; 1x1 conv weights/biases bf16 -> (rpn_cls_prob [1,2A,h,w],
    rpn_bbox_pred [1,4A,h,w]) fp32 NCHW, the {bg, fg} softmax applied (rn_rpn_head_fwd)."""
    ts = (r, Wcls, bcls, Wbbox, bbbox)
    if not all(t.is_cuda and t.dtype == torch.bfloat16 for t in ts) or r.dim() != 4 or r.shape[0] != 1 \
            or not r.is_contiguous(memory_format=torch.channels_last):
        raise L.RelnetError('rpn_head: bf16 CUDA tensors, r channels-last [1,Cin,h,w]')
    _, Cin, h, w = r.shape
    A = Wcls.shape[0] // 2
    if Wcls.numel() != 2 * A * Cin or Wbbox.numel() != 4 * A * Cin or bcls.numel() != 2 * A or bbbox.numel() != 4 * A:
        raise L.RelnetError('rpn_head: inconsistent weight shapes')
    prob = torch.empty((1, 2 * A, h, w), dtype=torch.float32, device=r.device)
    bbox = torch.empty((1, 4 * A, h, w), dtype=torch.float32, device=r.device)
    lib = L.lib()
    if device_info()['sm100'] and Cin % 8 == 0 and not simt:
        # tensor-core form: one bf16 tcgen05 GEMM over the concatenated heads + a softmax / layout kernel
        Wc, Wb = Wcls.contiguous(), Wbbox.contiguous()

        def pack(buf):
            L.check(lib.rn_rpn_head_pack(_ptr(Wc), _ptr(bcls), _ptr(Wb), _ptr(bbbox), Cin, A, _ptr(buf), _stream()), 'rn_rpn_head_pack')
        packed = _packs.get((Wcls, bcls, Wbbox, bbbox), lib.rn_rpn_head_packed_bytes(Cin, A), pack, tag=('rpn_head',))
        ws = _workspace(lib.rn_rpn_head_workspace_bytes(h * w, Cin, A), r.device)
        L.check(lib.rn_rpn_head_packed_fwd(_ptr(r), h * w, Cin, A, _ptr(packed), _ptr(prob), _ptr(bbox), _ptr(ws), ws.numel(),
                                           _stream()), 'rn_rpn_head_packed_fwd')
        return prob, bbox
    L.check(L.lib().rn_rpn_head_fwd(_ptr(r), h * w, Cin, A, _ptr(Wcls.contiguous()), _ptr(bcls), _ptr(Wbbox.contiguous()),
                                    _ptr(bbbox), _ptr(prob), _ptr(bbox), _stream()), 'rn_rpn_head_fwd')
    return prob, bbox


def image_s2d(image, pad=3):
    """fp32 [1,3,H,W] (or [3,H,W]) image -> bf16 NCHW-shaped, channels-last-strided [1,16,(H+2pad)/2,(W+2pad)/2]: the
    space-to-depth(2) form of the zero-padded image that turns the 7x7/2 stem conv into a 4x4/1 conv (rn_image_s2d_bf16)."""
    image = _f32(image, 'image')
    H, W = image.shape[-2], image.shape[-1]
    if image.numel() != 3 * H * W or (H + 2 * pad) % 2 or (W + 2 * pad) % 2:
        raise L.RelnetError('image_s2d: need one 3-channel image with even padded size, got %s' % (tuple(image.shape),))
    Hs, Ws = (H + 2 * pad) // 2, (W + 2 * pad) // 2
    out = torch.empty((1, Hs, Ws, 16), dtype=torch.bfloat16, device=image.device)
    L.check(L.lib().rn_image_s2d_bf16(_ptr(image), H, W, pad, _ptr(out), _stream()), 'rn_image_s2d_bf16')
    return out.permute(0, 3, 1, 2)


def maxpool3x3s2_nhwc(x):
    """3x3 / stride 2 / ceil-mode max pool of a channels-last bf16 [1,C,H,W] map (rn_maxpool3x3s2_nhwc_bf16)."""
    if not (x.is_cuda and x.dtype == torch.bfloat16 and x.dim() == 4 and x.shape[0] == 1
            and x.is_contiguous(memory_format=torch.channels_last) and x.shape[1] % 8 == 0):
        raise L.RelnetError('maxpool3x3s2_nhwc: need a channels-last bf16 CUDA tensor [1,C,H,W] with C % 8 == 0')
    _, Cc, H, W = x.shape
    Ho, Wo = (H - 2) // 2 + 1, (W - 2) // 2 + 1
    out = torch.empty((1, Ho, Wo, Cc), dtype=torch.bfloat16, device=x.device)
    L.check(L.lib().rn_maxpool3x3s2_nhwc_bf16(_ptr(x), H, W, Cc, _ptr(out), _stream()), 'rn_maxpool3x3s2_nhwc_bf16')
    return out.permute(0, 3, 1, 2)


def gemm_tf32(A, B, transA=False, transB=False, alpha=1.0, beta=0.0, out=None):
    """C = alpha * op(A) . op(B) + beta * C on the tcgen05 tf32 GEMM (rn_gemm_tf32).  A, B: fp32 CUDA tensors, 2-D or batched
    [..., rows, cols] with up to two leading batch dimensions (outer, inner); any strides the C ABI accepts."""
    for t, n in ((A, 'A'), (B, 'B')):          # strided views are the point: no .contiguous() here
        if not isinstance(t, torch.Tensor) or not t.is_cuda or t.dtype != torch.float32:
            raise L.RelnetError('%s must be a float32 CUDA tensor (relnet_b200 has no CPU path)' % n)
    def norm(t):
        while t.dim() < 4:
            t = t.unsqueeze(0)
        return t
    A4, B4 = norm(A), norm(B)
    outer, inner = A4.shape[0], A4.shape[1]
    M = A4.shape[3] if transA else A4.shape[2]
    K = A4.shape[2] if transA else A4.shape[3]
    N = B4.shape[2] if transB else B4.shape[3]
    if out is None:
        out = torch.zeros((outer, inner, M, N), dtype=torch.float32, device=A.device)
        ret = out.reshape(A.shape[:-2] + (M, N))
    else:
        ret = out
        out = norm(out)
    for t in (A4, B4, out):
        if t.stride(3) != 1:
            raise L.RelnetError('gemm_tf32: innermost dimension must be contiguous')
    L.check(L.lib().rn_gemm_tf32(int(transA), int(transB), M, N, K, float(alpha), _ptr(A4), A4.stride(2), A4.stride(0), A4.stride(1),
                                 _ptr(B4), B4.stride(2), B4.stride(0), B4.stride(1), float(beta), _ptr(out), out.stride(2),
                                 out.stride(0), out.stride(1), outer, inner, _stream()), 'rn_gemm_tf32')
    return ret


def umma_selftest(a, b, p, v):
    """tcgen05/TMA self-test: a,b [128,64], p [128,128], v [128,64] fp16 -> (a b^T [128,128], p v [128,64]) fp32."""
    for t in (a, b, p, v):
        assert t.is_cuda and t.dtype == torch.float16 and t.is_contiguous()
    s = torch.zeros((128, 128), dtype=torch.float32, device=a.device)
    o = torch.zeros((128, 64), dtype=torch.float32, device=a.device)
    L.check(L.lib().rn_umma_selftest(_ptr(a), _ptr(b), _ptr(p), _ptr(v), _ptr(s), _ptr(o), _stream()), 'rn_umma_selftest')
    return s, o
