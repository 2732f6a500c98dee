"""GPU: rn_relation_bwd against autograd through the torch restatement of the module (oracle/relation_torch.py).

The oracle's forward is pinned to the numpy oracle (tests/test_oracle_golden.py), which is pinned by reference execution;
the reference's own backward IS autograd over the same graph (MXNet), so autograd over the pinned forward is the
reference gradient.  Target: the FLOAT32 autograd gradients (the reference's arithmetic type), max-norm relative error
<= 2e-3 per gradient tensor; the float64 gradients are printed for information (the float32 geometry is ill-conditioned
for near-concentric boxes, see test_gpu_parity.py)."""
import numpy as np
import pytest
import torch
from conftest import rel_err
from oracle import relation_np as R, relation_torch as RT

pytestmark = pytest.mark.gpu
NAMES = ('X', 'Wq', 'bq', 'Wk', 'bk', 'Wg', 'bg', 'Wout', 'bout')


@pytest.fixture(scope='module')
def ops(cuda_device):
    import __graft_entry__ as g
    g.build()
    import relnet_b200
    torch.cuda.set_device(cuda_device)
    return relnet_b200.ops


def autograd_grads(c, dOut, dtype, **kw):
    t = {k: torch.tensor(c[k], dtype=dtype, requires_grad=(k != 'boxes')) for k in NAMES + ('boxes',)}
    out = RT.relation_forward(t['X'], t['boxes'], t['Wq'], t['bq'], t['Wk'], t['bk'], t['Wg'], t['bg'], t['Wout'], t['bout'], **kw)
    out.backward(torch.tensor(dOut, dtype=dtype))
    return {k: t[k].grad.numpy() for k in NAMES}, out.detach().numpy()


# Contraction engines of the backward (ops.relation_backward(precision=...)): 'fp32' = cuBLAS fp32 (parity mode, held to 2e-3
# of float32 autograd, measured 2e-7..4e-5); 'f16' = the library's tcgen05 tf32 GEMM on the fp32 operands (10-bit mantissa
# operands as the tensor core truncates them from fp32, fp32 accumulate; the recomputed forward runs on it too: per-tensor
# tolerance 1e-2, written here; measured 1e-4 .. 6e-3, the largest on the gradients that sum over all 8000 learn-NMS rows).
# KINK: bound for comparisons where the two sides may take different relu branches on a few units (see the learn-NMS test).
GRAD_TOL = {'fp32': 2e-3, 'f16': 1e-2}
KINK = 0.25
ZERO_TOL = {'fp32': 1e-4, 'f16': 3e-3}       # gradients that are exactly zero in exact arithmetic, relative to a sibling's scale


def grad_precisions(ops):
    return ['fp32', 'f16'] if ops.device_info()['sm100'] else ['fp32']


@pytest.mark.parametrize('tA,tB,M,N,K,outer,inner', [
    (False, True, 300, 1024, 1024, 1, 1),     # dX = dQ Wq^T-like (both K-major)
    (True, False, 1024, 1024, 300, 1, 1),     # dW = dQ^T X (both MN-major), K tail 300 = 9 x 32 + 12
    (False, False, 300, 64, 300, 1, 16),      # dQ_h = dS_h K_h, heads as the inner batch of column slices
    (True, False, 300, 64, 300, 2, 16),       # dV'_h = P_h^T dO_h, two-level batch
    (False, True, 100, 100, 8, 80, 16),       # learn-NMS dP = dO V'^T with d_v = 8: 1280 problems, K = 8
    (True, False, 100, 8, 100, 80, 16),       # learn-NMS dV' (N = 8 < one 32-wide atom)
    (False, True, 131, 77, 45, 1, 3),         # ragged everything
])
def test_gemm_tf32_matches_float64(ops, tA, tB, M, N, K, outer, inner):
    """rn_gemm_tf32 (the training side's contraction engine) vs float64 matmul of the same fp32 operands: relative error of a
    tf32 product sum (operands rounded to 10-bit mantissas, fp32 accumulate) <= 2e-3 of the result's max; alpha / beta; head
    slices of a wider matrix as the inner batch (strides that are NOT the matrix size)."""
    g = torch.Generator(device='cpu').manual_seed(M * 7 + N * 3 + K)
    def operand(rows, cols):
        # inner batch = column slices of one wide matrix (like the heads of Q / dO), outer = separate matrices
        wide = torch.randn((outer, rows, inner * ((cols + 3) // 4 * 4)), generator=g).cuda()
        cw = (cols + 3) // 4 * 4
        return wide.as_strided((outer, inner, rows, cols), (wide.stride(0), cw, wide.stride(1), 1))
    A = operand(K, M) if tA else operand(M, K)
    B = operand(N, K) if tB else operand(K, N)
    C0 = torch.randn((outer, inner, M, N), generator=g).cuda()
    ref = 0.5 * torch.matmul((A.transpose(-1, -2) if tA else A).double(), (B.transpose(-1, -2) if tB else B).double()) + 2.0 * C0.double()
    out = C0.clone()
    ops.gemm_tf32(A, B, transA=tA, transB=tB, alpha=0.5, beta=2.0, out=out)
    torch.cuda.synchronize()
    e = float((out.double() - ref).abs().max() / ref.abs().max())
    print('gemm_tf32 tA=%d tB=%d %dx%dx%d x(%d,%d): rel err %.2e' % (tA, tB, M, N, K, outer, inner, e))
    assert e <= 2e-3, e
    out2 = ops.gemm_tf32(A, B, transA=tA, transB=tB)          # beta = 0 never reads C
    ref2 = torch.matmul((A.transpose(-1, -2) if tA else A).double(), (B.transpose(-1, -2) if tB else B).double())
    assert float((out2.double() - ref2).abs().max() / ref2.abs().max()) <= 2e-3


@pytest.mark.parametrize('prec', ['fp32', 'f16'])
@pytest.mark.parametrize('seed,N,d,H,M,kidx,res', [
    (11, 70, 256, 4, 50, False, True),        # key prefix (FPN form, SYM_REL fpn :104-151), residual + relu
    (12, 300, 1024, 16, None, False, True),   # the module at its training size
    (13, 131, 256, 16, None, True, False),    # permuted key_index (learn-NMS form), no residual
])
def test_relation_backward_matches_autograd(ops, seed, N, d, H, M, kidx, res, prec):
    if prec not in grad_precisions(ops):
        pytest.skip('tcgen05 needs sm_100')
    c = R.make_relation_case(seed, N, d, H, M=M)
    rng = np.random.RandomState(seed)
    key_index = rng.permutation(N)[:97].astype(np.int32) if kidx else None
    dOut = rng.randn(N, d).astype(np.float32)
    kw = dict(group=H, residual_relu=res)
    kw['key_index'] = key_index if kidx else (M if M is not None else None)
    g32, out32 = autograd_grads(c, dOut, torch.float32, **kw)
    g64, _ = autograd_grads(c, dOut, torch.float64, **kw)
    dev = {k: torch.from_numpy(np.ascontiguousarray(c[k])).cuda() for k in NAMES + ('boxes',)}
    kd = torch.from_numpy(key_index).cuda() if kidx else None
    wargs = [dev[k] for k in ('Wq', 'bq', 'Wk', 'bk', 'Wg', 'bg', 'Wout', 'bout')]
    # autograd semantics: the relu mask is the sign pattern of the forward that was executed (here the fp32 forward, equal to
    # the oracle's to 1e-6).  Without it the mask comes from the backward's own recomputed forward, and under the tf32 engine
    # the few units within 1e-3 of the kink flip -- each flip moves a whole dOut entry (measured: 2 of 17920 units, max-norm
    # error 0.3 in dX from those two alone; tools/bwd_probe.py).
    fwd = ops.relation(dev['X'], dev['boxes'], *wargs, key_index=kd, M=M, group=H, residual_relu=res, precision='fp32') if res else None
    got = ops.relation_backward(torch.from_numpy(dOut).cuda(), dev['X'], dev['boxes'], *wargs, key_index=kd, M=M, group=H,
                                residual_relu=res, precision=prec, forward_out=fwd)
    torch.cuda.synchronize()
    bad = []
    for k in NAMES:
        a = got[k].cpu().numpy().reshape(g32[k].shape)
        if k == 'bk':
            # exactly zero in exact arithmetic (a key bias shifts every logit of a row alike; softmax is shift-invariant):
            # both sides are rounding noise, so hold it to the scale of the query-bias gradient instead
            scale = np.abs(g32['bq']).max()
            print('dbk    |got| %.2e  |float32 autograd| %.2e  (dbq scale %.2e)' % (np.abs(a).max(), np.abs(g32[k]).max(), scale))
            if np.abs(a).max() > ZERO_TOL[prec] * scale:
                bad.append((k, float(np.abs(a).max() / scale)))
            continue
        e32 = rel_err(a, g32[k])
        e64 = rel_err(a, g64[k])
        print('[%s] d%-5s vs float32 autograd %.2e   vs float64 %.2e' % (prec, k, e32, e64))
        if e32 > GRAD_TOL[prec]:
            bad.append((k, e32))
    assert not bad, bad


def test_relation_backward_batched_and_loud(ops):
    """batch = 2 problems == two single calls with the parameter gradients summed; bad shapes raise."""
    import relnet_b200
    H, N, d = 4, 40, 128
    cs = [R.make_relation_case(s, N, d, H) for s in (21, 22)]
    P = {k: torch.from_numpy(cs[0][k]).cuda() for k in NAMES if k != 'X'}
    X = torch.stack([torch.from_numpy(c['X']) for c in cs]).cuda()
    boxes = torch.stack([torch.from_numpy(c['boxes']) for c in cs]).cuda()
    dOut = torch.randn(2, N, d, device='cuda')
    args = [P[k] for k in ('Wq', 'bq', 'Wk', 'bk', 'Wg', 'bg', 'Wout', 'bout')]
    both = ops.relation_backward(dOut, X, boxes, *args, group=H, residual_relu=True)
    singles = [ops.relation_backward(dOut[b], X[b], boxes[b], *args, group=H, residual_relu=True) for b in range(2)]
    for k in NAMES:
        want = torch.stack([s['X'] for s in singles]) if k == 'X' else singles[0][k] + singles[1][k]
        if k == 'bk':
            assert both[k].abs().max().item() <= 3e-3 * both['bq'].abs().max().item()
            continue
        assert rel_err(both[k].cpu().numpy(), want.cpu().numpy()) <= 1e-5, k
    with pytest.raises(relnet_b200._lib.RelnetError):
        ops.relation_backward(dOut[0, :, :7], X[0], boxes[0], *args, group=H)


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_roi_pool_backward_matches_oracle(ops):
    from oracle import rois_np as RO
    rng = np.random.RandomState(31)
    data = rng.randn(2, 64, 38, 63).astype(np.float32)
    R = 200
    x1 = rng.uniform(0, 800, R); y1 = rng.uniform(0, 450, R)
    rois = np.stack([rng.randint(0, 2, R), x1, y1, np.minimum(x1 + rng.uniform(8, 500, R), 999),
                     np.minimum(y1 + rng.uniform(8, 400, R), 599)], 1).astype(np.float32)
    out, arg = ops.roi_pool(T(data), T(rois), (7, 7), 0.0625, return_argmax=True)
    o_ref, a_ref = RO.roi_pool(data, rois, (7, 7), 0.0625)
    assert np.array_equal(arg.cpu().numpy(), a_ref)
    dout = rng.randn(*o_ref.shape).astype(np.float32)
    got = ops.roi_pool_backward(T(dout), arg, T(rois), data.shape).cpu().numpy()
    want = RO.roi_pool_backward(dout, a_ref, rois, data.shape)
    assert rel_err(got, want) <= 1e-5
    assert ops.roi_pool_backward(T(dout[:0]), arg[:0], T(rois[:0]), data.shape).abs().max().item() == 0.0   # no rois


@pytest.mark.parametrize('with_trans', [False, True])
def test_deform_psroi_pool_backward_matches_oracle(ops, with_trans):
    from oracle import rois_np as RO
    rng = np.random.RandomState(32)
    B, C, H, W, R = 1, 256, 38, 63, 128
    data = rng.randn(B, C, H, W).astype(np.float32)
    x1 = rng.uniform(0, 800, R); y1 = rng.uniform(0, 450, R)
    rois = np.stack([np.zeros(R), x1, y1, np.minimum(x1 + rng.uniform(8, 500, R), 999),
                     np.minimum(y1 + rng.uniform(8, 400, R), 599)], 1).astype(np.float32)
    trans = (rng.randn(R, 2, 7, 7)).astype(np.float32) if with_trans else None
    kw = dict(spatial_scale=0.0625, output_dim=C, group_size=1, pooled_size=7, sample_per_part=4, trans_std=0.1)
    out, cnt = RO.deform_psroi_pool(data, rois, trans, **kw)
    dout = rng.randn(*out.shape).astype(np.float32)
    want_d, want_t = RO.deform_psroi_pool_backward(dout, cnt, data, rois, trans, **kw)
    o, c = ops.deform_psroi_pool(T(data), T(rois), T(trans) if with_trans else None, return_count=True, **kw)
    assert np.array_equal(c.cpu().numpy(), cnt)
    dd, dt = ops.deform_psroi_pool_backward(T(dout), c, T(data), T(rois), T(trans) if with_trans else None, **kw)
    assert rel_err(dd.cpu().numpy(), want_d) <= 1e-5        # atomics: summation order only
    if with_trans:
        assert rel_err(dt.cpu().numpy(), want_t) <= 1e-4
    else:
        assert dt is None


@pytest.mark.parametrize('deformed', [False, True])
def test_deform_conv_backward_matches_oracle(ops, deformed):
    from oracle import rois_np as RO
    rng = np.random.RandomState(33)
    B, C, H, W, Co, dg = 2, 32, 19, 23, 24, 4
    data = rng.randn(B, C, H, W).astype(np.float32)
    offset = (rng.randn(B, dg * 18, H, W) * 1.5).astype(np.float32)      # many samples leave the map / hit the borders
    weight = (rng.randn(Co, C, 3, 3) * 0.1).astype(np.float32)
    dout = rng.randn(B, Co, H, W).astype(np.float32)
    want = RO.deform_conv_backward(dout, data, offset, weight, num_deformable_group=dg, weight_grad_deformed=deformed,
                                   has_bias=True)
    got = ops.deform_conv_backward(T(dout), T(data), T(offset), T(weight), num_deformable_group=dg,
                                   weight_grad_deformed=deformed, has_bias=True)
    for g, w, name in zip(got, want, ('data', 'offset', 'weight', 'bias')):
        e = rel_err(g.cpu().numpy(), w)
        print('deform_conv d%-6s %.2e' % (name, e))
        assert e <= 2e-5, name


def test_deform_conv_weight_grad_modes_differ(ops):
    """the reference's plain-im2col dWeight and the deformed one are different functions once offsets are non-zero, and
    coincide for zero offsets (the reference's offset convs are zero-initialised, SYM_DCN_REL_NMS:1525-1530)"""
    rng = np.random.RandomState(35)
    data = T(rng.randn(1, 8, 12, 14).astype(np.float32)); weight = T((rng.randn(6, 8, 3, 3) * 0.1).astype(np.float32))
    dout = T(rng.randn(1, 6, 12, 14).astype(np.float32))
    off = T((rng.randn(1, 36, 12, 14)).astype(np.float32))
    a = ops.deform_conv_backward(dout, data, off, weight, num_deformable_group=2, weight_grad_deformed=False)[2]
    b = ops.deform_conv_backward(dout, data, off, weight, num_deformable_group=2, weight_grad_deformed=True)[2]
    assert rel_err(a.cpu().numpy(), b.cpu().numpy()) > 1e-2
    z = torch.zeros_like(off)
    a = ops.deform_conv_backward(dout, data, z, weight, num_deformable_group=2, weight_grad_deformed=False)[2]
    b = ops.deform_conv_backward(dout, data, z, weight, num_deformable_group=2, weight_grad_deformed=True)[2]
    assert rel_err(a.cpu().numpy(), b.cpu().numpy()) <= 1e-6


def test_deform_conv_backward_grouped(ops):
    from oracle import rois_np as RO
    rng = np.random.RandomState(34)
    B, C, H, W, Co, dg, G = 1, 16, 10, 12, 8, 2, 2
    data = rng.randn(B, C, H, W).astype(np.float32)
    offset = (rng.randn(B, dg * 18, H, W) * 0.7).astype(np.float32)
    weight = (rng.randn(Co, C // G, 3, 3) * 0.1).astype(np.float32)
    dout = rng.randn(B, Co, H, W).astype(np.float32)
    want = RO.deform_conv_backward(dout, data, offset, weight, num_deformable_group=dg, num_group=G)
    got = ops.deform_conv_backward(T(dout), T(data), T(offset), T(weight), num_deformable_group=dg, num_group=G)
    for g, w, name in zip(got[:3], want, ('data', 'offset', 'weight')):
        assert rel_err(g.cpu().numpy(), w) <= 2e-5, name


@pytest.mark.parametrize('prec,shift', [('fp32', 0.0), ('f16', 0.0), ('f16', 50.0)])
@pytest.mark.parametrize('seed,R,C,d,n,nongt', [(41, 60, 8, 256, 20, 50), (42, 300, 80, 1024, 100, 300)])
def test_learn_nms_backward_matches_autograd(ops, seed, R, C, d, n, nongt, prec, shift):
    """shift: added to nms_linear_out_1_bias.  The head's inner relation ends in relu(f + o): under the tf32 engine the units
    within ~1e-3 of that kink take the other branch than the float32 oracle (the head recomputes its own forward; there is
    no executed-forward output to take the mask from), and each flip switches a whole gradient entry.  shift = 50 puts every
    unit on the active branch, so that case holds the tf32 engine to its strict per-tensor tolerance; shift = 0 (the real
    head) is held to it under the fp32 engine, and under tf32 only to a plumbing-level bound with the numbers printed."""
    if prec not in grad_precisions(ops):
        pytest.skip('tcgen05 needs sm_100')
    """rn_learn_nms_bwd vs autograd through oracle/learn_nms_torch.py (forward pinned to the numpy oracle / reference)."""
    from oracle import learn_nms_np as LN, learn_nms_torch as LT
    c = LN.make_learn_nms_case(seed, R=R, C=C, d=d)
    if d != 1024:
        c['P']['roi_feat_embedding_weight'] = c['P']['roi_feat_embedding_weight'][:, :d].copy()
    c['P']['nms_linear_out_1_bias'] = (c['P']['nms_linear_out_1_bias'] + shift).astype(np.float32)
    tol = GRAD_TOL[prec] if (prec == 'fp32' or shift > 0) else KINK
    rng = np.random.RandomState(seed)
    d_multi = rng.randn(n, C, 5).astype(np.float32)
    means, stds = (0.0, 0.0, 0.0, 0.0), (0.1, 0.1, 0.2, 0.2)

    def autograd(dtype):
        P = {k: torch.tensor(v, dtype=dtype, requires_grad=True) for k, v in c['P'].items()}
        cs = torch.tensor(c['cls_score'], dtype=dtype, requires_grad=True)
        ft = torch.tensor(c['feat'], dtype=dtype, requires_grad=True)
        multi, _, _ = LT.learn_nms_forward(cs, c['bbox_pred'], c['rois'], c['im_info'], ft, P, first_n=n, num_fg_classes=C,
                                           class_thresh=0.0, means=means, stds=stds, nongt_dim=nongt)
        multi.backward(torch.tensor(d_multi, dtype=dtype))
        return {k: v.grad.numpy() for k, v in P.items()}, cs.grad.numpy(), ft.grad.numpy(), multi.detach().numpy()
    gP, gcs, gft, multi_ref = autograd(torch.float32)
    gP64 = autograd(torch.float64)[0]
    W = {k: T(v) for k, v in c['P'].items()}
    args = (T(c['cls_score']), T(c['bbox_pred']), T(c['rois']), T(c['im_info']), T(c['feat']), W)
    kw = dict(first_n=n, class_thresh=0.0, means=means, stds=stds, nongt_dim=nongt)
    multi = ops.learn_nms(*args, precision='fp32', **kw)[0]
    assert rel_err(multi.cpu().numpy(), multi_ref) <= 1e-4
    grads, d_cls, d_feat = ops.learn_nms_backward(T(d_multi), *args, precision=prec, **kw)
    torch.cuda.synchronize()
    worst = 0.0
    for k in gP:
        want = gP[k]
        got = grads[k].cpu().numpy().reshape(want.shape)
        if k == 'nms_key_1_bias':          # exactly zero in exact arithmetic (softmax shift invariance): rounding noise
            assert np.abs(got).max() <= ZERO_TOL[prec] * np.abs(gP['nms_query_1_bias']).max()
            continue
        e = rel_err(got, want)
        e64, ref64 = rel_err(got, gP64[k]), rel_err(want, gP64[k])
        print('[%s shift %g] d%-28s vs float32 autograd %.2e | vs float64 %.2e (float32 autograd itself: %.2e)' % (prec, shift, k, e, e64, ref64))
        worst = max(worst, e)
        # the geometry-FC gradient carries 1/g weights up to 1e6 next to the 1e-6 clamp: float32 evaluations of it differ
        # among themselves, so it is held to the float32 oracle's OWN distance from exact arithmetic instead
        assert e <= tol or e64 <= 3.0 * ref64 + 1e-4, (k, e, e64, ref64)
    e_cs, e_ft = rel_err(d_cls.cpu().numpy(), gcs), rel_err(d_feat.cpu().numpy(), gft)
    print('[%s] d_cls_score %.2e  d_feat %.2e' % (prec, e_cs, e_ft))
    assert e_cs <= tol and e_ft <= tol
    assert np.abs(d_cls.cpu().numpy()[nongt:]).max(initial=0.0) == 0.0      # gt rows are outside the non-gt slice


def test_nms_loss_and_ohem_match_oracle(ops):
    from conftest import golden
    from oracle import train_np as TN
    rng = np.random.default_rng(5)
    m = rng.uniform(0.0, 1.0, (100, 80, 5)).astype(np.float32); t = (rng.random((100, 80, 5)) < 0.05).astype(np.float32)
    m[0, 0, 0] = 0.0; m[0, 0, 1] = 1.0                          # the eps guards
    pos, neg, d = ops.nms_loss(T(m), T(t), loss_scale=1.0, pos_grad_scale=4.0)
    wp, wn, wd = TN.nms_loss(m, t, 100, 5)
    assert rel_err(pos.cpu().numpy(), wp) <= 1e-5 and rel_err(neg.cpu().numpy(), wn) <= 1e-5
    assert rel_err(d.cpu().numpy(), wd) <= 1e-5
    g = golden('box_annotator_ohem')
    lab, w, loss = ops.box_annotator_ohem(T(g['cls_score']), T(g['bbox_pred']), T(g['labels']), T(g['bbox_targets']),
                                          T(g['bbox_weights']), 81, 2, int(g['roi_per_img']), return_loss=True)
    assert np.array_equal(lab.cpu().numpy(), g['labels_ohem'])               # bit-exact vs the reference execution
    assert np.array_equal(w.cpu().numpy(), g['bbox_weights_ohem'])
    assert rel_err(loss.cpu().numpy(), TN.box_annotator_ohem(g['cls_score'], g['bbox_pred'], g['labels'], g['bbox_targets'],
                                                             g['bbox_weights'], 128)[2]) <= 1e-6


def test_backward_writes_into_gradient_bucket(ops):
    """the backward kernels fill windows of one flat bucket in place (what the NCCL allreduce then sums)"""
    from relnet_b200 import replicas
    H, N, d = 4, 40, 128
    c = R.make_relation_case(51, N, d, H)
    dev = {k: T(c[k]) for k in NAMES + ('boxes',)}
    names = ('Wq', 'bq', 'Wk', 'bk', 'Wg', 'bg', 'Wout', 'bout')
    bucket = replicas.GradientBucket({k: tuple(c[k].shape) for k in names}, device='cuda')
    dOut = torch.randn(N, d, device='cuda')
    args = [dev[k] for k in names]
    ref = ops.relation_backward(dOut, dev['X'], dev['boxes'], *args, group=H, residual_relu=True)
    got = ops.relation_backward(dOut, dev['X'], dev['boxes'], *args, group=H, residual_relu=True, out=bucket.views)
    torch.cuda.synchronize()
    for k in names:
        assert got[k].data_ptr() == bucket.views[k].data_ptr()
        assert torch.equal(bucket.views[k], ref[k])
    assert bucket.allreduce() is None                     # single process: no group, nothing to do


@pytest.mark.parametrize('gprec', ['fp32', 'f16'])
def test_autograd_bindings_train_a_small_head(ops, gprec):
    """torch.autograd Functions over the C-ABI fwd/bwd pairs: a 2FC + relation + learn-NMS head differentiates end to end
    and agrees with autograd through the torch oracles (float32).  gprec = contraction engine of the backward: fp32 is held
    to 5e-3; the tf32 engine to a plumbing-level bound KINK (the learn-NMS head's inner relu kink, see
    test_learn_nms_backward_matches_autograd) with the per-tensor numbers printed."""
    if gprec not in grad_precisions(ops):
        pytest.skip('tcgen05 needs sm_100')
    from relnet_b200 import autograd as AG
    from oracle import learn_nms_np as LN, learn_nms_torch as LT
    H, R_, d, C, n = 4, 50, 256, 6, 16
    c = R.make_relation_case(61, R_, d, H)
    l = LN.make_learn_nms_case(62, R=R_, C=C, d=d)
    l['P']['roi_feat_embedding_weight'] = l['P']['roi_feat_embedding_weight'][:, :d].copy()
    rng = np.random.RandomState(63)
    Wc = (rng.randn(C + 1, d) * 0.05).astype(np.float32)
    tgt = (rng.rand(n, C, 5) < 0.1).astype(np.float32)
    rel_names = ('Wq', 'bq', 'Wk', 'bk', 'Wg', 'bg', 'Wout', 'bout')

    def run(device, fwd):
        f = torch.float32
        X = torch.tensor(c['X'], dtype=f, device=device, requires_grad=True)
        boxes = torch.tensor(l['rois'][:, 1:], dtype=f, device=device)
        P = {k: torch.tensor(c[k], dtype=f, device=device, requires_grad=True) for k in rel_names}
        Wcls = torch.tensor(Wc, dtype=f, device=device, requires_grad=True)
        Q = {k: torch.tensor(v, dtype=f, device=device, requires_grad=True) for k, v in l['P'].items()}
        A = fwd['relation'](X, boxes, *[P[k] for k in rel_names])
        cls_score = A @ Wcls.T + torch.tensor(l['cls_score'], dtype=f, device=device)    # keeps the peaked classes
        multi = fwd['learn_nms'](cls_score, A, Q)
        t = torch.tensor(tgt, dtype=f, device=device)
        loss = (4.0 * -(t * torch.log(multi + 1e-8)) - (1 - t) * torch.log(1 - multi + 1e-8)).sum() / (n * 5)
        loss.backward()
        g = {'X': X.grad, 'Wcls': Wcls.grad}
        g.update({k: P[k].grad for k in rel_names}); g.update({k: Q[k].grad for k in Q})
        return float(loss), {k: v.detach().cpu().numpy() for k, v in g.items()}

    bp, rois, info = l['bbox_pred'], l['rois'], l['im_info']
    ours = dict(relation=lambda X, b, *w: AG.relation(X, b, *w, group=H, residual_relu=True, grad_precision=gprec),
                learn_nms=lambda cs, A, Q: AG.learn_nms(cs, T(bp), T(rois), T(info), A, Q, first_n=n, grad_precision=gprec)[0])
    oracle = dict(relation=lambda X, b, *w: RT.relation_forward(X, b, *w, group=H, residual_relu=True),
                  learn_nms=lambda cs, A, Q: LT.learn_nms_forward(cs, bp, rois, info, A, Q, first_n=n, num_fg_classes=C)[0])
    loss_g, g_g = run('cuda', ours)
    loss_o, g_o = run('cpu', oracle)
    assert abs(loss_g - loss_o) <= 1e-4 * abs(loss_o)
    for k in g_o:
        if k in ('bk', 'nms_key_1_bias'):
            continue                                    # identically zero in exact arithmetic
        e = rel_err(g_g[k], g_o[k])
        print('[grad engine %s] %-28s %.2e' % (gprec, k, e))
        assert e <= (5e-3 if gprec == 'fp32' else KINK), (k, e)
