"""GPU: ragged / degenerate / boundary inputs of the hot path (tile boundaries of the tcgen05 kernels, single rows,
padding branches, empty inputs).  Same oracle, same tolerances as test_gpu_parity.py."""
import numpy as np
import pytest
import torch
from conftest import rel_err
from oracle import relation_np as R, proposal_np as P, rois_np as RO, learn_nms_np as L

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def ops(cuda_device):
    import __graft_entry__ as g
    g.build()
    import relnet_b200
    torch.cuda.set_device(cuda_device)
    return relnet_b200.ops


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def precisions(ops):
    return ['fp32', 'f16'] if ops.device_info()['sm100'] else ['fp32']


def rel_args(c):
    return [c[k] for k in ('X', 'boxes', 'Wq', 'bq', 'Wk', 'bk', 'Wg', 'bg', 'Wout', 'bout')]


@pytest.mark.parametrize('N,M', [(1, 1), (2, 1), (127, 127), (128, 128), (129, 129), (130, 257), (300, 1), (513, 512),
                                 (640, 513)])
def test_relation_tile_boundaries(ops, N, M):
    """N, M around the 128-row tiles, one key, one query, M > N is impossible for a prefix so keys = first min(M, N)."""
    M = min(M, N)
    c = R.make_relation_case(1000 + 7 * N + M, N, 256, 4)
    args = rel_args(c)
    ref = R.relation_forward(*args, key_index=M, group=4, residual_relu=True, dtype=np.float32)
    for prec in precisions(ops):
        out = ops.relation(*[T(a) for a in args], M=M, group=4, residual_relu=True, precision=prec).cpu().numpy()
        assert np.isfinite(out).all()
        assert rel_err(out, ref) < 1e-3, (N, M, prec, rel_err(out, ref))


def test_relation_degenerate_boxes(ops):
    """identical boxes (eps = log 1e-3), zero-size boxes, boxes far apart; and an all-dead geometry row (g == 1e-6)."""
    c = R.make_relation_case(77, 64, 256, 4)
    b = c['boxes']
    b[:8] = b[0]                                   # identical
    b[8:12, 2:] = b[8:12, :2]                      # w = h = 1
    b[12] = [0, 0, 999, 599]; b[13] = [998, 598, 999, 599]
    c['bg'][:] = -10.0                             # relu kills every geometry unit -> uniform 1e-6 weights
    args = rel_args(c)
    ref = R.relation_forward(*args, group=4, dtype=np.float32)
    for prec in precisions(ops):
        out = ops.relation(*[T(a) for a in args], group=4, precision=prec).cpu().numpy()
        assert rel_err(out, ref) < 1e-3


def test_proposal_padding_branch_and_tiny_maps(ops):
    """fewer than post_nms_top_n survivors -> deterministic padding keep[i % kept] (the reference pads randomly)."""
    rng = np.random.default_rng(11)
    H, W, A = 3, 4, 12
    fg = (rng.permutation(A * H * W).astype(np.float64) + 0.5) / (A * H * W)
    cls_prob = np.concatenate([1 - fg, fg]).reshape(1, 2 * A, H, W).astype(np.float32)
    bbox_pred = (rng.standard_normal((1, 4 * A, H, W)) * 0.01).astype(np.float32)       # heavy overlap -> few survivors
    info = np.array([[48.0, 64.0, 1.0]], np.float32)
    rois, scores, nk = ops.proposal(T(cls_prob), T(bbox_pred), T(info), pre_nms_top_n=6000, post_nms_top_n=300,
                                    return_num_kept=True)
    o_rois, o_sc, aux = P.proposal_forward(cls_prob, bbox_pred, info, return_aux=True)
    assert aux['n_kept'] < 300 and int(nk.item()) == aux['n_kept']
    np.testing.assert_array_equal(rois.cpu().numpy(), o_rois)
    np.testing.assert_array_equal(scores.cpu().numpy(), o_sc)
    # min_size filter active + feature map larger than the image (the _clip_pad branch, proposal.py:184-197)
    info2 = np.array([[40.0, 50.0, 2.0]], np.float32)
    r2, s2 = ops.proposal(T(cls_prob), T(bbox_pred), T(info2), post_nms_top_n=20, min_size=16)
    o2, os2 = P.proposal_forward(cls_prob, bbox_pred, info2, post_nms_top_n=20, min_size=16)
    np.testing.assert_array_equal(r2.cpu().numpy(), o2)
    np.testing.assert_array_equal(s2.cpu().numpy(), os2)


def test_nms_and_overlaps_empty_and_single(ops):
    keep, num = ops.nms(torch.zeros((0, 5), device='cuda'), 0.7, max_keep=4)
    assert int(num.item()) == 0
    one = T(np.array([[0, 0, 10, 10, 0.9]], np.float32))
    keep, num = ops.nms(one, 0.7)
    assert int(num.item()) == 1 and int(keep[0].item()) == 0
    ov = ops.bbox_overlaps(torch.zeros((0, 4), device='cuda', dtype=torch.float64), T(np.zeros((3, 4))))
    assert tuple(ov.shape) == (0, 3)
    # touching boxes (iw == 0) and contained boxes
    b = np.array([[0, 0, 9, 9], [10, 0, 19, 9], [2, 2, 5, 5]], np.float64)
    np.testing.assert_allclose(ops.bbox_overlaps(T(b), T(b)).cpu().numpy(), P.bbox_overlaps(b, b), rtol=1e-15)


def test_proposal_target_single_gt_and_no_fg(ops):
    rng = np.random.default_rng(2)
    rois = np.hstack([np.zeros((20, 1), np.float32), R.make_boxes(rng, 20)])
    gt = np.array([[5000, 5000, 5100, 5100, 7]], np.float32)            # overlaps nothing -> only the gt row is fg
    ro, lab, bt, bw = ops.proposal_target(T(rois), T(gt))
    o = P.proposal_target_forward(rois, gt)
    np.testing.assert_array_equal(lab.cpu().numpy(), o[1])
    np.testing.assert_allclose(bt.cpu().numpy(), o[2], rtol=2e-6, atol=2e-6)
    assert int((lab > 0).sum().item()) == 1


def test_roi_ops_edge_rois(ops):
    rng = np.random.default_rng(4)
    data = rng.standard_normal((2, 8, 10, 12)).astype(np.float32)
    rois = np.array([[0, -50, -50, -10, -10],        # entirely outside -> empty bins -> 0
                     [1, 0, 0, 191, 159],            # whole image, second batch item
                     [0, 100, 80, 100, 80],          # single pixel
                     [1, 300, 300, 400, 400]],       # outside on the far side
                    np.float32)
    out, arg = ops.roi_pool(T(data), T(rois), return_argmax=True)
    o_ref, a_ref = RO.roi_pool(data, rois)
    np.testing.assert_array_equal(out.cpu().numpy(), o_ref)
    np.testing.assert_array_equal(arg.cpu().numpy(), a_ref)
    ps, cnt = ops.deform_psroi_pool(T(data), T(rois), output_dim=8, return_count=True)
    p_ref, c_ref = RO.deform_psroi_pool(data, rois, output_dim=8)
    np.testing.assert_array_equal(cnt.cpu().numpy(), c_ref)
    np.testing.assert_allclose(ps.cpu().numpy(), p_ref, rtol=1e-5, atol=1e-6)
    assert tuple(ops.roi_pool(T(data), torch.zeros((0, 5), device='cuda')).shape) == (0, 8, 7, 7)


def test_learn_nms_single_valid_class_and_small_n(ops):
    """one dominant class (every other class pruned), first_n == number of rois."""
    c = L.make_learn_nms_case(5, R=40, C=6, init='fan_in', n_peaky=1)
    w = {k: T(v) for k, v in c['P'].items()}
    ref = L.learn_nms_forward(c['cls_score'], c['bbox_pred'], c['rois'], c['im_info'], c['feat'], c['P'], first_n=40,
                              num_fg_classes=6, nongt_dim=40)
    for prec in precisions(ops):
        multi, sbbox, sscore, final = ops.learn_nms(T(c['cls_score']), T(c['bbox_pred']), T(c['rois']), T(c['im_info']),
                                                   T(c['feat']), w, first_n=40, nongt_dim=40, precision=prec)
        m = multi.cpu().numpy()
        assert np.array_equal(m.max(axis=(0, 2)) > 0, ref[0].max(axis=(0, 2)) > 0)
        assert rel_err(m, ref[0]) < 1e-3
        np.testing.assert_allclose(sscore.cpu().numpy(), ref[2], rtol=2e-5, atol=1e-8)


def test_errors_are_loud(ops):
    import relnet_b200
    Err = relnet_b200._lib.RelnetError
    c = R.make_relation_case(1, 16, 256, 4)
    a = [T(x) for x in rel_args(c)]
    with pytest.raises(Err):
        ops.relation(a[0], a[1], a[2][:, :100].contiguous(), *a[3:], group=4)        # Wq inner dim mismatch -> dq % H
    with pytest.raises(Err):
        ops.proposal(torch.zeros((1, 24, 200, 200), device='cuda'), torch.zeros((1, 48, 200, 200), device='cuda'),
                     torch.tensor([[3200.0, 3200.0, 1.0]], device='cuda'))            # 480000 anchors > sort capacity


def test_nms_multi_target_matches_reference_execution(ops):
    from conftest import golden
    g = golden('nms_multi_target')
    out = ops.nms_multi_target(T(g['bbox']), T(g['gt_box']), T(g['score']), list(g['target_thresh'])).cpu().numpy()
    np.testing.assert_array_equal(out, g['target'])
    # random second case against the oracle, incl. classes without gt and gts nobody overlaps
    rng = np.random.default_rng(3)
    bbox = R.make_boxes(rng, 150 * 5).reshape(150, 5, 4)
    gt = np.hstack([R.make_boxes(rng, 12), rng.integers(1, 7, (12, 1))]).astype(np.float32)
    bbox[:12, 2] = gt[:, :4] + rng.normal(0, 5, (12, 4)).astype(np.float32)
    score = rng.random((150, 5)).astype(np.float32)
    th = [0.5, 0.6, 0.7, 0.8, 0.9]
    out = ops.nms_multi_target(T(bbox), T(gt), T(score), th).cpu().numpy()
    np.testing.assert_array_equal(out, L.nms_multi_target(bbox, gt[None], score, th))
    import relnet_b200
    from relnet_b200 import compat
    o2 = compat.Custom(op_type='nms_multi_target', bbox=T(g['bbox']), gt_bbox=T(g['gt_box']), score=T(g['score']),
                       target_thresh='[0.5 0.6 0.7 0.8 0.9]')
    np.testing.assert_array_equal(o2.cpu().numpy(), g['target'])
