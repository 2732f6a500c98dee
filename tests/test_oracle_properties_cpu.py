"""Size-independent properties of the path, checked on the CPU oracle (small cases; the GPU suite checks the same properties on
the CUDA path at BASELINE's full sizes where no oracle run is affordable).  They follow from the reference's formulas:
SYM_REL:47-83 (geometry from box differences and ratios), SYM_REL:104-151 (softmax over keys, P.V, grouped output projection),
lib/nms/nms_kernel.cu (greedy sweep), operator_py/proposal.py:51-168 (decode, clip, sort, NMS, pad)."""
import numpy as np
import pytest

from oracle import relation_np as R, proposal_np as P, learn_nms_np as L


def case(seed=0, N=40, d=64, H=4, M=None):
    c = R.make_relation_case(seed, N, d, H, M=M)
    args = [c[k] for k in ('X', 'boxes', 'Wq', 'bq', 'Wk', 'bk', 'Wg', 'bg', 'Wout', 'bout')]
    return c, args


def fwd(args, **kw):
    return R.relation_forward(*args, dtype=np.float64, **kw)


def test_softmax_rows_sum_to_one_and_geometry_is_floored():
    c, args = case(1)
    r = fwd(args, group=4, return_all=True)
    np.testing.assert_allclose(r['softmax'].sum(axis=2), 1.0, rtol=1e-12)
    assert r['geom'].min() >= 1e-6 and r['softmax'].min() >= 0.0          # a dead unit is a soft mask (log 1e-6), never -inf


def test_geometry_is_translation_invariant():
    # eps uses centre DIFFERENCES over widths and width / height RATIOS: moving every box by the same vector changes nothing
    c, args = case(2)
    g0 = R.geometry_weight(c['boxes'], c['Wg'], c['bg'], dtype=np.float64)
    shifted = c['boxes'].astype(np.float64) + np.array([37.0, -11.0, 37.0, -11.0])
    g1 = R.geometry_weight(shifted, c['Wg'], c['bg'], dtype=np.float64)
    np.testing.assert_allclose(g1, g0, rtol=1e-9, atol=1e-12)


def test_module_is_equivariant_to_a_joint_permutation_of_the_rois():
    # N = M: permuting the rois permutes the outputs (keys are a set: the softmax sums over them)
    c, args = case(3)
    perm = np.random.default_rng(3).permutation(c['X'].shape[0])
    y0 = fwd(args, group=4, residual_relu=True)
    a2 = list(args); a2[0] = args[0][perm]; a2[1] = args[1][perm]
    y1 = fwd(a2, group=4, residual_relu=True)
    np.testing.assert_allclose(y1, y0[perm], rtol=1e-9, atol=1e-11)


def test_key_order_does_not_matter_and_index_list_equals_prefix():
    # FPN form: keys = take(rois, non_gt_index); an index list 0..M-1 is the `nongt_dim = M` slice, and its order is irrelevant
    c, args = case(4, N=30, M=22)
    y_prefix = fwd(args, group=4, key_index=22)
    y_list = fwd(args, group=4, key_index=np.arange(22))
    np.testing.assert_allclose(y_list, y_prefix, rtol=1e-12, atol=0)
    y_perm = fwd(args, group=4, key_index=np.random.default_rng(4).permutation(22))
    np.testing.assert_allclose(y_perm, y_prefix, rtol=1e-9, atol=1e-11)


def test_rows_are_separable_given_the_keys():
    # a query row's output depends on that row and on the key set only (what lets the kernels tile over queries)
    c, args = case(5)
    full = fwd(args, group=4)
    rows = np.array([0, 7, 19, 39])
    part = fwd(args, group=4, query_index=rows)
    np.testing.assert_allclose(part, full[rows], rtol=1e-12, atol=0)


def test_output_projection_is_linear_in_wout_and_bout():
    c, args = case(6)
    a2 = list(args); a2[8] = 2.0 * args[8]; a2[9] = 2.0 * args[9]
    np.testing.assert_allclose(fwd(a2, group=4), 2.0 * fwd(args, group=4), rtol=1e-12, atol=1e-14)


def test_reordered_form_equals_the_form_as_written():
    # project the keys first (V' = X_keys Wout^T) == P.V over the full d then the grouped 1x1 convolution
    c, args = case(7)
    a = R.relation_forward(*args, group=4, dtype=np.float64)
    b = R.relation_forward_reordered(*args, group=4, dtype=np.float64)
    np.testing.assert_allclose(b, a, rtol=1e-9, atol=1e-11)


# ------------------------------------------------------------------------------------------------ NMS / proposal
def _dets(seed, n):
    rng = np.random.default_rng(seed)
    b = R.make_boxes(rng, n)
    b[n // 2:] = b[:n - n // 2] + rng.normal(0, 2.0, (n - n // 2, 4)).astype(np.float32)      # near duplicates
    s = rng.permutation(n).astype(np.float32) / n
    o = np.argsort(-s, kind='stable')
    return np.hstack([b, s[:, None]])[o].astype(np.float32)


@pytest.mark.parametrize('thresh', [0.3, 0.7])
def test_nms_is_idempotent_and_its_survivors_are_pairwise_below_threshold(thresh):
    d = _dets(8, 400)
    keep = P.nms_sorted(d, thresh)
    kept = d[keep]
    np.testing.assert_array_equal(P.nms_sorted(kept, thresh), np.arange(len(keep)))            # nms(nms(x)) == nms(x)
    for i in range(len(keep) - 1):
        assert (P.iou_f32(kept[i, :4], kept[i + 1:, :4]) <= np.float32(thresh)).all()
    removed = np.setdiff1d(np.arange(len(d)), keep)
    for j in removed[:50]:                                                                     # every removed box has a kept, higher-ranked suppressor
        earlier = keep[keep < j]
        assert (P.iou_f32(d[j, :4], d[earlier, :4]) > np.float32(thresh)).any()
    assert keep[0] == 0                                                                        # the top-scoring box always survives


def test_nms_max_keep_is_a_prefix_and_threshold_is_monotone():
    d = _dets(9, 300)
    full = P.nms_sorted(d, 0.5)
    for mk in (1, 10, 10_000):
        np.testing.assert_array_equal(P.nms_sorted(d, 0.5, max_keep=mk), full[:mk])
    assert len(P.nms_sorted(d, 0.3)) <= len(full) <= len(P.nms_sorted(d, 0.7))


def test_proposal_outputs_are_clipped_sorted_and_padded_from_the_kept_set():
    cls_prob, bbox_pred, info = P.make_proposal_case(10, H=10, W=12, A=6, im_info=(160.0, 200.0, 1.0))
    rois, sc, aux = P.proposal_forward(cls_prob, bbox_pred, info, scales=(8, 16), pre_nms_top_n=200, post_nms_top_n=64, return_aux=True)
    k = aux['n_kept']
    assert rois.shape == (64, 5) and sc.shape == (64, 1) and (rois[:, 0] == 0).all()
    assert (rois[:, 1] >= 0).all() and (rois[:, 2] >= 0).all() and (rois[:, 3] <= 199).all() and (rois[:, 4] <= 159).all()
    assert (np.diff(sc[:k, 0]) <= 0).all()                                                     # score-descending among the kept
    np.testing.assert_array_equal(rois[k:], rois[np.arange(k, 64) % k])                        # padding = keep[i % kept]
    keep_again = P.nms_sorted(np.hstack([rois[:k, 1:], sc[:k]]), 0.7)
    np.testing.assert_array_equal(keep_again, np.arange(k))                                    # the kept set is NMS-stable


def test_bbox_overlaps_is_symmetric_bounded_and_one_on_the_diagonal():
    rng = np.random.default_rng(11)
    b = R.make_boxes(rng, 50).astype(np.float64)
    ov = P.bbox_overlaps(b, b)
    np.testing.assert_allclose(ov, ov.T, rtol=1e-14)
    np.testing.assert_allclose(np.diag(ov), 1.0, rtol=1e-14)
    assert ov.min() >= 0.0 and ov.max() <= 1.0 + 1e-14


def test_encode_then_decode_returns_the_target_box():
    rng = np.random.default_rng(12)
    ex, gt = R.make_boxes(rng, 64).astype(np.float64), R.make_boxes(rng, 64).astype(np.float64)
    back = P.decode_boxes(ex, P.encode_boxes(ex, gt))
    np.testing.assert_allclose(back, gt, rtol=0, atol=1e-4)          # decode is float32 inside (proposal.py path): sub-pixel round trip


# ------------------------------------------------------------------------------------------------ learn-NMS head
def test_learn_nms_outputs_are_rank_sorted_probabilities_and_pruned_classes_are_zero():
    c = L.make_learn_nms_case(13, R=40, C=6, init='fan_in', n_peaky=3)
    out = L.learn_nms_forward(c['cls_score'], c['bbox_pred'], c['rois'], c['im_info'], c['feat'], c['P'], first_n=20, num_fg_classes=6)
    multi, sbbox, sscore = out[0], out[1], out[2]
    assert multi.shape == (20, 6, 5) and sbbox.shape == (20, 6, 4) and sscore.shape == (20, 6)
    assert (np.diff(sscore, axis=0) <= 0).all()                                               # per class, score-descending ranks
    assert multi.min() >= 0.0 and (multi <= sscore[:, :, None] + 1e-7).all()                  # s1 = sigmoid(.) * s0 <= s0
    dead = multi.max(axis=(0, 2)) == 0
    assert (sscore[0, dead] < 0.01 + 1e-7).all()                                              # pruned = no rank reaches class_thresh (LNMS:298-303)
    assert (sbbox[..., 0] <= sbbox[..., 2]).all() and (sbbox[..., 1] <= sbbox[..., 3]).all()
