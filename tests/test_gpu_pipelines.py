"""GPU: the assembled test-time pipelines of BASELINE.json configs[2] (Deformable Faster-RCNN head) and configs[3] (FPN head)
at their real sizes against the numpy / C pipeline oracle (oracle/pipeline_np.py), plus a TIE-AWARE comparison of the
learn-NMS outputs for configs[1].

The learn-NMS outputs are indexed by (rank, class): two pipelines whose class scores differ in the last bits can order
near-tied rois differently, which permutes rows without changing any (roi, class) value.  `rank_aligned_err` therefore
re-identifies every output row by its sorted box (the roi it belongs to) before comparing scores, and reports the fraction
of rows it could align -- no sorting of the values themselves.
"""
import numpy as np
import pytest
import torch
from conftest import rel_err
from oracle import learn_nms_np as L, pipeline_np, proposal_np as P, relation_np as R

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def ops(cuda_device):
    import __graft_entry__ as g
    g.build()
    import relnet_b200
    torch.cuda.set_device(cuda_device)
    return relnet_b200.ops


def precisions(ops):
    return ['fp32', 'f16'] if ops.device_info()['sm100'] else ['fp32']


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def rank_aligned_err(final, sbbox, rfinal, rsbbox, box_tol=0.05):
    """final / rfinal [n,C]; sbbox / rsbbox [n,C,4].  For every class: row i of `final` is matched to the reference row
    holding the same box (L-inf distance < box_tol px).  Returns (max |score - ref| / max |ref| over matched rows,
    fraction of rows matched)."""
    n, C = final.shape
    worst, matched = 0.0, 0
    scale = max(float(np.abs(rfinal).max()), 1e-30)
    for c in range(C):
        d = np.abs(sbbox[:, None, c, :] - rsbbox[None, :, c, :]).max(-1)          # [n, n]
        j = d.argmin(1)
        ok = d[np.arange(n), j] < box_tol
        matched += int(ok.sum())
        if ok.any():
            worst = max(worst, float(np.abs(final[ok, c] - rfinal[j[ok], c]).max()))
    return worst / scale, matched / float(n * C)


def _params(seed):
    import relnet_b200
    from relnet_b200.pipeline import init_head_params
    return init_head_params(seed, 'cpu')


def _rois(seed, n, W=1000.0, H=600.0):
    rng = np.random.default_rng(seed)
    b = R.make_boxes(rng, n, canvas=(W, H))
    return np.hstack([np.zeros((n, 1), np.float32), b]).astype(np.float32)


def stagewise_learn_nms(out, prm, info, first_n, nongt_dim=None, non_gt_index=None, gap=1e-6):
    """The learn-NMS stage checked on ITS OWN inputs: the oracle consumes the product's cls_score / bbox_pred / fc_all_2_relu
    (each already compared with the oracle chain), so the two sides sort the same scores.  What remains ambiguous is a pair
    of scores closer than float32 softmax rounding: classes whose top first_n+1 sorted probabilities hold a relative gap
    below `gap` are excluded (every class is an independent problem of the head: a swap only moves rows of its own class).
    Returns (max |final - ref| / max |ref| over the compared classes, max box error, fraction of classes compared)."""
    cs = out['cls_score'].cpu().numpy(); bp = out['bbox_pred'].cpu().numpy()
    kw = dict(nongt_dim=nongt_dim) if non_gt_index is None else dict(non_gt_index=non_gt_index)
    ref = L.learn_nms_forward(cs, bp, out['rois'].cpu().numpy(), info, out['fc_all_2_relu'].cpu().numpy(),
                              {k: prm[k].numpy() for k in pipeline_np.NMS_NAMES}, first_n=first_n, class_thresh=0.01,
                              return_all=True, **kw)
    e = np.exp(cs - cs.max(axis=1, keepdims=True))
    prob = -np.sort(-(e / e.sum(axis=1, keepdims=True))[:, 1:], axis=0)[:first_n + 1]
    rgap = ((prob[:-1] - prob[1:]) / np.maximum(prob[:-1], 1e-30)).min(axis=0)            # [C]
    ok = rgap > gap
    fin = out['nms_final_score_output'].cpu().numpy(); box = out['learn_nms_sorted_bbox'].cpu().numpy()
    scale = max(float(np.abs(ref['final_score']).max()), 1e-30)
    e_fin = float(np.abs(fin[:, ok] - ref['final_score'][:, ok]).max()) / scale
    e_box = float(np.abs(box[:, ok] - ref['sorted_bbox'][:, ok]).max())
    return e_fin, e_box, float(ok.mean())


def _check(out, ref, prec, tag, prm, info, first_n=100, **nkw):
    e_feat = rel_err(out['fc_all_2_relu'].cpu().numpy(), ref['fc_all_2_relu'])
    e_cls = rel_err(out['cls_score'].cpu().numpy(), ref['cls_score'])
    e_chain, frac = rank_aligned_err(out['nms_final_score_output'].cpu().numpy(), out['learn_nms_sorted_bbox'].cpu().numpy(),
                                     ref['nms_final_score_output'], ref['learn_nms_sorted_bbox'])
    e_fin, e_box, cfrac = stagewise_learn_nms(out, prm, info, first_n, **nkw)
    print('%s[%s]: fc_all_2 %.2e cls_score %.2e | learn-NMS stage on its own inputs: final score %.2e, boxes %.2e px, %.0f%% of '
          'classes free of float32 near-ties | whole chain, rows aligned by roi (reported): %.2e, %.1f%% aligned'
          % (tag, prec, e_feat, e_cls, e_fin, e_box, 100 * cfrac, e_chain, 100 * frac))
    tol = 3e-4 if prec == 'fp32' else 3e-3
    assert e_feat < tol and e_cls < tol
    assert cfrac > 0.5
    assert e_fin < (1e-3 if prec == 'fp32' else 5e-3) and e_box < 1e-2


def test_faster_pipeline_tie_aware(ops):
    """configs[1] head (proposal .. learn_nms) with the learn-NMS outputs compared per (roi, class)."""
    from relnet_b200.pipeline import RelationHead
    prm = _params(3)
    cls_prob, bbox_pred, info = P.make_proposal_case(5)
    feat = np.maximum(np.random.default_rng(5).standard_normal((1, 256, 38, 63)), 0).astype(np.float32)
    ref = pipeline_np.head_forward({k: v.numpy() for k, v in prm.items()}, cls_prob, bbox_pred, feat, info)
    for prec in precisions(ops):
        head = RelationHead({k: v.cuda() for k, v in prm.items()}, precision=prec)
        out = head.forward(T(cls_prob), T(bbox_pred), T(feat), T(info))
        np.testing.assert_array_equal(out['rois'].cpu().numpy(), ref['rois'])
        _check(out, ref, prec, 'faster', prm, info, nongt_dim=300)


def test_deformable_pipeline_full_size(ops):
    """configs[2] head: PS-ROI pool (no_trans) -> offset FC -> PS-ROI pool (trans) -> fc_new_1 -> tail, R = 300, 256 ch."""
    from relnet_b200.pipeline import DeformableRelationHead
    prm = _params(4)
    g = torch.Generator().manual_seed(11)
    prm['offset_weight'] = torch.randn((98, 12544), generator=g) * 0.01
    prm['offset_bias'] = torch.zeros(98)
    rois = _rois(6, 300)
    feat = np.maximum(np.random.default_rng(6).standard_normal((1, 256, 38, 63)), 0).astype(np.float32)
    info = np.array([[600.0, 1000.0, 1.0]], np.float32)
    ref = pipeline_np.head_forward_dcn({k: v.numpy() for k, v in prm.items()}, rois, feat, info)
    for prec in precisions(ops):
        head = DeformableRelationHead({k: v.cuda() for k, v in prm.items()}, precision=prec)
        out = head.detect(T(rois), T(feat), T(info))
        _check(out, ref, prec, 'deformable', prm, info, nongt_dim=300)
        if prec == 'f16':        # the trunk's layout: bf16 channels_last map (values rounded to bf16 -> looser check vs the fp32 map)
            fb = T(feat).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
            ref_b = pipeline_np.head_forward_dcn({k: v.numpy() for k, v in prm.items()}, rois, fb.float().cpu().numpy(), info)
            _check(head.detect(T(rois), fb, T(info)), ref_b, prec, 'deformable(bf16 channels_last map)', prm, info, nongt_dim=300)


def _rois_fpn(seed, n, W, H, smin=16.0, smax=None):
    """rois with log-uniform sizes so that several pyramid levels are populated (level = floor(2 + log2(sqrt(wh)/224)))"""
    rng = np.random.default_rng(seed)
    smax = smax or 0.95 * min(W, H)
    s = np.exp(rng.uniform(np.log(smin), np.log(smax), n)); ar = np.exp(rng.uniform(np.log(0.5), np.log(2.0), n))
    w = np.minimum(s * np.sqrt(ar), W - 2); h = np.minimum(s / np.sqrt(ar), H - 2)
    x1 = rng.uniform(0, 1, n) * (W - 1 - w); y1 = rng.uniform(0, 1, n) * (H - 1 - h)
    return np.stack([np.zeros(n), x1, y1, x1 + w, y1 + h], 1).astype(np.float32)


def _fpn_case(ops, seed, n, W, H, want_levels):
    from relnet_b200.pipeline import FPNRelationHead
    prm = _params(5)
    rois = _rois_fpn(seed, n, float(W), float(H))
    rng = np.random.default_rng(seed + 1)
    feats = [np.maximum(rng.standard_normal((1, 256, H // s, W // s), dtype=np.float32), 0) for s in (4, 8, 16, 32)]
    info = np.array([[float(H), float(W), 1.0]], np.float32)
    lvl = pipeline_np.fpn_level(rois)
    order = np.argsort(lvl, kind='stable')
    rois_sorted, counts = rois[order], np.bincount(lvl, minlength=4).tolist()
    assert sum(1 for c in counts if c > 0) >= want_levels, counts
    first_n = min(150, n)
    ref = pipeline_np.head_forward_fpn({k: v.numpy() for k, v in prm.items()}, rois_sorted, counts, feats, info, first_n=first_n)
    for prec in precisions(ops):
        head = FPNRelationHead({k: v.cuda() for k, v in prm.items()}, precision=prec, first_n=first_n)
        rs, cn = head.split_rois(T(rois))
        np.testing.assert_array_equal(rs.cpu().numpy(), rois_sorted)
        assert cn == counts
        out = head.detect(rs, cn, [T(f) for f in feats], T(info))
        _check(out, ref, prec, 'fpn %dx%d levels=%s' % (H, W, counts), prm, info, first_n=first_n, nongt_dim=n)


def test_fpn_pipeline_full_size(ops):
    """configs[3] head at test time: 1000 rois dispatched over the pyramid of a 608 x 1024 input (strides 4..32; at this image
    size rois reach levels 0 and 1 -- level 2 needs sqrt(wh) >= 896), keys = all rois, first_n = 150."""
    _fpn_case(ops, 7, 1000, 1024, 608, want_levels=2)


def test_fpn_pipeline_all_four_levels(ops):
    """every pyramid level populated: 160 rois on a 2048 x 2048 input (level 3 needs sqrt(wh) >= 1792)"""
    _fpn_case(ops, 9, 160, 2048, 2048, want_levels=4)
