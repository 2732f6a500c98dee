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
from oracle import pipeline_np, proposal_np as P, relation_np as R

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


def _check(out, ref, prec, tag):
    e_feat = rel_err(out['fc_all_2_relu'].cpu().numpy(), ref['fc_all_2_relu'])
    e_cls = rel_err(out['cls_score'].cpu().numpy(), ref['cls_score'])
    e_fin, frac = rank_aligned_err(out['nms_final_score_output'].cpu().numpy(), out['learn_nms_sorted_bbox'].cpu().numpy(),
                                   ref['nms_final_score_output'], ref['learn_nms_sorted_bbox'])
    print('%s[%s]: fc_all_2 %.2e cls_score %.2e | learn-NMS final score, rows aligned by roi: %.2e (%.1f%% of rows aligned)'
          % (tag, prec, e_feat, e_cls, e_fin, 100 * frac))
    tol = 3e-4 if prec == 'fp32' else 3e-3
    assert e_feat < tol and e_cls < tol
    assert e_fin < (1e-3 if prec == 'fp32' else 5e-3) and frac > (0.999 if prec == 'fp32' else 0.97)


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
        _check(out, ref, prec, 'faster')


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
        _check(out, ref, prec, 'deformable')
        if prec == 'f16':        # the trunk's layout: bf16 channels_last map (values rounded to bf16 -> looser check vs the fp32 map)
            fb = T(feat).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
            ref_b = pipeline_np.head_forward_dcn({k: v.numpy() for k, v in prm.items()}, rois, fb.float().cpu().numpy(), info)
            _check(head.detect(T(rois), fb, T(info)), ref_b, prec, 'deformable(bf16 channels_last map)')


def test_fpn_pipeline_full_size(ops):
    """configs[3] head at test time: 1000 rois dispatched to 4 levels (strides 4..32 of a 608 x 1024 input), keys = all rois,
    first_n = 150."""
    from relnet_b200.pipeline import FPNRelationHead
    prm = _params(5)
    rois = _rois(7, 1000, W=1024.0, H=608.0)
    rng = np.random.default_rng(8)
    feats = [np.maximum(rng.standard_normal((1, 256, 608 // s, 1024 // s)), 0).astype(np.float32) for s in (4, 8, 16, 32)]
    info = np.array([[608.0, 1024.0, 1.0]], np.float32)
    lvl = pipeline_np.fpn_level(rois)
    order = np.argsort(lvl, kind='stable')
    rois_sorted, counts = rois[order], np.bincount(lvl, minlength=4).tolist()
    assert min(counts) > 0, counts
    ref = pipeline_np.head_forward_fpn({k: v.numpy() for k, v in prm.items()}, rois_sorted, counts, feats, info)
    for prec in precisions(ops):
        head = FPNRelationHead({k: v.cuda() for k, v in prm.items()}, precision=prec)
        rs, cn = head.split_rois(T(rois))
        np.testing.assert_array_equal(rs.cpu().numpy(), rois_sorted)
        assert cn == counts
        out = head.detect(rs, cn, [T(f) for f in feats], T(info))
        _check(out, ref, prec, 'fpn')
