"""Pins for the MXNet op semantics restated in oracle/mxshim.py (the numpy stand-in the goldens are generated on).

MXNet 1.1.0 itself cannot be imported here, so two independent sources stand in for it:
  * the worked examples of MXNet's own operator documentation for ``Reshape`` (special codes 0 / -1 / -2 / -3 / -4), ``slice_axis``,
    ``take``, ``pick``, ``broadcast_to``, ``batch_dot``, ``SliceChannel`` -- published known answers, quoted shape by shape below;
  * torch's CPU implementations of the ops that exist in both libraries (grouped convolution, linear, bmm, softmax, sort,
    gather, smooth-L1), on the call shapes the reference uses (file:line of one call site each).
The composition of these ops is the reference's own code (oracle/refexec.py); this file is about the single ops.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import mxshim as mx


def rnd(*shape, seed=0):
    return np.random.default_rng(seed).standard_normal(shape).astype(np.float32)


# ------------------------------------------------------------------------------------------------ Reshape codes (MXNet docs examples)
@pytest.mark.parametrize('src, code, want', [
    ((2, 3, 4), (4, 0, 2), (4, 3, 2)),              # 0 copies this dimension from the input
    ((2, 3, 4), (2, 0, 0), (2, 3, 4)),
    ((2, 3, 4), (6, 1, -1), (6, 1, 4)),             # -1 infers
    ((2, 3, 4), (3, -1, 8), (3, 1, 8)),
    ((2, 3, 4), (-1,), (24,)),
    ((2, 3, 4), (-2,), (2, 3, 4)),                  # -2 copies all remaining dimensions
    ((2, 3, 4), (2, -2), (2, 3, 4)),
    ((2, 3, 4), (-2, 1, 1), (2, 3, 4, 1, 1)),
    ((2, 3, 4), (-3, 4), (6, 4)),                   # -3 merges two consecutive dimensions
    ((2, 3, 4, 5), (-3, -3), (6, 20)),
    ((2, 3, 4), (0, -3), (2, 12)),
    ((2, 3, 4), (-3, -2), (6, 4)),
    ((2, 3, 4), (-4, 1, 2, -2), (1, 2, 3, 4)),      # -4 splits one dimension into the two that follow
    ((2, 3, 4), (2, -4, -1, 3, -2), (2, 1, 3, 4)),
])
def test_reshape_special_codes_match_the_documented_examples(src, code, want):
    x = np.arange(int(np.prod(src)), dtype=np.float32).reshape(src)
    y = mx.Reshape(mx.ND(x), shape=code).a
    assert y.shape == want
    np.testing.assert_array_equal(y.ravel(), x.ravel())          # Reshape never moves data


def test_reshape_codes_at_the_reference_call_sites():
    # SYM_REL:123-124  Reshape(q_data, shape=(-1, group, dim_group[0]))  then  transpose(axes=(1, 0, 2))
    q = rnd(300, 1024)
    qb = mx.transpose(mx.Reshape(mx.ND(q), shape=(-1, 16, 64)), axes=(1, 0, 2)).a
    np.testing.assert_array_equal(qb, torch.from_numpy(q).view(300, 16, 64).permute(1, 0, 2).numpy())
    # SYM_REL:107  Reshape(position_embedding, shape=(-3, -2)):  [N, M, 64] -> [N*M, 64]  (also :142, LNMS:67,91,98,106)
    pe = rnd(7, 5, 64)
    assert mx.Reshape(mx.ND(pe), shape=(-3, -2)).shape == (35, 64)
    # SYM_REL:114,116  Reshape(relu(fc), shape=(-1, nongt_dim, fc_dim))  then  transpose(axes=(0, 2, 1))
    fc = rnd(35, 16)
    aff = mx.transpose(mx.Reshape(mx.ND(fc), shape=(-1, 5, 16)), axes=(0, 2, 1)).a
    np.testing.assert_array_equal(aff, torch.from_numpy(fc).view(7, 5, 16).permute(0, 2, 1).numpy())
    # SYM_REL:146  Reshape(output_t, shape=(-1, fc_dim * feat_dim, 1, 1)) feeding the grouped 1x1 convolution
    o = rnd(7, 16, 32)
    assert mx.Reshape(mx.ND(o), shape=(-1, 16 * 32, 1, 1)).shape == (7, 512, 1, 1)
    # SYM_REL:43  Reshape(embedding, shape=(0, 0, feat_dim)):  [N, M, 4, 16] -> [N, M, 64]  (the trailing dims are absorbed)
    t = rnd(4, 3, 4, 16)
    np.testing.assert_array_equal(mx.Reshape(mx.ND(t), shape=(0, 0, 64)).a, t.reshape(4, 3, 64))
    # LNMS:111  Reshape(aff_softmax, shape=(-1, fc_dim[1] * num_rois, 0)): leading dims regrouped, the 0 copies input dim 2
    u = rnd(6, 8, 5)
    np.testing.assert_array_equal(mx.Reshape(mx.ND(u), shape=(-1, 16, 0)).a, u.reshape(3, 16, 5))
    # LNMS:183  Reshape(bbox_delta, shape=(0, -1, 4)):  [R, 4*C] -> [R, C, 4]
    v = rnd(5, 12)
    np.testing.assert_array_equal(mx.Reshape(mx.ND(v), shape=(0, -1, 4)).a, v.reshape(5, 3, 4))


# ------------------------------------------------------------------------------------------------ layers against torch
@pytest.mark.parametrize('B, g, cin_g, cout_g', [(7, 16, 32, 64), (5, 4, 8, 16), (3, 1, 12, 6)])
def test_grouped_1x1_convolution_matches_torch(B, g, cin_g, cout_g):
    # SYM_REL:147-149: Convolution(kernel=(1,1), num_filter=dim[2], num_group=fc_dim) on [N, fc_dim*feat_dim, 1, 1]
    x, w, b = rnd(B, g * cin_g, 1, 1, seed=1), rnd(g * cout_g, cin_g, 1, 1, seed=2), rnd(g * cout_g, seed=3)
    y = mx.Convolution(data=mx.ND(x), weight=mx.ND(w), bias=mx.ND(b), kernel=(1, 1), num_filter=g * cout_g, num_group=g).a
    ref = F.conv2d(torch.from_numpy(x).double(), torch.from_numpy(w).double(), torch.from_numpy(b).double(), groups=g).numpy()
    np.testing.assert_allclose(y, ref, rtol=2e-5, atol=2e-5)
    # spatial extent > 1 as well (the op is position-wise)
    x2 = rnd(2, g * cin_g, 3, 2, seed=4)
    y2 = mx.Convolution(data=mx.ND(x2), weight=mx.ND(w), bias=mx.ND(b), kernel=(1, 1), num_filter=g * cout_g, num_group=g).a
    ref2 = F.conv2d(torch.from_numpy(x2).double(), torch.from_numpy(w).double(), torch.from_numpy(b).double(), groups=g).numpy()
    np.testing.assert_allclose(y2, ref2, rtol=2e-5, atol=2e-5)


def test_fully_connected_flattens_and_uses_out_in_weights():
    # SYM_REL:261 FullyConnected(name='fc_new_1', data=roi_pool [R, 256, 7, 7], num_hidden=1024): flatten=True, weight [out, in]
    x, w, b = rnd(6, 4, 3, 3, seed=5), rnd(10, 36, seed=6), rnd(10, seed=7)
    y = mx.FullyConnected(data=mx.ND(x), weight=mx.ND(w), bias=mx.ND(b), num_hidden=10).a
    ref = F.linear(torch.from_numpy(x).double().flatten(1), torch.from_numpy(w).double(), torch.from_numpy(b).double()).numpy()
    np.testing.assert_allclose(y, ref, rtol=2e-5, atol=2e-5)
    y0 = mx.FullyConnected(data=mx.ND(x), weight=mx.ND(w), num_hidden=10, no_bias=True).a
    np.testing.assert_allclose(y0, ref - b, rtol=2e-5, atol=2e-5)
    # name-addressed parameters (sym API): '<name>_weight' / '<name>_bias'
    mx.PARAMS.update({'q_weight': w, 'q_bias': b})
    try:
        np.testing.assert_array_equal(mx.FullyConnected(name='q', data=mx.ND(x), num_hidden=10).a, y)
    finally:
        mx.PARAMS.pop('q_weight'); mx.PARAMS.pop('q_bias')


@pytest.mark.parametrize('ta, tb', [(False, False), (False, True), (True, False), (True, True)])
def test_batch_dot_matches_bmm(ta, tb):
    # SYM_REL:132 batch_dot(lhs=q_data_batch, rhs=k_data_batch, transpose_a=False, transpose_b=True)
    a = rnd(*((4, 5, 7) if not ta else (4, 7, 5)), seed=8)
    b = rnd(*((4, 7, 6) if not tb else (4, 6, 7)), seed=9)
    y = mx.batch_dot(lhs=mx.ND(a), rhs=mx.ND(b), transpose_a=ta, transpose_b=tb).a
    A, B = torch.from_numpy(a).double(), torch.from_numpy(b).double()
    ref = torch.bmm(A.transpose(1, 2) if ta else A, B.transpose(1, 2) if tb else B).numpy()
    assert y.shape == (4, 5, 6)                      # docs: x (B,N,M), y (B,M,K) -> (B,N,K)
    np.testing.assert_allclose(y, ref, rtol=2e-5, atol=2e-5)
    d = mx.dot(lhs=mx.ND(a[0]), rhs=mx.ND(b[0]), transpose_a=ta, transpose_b=tb).a
    np.testing.assert_allclose(d, ref[0], rtol=2e-5, atol=2e-5)


@pytest.mark.parametrize('axis', [2, -1, 1, 0])
def test_softmax_axis_matches_torch(axis):
    # SYM_REL:140 softmax(data=weighted_aff, axis=2)
    x = rnd(5, 16, 9, seed=10) * 4
    y = mx.softmax(data=mx.ND(x), axis=axis).a
    np.testing.assert_allclose(y, torch.softmax(torch.from_numpy(x).double(), dim=axis).numpy(), rtol=2e-6, atol=1e-7)
    np.testing.assert_allclose(y.sum(axis=axis), 1.0, rtol=1e-5)


def test_softmax_activation_is_softmax_over_axis_1():
    # proposal.py consumes SoftmaxActivation(mode='channel') of the reshaped rpn scores; 'instance' on 2-D = axis 1
    x = rnd(11, 81, seed=11) * 3
    np.testing.assert_allclose(mx.SoftmaxActivation(data=mx.ND(x)).a, torch.softmax(torch.from_numpy(x).double(), dim=1).numpy(),
                               rtol=3e-6, atol=1e-7)


@pytest.mark.parametrize('sigma', [1.0, 3.0])
def test_smooth_l1_matches_torch_with_beta(sigma):
    # SYM_REL:291,298 smooth_l1(scalar=1.0) (RCNN bbox loss), :190 rpn: scalar=3.0;  f = 0.5 (sigma x)^2 / |x| - 0.5/sigma^2, knee at 1/sigma^2
    x = np.concatenate([rnd(200, seed=12), np.float32([0.0, 1.0 / sigma ** 2, -1.0 / sigma ** 2, 1e-4, -3.0])])
    y = mx.smooth_l1(data=mx.ND(x), scalar=sigma).a
    ref = F.smooth_l1_loss(torch.from_numpy(x).double(), torch.zeros(x.size, dtype=torch.float64), reduction='none',
                           beta=1.0 / sigma ** 2).numpy()
    np.testing.assert_allclose(y, ref, rtol=2e-6, atol=1e-7)


# ------------------------------------------------------------------------------------------------ sort / gather family
def test_sort_and_argsort_descending_match_torch_stable():
    # LNMS:291,306 sort / argsort(axis=0, is_ascend=False) of the class scores; ties: lower index first (defined)
    x = np.round(rnd(50, 6, seed=13), 1)                 # rounding creates ties
    v = mx.sort(data=mx.ND(x), axis=0, is_ascend=False).a
    i = mx.argsort(data=mx.ND(x), axis=0, is_ascend=False).a
    tv, ti = torch.sort(torch.from_numpy(x), dim=0, descending=True, stable=True)
    np.testing.assert_array_equal(v, tv.numpy())
    np.testing.assert_array_equal(i, ti.numpy().astype(np.float32))
    assert i.dtype == np.float32                          # MXNet returns float indices
    va = mx.sort(data=mx.ND(x), axis=1).a                 # default: ascending, last axis
    np.testing.assert_array_equal(va, torch.sort(torch.from_numpy(x), dim=1, stable=True)[0].numpy())


def test_take_clips_out_of_range_indices_like_the_docs():
    # docs: x = [[1,2],[3,4],[5,6]]; take(x, [[0,1],[1,2]]) -> [[[1,2],[3,4]],[[3,4],[5,6]]]; mode='clip' is the default
    x = np.float32([[1, 2], [3, 4], [5, 6]])
    y = mx.take(a=mx.ND(x), indices=mx.ND(np.float32([[0, 1], [1, 2]]))).a
    np.testing.assert_array_equal(y, np.float32([[[1, 2], [3, 4]], [[3, 4], [5, 6]]]))
    z = mx.take(a=mx.ND(x), indices=mx.ND(np.float32([-2, 0, 7]))).a
    np.testing.assert_array_equal(z, x[[0, 0, 2]])
    ref = torch.from_numpy(x).index_select(0, torch.tensor([0, 0, 2])).numpy()
    np.testing.assert_array_equal(z, ref)


def test_pick_matches_the_docs_and_torch_gather():
    # docs: x = [[1,2],[3,4],[5,6]]; pick(x, y=[0,1], 0) -> [1,4]; pick(x, y=[0,1,0], 1) -> [1,4,5]; default axis=-1, clip
    x = np.float32([[1, 2], [3, 4], [5, 6]])
    np.testing.assert_array_equal(mx.pick(data=mx.ND(x), index=mx.ND(np.float32([0, 1])), axis=0).a, np.float32([1, 4]))
    np.testing.assert_array_equal(mx.pick(data=mx.ND(x), index=mx.ND(np.float32([0, 1, 0])), axis=1).a, np.float32([1, 4, 5]))
    np.testing.assert_array_equal(mx.pick(data=mx.ND(x), index=mx.ND(np.float32([0, 1, 0]))).a, np.float32([1, 4, 5]))
    np.testing.assert_array_equal(mx.pick(data=mx.ND(x), index=mx.ND(np.float32([0, 1, 0])), axis=1, keepdims=True).a, np.float32([[1], [4], [5]]))
    np.testing.assert_array_equal(mx.pick(data=mx.ND(x), index=mx.ND(np.float32([9, -3, 1])), axis=1).a, np.float32([2, 3, 6]))     # clip
    big, idx = rnd(40, 9, seed=14), np.random.default_rng(15).integers(0, 9, 40)
    ref = torch.gather(torch.from_numpy(big), 1, torch.from_numpy(idx)[:, None]).squeeze(1).numpy()
    np.testing.assert_array_equal(mx.pick(data=mx.ND(big), index=mx.ND(idx.astype(np.float32)), axis=1).a, ref)


def test_slice_axis_split_tile_reverse_broadcast_to():
    x = np.float32([[1, 2, 3, 4], [5, 6, 7, 8], [9, 10, 11, 12]])
    # docs: slice_axis(x, axis=0, begin=1, end=3) = rows 1..2; (axis=1, begin=0, end=2); (axis=1, begin=-3, end=-1) = [[2,3],[6,7],[10,11]]
    np.testing.assert_array_equal(mx.slice_axis(data=mx.ND(x), axis=0, begin=1, end=3).a, x[1:3])
    np.testing.assert_array_equal(mx.slice_axis(data=mx.ND(x), axis=1, begin=0, end=2).a, x[:, 0:2])
    np.testing.assert_array_equal(mx.slice_axis(data=mx.ND(x), axis=1, begin=-3, end=-1).a, np.float32([[2, 3], [6, 7], [10, 11]]))
    np.testing.assert_array_equal(mx.slice_axis(data=mx.ND(x), axis=1, begin=2, end=None).a, x[:, 2:])       # LNMS:278,284,290 end=None
    # docs (SliceChannel): x (1,2,3)... num_outputs along axis=1 by default; squeeze_axis drops the axis of length 1
    z = rnd(5, 4, 6, seed=16)
    parts = mx.split(data=mx.ND(z), num_outputs=4, axis=1)
    tparts = torch.split(torch.from_numpy(z), 1, dim=1)
    assert len(parts) == 4 and parts[0].shape == (5, 1, 6)
    for p, t in zip(parts, tparts):
        np.testing.assert_array_equal(p.a, t.numpy())
    sq = mx.split(data=mx.ND(z), num_outputs=4, axis=1, squeeze_axis=1)
    assert sq[2].shape == (5, 6)
    np.testing.assert_array_equal(sq[2].a, z[:, 2])
    # SYM_REL:56-57 split(data=bbox, num_outputs=4, axis=1) of the [N, 4] boxes (bbox[:, k:k+1])
    bx = rnd(9, 4, seed=17)
    for k, p in enumerate(mx.split(data=mx.ND(bx), num_outputs=4, axis=1)):
        np.testing.assert_array_equal(p.a, bx[:, k:k + 1])
    np.testing.assert_array_equal(mx.tile(data=mx.ND(x), reps=(2, 3)).a, torch.from_numpy(x).repeat(2, 3).numpy())
    np.testing.assert_array_equal(mx.reverse(data=mx.ND(x), axis=1).a, torch.from_numpy(x).flip(1).numpy())
    # docs: broadcast_to(x (1,2,1) -> shape (2,2,3)); a 0 in shape keeps the input's size:  broadcast_to([[1,2,3]], shape=(2,0)) -> (2,3)
    r = np.float32([[1, 2, 3]])
    np.testing.assert_array_equal(mx.broadcast_to(data=mx.ND(r), shape=(2, 0)).a, np.float32([[1, 2, 3], [1, 2, 3]]))
    np.testing.assert_array_equal(mx.broadcast_to(data=mx.ND(r.reshape(1, 3, 1)), shape=(2, 3, 4)).a,
                                  torch.from_numpy(r.reshape(1, 3, 1)).expand(2, 3, 4).numpy())


def test_broadcast_arithmetic_and_scalar_ops_stay_float32():
    a, b = rnd(7, 1, 5, seed=18), rnd(1, 6, 5, seed=19)
    A, B = torch.from_numpy(a), torch.from_numpy(b)
    for name, ref in [('broadcast_add', A + B), ('broadcast_minus', A - B), ('broadcast_mul', A * B), ('broadcast_div', A / B),
                      ('broadcast_maximum', torch.maximum(A, B)), ('broadcast_minimum', torch.minimum(A, B))]:
        y = getattr(mx, name)(lhs=mx.ND(a), rhs=mx.ND(b)).a
        assert y.dtype == np.float32
        np.testing.assert_array_equal(y, ref.numpy())
    # SYM_REL:67,71 maximum(abs(delta), 1e-3): the python scalar is cast to float32 (mx *_scalar ops)
    d = mx.maximum(left=mx.abs(mx.ND(a) / 3.0), right=1e-3).a
    np.testing.assert_array_equal(d, torch.clamp_min((A / 3.0).abs(), 1e-3).numpy())
    # SYM_REL:32-34 broadcast_power(lhs=full((1,), wave_length), rhs=(8. / feat_dim) * arange(0, feat_dim / 8))
    fr = mx.arange(0, 8)
    p = mx.broadcast_power(lhs=mx.full((1,), 1000.0), rhs=fr * (8.0 / 64.0)).a
    np.testing.assert_allclose(p, torch.pow(torch.tensor(1000.0), torch.arange(8, dtype=torch.float32) * 0.125).numpy(), rtol=2e-7)
    assert fr.a.dtype == np.float32
