"""CPU: host-side logic that needs no device -- the fp16 operand cache's identity / version / invalidate semantics,
and the Python mirror of the kernels' head-shape rule against the C library's own answer (host-only entry points)."""
import ctypes as C
import pytest
import torch


@pytest.fixture(scope='module')
def pkg():
    import __graft_entry__ as g
    g.build()
    import relnet_b200
    return relnet_b200


def test_pack_cache_identity_version_invalidate(pkg):
    cache = pkg.ops._PackCache()
    w = torch.zeros(4, 8)
    calls = []

    def pack(buf):
        calls.append(buf.data_ptr())
    a = cache.get((w,), 64, pack, tag=('relation', 16, 1024, 1024))
    b = cache.get((w,), 64, pack, tag=('relation', 16, 1024, 1024))
    assert a.data_ptr() == b.data_ptr() and len(calls) == 1                 # same identity + version: reused
    c = cache.get((w,), 64, pack, tag=('relation', 4, 1024, 1024))
    assert c.data_ptr() != a.data_ptr() and len(calls) == 2                 # other descriptor: its own block
    w.add_(1.0)                                                             # in-place update bumps _version
    d = cache.get((w,), 64, pack, tag=('relation', 16, 1024, 1024))
    assert d.data_ptr() == a.data_ptr() and len(calls) == 3                 # repacked IN PLACE: no dead buffers pile up
    assert len(cache.d) == 2
    w.data.mul_(2.0)                                                        # .data write: version counter does not move ...
    e = cache.get((w,), 64, pack, tag=('relation', 16, 1024, 1024))
    assert len(calls) == 3 and e.data_ptr() == a.data_ptr()                 # ... so the stale block would be reused:
    cache.invalidate((w,))                                                  # the documented hook
    assert len(cache.d) == 0
    cache.get((w,), 64, pack, tag=('relation', 16, 1024, 1024))
    assert len(calls) == 4
    for i in range(cache.MAX + 5):                                          # bounded: oldest identities are evicted
        cache.get((torch.zeros(1),), 16, pack, tag=('x', i))
    assert len(cache.d) <= cache.MAX


@pytest.mark.parametrize('fused', [1, 0])
def test_python_shape_rule_matches_the_library(pkg, fused):
    """ops.relation_tc_supported mirrors relation_tc.cu:head_chunks (+ 'wide heads need the fused kernel'): compare with what the
    library itself says (rn_relation_packed_bytes > 0 is its shape test; host-only)."""
    ops, L = pkg.ops, pkg._lib
    prev = ops.relation_fused_enable(fused)
    try:
        for H in (1, 2, 3, 4, 8, 16):
            for dk in (8, 16, 48, 64, 96, 128, 192, 256, 512):
                for dv in (8, 64, dk):
                    dq, dout = H * dk, H * dv
                    d = L.RelationDesc(1, 300, 300, 1024, dq, dout, H, 64, 1000.0, 0, 1)
                    lib_ok = L.lib().rn_relation_packed_bytes(C.byref(d)) > 0
                    wide = dk > 64 or dv > 64
                    want = lib_ok and (not wide or bool(fused))
                    assert ops.relation_tc_supported(dq, dout, H) == want, (H, dk, dv, lib_ok, fused)
    finally:
        ops.relation_fused_enable(prev)
    assert not ops.relation_tc_supported(1024, 1024, 16, return_softmax=True)


def test_new_workspace_queries_are_host_only(pkg):
    lib = pkg._lib.lib()
    assert lib.rn_rpn_head_packed_bytes(512, 12) >= 72 * 512 * 2 + 72 * 4
    assert lib.rn_rpn_head_workspace_bytes(38 * 63, 512, 12) >= 38 * 63 * 72 * 4
    assert lib.rn_rpn_head_packed_bytes(0, 12) == 0
