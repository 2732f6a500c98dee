"""Static check of the built library (no GPU): the kernels DESIGN.md puts on the 5th-generation tensor cores really contain
tcgen05 / TMEM / TMA instructions in their sm_100a SASS (UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG = cp.async.bulk.tensor,
UTCBAR = tcgen05.commit), and nothing in the library was compiled for another architecture."""
import collections
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'relation-networks-for-object-detection_b200', 'librelnet_b200.so')
CUOBJDUMP = shutil.which('cuobjdump') or '/usr/local/cuda/bin/cuobjdump'

pytestmark = pytest.mark.skipif(not os.path.exists(CUOBJDUMP), reason='cuobjdump not available')


@pytest.fixture(scope='module')
def sass():
    import __graft_entry__ as entry
    entry.build()
    out = subprocess.run([CUOBJDUMP, '-sass', LIB], capture_output=True, text=True, check=True).stdout
    per, kern = collections.defaultdict(collections.Counter), None
    archs = set(re.findall(r'arch = (sm_\w+)', out))
    for line in out.splitlines():
        m = re.match(r'\s*Function : (\S+)', line)
        if m:
            kern = m.group(1)
        elif kern:
            for t in re.findall(r'\b(UTCHMMA|LDTM|UTMALDG|UBLKCP|UTCBAR|HMMA|SYNCS)\b', line):
                per[kern][t] += 1
    return per, archs


def _sum(per, needle):
    tot = collections.Counter()
    hit = [k for k in per if needle in k]
    for k in hit:
        tot.update(per[k])
    return hit, tot


def test_only_sm_100a_code_is_shipped(sass):
    _, archs = sass
    assert archs == {'sm_100a'}, archs


@pytest.mark.parametrize('needle, need', [
    ('relation_fused_kernel', ('UTCHMMA', 'LDTM', 'UTMALDG', 'UTCBAR')),        # pair FC + QK^T + P.V' on tcgen05, Q/K/V' by TMA
    ('gemm_f16_tc_kernel', ('UTCHMMA', 'LDTM', 'UTMALDG', 'UTCBAR')),           # projections, head GEMMs, deformable-conv GEMM, RPN head
    ('gemm_tf32_tc_kernel', ('UTCHMMA', 'LDTM', 'UTMALDG', 'UTCBAR')),          # every backward contraction (kind::tf32)
    ('relation_attn_tile_kernel', ('UTCHMMA', 'LDTM', 'UTMALDG')),              # learn-NMS per-class attention
    ('geom_weight_tc', ('UTCHMMA', 'LDTM')),                                    # roi-level geometry table of the learn-NMS head
    ('nms_sweep_kernel', ('UBLKCP',)),                                          # bulk-copy staging of the mask rows
])
def test_kernel_uses_the_blackwell_units(sass, needle, need):
    per, _ = sass
    hit, tot = _sum(per, needle)
    assert hit, 'kernel %s not found in the library' % needle
    for mnem in need:
        assert tot[mnem] > 0, (needle, mnem, dict(tot))


def test_no_legacy_mma_sync_in_the_contraction_kernels(sass):
    per, _ = sass
    for needle in ('relation_fused_kernel', 'gemm_f16_tc_kernel', 'gemm_tf32_tc_kernel'):
        _, tot = _sum(per, needle)
        assert tot['HMMA'] == 0, (needle, tot['HMMA'])
