"""CPU: the C-ABI library builds for sm_100a, loads, exports every symbol include/relnet_b200.h declares, and the
product refuses to run without CUDA tensors (no fallback).  No compute calls here (no GPU in this container)."""
import ctypes
import os
import re
import numpy as np
import pytest
import torch
from conftest import ROOT


@pytest.fixture(scope='module')
def built():
    import __graft_entry__ as g
    g.build()
    import relnet_b200
    return relnet_b200


def test_header_symbols_exported(built):
    hdr = open(os.path.join(ROOT, 'include', 'relnet_b200.h')).read()
    declared = set(re.findall(r'\b(rn_[a-z0-9_]+)\s*\(', hdr))
    declared = {d for d in declared if not d.endswith('_desc') and not d.endswith('_weights')}
    assert len(declared) >= 20
    lib = ctypes.CDLL(built._lib.LIB_PATH)
    missing = [d for d in sorted(declared) if not hasattr(lib, d)]
    assert not missing, missing
    assert declared == set(built._lib.SIGNATURES), declared ^ set(built._lib.SIGNATURES)


def test_library_reports_version_and_no_device(built):
    lib = built._lib.lib()
    assert lib.rn_version() >= 100
    assert isinstance(lib.rn_last_error(), bytes)


def test_no_cpu_fallback(built):
    x = torch.zeros(4, 8)
    with pytest.raises(built._lib.RelnetError):
        built.ops.linear(x, torch.zeros(3, 8))
    with pytest.raises(built._lib.RelnetError):
        built.ops.roi_pool(torch.zeros(1, 1, 4, 4), torch.zeros(1, 5))


def test_workspace_queries_are_host_only(built):
    L = built._lib
    d = L.RelationDesc(1, 300, 300, 1024, 1024, 1024, 16, 64, 1000.0, 1, 0)
    assert L.lib().rn_relation_workspace_bytes(ctypes.byref(d)) > 300 * 1024 * 4 * 3
    p = L.ProposalDesc(38, 63, 16, 4, 3, 6000, 300, 0.7, 0.0)
    assert L.lib().rn_proposal_workspace_bytes(ctypes.byref(p)) > 6000 * 94 * 8


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, 'relation-networks-for-object-detection_b200')
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(dp, f)).read()
                assert 'import oracle' not in src and 'from oracle' not in src and 'oracle/' not in src.replace(
                    'oracle/oracle_c.c', '').replace('oracle/proposal_np.py', ''), (f, 'product must not use the oracle')
