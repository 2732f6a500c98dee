"""CPU: the MXNet .params reader/writer (relnet_b200/params.py) -- round trip, the 'arg:'/'aux:' split and *_test handling
of lib/utils/load_model.py:12-67, the fold of core/callback.py:54-61, V1 / pre-V1 records built by hand, error paths."""
import struct
import numpy as np
import pytest


@pytest.fixture(scope='module')
def params():
    import relnet_b200            # noqa: F401
    from relnet_b200 import params
    return params


def test_round_trip_and_checkpoint_split(params, tmp_path):
    rng = np.random.default_rng(0)
    arg = {'query_1_weight': rng.standard_normal((16, 8)).astype(np.float32), 'query_1_bias': np.zeros(16, np.float32),
           'bbox_pred_weight': rng.standard_normal((8, 16)).astype(np.float32),
           'bbox_pred_bias': rng.standard_normal(8).astype(np.float32),
           'linear_out_1_weight': rng.standard_normal((16, 16, 1, 1)).astype(np.float32)}
    aux = {'bn_conv1_moving_mean': rng.standard_normal(4).astype(np.float32), 'half': np.arange(6, dtype=np.float16).reshape(2, 3)}
    means, stds = (0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0), (0.1, 0.1, 0.2, 0.2, 0.1, 0.1, 0.2, 0.2)
    folded = params.fold_bbox_test(arg, means, stds)
    assert np.allclose(folded['bbox_pred_weight_test'], arg['bbox_pred_weight'] * np.asarray(stds, np.float32)[:, None])
    prefix = str(tmp_path / 'rcnn_coco')
    params.save_checkpoint(prefix, 8, folded, aux)
    a, x = params.load_checkpoint(prefix, 8)
    assert set(a) == set(folded) and set(x) == set(aux)
    for k in folded:
        assert a[k].dtype == folded[k].dtype and np.array_equal(a[k], folded[k])
    assert x['half'].dtype == np.float16 and np.array_equal(x['half'], aux['half'])
    a2, _ = params.load_param(prefix, 8, process=True)                 # *_test replaces the trained entries
    assert 'bbox_pred_weight_test' not in a2 and np.array_equal(a2['bbox_pred_weight'], folded['bbox_pred_weight_test'])
    assert np.array_equal(a2['bbox_pred_bias'], folded['bbox_pred_bias_test'])


def test_v1_and_legacy_records(params):
    data = np.arange(6, dtype=np.float32).reshape(2, 3)
    v1 = struct.pack('<II2qiii', params.V1_MAGIC, 2, 2, 3, 1, 0, 0) + data.tobytes()
    legacy = struct.pack('<I2Iiii', 2, 2, 3, 1, 0, 0) + data.tobytes()
    name = b'arg:w'
    for rec in (v1, legacy):
        buf = struct.pack('<QQQ', params.LIST_MAGIC, 0, 1) + rec + struct.pack('<QQ', 1, len(name)) + name
        arrays, names = params.loads(buf)
        assert names == ['arg:w'] and np.array_equal(arrays[0], data)
    # unnamed list + a "none" array (ndim 0)
    buf = struct.pack('<QQQ', params.LIST_MAGIC, 0, 2) + v1 + struct.pack('<IiI', params.V2_MAGIC, 0, 0) + struct.pack('<Q', 0)
    arrays, names = params.loads(buf)
    assert names == [] and arrays[1] is None and np.array_equal(arrays[0], data)


def test_error_paths(params):
    with pytest.raises(params.ParamsError):
        params.loads(struct.pack('<QQQ', 0x113, 0, 0))                                 # wrong magic
    good = params.dumps({'arg:a': np.ones((3, 3), np.float32)})
    with pytest.raises(params.ParamsError):
        params.loads(good[:-20])                                                       # truncated
    sparse = struct.pack('<QQQ', params.LIST_MAGIC, 0, 1) + struct.pack('<Ii', params.V2_MAGIC, 1)
    with pytest.raises(params.ParamsError):
        params.loads(sparse + b'\0' * 64)
    with pytest.raises(params.ParamsError):
        params.dumps({'x': np.ones(2, np.complex64)})
