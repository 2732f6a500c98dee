"""MXNet ``.params`` checkpoint reader / writer and the ``*_test`` folding -- lets released Relation-Networks weights be
loaded into this library's heads.

Reference call sites: lib/utils/load_model.py:12-67 (``mx.nd.load`` -> split the 'arg:' / 'aux:' name prefixes ->
``process=True`` renames ``*_test`` entries over the trained ones) and relation_rcnn/core/callback.py:54-61 (how
``bbox_pred_{weight,bias}_test`` are produced at save time: weight rows scaled by BBOX_STDS, bias * stds + means).

The file format belongs to Apache MXNet 1.1.0 (pinned by the reference README; not vendored), src/ndarray/ndarray.cc
``NDArray::Save`` / ``NDArray::Load`` -- restated from its published layout, all little-endian:

    uint64 0x112 (kMXAPINDArrayListMagic) | uint64 reserved = 0
    uint64 n_arrays | n_arrays x NDArray
    uint64 n_names  | n_names x (uint64 length | bytes)                 (n_names is 0 or n_arrays)
  NDArray (V2, magic 0xF993FAC9):
    uint32 magic | int32 storage_type (0 = dense; sparse checkpoints are rejected)
    shape: uint32 ndim | int64 dims[ndim]            (ndim == 0: "none" array, nothing follows)
    context: int32 dev_type | int32 dev_id
    int32 type_flag (0 f32, 1 f64, 2 f16, 3 u8, 4 i32, 5 i8, 6 i64) | raw data, C order
  V1 (magic 0xF993FAC8): the same without storage_type.  Pre-V1: the first uint32 IS ndim and dims are uint32.
This is host-side glue (pure numpy); tensors go to the GPU through ``to_torch``.  ``parity unpinned``: no MXNet here to
cross-check a real file -- the reader is tested by round-tripping files written by ``save`` and by hand-built V1 / legacy
byte strings (tests/test_params_cpu.py)."""
import struct
import numpy as np

LIST_MAGIC = 0x112
V1_MAGIC = 0xF993FAC8
V2_MAGIC = 0xF993FAC9
DTYPES = {0: np.float32, 1: np.float64, 2: np.float16, 3: np.uint8, 4: np.int32, 5: np.int8, 6: np.int64}
FLAGS = {np.dtype(v): k for k, v in DTYPES.items()}


class ParamsError(ValueError):
    pass


class _Reader(object):
    def __init__(self, buf):
        self.b, self.o = memoryview(buf), 0

    def take(self, fmt):
        n = struct.calcsize(fmt)
        if self.o + n > len(self.b):
            raise ParamsError('truncated .params file at byte %d' % self.o)
        v = struct.unpack_from('<' + fmt, self.b, self.o)
        self.o += n
        return v if len(v) > 1 else v[0]

    def raw(self, n):
        if self.o + n > len(self.b):
            raise ParamsError('truncated .params file at byte %d (need %d more)' % (self.o, n))
        v = self.b[self.o:self.o + n]
        self.o += n
        return v


def _read_ndarray(r):
    magic = r.take('I')
    if magic == V2_MAGIC:
        stype = r.take('i')
        if stype != 0:
            raise ParamsError('sparse NDArray (storage type %d) is not supported' % stype)
        ndim = r.take('I')
        shape = [r.take('q') for _ in range(ndim)]
    elif magic == V1_MAGIC:
        ndim = r.take('I')
        shape = [r.take('q') for _ in range(ndim)]
    else:                                   # pre-V1: magic is ndim, uint32 dims
        ndim = magic
        if ndim > 32:
            raise ParamsError('bad NDArray header 0x%08x' % magic)
        shape = [r.take('I') for _ in range(ndim)]
    if ndim == 0:
        return None
    r.take('ii')                            # context (dev_type, dev_id): irrelevant for a file
    flag = r.take('i')
    if flag not in DTYPES:
        raise ParamsError('unknown type flag %d' % flag)
    dt = np.dtype(DTYPES[flag]).newbyteorder('<')
    count = int(np.prod(shape, dtype=np.int64))
    return np.frombuffer(r.raw(count * dt.itemsize), dtype=dt, count=count).reshape(shape).copy()


def loads(buf):
    """bytes -> (list of arrays, list of names)"""
    r = _Reader(buf)
    magic, _reserved = r.take('QQ')
    if magic != LIST_MAGIC:
        raise ParamsError('not an MXNet NDArray list (magic 0x%x)' % magic)
    arrays = [_read_ndarray(r) for _ in range(r.take('Q'))]
    names = []
    for _ in range(r.take('Q')):
        names.append(bytes(r.raw(r.take('Q'))).decode('utf-8'))
    if names and len(names) != len(arrays):
        raise ParamsError('%d names for %d arrays' % (len(names), len(arrays)))
    return arrays, names


def dumps(named):
    """ordered mapping name -> numpy array -> bytes (V2 records, cpu context)"""
    out = [struct.pack('<QQQ', LIST_MAGIC, 0, len(named))]
    for a in named.values():
        a = np.ascontiguousarray(a)
        if a.dtype not in FLAGS:
            raise ParamsError('dtype %s has no MXNet type flag' % a.dtype)
        out.append(struct.pack('<IiI', V2_MAGIC, 0, a.ndim))
        out.append(struct.pack('<%dq' % a.ndim, *a.shape))
        out.append(struct.pack('<iii', 1, 0, FLAGS[a.dtype]))
        out.append(a.astype(a.dtype.newbyteorder('<'), copy=False).tobytes())
    out.append(struct.pack('<Q', len(named)))
    for n in named:
        b = n.encode('utf-8')
        out.append(struct.pack('<Q', len(b)) + b)
    return b''.join(out)


def load(path):
    """``mx.nd.load``: dict name -> array (or a list when the file carries no names)"""
    with open(path, 'rb') as f:
        arrays, names = loads(f.read())
    return dict(zip(names, arrays)) if names else arrays


def save(path, named):
    with open(path, 'wb') as f:
        f.write(dumps(named))


def load_checkpoint(prefix, epoch):
    """lib/utils/load_model.py:12-32 -> (arg_params, aux_params)"""
    save_dict = load('%s-%04d.params' % (prefix, epoch))
    arg_params, aux_params = {}, {}
    for k, v in save_dict.items():
        tp, name = k.split(':', 1)
        if tp == 'arg':
            arg_params[name] = v
        if tp == 'aux':
            aux_params[name] = v
    return arg_params, aux_params


def load_param(prefix, epoch, process=False):
    """lib/utils/load_model.py:47-67: with process=True every ``*_test`` entry replaces its trained twin"""
    arg_params, aux_params = load_checkpoint(prefix, epoch)
    if process:
        for test in [k for k in arg_params if '_test' in k]:
            arg_params[test.replace('_test', '')] = arg_params.pop(test)
    return arg_params, aux_params


def fold_bbox_test(arg_params, means, stds):
    """relation_rcnn/core/callback.py:54-61: the std/mean-folded copies written next to the trained bbox_pred layer"""
    stds = np.asarray(stds, np.float32); means = np.asarray(means, np.float32)
    w, b = np.asarray(arg_params['bbox_pred_weight']), np.asarray(arg_params['bbox_pred_bias'])
    out = dict(arg_params)
    out['bbox_pred_weight_test'] = (w.T * stds).T.astype(w.dtype)
    out['bbox_pred_bias_test'] = (b * stds + means).astype(b.dtype)
    return out


def save_checkpoint(prefix, epoch, arg_params, aux_params):
    """``mx.model.save_checkpoint`` without the symbol json: 'arg:' / 'aux:' prefixed NDArray list"""
    named = {'arg:%s' % k: np.asarray(v) for k, v in arg_params.items()}
    named.update({'aux:%s' % k: np.asarray(v) for k, v in aux_params.items()})
    save('%s-%04d.params' % (prefix, epoch), named)


def to_torch(params, device='cuda', names=None):
    """numpy dict -> float32 torch tensors on `device` (optionally only `names`)"""
    import torch
    return {k: torch.from_numpy(np.ascontiguousarray(v, np.float32)).to(device) for k, v in params.items()
            if names is None or k in names}
