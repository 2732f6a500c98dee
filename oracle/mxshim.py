"""numpy-backed stand-in for the subset of ``mxnet`` (v1.1.0 semantics) the reference hot path uses.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package may import this file; it is used by
``oracle/refexec.py`` to *execute the reference's own Python* (``/root/reference/relation_rcnn/operator_py/
learn_nms.py``, ``symbols/*attention*.py`` ...) eagerly on numpy float32 arrays, in this container, to produce
the golden vectors under ``tests/golden/``.

MXNet 1.1.0 is not vendored in the reference (README.md:15,68) and cannot be imported here, so the op
semantics below are a restatement of the MXNet operator documentation, one function per ``mx.nd.*`` /
``mx.sym.*`` call the reference makes (the *composition* is the reference's own code).  MXNet itself was never
executed against them: what pins the single ops is ``tests/test_mxshim_cpu.py`` -- the worked examples of
MXNet's operator documentation (Reshape codes, take / pick / slice_axis / broadcast_to) and torch's independent
CPU implementations of the ops both libraries have (grouped convolution, linear, bmm, softmax, sort, gather,
smooth-L1).  Label for this file: parity PARTIAL -- pinned by published examples and an independent implementation, not by
an execution of MXNet.  float32 everywhere, like MXNet's default dtype; ``sort/argsort/arange`` return float32 like MXNet.

Symbols are evaluated eagerly: ``mx.sym.FullyConnected(name='query_1', data=x, num_hidden=n)`` with no
explicit weight looks ``query_1_weight`` / ``query_1_bias`` up in ``mxshim.PARAMS`` (a dict the caller fills),
which is how the reference addresses weights by name.
"""
import builtins
import math
import types
import numpy as np

F32 = np.float32
PARAMS = {}          # name -> np.ndarray, consulted by name-addressed layers (sym API)


class Context(object):
    def __init__(self, device_id=0):
        self.device_id = device_id


class ND(object):
    """Minimal NDArray: wraps a float32 numpy array."""
    __array_priority__ = 100

    def __init__(self, a, ctx=None):
        if isinstance(a, ND):
            a = a.a
        self.a = np.asarray(a, dtype=F32)
        self.context = ctx or Context(0)

    # --- numpy bridge
    def asnumpy(self):
        return self.a.copy()

    @property
    def shape(self):
        return tuple(self.a.shape)

    def __len__(self):
        return self.a.shape[0]

    def __getitem__(self, k):
        return ND(self.a[k])

    def __setitem__(self, k, v):
        self.a[k] = v.a if isinstance(v, ND) else v

    # --- arithmetic (scalars are cast to float32 like MXNet's *_scalar ops)
    def _b(self, o):
        return o.a if isinstance(o, ND) else F32(o)

    def __add__(self, o): return ND(self.a + self._b(o))
    __radd__ = __add__
    def __sub__(self, o): return ND(self.a - self._b(o))
    def __rsub__(self, o): return ND(self._b(o) - self.a)
    def __mul__(self, o): return ND(self.a * self._b(o))
    __rmul__ = __mul__
    def __truediv__(self, o): return ND(self.a / self._b(o))
    __div__ = __truediv__
    def __rtruediv__(self, o): return ND(self._b(o) / self.a)
    def __neg__(self): return ND(-self.a)

    # --- methods the reference calls on arrays
    def max(self, axis=None): return ND(self.a.max(axis=axis))
    def mean(self, axis=None): return ND(self.a.mean(axis=axis, dtype=F32))
    def transpose(self, axes=None): return ND(np.transpose(self.a, axes))
    def take(self, indices): return take(self, indices)
    def reshape(self, shape): return Reshape(self, shape=shape)


def _a(x):
    return x.a if isinstance(x, ND) else np.asarray(x, dtype=F32)


def _ival(v):
    # py2 integer division in the reference becomes float under py3 (e.g. dim[0] / group == 64.0)
    iv = int(v)
    assert iv == v, v
    return iv


# ----------------------------------------------------------------------------------------------
# creation
def array(x, ctx=None, dtype=None): return ND(np.asarray(x), ctx)
def zeros(shape, ctx=None, dtype=None): return ND(np.zeros(tuple(_ival(s) for s in np.atleast_1d(shape)), F32), ctx)
def zeros_like(x): return ND(np.zeros_like(_a(x)))
def full(shape, val, ctx=None): return ND(np.full(tuple(np.atleast_1d(shape)), val, F32))
def arange(start, stop=None, step=1.0, repeat=1, ctx=None, dtype=None):
    return ND(np.arange(start, stop, step).astype(F32))


# ----------------------------------------------------------------------------------------------
# shape ops
def Reshape(data, shape=None, name=None, **kw):
    """MXNet Reshape with special codes 0 (copy), -1 (infer), -2 (copy rest), -3 (merge two), -4 (split)."""
    x = _a(data)
    src = list(x.shape)
    out = []
    i = 0
    shape = [_ival(s) for s in shape]
    j = 0
    while j < len(shape):
        s = shape[j]
        if s == 0:
            out.append(src[i]); i += 1
        elif s == -1:
            out.append(-1); i += 1
        elif s == -2:
            out.extend(src[i:]); i = len(src)
        elif s == -3:
            out.append(src[i] * src[i + 1]); i += 2
        elif s == -4:
            d1, d2 = shape[j + 1], shape[j + 2]
            if d1 == -1: d1 = src[i] // d2
            if d2 == -1: d2 = src[i] // d1
            out.extend([d1, d2]); i += 1; j += 2
        else:
            out.append(s); i += 1
        j += 1
    return ND(x.reshape(out))


reshape = Reshape


def expand_dims(data, axis): return ND(np.expand_dims(_a(data), axis))
def transpose(data=None, axes=None, name=None):
    x = _a(data)
    if axes is None or len(axes) == 0:
        return ND(x.T)
    return ND(np.transpose(x, axes))
def concat(*args, **kw):
    dim = kw.get('dim', 1)
    return ND(np.concatenate([_a(x) for x in args], axis=dim))
Concat = concat
def split(data=None, num_outputs=None, axis=1, squeeze_axis=0, name=None):
    parts = np.split(_a(data), num_outputs, axis=axis)
    if squeeze_axis:
        parts = [p.squeeze(axis=axis) for p in parts]
    return [ND(p) for p in parts]
SliceChannel = split
def slice_axis(data=None, axis=0, begin=0, end=None, name=None):
    x = _a(data)
    sl = [builtins.slice(None)] * x.ndim
    sl[axis] = builtins.slice(begin, end)
    return ND(x[tuple(sl)])
def slice(data=None, begin=None, end=None, name=None):   # noqa: A001 (mirrors mx.sym.slice)
    x = _a(data)
    sl = tuple(builtins.slice(b, e) for b, e in zip(begin, end))
    return ND(x[sl])
def tile(data=None, reps=None): return ND(np.tile(_a(data), reps))
def reverse(data=None, axis=0): return ND(np.flip(_a(data), axis=axis))
def broadcast_to(data=None, shape=None):
    x = _a(data)
    shp = tuple(x.shape[i] if s == 0 else s for i, s in enumerate(shape))
    return ND(np.broadcast_to(x, shp).copy())
def BlockGrad(data=None, name=None): return ND(_a(data))
def identity(data=None): return ND(_a(data))


# ----------------------------------------------------------------------------------------------
# elementwise
def broadcast_add(lhs=None, rhs=None): return ND(_a(lhs) + _a(rhs))
def broadcast_minus(lhs=None, rhs=None): return ND(_a(lhs) - _a(rhs))
broadcast_sub = broadcast_minus
def broadcast_mul(lhs=None, rhs=None): return ND(_a(lhs) * _a(rhs))
def broadcast_div(lhs=None, rhs=None): return ND(_a(lhs) / _a(rhs))
def broadcast_power(lhs=None, rhs=None): return ND(np.power(_a(lhs), _a(rhs)))
def broadcast_maximum(lhs=None, rhs=None): return ND(np.maximum(_a(lhs), _a(rhs)))
def broadcast_minimum(lhs=None, rhs=None): return ND(np.minimum(_a(lhs), _a(rhs)))
def maximum(left=None, right=None):
    l = _a(left) if isinstance(left, ND) else F32(left)
    r = _a(right) if isinstance(right, ND) else F32(right)
    return ND(np.maximum(l, r))
def minimum(left=None, right=None):
    l = _a(left) if isinstance(left, ND) else F32(left)
    r = _a(right) if isinstance(right, ND) else F32(right)
    return ND(np.minimum(l, r))
def sin(data=None): return ND(np.sin(_a(data)))
def cos(data=None): return ND(np.cos(_a(data)))
def log(data=None): return ND(np.log(_a(data)))
def exp(data=None): return ND(np.exp(_a(data)))
def abs(data=None): return ND(np.abs(_a(data)))   # noqa: A001
def relu(data=None): return ND(np.maximum(_a(data), F32(0)))
def sigmoid(data=None): return ND(F32(1) / (F32(1) + np.exp(-_a(data))))
def Activation(data=None, act_type=None, name=None):
    return {'relu': relu, 'sigmoid': sigmoid}[act_type](data)
def mean(data=None, axis=None, name=None): return ND(_a(data).mean(axis=axis, dtype=F32))
def max(data=None, axis=None, name=None): return ND(_a(data).max(axis=axis))   # noqa: A001
def sum(data=None, axis=None, name=None):   # noqa: A001
    x = _a(data)
    if axis == 1 and x.ndim == 2:       # sequential float32 accumulation (MXNet CPU reduce order), not numpy's pairwise
        acc = np.zeros(x.shape[0], F32)
        for j in range(x.shape[1]):
            acc = (acc + x[:, j].astype(F32)).astype(F32)
        return ND(acc)
    return ND(x.sum(axis=axis, dtype=F32))


# ----------------------------------------------------------------------------------------------
# layers
def _param(name, suffix, given):
    if given is not None:
        return _a(given)
    return np.asarray(PARAMS[name + suffix], dtype=F32)


def FullyConnected(data=None, weight=None, bias=None, num_hidden=None, name=None, no_bias=False, flatten=True):
    """y = flatten(x) . W^T + b, W is [num_hidden, in] (MXNet layout)."""
    x = _a(data)
    x2 = x.reshape(x.shape[0], -1)
    w = _param(name, '_weight', weight)
    assert w.shape == (num_hidden, x2.shape[1]), (w.shape, num_hidden, x2.shape)
    y = x2 @ w.T
    if not no_bias:
        y = y + _param(name, '_bias', bias)
    return ND(y.astype(F32))


def Convolution(data=None, weight=None, bias=None, kernel=None, num_filter=None, num_group=1, name=None,
                no_bias=False, **kw):
    """Only the 1x1 grouped convolution the relation module uses (SYM_REL:147-149, LNMS:118-120)."""
    assert tuple(kernel) == (1, 1)
    x = _a(data)                                   # [B, Cin, H, W]
    w = _param(name, '_weight', weight)            # [Cout, Cin/g, 1, 1]
    B, Cin, H, W = x.shape
    g = num_group
    cout_g, cin_g = num_filter // g, Cin // g
    assert w.shape[:2] == (num_filter, cin_g), (w.shape, num_filter, cin_g)
    w2 = w.reshape(g, cout_g, cin_g)
    xg = x.reshape(B, g, cin_g, H * W)
    y = np.einsum('goc,bgcs->bgos', w2, xg, optimize=True).reshape(B, num_filter, H, W)
    if not no_bias:
        y = y + _param(name, '_bias', bias).reshape(1, -1, 1, 1)
    return ND(y.astype(F32))


def batch_dot(lhs=None, rhs=None, transpose_a=False, transpose_b=False, name=None):
    a, b = _a(lhs), _a(rhs)
    if transpose_a: a = np.swapaxes(a, 1, 2)
    if transpose_b: b = np.swapaxes(b, 1, 2)
    return ND(np.matmul(a, b).astype(F32))


def dot(lhs=None, rhs=None, transpose_a=False, transpose_b=False, name=None):
    a, b = _a(lhs), _a(rhs)
    if transpose_a: a = a.T
    if transpose_b: b = b.T
    return ND((a @ b).astype(F32))


def softmax(data=None, axis=-1, name=None):
    x = _a(data)
    m = x.max(axis=axis, keepdims=True)
    e = np.exp(x - m)
    return ND(e / e.sum(axis=axis, keepdims=True, dtype=F32))


def SoftmaxActivation(data=None, mode='instance', name=None):
    """softmax over axis 1 (mode='instance'); exp(x - max) summed sequentially in float32 like mshadow's Softmax"""
    x = _a(data).astype(F32)
    e = np.exp(x - x.max(axis=1, keepdims=True))
    s = np.zeros(x.shape[0], F32)
    for c in range(x.shape[1]):
        s = (s + e[:, c]).astype(F32)
    return ND((e / s[:, None]).astype(F32))


def smooth_l1(data=None, scalar=1.0, name=None):
    """f(x) = 0.5 (sigma x)^2 if |x| < 1/sigma^2 else |x| - 0.5/sigma^2   (mshadow_op::smooth_l1_loss)"""
    x = _a(data).astype(F32)
    b = F32(scalar); bsq = F32(b * b); ibsq = F32(1.0) / bsq
    inner = F32(0.5) * ((x * b) * (x * b))
    return ND(np.where(x > ibsq, x - F32(0.5) * ibsq, np.where(x < -ibsq, -x - F32(0.5) * ibsq, inner)).astype(F32))


def sort(data=None, axis=-1, is_ascend=True, name=None):
    x = _a(data)
    idx = _argsort(x, axis, is_ascend)
    return ND(np.take_along_axis(x, idx, axis=axis))


def _argsort(x, axis, is_ascend):
    # tie order is unspecified in MXNet; we fix "lower index first" (stable) in both directions
    return np.argsort(x if is_ascend else -x, axis=axis, kind='stable')


def argsort(data=None, axis=-1, is_ascend=True, name=None):
    return ND(_argsort(_a(data), axis, is_ascend).astype(F32))     # float32 indices, like MXNet


def take(a=None, indices=None, axis=0, mode='clip', name=None):
    x = _a(a)
    idx = np.clip(_a(indices).astype(np.int64), 0, x.shape[0] - 1)
    return ND(x[idx])


def pick(data=None, index=None, axis=-1, keepdims=False, name=None):
    x = _a(data)
    idx = np.clip(_a(index).astype(np.int64), 0, x.shape[axis] - 1)
    out = np.take_along_axis(x, np.expand_dims(idx, axis), axis=axis)
    return ND(out if keepdims else out.squeeze(axis))


# ----------------------------------------------------------------------------------------------
# operator protocol (mx.operator.CustomOp / CustomOpProp / register)
class CustomOp(object):
    def assign(self, dst, req, src):
        if req == 'null':
            return
        s = src.a if isinstance(src, ND) else np.asarray(src, dtype=F32)
        if req == 'add':
            dst.a[...] = dst.a + s
        else:
            if dst.a.shape != np.shape(s):
                dst.a = np.array(s, dtype=F32)       # shim convenience: shape comes from the op
            else:
                dst.a[...] = s


class CustomOpProp(object):
    def __init__(self, need_top_grad=False):
        self.need_top_grad = need_top_grad


REGISTRY = {}


def register(name):
    def deco(cls):
        REGISTRY[name] = cls
        return cls
    return deco


def build_module():
    """Return an object usable as ``mx`` (``mx.nd``, ``mx.sym``, ``mx.symbol``, ``mx.operator``, ``mx.contrib``)."""
    import sys
    me = sys.modules[__name__]
    mx = types.ModuleType('mxnet')
    nd = types.ModuleType('mxnet.nd')
    for k in dir(me):
        if not k.startswith('__'):
            setattr(nd, k, getattr(me, k))
    nd.NDArray = ND
    operator = types.ModuleType('mxnet.operator')
    operator.CustomOp = CustomOp
    operator.CustomOpProp = CustomOpProp
    operator.register = register
    mx.nd = nd
    mx.ndarray = nd
    mx.sym = nd
    mx.symbol = nd
    mx.operator = operator
    mx.cpu = lambda i=0: Context(i)
    mx.gpu = lambda i=0: Context(i)
    return mx, nd, operator
