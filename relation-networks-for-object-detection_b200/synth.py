"""Seeded synthetic inputs of the hot path for bench.py / tools (SURVEY.md section 8d generators).  Product-side copy of
the generators so that nothing outside tests / the CPU arm has to import ``oracle``; tests/test_oracle_golden.py checks the
two stay identical."""
import numpy as np


def make_boxes(rng, N, canvas=(1000.0, 600.0)):
    """x1~U(0,900), y1~U(0,500), w,h~U(16,400), clipped to the image."""
    x1 = rng.uniform(0, 900, N); y1 = rng.uniform(0, 500, N)
    w = rng.uniform(16, 400, N); h = rng.uniform(16, 400, N)
    x2 = np.minimum(x1 + w, canvas[0] - 1); y2 = np.minimum(y1 + h, canvas[1] - 1)
    return np.stack([x1, y1, x2, y2], 1).astype(np.float32)


def make_relation_case(seed, N, d, H, E=64, init='fan_in', bias_g=True, M=None, dq=None, dout=None):
    """Weights: 'ref' = N(0,0.01) (SYM_REL:327-344) or 'fan_in' = N(0,1/sqrt(fan_in)) so logits are O(1)."""
    rng = np.random.default_rng(seed)
    dout = dout or d
    boxes = make_boxes(rng, N)
    X = (rng.standard_normal((N, d)) * 0.5).astype(np.float32)
    sd = (lambda fan: 0.01) if init == 'ref' else (lambda fan: 1.0 / np.sqrt(fan))
    dq = dq or d
    p = dict(
        X=X, boxes=boxes,
        Wq=(rng.standard_normal((dq, d)) * sd(d)).astype(np.float32), bq=np.zeros(dq, np.float32),
        Wk=(rng.standard_normal((dq, d)) * sd(d)).astype(np.float32), bk=np.zeros(dq, np.float32),
        Wg=(rng.standard_normal((H, E)) * (0.01 if init == 'ref' else 0.125)).astype(np.float32),
        bg=(rng.uniform(0, 0.5, H) if bias_g else np.zeros(H)).astype(np.float32),
        Wout=(rng.standard_normal((dout, d)) * sd(d)).astype(np.float32),
        bout=(rng.standard_normal(dout) * 0.01).astype(np.float32))
    if init != 'ref':
        p['bq'] = (rng.standard_normal(dq) * 0.1).astype(np.float32)
        p['bk'] = (rng.standard_normal(dq) * 0.1).astype(np.float32)
    if M is not None:
        p['M'] = M
    return p
