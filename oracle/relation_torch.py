"""torch float64 restatement of the relation module with autograd -- the oracle for rn_relation_bwd (TEST INFRASTRUCTURE).

Same math as oracle/relation_np.py:relation_forward (SYM_REL:30-151, :267-268), written with torch ops so that
gradients come from autograd; the forward is checked against the numpy oracle (itself pinned by reference execution)
in tests/test_oracle_golden.py, so the gradients are those of the pinned function.  Boxes carry no gradient (the rois
come from zero-gradient CustomOps: proposal.py:170-173)."""
import math
import torch


def relation_forward(X, boxes, Wq, bq, Wk, bk, Wg, bg, Wout, bout, key_index=None, group=16, wave_length=1000.0,
                     residual_relu=False):
    f = X.dtype
    N, d = X.shape
    kidx = torch.arange(N) if key_index is None else (torch.arange(int(key_index)) if isinstance(key_index, int)
                                                      else torch.as_tensor(key_index, dtype=torch.long))
    b = boxes.to(f)
    w = b[:, 2] - b[:, 0] + 1; h = b[:, 3] - b[:, 1] + 1
    cx = 0.5 * (b[:, 0] + b[:, 2]); cy = 0.5 * (b[:, 1] + b[:, 3])
    eps = torch.stack([
        torch.log(torch.clamp(((cx[:, None] - cx[None, kidx]) / w[:, None]).abs(), min=1e-3)),
        torch.log(torch.clamp(((cy[:, None] - cy[None, kidx]) / h[:, None]).abs(), min=1e-3)),
        torch.log(w[:, None] / w[None, kidx]), torch.log(h[:, None] / h[None, kidx])], dim=2)          # [N,M,4]
    E = Wg.shape[1]
    nf = E // 8
    dim = torch.pow(torch.tensor(wave_length, dtype=f), (8.0 / E) * torch.arange(nf, dtype=f))
    div = (100.0 * eps)[..., None] / dim
    phi = torch.cat([torch.sin(div), torch.cos(div)], dim=-1).reshape(N, kidx.numel(), E)
    H = group
    g = torch.clamp(torch.relu(phi @ Wg.T + bg), min=1e-6).permute(0, 2, 1)                            # [N,H,M]
    dq = Wq.shape[0]; dk = dq // H
    Xk = X[kidx]
    q = (X @ Wq.T + bq).reshape(N, H, dk).permute(1, 0, 2)
    k = (Xk @ Wk.T + bk).reshape(-1, H, dk).permute(1, 0, 2)
    aff = torch.matmul(q, k.transpose(1, 2)) / math.sqrt(dk)                                            # [H,N,M]
    p = torch.softmax(torch.log(g) + aff.permute(1, 0, 2), dim=2)                                      # [N,H,M]
    out_t = (p.reshape(N * H, -1) @ Xk).reshape(N, H, d)
    dout = Wout.shape[0]; do = dout // H
    o = torch.einsum('nhc,hoc->nho', out_t, Wout.reshape(H, do, d)).reshape(N, dout) + bout
    return torch.relu(X + o) if residual_relu else o
