"""Parity of the fused optimizer tail (clip_grad_norm_ + AdamW + EMA, ref:train_byol.py:67-71) with the oracle and
with the fixture produced by the real torch.optim.AdamW + the reference's EMA (fp32, 1e-6)."""
import numpy as np
import pytest
import torch

import oracle
from conftest import rel_err
from nrse_b200 import ops
from nrse_b200.train import FusedAdamWEma
from nrse_b200.utils import synthetic
from test_oracle_golden import _optim_fixture

pytestmark = pytest.mark.gpu
TOL = 1e-6  # north star: EMA / fp32 elementwise within 1e-6 relative (norm-relative per tensor)


def _close(a, b, tol=TOL, floor=0.0):
    """max|a-b| <= tol * max(max|b|, floor).  ``floor`` = the magnitude of the terms that were summed: a moment of a
    one-element tensor can cancel to far below the gradients it was built from, and its error cannot."""
    a = a.detach().cpu().numpy().astype(np.float64)
    b = (b if isinstance(b, np.ndarray) else b.numpy()).astype(np.float64)
    if a.size == 0:
        return True
    return float(np.abs(a - b).max()) <= tol * max(float(np.abs(b).max()), floor)


def test_golden_three_steps(dev, golden):
    g = golden("optim_step")
    names, params, targets, has_grad, grads = _optim_fixture(g)
    P = [torch.nn.Parameter(p.to(dev)) for p in params]
    T = [None if t is None else t.to(dev) for t in targets]
    opt = FusedAdamWEma(P, lr=float(g["lr"]), weight_decay=float(g["weight_decay"]), max_grad_norm=1.0,
                        ema_pairs=[(p, t) for p, t in zip(P, T) if t is not None], ema_decay=float(g["ema_decay"]))
    for k in range(int(g["steps"])):
        for p, gk in zip(P, grads(k)):
            p.grad = None if gk is None else gk.to(dev)
        opt.step()
        assert abs(float(opt.last_grad_norm) - float(g["norms"][k])) <= 2e-6 * float(g["norms"][k])
    opt.sync_state()  # step counts are kept per bucket and written to state[p]['step'] on demand
    for i, name in enumerate(names):
        assert _close(P[i], g[f"p_{i}"]), name
        if T[i] is not None:
            assert _close(T[i], g[f"t_{i}"]), name
        if has_grad[i]:
            st = opt.state[P[i]]
            assert _close(st["exp_avg"], g[f"m_{i}"]) and _close(st["exp_avg_sq"], g[f"v_{i}"]), name
            assert int(st["step"]) == int(g["steps"])
        else:
            assert len(opt.state[P[i]]) == 0  # torch skips parameters without a gradient; so do we


@pytest.mark.parametrize("max_norm,wd", [(1.0, 1e-2), (0.0, 0.0), (0.05, 1e-5)])
def test_ragged_sizes_vs_oracle(dev, max_norm, wd):
    rs = np.random.RandomState(7)
    shapes = [(1,), (3,), (8,), (4097,), (16384,), (16385,), (16384 * 2 + 5,), (128, 65), (1 << 18,)]
    p0 = [rs.standard_normal(s).astype(np.float32) for s in shapes]
    t0 = [rs.standard_normal(s).astype(np.float32) if i % 3 != 1 else None for i, s in enumerate(shapes)]
    cp = [torch.from_numpy(p.copy()) for p in p0]
    ct = [None if t is None else torch.from_numpy(t.copy()) for t in t0]
    cm = [torch.zeros_like(p) for p in cp]
    cv = [torch.zeros_like(p) for p in cp]
    P = [torch.nn.Parameter(torch.from_numpy(p.copy()).to(dev)) for p in p0]
    T = [None if t is None else torch.from_numpy(t.copy()).to(dev) for t in t0]
    opt = FusedAdamWEma(P, lr=3e-3, betas=(0.9, 0.98), eps=1e-6, weight_decay=wd, max_grad_norm=max_norm,
                        ema_pairs=[(p, t) for p, t in zip(P, T) if t is not None], ema_decay=0.99)
    for k in range(4):
        gs = [(0.02 * rs.standard_normal(s)).astype(np.float32) for s in shapes]
        if k == 2:
            gs[4] = None  # a parameter without a gradient this step: AdamW skips it, its EMA twin still moves
        for p, gk in zip(P, gs):
            p.grad = None if gk is None else torch.from_numpy(gk).to(dev)
        opt.step()
        # oracle: global clip, then per-tensor AdamW (tensor 4 is one step behind after k == 2), then the EMA
        cg = [None if gk is None else torch.from_numpy(gk.copy()) for gk in gs]
        with torch.no_grad():
            if max_norm > 0:
                oracle.clip_grad_norm(cg, max_norm)
            for i in range(len(shapes)):
                if cg[i] is None:
                    continue
                step_i = k + 1 if (i != 4 or k < 2) else k
                oracle.adamw_step([cp[i]], [cg[i]], [cm[i]], [cv[i]], step_i, 3e-3, (0.9, 0.98), 1e-6, wd)
            idx = [i for i, t in enumerate(ct) if t is not None]
            new = oracle.ema_update([cp[i] for i in idx], [ct[i] for i in idx], 0.99)
            for i, t in zip(idx, new):
                ct[i].copy_(t)
    # With clipping on, the moments inherit the error of the gradient-norm REDUCTION, and that is the reference's, not the
    # kernel's: torch's fp32 CPU norm of these gradients is 1.0e-6 .. 1.8e-6 below the float64 value (the kernel
    # accumulates in fp64 and is exact to fp32 rounding), so exp_avg (~ coef) may differ by 2e-6 and exp_avg_sq
    # (~ coef^2) by 4e-6 -- per element, which is what a one-element tensor measures.  Without clipping: 1e-6.
    tol_m, tol_v = (3e-6, 5e-6) if max_norm > 0 else (TOL, TOL)
    for i in range(len(shapes)):
        assert _close(P[i], cp[i]), i
        assert _close(opt.state[P[i]]["exp_avg"], cm[i], tol=tol_m, floor=0.02 if max_norm <= 0 else 0.002), i
        assert _close(opt.state[P[i]]["exp_avg_sq"], cv[i], tol=tol_v), i
        if T[i] is not None:
            assert _close(T[i], ct[i]), i


def test_misaligned_views_take_the_scalar_path(dev):
    rs = np.random.RandomState(3)
    base = [torch.from_numpy(rs.standard_normal(6000).astype(np.float32)).to(dev) for _ in range(5)]
    p, g, m, v, t = base[0][1:4098], base[1][3:4100], base[2][2:4099], base[3][1:4098], base[4][5:4102]
    m.zero_(); v.zero_()
    cp, cg, cm, cv, ct = (x.cpu().clone() for x in (p, g, m, v, t))
    tab = ops.OptimChunkTable()
    tab.update([p], [g], [m], [v], [t])
    part = torch.zeros(ops.OptimChunkTable.partials_count(), dtype=torch.float64, device=dev)
    norm = torch.zeros(1, device=dev)
    tab.grad_sqnorm(part)
    tab.clip_adamw_ema(lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, step=1, max_grad_norm=1.0,
                       ema_decay=0.996, partials=part, norm_out=norm)
    want = oracle.clip_adamw_ema_step([cp], [cg], [cm], [cv], [ct], 1, 1e-3, max_norm=1.0, ema_decay=0.996)
    assert abs(float(norm) - float(want)) < 1e-6 * float(want)
    for a, b in ((p, cp), (m, cm), (v, cv), (t, ct)):
        assert _close(a, b)
    # neighbours of the views are untouched
    assert float(base[0][0]) == float(base[0].cpu()[0]) and torch.equal(base[0][4098:].cpu(), base[0].cpu()[4098:])


def test_state_dict_interchange_with_torch_adamw(dev):
    """A checkpoint written by the reference (torch.optim.AdamW state) resumes under FusedAdamWEma and vice versa."""
    torch.manual_seed(0)
    w0 = [torch.randn(300, 17), torch.randn(1025)]
    A = [torch.nn.Parameter(w.clone().to(dev)) for w in w0]
    B = [torch.nn.Parameter(w.clone().to(dev)) for w in w0]
    ref = torch.optim.AdamW(A, lr=1e-3, weight_decay=1e-2)
    mine = FusedAdamWEma(B, lr=1e-3, weight_decay=1e-2)
    gs = [[torch.randn_like(w).to(dev) * 0.1 for w in w0] for _ in range(4)]
    for k in range(2):
        for p, q, g in zip(A, B, gs[k]):
            p.grad, q.grad = g.clone(), g.clone()
        ref.step(); mine.step()
    sd_ref, sd_mine = ref.state_dict(), mine.state_dict()
    assert sd_ref["param_groups"][0].keys() == sd_mine["param_groups"][0].keys()
    assert sd_ref["state"][0].keys() == sd_mine["state"][0].keys()
    ref.load_state_dict(sd_mine); mine.load_state_dict(sd_ref)   # swap the two histories
    for k in range(2, 4):
        for p, q, g in zip(A, B, gs[k]):
            p.grad, q.grad = g.clone(), g.clone()
        ref.step(); mine.step()
    for p, q in zip(A, B):
        assert _close(q, p.detach().cpu())


def test_large_tensor_against_stock_torch_on_the_same_gpu(dev):
    """8 M + 4 M + ragged elements, 3 steps: fused kernel vs clip_grad_norm_ + torch.optim.AdamW + the per-tensor EMA
    expression, all on the GPU (size-independent cross-check; the CPU oracle covers the small cases)."""
    torch.manual_seed(1)
    sizes = [1 << 23, 1 << 22, 123457]
    A = [torch.nn.Parameter(torch.randn(n, device=dev)) for n in sizes]
    B = [torch.nn.Parameter(a.detach().clone()) for a in A]
    TA = [torch.randn(n, device=dev) for n in sizes]
    TB = [t.clone() for t in TA]
    ref = torch.optim.AdamW(A, lr=1e-4, weight_decay=1e-5, foreach=False)
    mine = FusedAdamWEma(B, lr=1e-4, weight_decay=1e-5, max_grad_norm=1.0, ema_pairs=zip(B, TB), ema_decay=0.996)
    for k in range(3):
        for p, q in zip(A, B):
            g = torch.randn_like(p) * (1e-3 if k < 2 else 1e-5)
            p.grad, q.grad = g, g.clone()
        n_ref = torch.nn.utils.clip_grad_norm_(A, 1.0)
        ref.step()
        for i in range(len(A)):
            TA[i] = 0.996 * TA[i] + (1 - 0.996) * A[i].data
        mine.step()
        assert abs(float(n_ref) - float(mine.last_grad_norm)) < 1e-5 * float(n_ref)
    for p, q, ta, tb in zip(A, B, TA, TB):
        assert _close(q, p.detach().cpu()) and _close(tb, ta.cpu())
