"""Parity of the multi-tensor EMA (bit-exact) and the fused BYOL loss fwd/bwd with the oracle / reference fixtures."""
import numpy as np
import pytest
import torch

import oracle
from conftest import rel_err
from nrse_b200 import ops
from nrse_b200.utils import synthetic

pytestmark = pytest.mark.gpu


def test_ema_golden_bit_exact(dev, golden):
    g = golden("ema")
    n = int(g["n"])
    online = [torch.from_numpy(g[f"o{i}"]).to(dev) for i in range(n)]
    for decay in (0.996, 0.997):
        target = [torch.from_numpy(g[f"t{i}"]).to(dev) for i in range(n)]
        plan = ops.EmaPlan(online, target)
        plan.step(decay)
        plan.step(decay)
        for i in range(n):
            assert np.array_equal(target[i].cpu().numpy(), g[f"r{int(decay * 1000)}_{i}"]), (decay, i)


def test_ema_ragged_sizes_and_identities(dev):
    rs = np.random.RandomState(0)
    shapes = [(1,), (3,), (8,), (4097,), (16384,), (16385,), (16384 * 3 + 5,), (128, 65), (0,), (1 << 20,)]
    online = [rs.standard_normal(s).astype(np.float32) for s in shapes]
    target = [rs.standard_normal(s).astype(np.float32) for s in shapes]
    ref = oracle.ema_update([torch.from_numpy(o) for o in online], [torch.from_numpy(t) for t in target], 0.996)
    t_dev = [torch.from_numpy(t).to(dev) for t in target]
    o_dev = [torch.from_numpy(o).to(dev) for o in online]
    ops.EmaPlan(o_dev, t_dev).step(0.996)
    for a, b in zip(t_dev, ref):
        assert np.array_equal(a.cpu().numpy(), b.numpy())
    # decay 1.0 is the identity, decay 0.0 copies the online weights
    t1 = [torch.from_numpy(t).to(dev) for t in target]
    ops.EmaPlan(o_dev, t1).step(1.0)
    for a, b in zip(t1, target):
        assert np.array_equal(a.cpu().numpy(), b)
    ops.EmaPlan(o_dev, t1).step(0.0)
    for a, b in zip(t1, online):
        assert np.array_equal(a.cpu().numpy(), b + 0.0 * 0)  # 0*t + 1*o
    # misaligned views (odd offsets into one storage) take the scalar path
    base_o = torch.from_numpy(rs.standard_normal(5000).astype(np.float32)).to(dev)
    base_t = torch.from_numpy(rs.standard_normal(5000).astype(np.float32)).to(dev)
    o_v, t_v = base_o[1:4098], base_t[3:4100]
    want = oracle.ema_update([o_v.cpu()], [t_v.cpu()], 0.99)[0]
    ops.EmaPlan([o_v], [t_v]).step(0.99)
    assert np.array_equal(t_v.cpu().numpy(), want.numpy())


def test_ema_full_model_size_linearity(dev):
    """317.5 M parameters (WavLM-large encoder + projector sizes): EMA twice with decay d equals the closed form
    on a sampled subset, and untouched neighbours stay untouched (checksum over a guard tensor)."""
    sizes = [8] * 24 + [16] * 24 + [128] + [512] * 40 + [1024] * 227 + [4096] * 24 + [5120] * 2 + [524288] * 3 + \
            [786432] * 4 + [1048576] * 98 + [4194304] * 48 + [8388608]
    assert sum(sizes) == 317556416
    gen = torch.Generator(device=dev).manual_seed(0)
    online = [torch.randn(s, device=dev, generator=gen) for s in sizes]
    target = [torch.randn(s, device=dev, generator=gen) for s in sizes]
    guard = torch.full((1024,), 7.0, device=dev)
    before = [t[:64].clone() for t in target]
    plan = ops.EmaPlan(online, target)
    assert plan.numel == 317556416
    plan.step(0.996)
    for t0, o, t in zip(before, online, target):
        want = oracle.ema_update([o[:64].cpu()], [t0.cpu()], 0.996)[0]
        assert np.array_equal(t[:64].cpu().numpy(), want.numpy())
    last = target[-1]
    want_tail = oracle.ema_update([online[-1][-64:].cpu()], [last[-64:].cpu()], 0.996)[0]
    plan.step(0.996)
    assert np.array_equal(last[-64:].cpu().numpy(), want_tail.numpy())
    assert float(guard.sum()) == 7.0 * 1024


@pytest.mark.parametrize("name", ["b64", "b2", "b5_d96"])
def test_loss_golden(dev, golden, name):
    g = golden("byol_loss")
    p = torch.from_numpy(g[f"{name}_p"]).to(dev).requires_grad_(True)
    z = torch.from_numpy(g[f"{name}_z"]).to(dev)
    loss = ops.byol_loss(p, z)
    loss.backward()
    assert loss.shape == () and loss.dtype == torch.float32
    assert abs(loss.item() - float(g[f"{name}_loss"])) <= 2e-6 * max(1.0, abs(float(g[f"{name}_loss"])))
    assert rel_err(p.grad.cpu().numpy(), g[f"{name}_grad"]) < 1e-5


def test_loss_properties_and_bf16(dev):
    p, z = synthetic.embeddings(37, 1024, seed=1)
    pd, zd = torch.from_numpy(p).to(dev), torch.from_numpy(z).to(dev)
    assert abs(ops.byol_loss(pd, pd).item()) < 1e-6            # identical inputs -> 0
    assert abs(ops.byol_loss(pd, -pd).item() - 4.0) < 1e-5     # opposite -> 4
    l = ops.byol_loss(pd, zd).item()
    assert 0.0 <= l <= 4.0
    ref = oracle.byol_loss(torch.from_numpy(p), torch.from_numpy(z)).item()
    assert abs(l - ref) < 1e-6 * max(1, abs(ref)) + 2e-6
    # bf16 inputs (autocast heads): tolerance 1e-2 relative (north star); measured far below
    pb = pd.bfloat16().requires_grad_(True)
    lb = ops.byol_loss(pb, zd.bfloat16())
    lb.backward()
    pr = pb.detach().float().cpu().requires_grad_(True)
    lr = oracle.byol_loss(pr, zd.bfloat16().float().cpu())
    lr.backward()
    assert abs(lb.item() - lr.item()) < 1e-2 * abs(lr.item())
    assert pb.grad.dtype == torch.bfloat16
    assert rel_err(pb.grad.float().cpu().numpy(), pr.grad.numpy()) < 1e-2
    # per-row cosine used by evaluate_byol.py
    rows = ops.cosine_rows(pd, zd).cpu().numpy()
    want = torch.nn.functional.cosine_similarity(torch.from_numpy(p), torch.from_numpy(z), dim=1).numpy()
    assert np.abs(rows - want).max() < 1e-5
    # degenerate rows: zeros give similarity 0 -> loss 2, gradient finite
    zp = torch.zeros(4, 64, device=dev, requires_grad=True)
    lz = ops.byol_loss(zp, torch.zeros(4, 64, device=dev))
    lz.backward()
    rz = torch.zeros(4, 64, requires_grad=True)
    lrz = oracle.byol_loss(rz, torch.zeros(4, 64))
    lrz.backward()
    assert abs(lz.item() - lrz.item()) < 1e-6
    assert torch.isfinite(zp.grad).all()
