"""Parity of the fused SNR-mix + peak-norm + z-norm kernel with the oracle / reference fixtures (fp32, 1e-6)."""
import numpy as np
import pytest
import torch

import oracle
from conftest import rel_err
from nrse_b200 import ops
from nrse_b200.utils import synthetic

pytestmark = pytest.mark.gpu
TOL = 1e-6  # north star: mixing within 1e-6 relative (norm-relative)


@pytest.fixture(params=[4, 5, 0], ids=["resident", "resident1024", "generic"], autouse=True)
def mix_variant(request):
    """Every test runs on both kernels: on-chip resident rows (default; both CTA shapes) and the generic re-read-from-L2
    kernel that unaligned / tiled-noise / very long rows fall back to."""
    ops.set_mix_variant(request.param)
    yield request.param
    ops.set_mix_variant(4)


def _run(dev, clean, noise, snr_idx, table, peak_norm=True):
    c, n, st = ops.mix_normalize(torch.from_numpy(np.ascontiguousarray(clean)).to(dev),
                                 torch.from_numpy(np.ascontiguousarray(noise)).to(dev),
                                 torch.from_numpy(np.asarray(snr_idx, dtype=np.int32)).to(dev),
                                 [float(v) for v in table], peak_norm)
    return (None if c is None else c.cpu().numpy()), n.cpu().numpy(), st.cpu().numpy()


def test_golden_byol(dev, golden):
    g = golden("mix_byol")
    c, n, st = _run(dev, g["clean"], g["noise"], g["snr_idx"], g["snr_table"])
    assert st.tolist() == [0] * len(st)
    assert rel_err(c, g["clean_out"]) < TOL
    assert rel_err(n, g["noisy_out"]) < TOL
    for row in range(len(st)):  # per-row as well: a batch-wide max would hide a bad quiet row
        assert rel_err(n[row], g["noisy_out"][row]) < TOL
        assert rel_err(c[row], g["clean_out"][row]) < TOL


def test_golden_emotion_short_and_long_noise(dev, golden):
    g = golden("mix_emotion")
    for rows, noise in ((slice(0, 2), g["noise_short"]), (slice(2, 4), g["noise_long"])):
        c, n, st = _run(dev, g["clean"][rows], noise, g["snr_idx"][rows], g["snr_table"], peak_norm=False)
        assert c is None and st.tolist() == [0, 0]
        assert rel_err(n, g["noisy_out"][rows]) < TOL


def test_golden_edge_statuses(dev, golden):
    g = golden("mix_edge")
    c_ref, n_ref, st_ref = oracle.mix_normalize_batch(g["clean"], g["noise"], np.zeros(8, np.int32), [10.0])
    c, n, st = _run(dev, g["clean"], g["noise"], np.zeros(8, np.int32), [10.0])
    assert st.tolist() == st_ref.tolist()
    assert [bool(s) for s in st[:6]] == [True] * 6 and st[6] == 0 and st[7] == 0
    for b in range(8):
        if st[b] != 0:
            assert not c[b].any() and not n[b].any()  # rejected rows are zero-filled
        else:
            assert rel_err(n[b], n_ref[b].numpy()) < TOL and rel_err(c[b], c_ref[b].numpy()) < TOL
    # emotion mode: a failed mix keeps the clean waveform (emotion_dataset.py:193-194)
    _, n_ref, st_ref = oracle.mix_normalize_batch(g["clean"], g["noise"], np.zeros(8, np.int32), [10.0], peak_norm=False)
    _, n, st = _run(dev, g["clean"], g["noise"], np.zeros(8, np.int32), [10.0], peak_norm=False)
    assert st.tolist() == st_ref.tolist()
    for b in range(8):
        r = n_ref[b].numpy()
        if np.isnan(r).any():
            assert np.isnan(n[b]).any()
        else:
            assert rel_err(n[b], r) < TOL


@pytest.mark.parametrize("B,L,Ln,peak", [(5, 4000, 4000, True), (3, 4001, 4001, True), (4, 3998, 1500, True),
                                         (3, 6000, 7003, False), (2, 401, 401, True), (7, 16000, 16000, False),
                                         (2, 64000, 64000, True), (2, 192000, 192000, True), (3, 32000, 32004, False),
                                         (1, 240000, 240000, True), (9, 8, 8, True),
                                         (1, 400000, 400000, True), (2, 80000, 80000, True),
                                         (3, 16388, 16388, False), (2, 40960, 40960, True), (2, 40964, 40964, True)])
def test_random_vs_oracle(dev, B, L, Ln, peak):
    clean, noise, snr_idx, table = synthetic.waveforms(B, L, seed=100 + L % 97, n_noise=Ln)
    c_ref, n_ref, st_ref = oracle.mix_normalize_batch(clean, noise, snr_idx, table, peak_norm=peak)
    c, n, st = _run(dev, clean, noise, snr_idx, table, peak)
    assert st.tolist() == st_ref.tolist()
    for b in range(B):
        assert rel_err(n[b], n_ref[b].numpy()) < TOL
        if peak:
            assert rel_err(c[b], c_ref[b].numpy()) < TOL


def test_dc_offset_and_quiet_rows(dev):
    """Large DC offset (var << mean^2) and very quiet rows stress the algebraic variance."""
    clean, noise, snr_idx, table = synthetic.waveforms(4, 8000, seed=3)
    clean[0] += 0.5
    noise[1] += 2.0
    clean[2] *= 1e-3
    noise[3] *= 1e-3
    c_ref, n_ref, st_ref = oracle.mix_normalize_batch(clean, noise, snr_idx, table)
    c, n, st = _run(dev, clean, noise, snr_idx, table)
    assert st.tolist() == st_ref.tolist() == [0, 0, 0, 0]
    for b in range(4):
        assert rel_err(n[b], n_ref[b].numpy()) < 2 * TOL
        assert rel_err(c[b], c_ref[b].numpy()) < 2 * TOL


def test_full_size_properties(dev):
    """BASELINE shape 64 x 64000: z-normed outputs have mean 0 / var 1, and the realised SNR of the mix
    (recovered from the normalised views) equals the requested one."""
    B, L = 64, 64000
    clean, noise, snr_idx, table = synthetic.waveforms(B, L, seed=1234)
    c, n, st = _run(dev, clean, noise, snr_idx, table)
    assert not st.any()
    assert np.abs(c.mean(1)).max() < 1e-5 and np.abs(n.mean(1)).max() < 1e-5
    assert np.abs(c.astype(np.float64).var(1) - 1).max() < 1e-4
    assert np.abs(n.astype(np.float64).var(1) - 1).max() < 1e-4
    # noisy_z = a*(clean + s*noise) + b  =>  regress out clean, the residual is the scaled noise
    for b in range(0, B, 9):
        cz, nz = clean[b].astype(np.float64), n[b].astype(np.float64)
        A = np.stack([cz, noise[b].astype(np.float64), np.ones(L)], 1)
        coef, *_ = np.linalg.lstsq(A, nz, rcond=None)
        snr = 10 * np.log10((coef[0] ** 2 * (cz ** 2).mean()) / (coef[1] ** 2 * (noise[b].astype(np.float64) ** 2).mean()))
        assert abs(snr - table[snr_idx[b]]) < 1e-3


@pytest.mark.parametrize("cs", [1, 2, 3, 4, 8])
def test_forced_cluster_sizes(dev, mix_variant, cs):
    """The resident kernel with every CTAs-per-row setting (incl. a non-power-of-two cluster): same result, so the
    automatic choice is a pure performance knob."""
    if mix_variant not in (4, 5):
        pytest.skip("the cluster knob applies to the resident kernel")
    clean, noise, snr_idx, table = synthetic.waveforms(3, 64000, seed=17)
    c_ref, n_ref, st_ref = oracle.mix_normalize_batch(clean, noise, snr_idx, table)
    ops.set_mix_cluster(cs)
    try:
        c, n, st = _run(dev, clean, noise, snr_idx, table)
    finally:
        ops.set_mix_cluster(0)
    assert st.tolist() == st_ref.tolist()
    for b in range(3):
        assert rel_err(n[b], n_ref[b].numpy()) < TOL and rel_err(c[b], c_ref[b].numpy()) < TOL


def test_device_side_retry_touches_only_rejected_rows(dev, mix_variant):
    """nrse_mix_normalize_retry_f32: rows with status 0 keep their outputs bit for bit; rejected rows are redone with
    the noise row AND the SNR draw of row (b + shift) % B -- the reference's next attempt re-draws both
    (ref:src/data/noisy_speech_dataset.py:69-81) -- and equal a fresh mix of (clean[b], noise[donor], snr[donor])."""
    B, L = 6, 8000
    clean, noise, snr_idx, table = synthetic.waveforms(B, L, seed=21)
    snr_idx = np.asarray([0, 1, 2, 3, 4, 0], np.int32) % len(table)
    noise[1] = 0.0          # status 4
    noise[4] = np.nan       # status 2
    cd, nd = torch.from_numpy(clean).to(dev), torch.from_numpy(noise).to(dev)
    sd = torch.from_numpy(snr_idx).to(dev)
    tab = [float(v) for v in table]
    c, n, st = ops.mix_normalize(cd, nd, sd, tab, True)
    assert st.tolist() == [0, 4, 0, 0, 2, 0]
    c0, n0 = c.clone(), n.clone()
    used = sd.clone()
    ops.mix_normalize_retry_(cd, nd, sd, tab, c, n, st, 2, True, used)   # row 1 <- (noise, snr)[3], row 4 <- [0]
    assert st.tolist() == [0] * B
    assert torch.equal(sd.cpu(), torch.from_numpy(snr_idx))              # the draw itself is never modified
    want_used = snr_idx.copy()
    want_used[1], want_used[4] = snr_idx[3], snr_idx[0]
    assert used.cpu().tolist() == want_used.tolist()
    for b in (0, 2, 3, 5):
        assert torch.equal(c[b], c0[b]) and torch.equal(n[b], n0[b])
    for b, donor in ((1, 3), (4, 0)):
        cr, nr, sr = oracle.mix_normalize_batch(clean[b:b + 1], noise[donor:donor + 1], snr_idx[donor:donor + 1], table)
        assert sr.tolist() == [0]
        assert rel_err(n[b].cpu().numpy(), nr[0].numpy()) < TOL and rel_err(c[b].cpu().numpy(), cr[0].numpy()) < TOL
    # a second retry on a healthy batch is a no-op
    c1, n1 = c.clone(), n.clone()
    ops.mix_normalize_retry_(cd, nd, sd, tab, c, n, st, 1, True, used)
    assert torch.equal(c, c1) and torch.equal(n, n1) and st.tolist() == [0] * B and used.cpu().tolist() == want_used.tolist()


def test_substitute_rows_replaces_permanently_bad_rows(dev, mix_variant):
    """A silent CLEAN crop can never recover by re-drawing noise: after the retries ``mix_substitute_rows_`` copies the
    nearest following good row over it (outputs and SNR index); good rows are untouched, the status keeps the reason."""
    B, L = 5, 4000
    clean, noise, snr_idx, table = synthetic.waveforms(B, L, seed=33)
    clean[1] = 0.0          # status 3 whatever the noise
    clean[4] = 0.0          # last row: wraps around to row 0
    cd, nd = torch.from_numpy(clean).to(dev), torch.from_numpy(noise).to(dev)
    sd = torch.from_numpy(snr_idx).to(dev)
    tab = [float(v) for v in table]
    c, n, st = ops.mix_normalize(cd, nd, sd, tab, True)
    used = sd.clone()
    for attempt in range(1, 5):
        ops.mix_normalize_retry_(cd, nd, sd, tab, c, n, st, attempt, True, used)
    assert st.tolist() == [0, 3, 0, 0, 3]
    assert not c[1].any() and not n[4].any()
    c0, n0 = c.clone(), n.clone()
    ops.mix_substitute_rows_(c, n, st, used)
    assert st.tolist() == [0, 3, 0, 0, 3]
    for b in (0, 2, 3):
        assert torch.equal(c[b], c0[b]) and torch.equal(n[b], n0[b])
    assert torch.equal(c[1], c0[2]) and torch.equal(n[1], n0[2]) and int(used[1]) == int(snr_idx[2])
    assert torch.equal(c[4], c0[0]) and torch.equal(n[4], n0[0]) and int(used[4]) == int(snr_idx[0])


@pytest.mark.parametrize("L", [4000, 64000, 6001])
@pytest.mark.parametrize("substitute", [True, False])
def test_mix_batch_equals_launch_per_attempt(dev, mix_variant, L, substitute):
    """nrse_mix_batch_f32 (three launches: mix, ONE retry launch that loops over the attempts inside the kernel, finish)
    against the launch-per-attempt sequence it replaces (mix, 4 x retry with shifts 1..4, substitute, torch gather of the
    labels, torch count): outputs, status, SNR indices, labels and the rejected-row count bit for bit.  The batch has rows
    that recover at the first retry, a row that needs THREE retries (its next two donors are bad as well), rows that never
    recover (silent clean crop) and, at L = 6001, takes the unaligned generic kernel."""
    B = 9
    clean, noise, snr_idx, table = synthetic.waveforms(B, L, seed=41)
    snr_idx = (np.arange(B, dtype=np.int32) * 2) % len(table)
    noise[2] = 0.0          # row 2: donors 3 (bad), 4 (bad), 5 (good): third retry
    noise[3] = np.nan       # row 3: donors 4 (bad), 5 (good)
    noise[4] = 0.0          # row 4: donor 5 (good)
    clean[6] = 0.0          # never recovers -> substituted by row 7
    clean[8] = 0.0          # last row: wraps around to row 0
    cd, nd = torch.from_numpy(clean).to(dev), torch.from_numpy(noise).to(dev)
    sd = torch.from_numpy(snr_idx).to(dev)
    tab = [float(v) for v in table]
    labels_tab = torch.tensor([int(round(v)) for v in tab], device=dev, dtype=torch.int64)
    # reference sequence
    c, n, st = ops.mix_normalize(cd, nd, sd, tab, True)
    assert st.tolist() == [0, 0, 4, 2, 4, 0, 3, 0, 3]
    used = sd.clone()
    for attempt in range(1, 5):
        ops.mix_normalize_retry_(cd, nd, sd, tab, c, n, st, attempt, True, used)
    assert st.tolist() == [0, 0, 0, 0, 0, 0, 3, 0, 2]   # row 8's last donor is row 3 (NaN noise): rejected as noise_nan
    assert used.cpu().tolist()[2:5] == [int(snr_idx[5])] * 3
    if substitute:
        ops.mix_substitute_rows_(c, n, st, used)
    want_labels = labels_tab[used.long()]
    # one call
    c2, n2, st2, used2, labels2, cnt2 = ops.mix_batch(cd, nd, sd, tab, labels_tab, True, 5, substitute)
    assert st2.tolist() == st.tolist() and used2.tolist() == used.tolist()
    assert torch.equal(c2, c) and torch.equal(n2, n)
    assert torch.equal(labels2, want_labels) and labels2.dtype == torch.int64
    assert cnt2.tolist() == [2]
    if substitute:
        assert torch.equal(c2[6], c2[7]) and torch.equal(n2[8], n2[0]) and c2[6].any()
    else:
        assert not c2[6].any() and not n2[8].any()
    # fewer attempts: row 2 (needs three retries) stays rejected with max_attempts = 3
    _, _, st3, _, _, cnt3 = ops.mix_batch(cd, nd, sd, tab, labels_tab, True, 3, substitute)
    assert st3.tolist() == [0, 0, 4, 0, 0, 0, 3, 0, 3] and cnt3.tolist() == [3]
    # a healthy batch: same as the plain mix, nothing rejected
    clean_h, noise_h, idx_h, _ = synthetic.waveforms(4, L, seed=42)
    ch, nh, ih = (torch.from_numpy(a).to(dev) for a in (clean_h, noise_h, idx_h))
    c4, n4, st4, used4, labels4, cnt4 = ops.mix_batch(ch, nh, ih, tab, labels_tab, True, 5, substitute)
    c5, n5, st5 = ops.mix_normalize(ch, nh, ih, tab, True)
    assert torch.equal(c4, c5) and torch.equal(n4, n5) and st4.tolist() == [0] * 4 == st5.tolist()
    assert used4.tolist() == idx_h.tolist() and cnt4.tolist() == [0] and torch.equal(labels4, labels_tab[ih.long()])


@pytest.mark.parametrize("max_attempts,substitute", [(5, True), (5, False), (3, True), (1, True)])
def test_mix_batch_vs_oracle_attempt_loop(dev, mix_variant, max_attempts, substitute):
    """nrse_mix_batch_f32 against the oracle's restatement of the reference's attempt loop
    (ref:src/data/noisy_speech_dataset.py:55-149 with explicit donors, ``oracle.mix_batch_attempts``): waveforms within 1e-6
    per row, statuses, SNR indices, labels and the rejected-row count exactly."""
    B, L = 9, 16000
    clean, noise, snr_idx, table = synthetic.waveforms(B, L, seed=43)
    snr_idx = (np.arange(B, dtype=np.int32) * 3 + 1) % len(table)
    noise[2] = 0.0
    noise[3] = np.nan
    noise[4] = 0.0
    clean[6] = 0.0
    clean[8, ::2] = np.inf      # speech power inf: scale invalid whatever the noise
    c_ref, n_ref, st_ref, used_ref, rejected = oracle.mix_batch_attempts(clean, noise, snr_idx, table, max_attempts, substitute)
    cd, nd, sd = (torch.from_numpy(a).to(dev) for a in (clean, noise, snr_idx))
    tab = [float(v) for v in table]
    labels_tab = torch.tensor([int(round(v)) for v in tab], device=dev, dtype=torch.int64)
    c, n, st, used, labels, cnt = ops.mix_batch(cd, nd, sd, tab, labels_tab, True, max_attempts, substitute)
    assert st.tolist() == st_ref.tolist()
    assert used.tolist() == used_ref.tolist() and labels.tolist() == [int(round(tab[i])) for i in used_ref]
    assert cnt.tolist() == [rejected]
    for b in range(B):
        if not n_ref[b].any():
            assert not n[b].any() and not c[b].any(), b
        else:
            assert rel_err(n[b].cpu().numpy(), n_ref[b].numpy()) < TOL and rel_err(c[b].cpu().numpy(), c_ref[b].numpy()) < TOL, b


def test_mix_batch_emotion_mode_never_retries(dev, mix_variant):
    """peak_norm = False (ref:src/data/emotion_dataset.py:177-203): a failed mix keeps the clean waveform and is NOT retried
    with another noise crop; mix_batch equals the plain launch whatever max_attempts says, the count reports the rows."""
    B, L = 5, 8000
    clean, noise, snr_idx, table = synthetic.waveforms(B, L, seed=51)
    noise[1] = 0.0
    noise[3] = np.nan
    cd, nd, sd = (torch.from_numpy(a).to(dev) for a in (clean, noise, snr_idx))
    tab = [float(v) for v in table]
    labels_tab = torch.tensor([int(round(v)) for v in tab], device=dev, dtype=torch.int64)
    _, n_ref, st_ref = ops.mix_normalize(cd, nd, sd, tab, False)
    c, n, st, used, labels, cnt = ops.mix_batch(cd, nd, sd, tab, labels_tab, False, 5, True)
    assert c is None and torch.equal(n, n_ref) and st.tolist() == st_ref.tolist() == [0, 4, 0, 2, 0]
    assert used.tolist() == snr_idx.tolist() and cnt.tolist() == [2] and torch.equal(labels, labels_tab[sd.long()])


def test_full_size_vs_oracle(dev):
    """BASELINE shape 64 x 64000 directly against the CPU oracle, 1e-6 per row (north star: mixing fp32 within 1e-6)."""
    B, L = 64, 64000
    clean, noise, snr_idx, table = synthetic.waveforms(B, L, seed=1234)
    c_ref, n_ref, st_ref = oracle.mix_normalize_batch(clean, noise, snr_idx, table)
    c, n, st = _run(dev, clean, noise, snr_idx, table)
    assert st.tolist() == st_ref.tolist() == [0] * B
    for b in range(B):
        assert rel_err(n[b], n_ref[b].numpy()) < TOL, b
        assert rel_err(c[b], c_ref[b].numpy()) < TOL, b
