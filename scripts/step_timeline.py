"""Kernel timeline (CUPTI through torch.profiler) of the headline step as bench.py runs it: GpuBatchMixer ->
B200FeatureEncoder.forward on both views, eager module calls.  Prints every kernel of two consecutive warm steps with its
start, duration and the gap to the previous kernel, then the per-step span against the sum of kernel durations -- the
difference is what launch gaps / host stalls cost; then per-step device and host-enqueue times of 30 consecutive steps."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from nrse_b200.data import GpuBatchMixer  # noqa: E402
from nrse_b200.models import B200FeatureEncoder, wavlm_large_config  # noqa: E402
from nrse_b200.utils import synthetic  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

dev = torch.device("cuda", 0)
clean_np, noise_np, snr_idx_np, snr_table = synthetic.waveforms(bench.BATCH, bench.N_SAMPLES, seed=1234)
layers = synthetic.frontend_weights("layer", seed=0)
raw = {"clean_wave": torch.from_numpy(clean_np).to(dev), "noise_wave": torch.from_numpy(noise_np).to(dev),
       "snr_idx": torch.from_numpy(snr_idx_np).to(dev),
       "snr": torch.from_numpy(snr_table[snr_idx_np].astype(np.int64)).to(dev)}
mixer = GpuBatchMixer(snr_table.tolist(), dev)
encoder = bench.load_frontend(B200FeatureEncoder(wavlm_large_config()), layers, dev).eval()


@torch.no_grad()
def step():
    batch = mixer(raw)
    return encoder(batch["clean_input_values"]), encoder(batch["noisy_input_values"]), batch["mix_status"]


for _ in range(5):
    step()
torch.cuda.synchronize()
run = step
N = 4
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(N):
        run()
    torch.cuda.synchronize()
ev = sorted((e for e in prof.events() if e.device_type.name == "CUDA"), key=lambda e: e.time_range.start)
per = len(ev) // N
ev = ev[per * (N - 2):]  # the last two steps
t0 = ev[0].time_range.start
prev_end = t0
for e in ev:
    print(f"  {e.time_range.start - t0:8.1f} dur {e.time_range.end - e.time_range.start:7.1f} gap {e.time_range.start - prev_end:6.1f}  {e.name[:70]}")
    prev_end = max(prev_end, e.time_range.end)
span = ev[-1].time_range.end - t0
busy = sum(e.time_range.end - e.time_range.start for e in ev)
print(f"two steps: span_us {span:.1f} sum_of_kernel_us {busy:.1f} kernels_per_step {per} gap_share {1 - busy / span:.3f}")
# without the profiler: per-step device time (CUDA events between steps) and host enqueue time (perf_counter, no sync) of 30
# consecutive steps started from an empty queue -- shows whether the first steps of a short timed window differ from the rest
import time
for rep in range(2):
    torch.cuda.synchronize()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(31)]
    host = []
    evs[0].record()
    for i in range(30):
        h0 = time.perf_counter()
        step()
        host.append((time.perf_counter() - h0) * 1e3)
        evs[i + 1].record()
    torch.cuda.synchronize()
    dev_ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(30)]
    print(f"rep {rep}: device ms/step first 5 {[round(v, 3) for v in dev_ms[:5]]} mean(20) {sum(dev_ms[:20]) / 20:.4f} "
          f"mean(last 10) {sum(dev_ms[20:]) / 10:.4f}")
    print(f"rep {rep}: host enqueue ms/step first 5 {[round(v, 3) for v in host[:5]]} mean {sum(host) / 30:.4f} max {max(host):.3f}")
