"""Which CUDA API call is slow in the second step after a device synchronisation?  torch.profiler (CUPTI) records every
runtime / driver API call with its host duration: prints the longest ones of 6-step windows started from an idle GPU."""
import os
import sys
import time

import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from nrse_b200.data import GpuBatchMixer  # noqa: E402
from nrse_b200.models import B200FeatureEncoder, wavlm_large_config  # noqa: E402
from nrse_b200.utils import synthetic  # noqa: E402

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
    out = step()
torch.cuda.synchronize()
# (1) plain host timing of whole steps incl. the release of the previous result, as bench.py's debug line measures it
for trial in range(6):
    torch.cuda.synchronize()
    host = []
    for i in range(6):
        h0 = time.perf_counter()
        out = step()
        host.append((time.perf_counter() - h0) * 1e3)
    torch.cuda.synchronize()
    print(f"plain trial {trial}: host ms per step " + " ".join(f"{v:6.2f}" for v in host))
# (2) the same under the profiler
for trial in range(3):
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        marks = []
        for i in range(6):
            marks.append(time.perf_counter())
            out = step()
        marks.append(time.perf_counter())
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type.name == "CPU" and (e.name.startswith("cuda") or e.name.startswith("cu"))]
    evs.sort(key=lambda e: -(e.time_range.end - e.time_range.start))
    t0 = min(e.time_range.start for e in prof.events())
    print(f"profiled trial {trial}: host ms per step " + " ".join(f"{(marks[i + 1] - marks[i]) * 1e3:6.2f}" for i in range(6)))
    for e in evs[:8]:
        print(f"    {e.name:40s} start {e.time_range.start - t0:9.1f} us  dur {e.time_range.end - e.time_range.start:9.1f} us")
