"""Where does the host spend its time in the first steps after a device synchronisation?  Runs the headline step (as
bench.py does) in windows of 6 steps from an idle GPU and prints the host time of every sub-stage of every step."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from nrse_b200 import ops  # noqa: E402
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
mode = sys.argv[1] if len(sys.argv) > 1 else "keep"


@torch.no_grad()
def step(t):
    t.append(time.perf_counter())
    batch = mixer(raw)
    t.append(time.perf_counter())
    a = encoder(batch["clean_input_values"])
    t.append(time.perf_counter())
    b = encoder(batch["noisy_input_values"])
    t.append(time.perf_counter())
    return a, b, batch["mix_status"]


for _ in range(5):
    step([])
torch.cuda.synchronize()
for trial in range(8):
    torch.cuda.synchronize()
    if mode == "sleep":
        time.sleep(0.05)
    rows = []
    out = None
    for i in range(6):
        t = []
        if mode == "drop":
            step(t)
        else:
            out = step(t)
        rows.append([(t[j + 1] - t[j]) * 1e3 for j in range(3)])
    torch.cuda.synchronize()
    print(f"trial {trial} [{mode}] host ms (mixer, enc view 1, enc view 2) per step: " +
          " | ".join(" ".join(f"{v:6.2f}" for v in r) for r in rows))
print("memory", torch.cuda.memory_allocated() >> 20, "MiB allocated", torch.cuda.memory_reserved() >> 20, "MiB reserved",
      "cudaMalloc calls", torch.cuda.memory_stats()["num_device_alloc"])
