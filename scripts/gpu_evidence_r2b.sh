#!/bin/bash
# round 2, final state: full ncu capture of the elementwise kernels of one backward pass that VERDICT r1 named (LayerNorm + GELU
# backward of layers 1 and 0, tensor-core layer-0 weight gradient), summarised ON THE BOX (a 26-launch .ncu-rep with sources
# exceeds what gpurun copies back)
mkdir -p gpurun_out
T=${TAG:-r2z}
PK="python scripts/profile_kernels.py"
# first backward pass: ln_gelu_bwd of layers 6..0 (7 launches) then layer0_wgrad_tc -> skip the five small ones
timeout 600 ncu --set full --clock-control none -k regex:"ln_gelu_bwd|layer0_wgrad" -s 5 -c 3 -f -o /tmp/${T}_prof_bwd $PK > gpurun_out/${T}_ncu4.log 2>&1; echo "backward capture rc=$?"
python scripts/ncu_summary.py /tmp/${T}_prof_bwd.ncu-rep > gpurun_out/${T}_ncu_full_backward.txt 2>&1
grep -E "^(void|nrse|unnamed)|duration|dram % of peak|dram read|dram write" gpurun_out/${T}_ncu_full_backward.txt | cut -c1-120
