#!/bin/bash
# mix_batch (one retry launch + finish launch): parity tests of the mix / data modules, then the headline step's timeline
mkdir -p gpurun_out
T=${TAG:-r2r}
timeout 600 python -m pytest tests/test_gpu_mix.py tests/test_gpu_model.py -q --timeout 300 -p no:cacheprovider -x > gpurun_out/${T}_tests_mix.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${T}_tests_mix.log
timeout 300 python scripts/step_timeline.py > gpurun_out/${T}_step_timeline.log 2>&1
tail -4 gpurun_out/${T}_tests_mix.log; tail -5 gpurun_out/${T}_step_timeline.log
