#!/bin/bash
# full GPU suite as the driver runs it, then smoke and the bench
mkdir -p gpurun_out
run() { name=$1; shift; timeout "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$?"; tail -n 8 gpurun_out/$name.log | cut -c1-1500; }
run t_all 1200 python -m pytest tests/ -x -q -m gpu
run smoke 300 python __graft_entry__.py smoke
run bench 600 python bench.py
