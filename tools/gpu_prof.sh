#!/bin/bash
# ncu full capture of the rollout kernel on a 4096-individual C5 slice
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
python bench.py --steps 1 --warmup 3 --no-secondary --no-cpu --population 8192 > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_rollout -s 6 -c 1 -f -o gpurun_out/prof_rollout \
    python bench.py --steps 1 --warmup 3 --no-secondary --no-cpu --population 8192 > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
