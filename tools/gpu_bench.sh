#!/bin/bash
# bench + ncu launch list (+ optional full capture of the rollout kernel)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench rc=$?"; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
python bench.py --steps 2 --warmup 3 --no-secondary --no-cpu --population 8192 > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-secondary --no-cpu --population 8192 > gpurun_out/ncu_launch.log 2>&1
echo "ncu launches rc=$?"
if [ "$1" == "full" ]; then
python bench.py --steps 1 --warmup 3 --no-secondary --no-cpu --population 4096 > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_rollout -s 3 -c 1 -f -o gpurun_out/prof_rollout \
    python bench.py --steps 1 --warmup 3 --no-secondary --no-cpu --population 4096 > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
fi
