#!/bin/bash
# Round-end measurement, one profiler invocation per gpurun call:
#   bash tools/gpu_final.sh bench      GPU tests, full bench.py line (+ reference arm), the ncu launch list of the same command
#   bash tools/gpu_final.sh rollout    ncu --set full capture of k_rollout at the bench size
#   bash tools/gpu_final.sh pointwise  ncu --set full capture of k_pointwise at the bench size
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
case "$1" in
bench)
  timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
  python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
  echo "bench rc=$?"; tail -2 gpurun_out/bench.err; cut -c1-400 gpurun_out/bench.json
  python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
  echo "reference rc=$?"; cut -c1-300 gpurun_out/bench_reference.json
  python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
  python bench.py --steps 2 --warmup 3 --no-cpu --no-secondary > gpurun_out/plain.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu --no-secondary > gpurun_out/ncu_launch.log 2>&1
  echo "ncu launches rc=$?" ;;
rollout)
  python bench.py --steps 1 --warmup 3 --no-secondary --no-cpu > gpurun_out/plain2.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:k_rollout -s 6 -c 1 -f -o gpurun_out/prof_rollout_full \
      python bench.py --steps 1 --warmup 3 --no-secondary --no-cpu > gpurun_out/ncu_full.log 2>&1
  echo "ncu rollout rc=$?" ;;
pointwise)
  python bench.py --only-secondary --steps 2 > gpurun_out/plain3.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:k_pointwise -s 6 -c 1 -f -o gpurun_out/prof_pointwise_full \
      python bench.py --only-secondary --steps 2 > gpurun_out/ncu_full2.log 2>&1
  echo "ncu pointwise rc=$?" ;;
*) echo "usage: gpu_final.sh bench|rollout|pointwise"; exit 2 ;;
esac
