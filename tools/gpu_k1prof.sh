#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
python bench.py --only-secondary --steps 2 > gpurun_out/k1.json 2> gpurun_out/k1.err && cut -c1-200 gpurun_out/k1.json &&
ncu --set full --clock-control none --import-source on -k regex:k_pointwise -s 6 -c 1 -f -o gpurun_out/prof_pointwise \
    python bench.py --only-secondary --steps 2 > gpurun_out/k1_ncu.log 2>&1
echo "rc=$?"
