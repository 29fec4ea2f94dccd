#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "pointwise or comb" 2>&1 | tail -3
python bench.py --only-secondary --steps 3 > gpurun_out/k1.json 2> gpurun_out/k1.err; tail -2 gpurun_out/k1.err; cut -c1-260 gpurun_out/k1.json
