#!/bin/bash
# first-contact run on the GPU box: parity tests, verbose
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -150
