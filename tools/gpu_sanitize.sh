#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_golden.py -m gpu -q -x > gpurun_out/memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/memcheck.log | tail -3
