#!/bin/bash
# (two launches per evaluation: fast / worst-case class; -s 10 lands on the fast-class launch of the 2nd C3 evaluation)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
python bench.py --only-notebooks > gpurun_out/nb.json 2> gpurun_out/nb.err && cat gpurun_out/nb.json | cut -c1-600 &&
ncu --set full --clock-control none --import-source on -k regex:k_rollout_rk -s 10 -c 1 -f -o gpurun_out/prof_rk \
    python bench.py --only-notebooks > gpurun_out/nb_ncu.log 2>&1
echo "rc=$?"
