#!/bin/bash
# N-GPU check: GPU test suite (incl. the 2-rank NCCL test), then the bench under torchrun at N ranks (strong scaling)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
N=${1:-2}
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
  bench.py --gpus $N --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
echo "bench rc=$?"; tail -3 gpurun_out/bench_${N}gpu.err
python - $N <<'PY'
import json,sys
d=json.load(open(f'gpurun_out/bench_{sys.argv[1]}gpu.json'))
print("N=%d value %.3e ms/step %.2f e2e %.3e gather_verified %s weak %s"%(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d.get('gather_verified'), d.get('weak')))
print("front_end_ms", d['roofline']['front_end_ms'], "kernel_ms", d['roofline']['kernel_ms'])
if 'secondary' in d: print("secondary: %.3e node-evals/s  %.2f ms"%(d['secondary']['value'], d['secondary']['ms_per_step']))
PY
