#!/bin/bash
# One N-GPU session: the C5 seed sweep sharded over N ranks, bench.py under torchrun, and (N=2 only) the hardware tests.
set -u
N=${1:-8}
TAG=${2:-r02}
mkdir -p gpurun_out
nvidia-smi -L | wc -l
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 scripts/c5_sweep.py \
  > gpurun_out/c5_${TAG}_n${N}.jsonl 2> gpurun_out/c5_${TAG}_n${N}.err
echo "c5 rc=$?"; cut -c1-260 gpurun_out/c5_${TAG}_n${N}.jsonl
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 3 --warmup 3 \
  > gpurun_out/bench_${TAG}_n${N}.json 2> gpurun_out/bench_${TAG}_n${N}.err
echo "bench rc=$?"
python - <<PY
import json
l=json.loads(open("gpurun_out/bench_${TAG}_n${N}.json").read().strip().splitlines()[-1])
print({k:l[k] for k in ("value","ms_per_step","n_gpus","population")}, l["e2e"]["value"], l["e2e"]["comm"], json.dumps(l.get("strong_scaling")))
PY
