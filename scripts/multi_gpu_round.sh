#!/bin/bash
# One multi-GPU session (gpurun --gpus N): hardware tests of every N > 1 code path, the CLI over N devices, and
# bench.py under torchrun.  usage: multi_gpu_round.sh N TAG
set -u
N=${1:-2}
TAG=${2:-r02}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/multi_${TAG}_n${N}.log
python -m pytest tests/test_multi_gpu.py tests/test_gpu_parity.py -m gpu -q -k "multi or device_counts or nccl" 2>&1 | tail -5 >> gpurun_out/multi_${TAG}_n${N}.log
# the C++ CLI over N devices: a short GA, same --ga-seed on 1 and N devices must print the same fitness trajectory
mkdir -p /tmp/cli1 /tmp/cliN
for D in 1 $N; do
  ( cd /tmp && /root/repo/evosops_b200/host/evosops -b agg -e ga -k "(6,4)" -k "(10,7)" -s 1 -s 2 --gran 10 -p 24 -g 4 --ga-seed 5 --rng verify --devices $D --out /tmp/cli$D -v ) > /tmp/cli_$D.txt 2>&1
  echo "evosops --devices $D rc=$?" >> gpurun_out/multi_${TAG}_n${N}.log
done
grep -h "Avg. Fitness" /tmp/cli1/*.log > /tmp/a1.txt; grep -h "Avg. Fitness" /tmp/cli$N/*.log > /tmp/aN.txt
if cmp -s /tmp/a1.txt /tmp/aN.txt; then echo "CLI: fitness trajectory identical on 1 and $N devices ($(wc -l < /tmp/a1.txt) generations)"; else echo "CLI: TRAJECTORIES DIFFER"; fi >> gpurun_out/multi_${TAG}_n${N}.log
cat /tmp/aN.txt >> gpurun_out/multi_${TAG}_n${N}.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 3 --warmup 3 \
  > gpurun_out/bench_${TAG}_n${N}.json 2> gpurun_out/bench_${TAG}_n${N}.err
echo "bench rc=$?" >> gpurun_out/multi_${TAG}_n${N}.log
cat gpurun_out/multi_${TAG}_n${N}.log
python - <<PY
import json
l=json.loads(open("gpurun_out/bench_${TAG}_n${N}.json").read().strip().splitlines()[-1])
print({k:l[k] for k in ("value","ms_per_step","n_gpus","population")}, l["e2e"], l.get("strong_scaling"))
PY
