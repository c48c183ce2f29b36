#!/bin/bash
# One GPU session: bench line, ncu launch list, ncu full capture of the chain kernel (bench + sweep probe).
set -u
mkdir -p gpurun_out
TAG=${1:-r01}
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; cat gpurun_out/bench_$TAG.json
CMD="python bench.py --steps 2 --warmup 3 --no-extra"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:sops_chain -s 3 -c 1 -o gpurun_out/prof_c2_$TAG -f $CMD > gpurun_out/ncu_full_c2.log 2>&1
echo "full c2 rc=$?"
CMD2="python scripts/perf_probe.py c5s fast"
$CMD2 > gpurun_out/plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:sops_chain -c 1 -o gpurun_out/prof_sweep_$TAG -f $CMD2 > gpurun_out/ncu_full_sweep.log 2>&1
echo "full sweep rc=$?"
cat gpurun_out/plain3.log
ls -la gpurun_out
