#!/bin/bash
# One GPU session of evidence: bench line, ncu launch list of the bench command, ncu --set full captures of the chain
# kernel on the C2 generation, a throughput-bound seed sweep and one generation each of separation / locomotion / coating.
# usage: gpu_round.sh TAG   (outputs under gpurun_out/; summarised into profiles/ by scripts/ncu_summary.py)
set -u
mkdir -p gpurun_out
TAG=${1:-r02}
if [ -z "${SKIP_BENCH:-}" ]; then python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/bench_$TAG.json; fi
CMD="python bench.py --steps 2 --warmup 3 --no-extra"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
full() {   # name, launches to skip, command...
  local name=$1 skip=$2; shift 2
  "$@" > gpurun_out/plain_$name.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:sops_chain -s $skip -c 1 -o gpurun_out/prof_${name}_$TAG -f "$@" > gpurun_out/ncu_full_$name.log 2>&1
  echo "full $name rc=$?"; tail -2 gpurun_out/plain_$name.log | cut -c1-300
  # only the raw metric table travels back (gpurun_out/ is capped at 64 MiB); KEEP_REP=name keeps that one report
  ncu -i gpurun_out/prof_${name}_$TAG.ncu-rep --page raw --csv > gpurun_out/raw_${name}_$TAG.csv 2>/dev/null
  # per-instruction stall samples of the two aggregation captures (scripts/sass_hot.py condenses them)
  case $name in c2|evo) ncu -i gpurun_out/prof_${name}_$TAG.ncu-rep --page source --csv --print-source sass > gpurun_out/src_${name}_$TAG.csv 2>/dev/null;; esac
  if [ "${KEEP_REP:-}" != "$name" ]; then rm -f gpurun_out/prof_${name}_$TAG.ncu-rep; fi
}
full c2 3 python bench.py --steps 2 --warmup 3 --no-extra
full evo 2 python scripts/perf_probe.py evo fast
full sweep 0 python scripts/perf_probe.py c5s fast
# (chains capped at 4e6 steps, a fifth of a (13,9) chain: the full generations are in the bench line)
full sep 1 python scripts/c3_probe.py sep cap=4000000
full loco 1 python scripts/c3_probe.py loco cap=4000000
full coat 1 python scripts/c3_probe.py coat cap=4000000
# a real GA run through the C++ CLI (config C2, 60 generations): wall time per generation as the population converges
if [ -z "${SKIP_GA:-}" ]; then
  make -C evosops_b200/host -s
  ( cd /tmp && time /root/repo/evosops_b200/host/evosops -b agg -e ga -p 50 -g 60 --gran 10 -m 0.08 -k "(6,4)" -k "(10,7)" -k "(13,9)" \
      -s 1 -s 2 -s 3 -s 4 -s 5 --ga-seed 1 --out /tmp/ga_out -v ) > gpurun_out/ga_c2_run_$TAG.log 2>&1
  echo "ga rc=$?"; grep -c "Wall Time" gpurun_out/ga_c2_run_$TAG.log; grep "Wall Time" gpurun_out/ga_c2_run_$TAG.log | awk 'NR%10==1'
fi
ls -la gpurun_out | tail -20
