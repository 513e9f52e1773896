#!/bin/bash
# usage: profile_round.sh <tag>
# 1) bench.py plain (its JSON line), 2) ncu launch list of the same command, 3) one ncu --set full
# capture of the episode kernel of the same command.  Each profiler pass runs only after the plain
# command exited 0.  Outputs under gpurun_out/.
TAG=$1
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err || { echo "bench failed"; tail -5 gpurun_out/bench_$TAG.err; exit 1; }
tail -1 gpurun_out/bench_$TAG.json
$CMD > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_l_$TAG.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:vsr_episode -s 3 -c 1 -o gpurun_out/full_$TAG $CMD > gpurun_out/ncu_f_$TAG.log 2>&1
echo "full capture rc=$?"
