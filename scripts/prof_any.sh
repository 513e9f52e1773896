#!/bin/bash
# usage: prof_any.sh <tag> <prof_run.py args...> : plain run, then ncu --set full of the same command (second launch)
TAG=$1; shift
CMD="python scripts/prof_run.py --reps 2 $*"
$CMD > gpurun_out/plain_$TAG.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:vsr_episode -s 1 -c 1 -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_$TAG.log 2>&1
cat gpurun_out/plain_$TAG.log; tail -2 gpurun_out/ncu_$TAG.log
