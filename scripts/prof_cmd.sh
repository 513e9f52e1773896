#!/bin/bash
# usage: prof_cmd.sh <tag> : plain run then ncu --set full of the same command (one launch)
TAG=$1
CMD="python scripts/prof_run.py --final-t 5 --n 4096 --reps 2"
$CMD > gpurun_out/plain_$TAG.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:vsr_episode -s 1 -c 1 -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_$TAG.log 2>&1
cat gpurun_out/plain_$TAG.log; tail -2 gpurun_out/ncu_$TAG.log
