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
# the report of the 48-instantiation library is above gpurun's 64 MiB merge limit: summarise it here, on the box, into
# gpurun_out/profiles_<tag>/ (copy those files into profiles/ afterwards) and drop the report
python scripts/ncu_to_profiles.py $TAG r2 > gpurun_out/ncu_to_profiles_$TAG.log 2>&1
{ echo "# $CMD ; ncu --set full --clock-control none --import-source on -k regex:vsr_episode -s 3 -c 1 (the headline kernel: MODE 3)"
  python scripts/ncu_lines.py gpurun_out/full_$TAG.ncu-rep
  echo "--- SASS segments between block barriers / returns"
  python scripts/ncu_stalls.py gpurun_out/full_$TAG.ncu-rep; } > profiles/r2_ncu_headline_kernel_lines.txt 2>&1
mkdir -p gpurun_out/profiles_$TAG
cp profiles/r2_* profiles/traffic.json gpurun_out/profiles_$TAG/
rm -f gpurun_out/full_$TAG.ncu-rep
tail -3 gpurun_out/ncu_to_profiles_$TAG.log
