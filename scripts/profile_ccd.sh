#!/bin/bash
# ncu --set full of the episode kernel with continuous detection on (config 2, 2 s), summarised ON THE BOX (the report
# of the 48-instantiation library is above gpurun's 64 MiB): gpurun_out/ccd_kernel_<tag>.txt
TAG=$1
python scripts/prof_ccd.py > gpurun_out/ccd_plain_$TAG.log 2>&1 || { echo "plain failed"; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:vsr_episode -s 1 -c 1 -o /tmp/ccd_$TAG python scripts/prof_ccd.py > gpurun_out/ccd_ncu_$TAG.log 2>&1
echo "capture rc=$?"
{ echo "# python scripts/prof_ccd.py (4096 bipeds, 120 ticks, continuous_detection = 1); ncu --set full --clock-control none --import-source on -k regex:vsr_episode -s 1 -c 1"
  python scripts/ncu_summary.py /tmp/ccd_$TAG.ncu-rep
  python scripts/ncu_lines.py /tmp/ccd_$TAG.ncu-rep 164400
  echo "--- SASS segments between block barriers / returns"
  python scripts/ncu_stalls.py /tmp/ccd_$TAG.ncu-rep; } > gpurun_out/ccd_kernel_$TAG.txt 2>&1
