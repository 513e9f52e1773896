#!/bin/bash
# wave/tail study: robots per launch vs throughput (148 SMs x 12 warps x 2 bipeds = 3552 per pass)
for n in 296 888 1776 3552 4096 7104 14208; do
  python scripts/prof_run.py --final-t 20 --n $n --reps 2 | tail -1
done
for w in 4 6 7 8 10; do echo "VSR_WPB=$w"; VSR_WPB=$w python scripts/prof_run.py --final-t 20 --n 4096 --reps 2 | tail -1; done
