#!/bin/bash
# robots per block of the big-robot kernel (needs the -DVSR_TUNING build: build/libvsr_b200_tuning.so)
cp build/libvsr_b200_tuning.so 2dhmsr_b200/libvsr_b200.so
for r in 1 2 4; do
  echo "== VSR_RPB=$r"; VSR_RPB=$r timeout 300 python scripts/cfg4_timing.py 32768 5 | tail -1
done
for shp in box-8x8 box-10x10 box-6x6 worm-10x7; do
  for r in 1 2 4; do echo "== $shp VSR_RPB=$r"; VSR_RPB=$r timeout 300 python scripts/prof_run.py --shape $shp --n 4096 --final-t 5 --reps 2 --terrain hilly-1-10-0 --sensors uniform-t+vxy+a-0 | tail -1; done
done
