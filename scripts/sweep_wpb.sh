#!/bin/bash
# warps-per-block sweep on a saturating and on the config-2 batch
for w in 1 2 4 6 12; do
  echo "== VSR_WPB=$w"; VSR_WPB=$w python scripts/prof_run.py --final-t 2 --n 65536 --reps 2 | tail -1
  VSR_WPB=$w python scripts/prof_run.py --final-t 2 --n 4096 --reps 2 | tail -1
done
