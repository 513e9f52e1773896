#!/bin/bash
python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "fp64_kernel_matches or effective_mass or mlp" 2>&1 | tail -3
for sy in 3 1 2 0; do
  echo "== VSR_SYNC=$sy"
  VSR_SYNC=$sy python scripts/prof_run.py --final-t 5 --n 65536 --reps 2 | tail -1
  VSR_SYNC=$sy python scripts/prof_run.py --final-t 5 --n 4096 --reps 2 | tail -1
done
