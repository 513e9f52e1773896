#!/bin/bash
python -m pytest tests/test_parity_gpu.py -x -q -m gpu 2>&1 | tail -3
python scripts/prof_run.py --final-t 5 --n 65536 --reps 2 | tail -1
python scripts/prof_run.py --final-t 5 --n 4096 --reps 2 | tail -1
python scripts/prof_run.py --final-t 20 --n 4096 --reps 2 | tail -1
