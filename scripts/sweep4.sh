for sy in 7 15; do
  echo "== VSR_SYNC=$sy"
  VSR_SYNC=$sy python scripts/prof_run.py --final-t 5 --n 65536 --reps 2 | tail -1
  VSR_SYNC=$sy python scripts/prof_run.py --final-t 5 --n 4096 --reps 2 | tail -1
done
