#!/bin/bash
# development: time config 2 with continuous detection under the build variants in build/dev_*.so
for v in "$@"; do
  cp build/dev_$v.so 2dhmsr_b200/libvsr_b200.so
  echo "== $v"; timeout 200 python scripts/ccd_cost.py 4096 2>&1 | tail -3
done
