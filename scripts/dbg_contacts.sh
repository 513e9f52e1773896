VSR_DEBUG=2:20:47 VSR_ORACLE_DEBUG=2:20:47 python scripts/f64_diag.py > gpurun_out/dbg.txt 2>&1
grep -A2 "^\[kernel\]" gpurun_out/dbg.txt | grep -v "^--" | sed 's/ pose.*//; s/\[kernel\]//' > /tmp/k.txt
grep -A2 "^\[oracle\]" gpurun_out/dbg.txt | grep -v "^--" | sed 's/\[oracle\]//' > /tmp/o.txt
wc -l /tmp/k.txt /tmp/o.txt
diff /tmp/k.txt /tmp/o.txt | head -40
