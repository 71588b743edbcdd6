#!/bin/bash
# Measurement helper (B200 box): ncu counters of conv_stack_kernel for a list of schedules "G2,G3".
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,sm__cycles_elapsed.max,lts__t_bytes.sum
for item in "$@"; do
  IFS=, read g2 g3 <<< "$item"
  export IVOSW_STACK_G2=$g2 IVOSW_STACK_G3=$g3
  ncu --metrics $M --clock-control none -k regex:conv_stack -s 4 -c 2 --csv --log-file gpurun_out/stack_ncu_${g2}_${g3}.csv \
      python bench.py --steps 2 --warmup 3 --no-ref-gpu --no-cpu-baseline --no-parity > /dev/null 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/stack_ncu_${g2}_${g3}.csv')) if len(r)>10]
hdr=rows[0]; i_n=hdr.index('Metric Name'); i_v=hdr.index('Metric Value'); i_id=hdr.index('ID')
d={}
for r in rows[1:]:
    d.setdefault(r[i_id],{})[r[i_n]]=r[i_v]
for k,v in d.items():
    print('G=$item', k, {a:b for a,b in v.items()})
PY
done
