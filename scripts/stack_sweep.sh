#!/bin/bash
# Measurement helper (B200 box): conv-stack stage time of bench.py for a list of schedules of conv_stack.cu.
# usage: scripts/stack_sweep.sh "G2,G3[,G4,G5]" ...    (STACK=0 as an item = one launch per layer)
out=gpurun_out/stack_sweep.txt
: > $out
for item in "$@"; do
  unset IVOSW_STACK IVOSW_STACK_G2 IVOSW_STACK_G3 IVOSW_STACK_G4 IVOSW_STACK_G5
  if [ "$item" = "STACK=0" ]; then export IVOSW_STACK=0; else
    IFS=, read g2 g3 g4 g5 <<< "$item"
    export IVOSW_STACK_G2=$g2 IVOSW_STACK_G3=$g3
    [ -n "$g4" ] && export IVOSW_STACK_G4=$g4
    [ -n "$g5" ] && export IVOSW_STACK_G5=$g5
  fi
  python bench.py --steps 10 --warmup 3 --no-ref-gpu --no-cpu-baseline --no-parity 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$item', 'round %.3f ms' % d['ms_per_step'], 'conv_stack %.3f ms' % d['roofline']['stage_ms_per_step']['conv_stack'], 'e2e %.2f' % d['e2e']['ms_per_step'], 'sm %s' % d['clocks']['sm_mhz'])" | tee -a $out
done
