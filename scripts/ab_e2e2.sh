run() { timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-ref-gpu --no-parity 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), d['e2e']['h2d_bytes_per_step'])"; }
run default
for sch in "8,16,20,20" "8,16,18,22" "6,14,20,24" "10,18,18,18" "8,14,14,14,14" "6,12,14,16,16" "8,16,16,24"; do IVOSW_E2E_SCHEDULE=$sch run "sched $sch"; done
