# e2e variants of the host-buffer path
run() { timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), d['e2e']['h2d_bytes_per_step'])"; }
run default
IVOSW_GRAPHS=0 run nographs
for sch in "8,16,16,16,8" "4,12,16,16,16" "8,12,16,16,12" "8,24,24,8" "4,8,12,16,16,8"; do IVOSW_E2E_SCHEDULE=$sch run "sched $sch"; done
