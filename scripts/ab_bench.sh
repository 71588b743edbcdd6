# A/B of two builds of the library on one box: IVOSW_LIB selects the build.  Put the build to compare against at
# ivos-w_b200/lib/libivosw_b200_old.so (e.g. `git stash; make -C ivos-w_b200/csrc; cp ...; git stash pop`).
[ -f ivos-w_b200/lib/libivosw_b200_old.so ] || { echo 'no libivosw_b200_old.so to compare against'; exit 1; }
run() {
  timeout 200 python bench.py --steps ${STEPS:-30} --warmup 4 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), {k: round(v,3) for k,v in d['roofline']['stage_ms_per_step'].items()}, d['clocks']['sm_mhz'], d['clocks'].get('power_w_max'))"
}
for cfg in "IVOSW_GRAPHS=1 IVOSW_PDL=1" "IVOSW_GRAPHS=0 IVOSW_PDL=1" "IVOSW_GRAPHS=0 IVOSW_PDL=0" "IVOSW_GRAPHS=1 IVOSW_PDL=0"; do
  for L in old new; do
    if [ $L = old ]; then export IVOSW_LIB=$PWD/ivos-w_b200/lib/libivosw_b200_old.so; else unset IVOSW_LIB; fi
    env $cfg bash -c "$(declare -f run); run '$cfg $L'"
  done
done
