[ -f ivos-w_b200/lib/libivosw_b200_old.so ] || { echo 'no libivosw_b200_old.so to compare against'; exit 1; }
# per-kernel durations of one warm round, caches left warm (single-pass metric), for two builds of the library
for L in old new; do
  if [ $L = old ]; then export IVOSW_LIB=$PWD/ivos-w_b200/lib/libivosw_b200_old.so; else unset IVOSW_LIB; fi
  IVOSW_GRAPHS=0 ncu --metrics gpu__time_duration.sum --cache-control none --clock-control none -k regex:"conv_tc|stem_tc" -s 106 -c 53 --csv --log-file gpurun_out/launches_warm_$L.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
done
