# Round-end measurement set (one GPU): GPU tests, bench (both arms), ncu launch list and per-layer metrics.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 600 gpurun_out/bench_final.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_final.json 2>> gpurun_out/bench_final.err; tail -c 400 gpurun_out/bench_ref_final.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_v10.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
IVOSW_GRAPHS=0 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"conv_tc|stem_tc" -s 106 -c 53 -o gpurun_out/prof_conv_v10 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_v10.log 2>&1
ncu -i gpurun_out/prof_conv_v10.ncu-rep --page raw --csv > gpurun_out/prof_conv_v10_raw.csv 2>/dev/null
rm -f gpurun_out/prof_conv_v10.ncu-rep
ls -la gpurun_out
