# Round-2 measurement set (one GPU): bench (both arms), ncu launch list of the bench command, per-layer ncu metrics of the
# conv stack (duration, tensor pipe, DRAM bytes), one --set full capture of the dominant kernel variant.
set -x
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; tail -c 400 gpurun_out/r2_bench_final.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref_final.json 2>> gpurun_out/r2_bench_final.err; tail -c 300 gpurun_out/r2_bench_ref_final.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-ref-gpu --no-parity > /dev/null 2>&1
IVOSW_GRAPHS=0 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"conv_tc|stem_tc" -s 98 -c 49 -o gpurun_out/r2_prof_conv python scripts/one_pass.py 64 > gpurun_out/r2_ncu.log 2>&1
ncu -i gpurun_out/r2_prof_conv.ncu-rep --page raw --csv > gpurun_out/r2_prof_conv_raw.csv 2>/dev/null
rm -f gpurun_out/r2_prof_conv.ncu-rep
python profiles/summarise_ncu_raw.py gpurun_out/r2_prof_conv_raw.csv 128 > gpurun_out/r2_ncu_conv_stack.md 2>&1; tail -5 gpurun_out/r2_ncu_conv_stack.md
IVOSW_GRAPHS=0 ncu --set full --clock-control none --import-source on -k conv_tc_kernel -s 121 -c 1 -o gpurun_out/r2_full_res3 python scripts/one_pass.py 64 >> gpurun_out/r2_ncu.log 2>&1
ncu -i gpurun_out/r2_full_res3.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]
keep=['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed','sm__cycles_elapsed.max','launch__registers_per_thread','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','lts__t_sector_hit_rate.pct']
for k in keep:
    if k in h: print(k, '=', rows[2][h.index(k)], rows[1][h.index(k)])
" > gpurun_out/r2_full_res3.txt; cat gpurun_out/r2_full_res3.txt
ls -la gpurun_out | tail -12
