# Measurement helper (B200 box): per-layer ncu table of one round (48 conv launches + stem) -> gpurun_out/<tag>_ncu_conv_stack.md
tag=${1:-r2}
IVOSW_GRAPHS=0 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none ${NCU_EXTRA} -k regex:"conv_tc|stem_tc" -s 98 -c 49 -o gpurun_out/${tag}_prof_conv python scripts/one_pass.py 64 > gpurun_out/${tag}_ncu.log 2>&1
ncu -i gpurun_out/${tag}_prof_conv.ncu-rep --page raw --csv > gpurun_out/${tag}_prof_conv_raw.csv 2>/dev/null
rm -f gpurun_out/${tag}_prof_conv.ncu-rep
python profiles/summarise_ncu_raw.py gpurun_out/${tag}_prof_conv_raw.csv 128 > gpurun_out/${tag}_ncu_conv_stack.md 2>&1
grep -E "conv3 |conv3\+ds|conv stack" gpurun_out/${tag}_ncu_conv_stack.md | cut -c1-110
