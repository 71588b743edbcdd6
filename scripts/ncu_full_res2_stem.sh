# Measurement helper (B200 box): ncu --set full of the res2 3x3 kernel (BN = 64 variant, res2.0.conv2 of the third round) and of the stem
mkdir -p gpurun_out
IVOSW_GRAPHS=0 ncu --set full --clock-control none -k regex:"conv_tc" -s 97 -c 1 -o gpurun_out/r2_full_res2conv2 python scripts/one_pass.py 64 > /dev/null 2>&1
IVOSW_GRAPHS=0 ncu --set full --clock-control none -k regex:"stem_tc" -s 2 -c 1 -o gpurun_out/r2_full_stem python scripts/one_pass.py 64 > /dev/null 2>&1
for f in r2_full_res2conv2 r2_full_stem; do
ncu -i gpurun_out/$f.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]
keep=['Kernel Name','gpu__time_duration.sum','sm__cycles_elapsed.max','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','launch__registers_per_thread','launch__grid_size','l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed','l1tex__data_bank_reads.avg.pct_of_peak_sustained_elapsed','l1tex__data_bank_writes.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__m_xbar2l1tex_read_bytes.sum','l1tex__m_xbar2l1tex_read_bytes.sum.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed','sm__inst_executed.avg.per_cycle_elapsed']
for k in keep:
    if k in h: print(k, '=', rows[2][h.index(k)], rows[1][h.index(k)])
" > gpurun_out/$f.txt; cat gpurun_out/$f.txt; echo; done
rm -f gpurun_out/r2_full_res2conv2.ncu-rep gpurun_out/r2_full_stem.ncu-rep
