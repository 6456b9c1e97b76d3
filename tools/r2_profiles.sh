#!/bin/bash
# Round-2 profile evidence on one B200 (one gpurun call; every ncu run follows a plain run of the same command):
#   launch list of a short cfg2 bench, full captures of the two pipeline kernels (warm caches) at cfg2,
#   cold-cache (--cache-control all) DRAM traffic of both kernels at cfg2 and cfg3.
mkdir -p gpurun_out
B="python bench.py --no-e2e --no-cpu-baseline --no-cfg3 --parity-frames 0 --steps 2 --warmup 3 --distinct-images 32"
C2="$B --workload cfg2 --frames-per-step 64"
C3="$B --workload cfg3 --frames-per-step 32"
$C2 > gpurun_out/r2p_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_cfg2.csv $C2 > gpurun_out/r2p_l.log 2>&1
$C2 > gpurun_out/r2p_plain2.log 2>&1 &&
ncu --set full --clock-control none --cache-control none --import-source on -k regex:"k_apply_chunk|k_expand" -s 12 -c 4 -o gpurun_out/r2_full_cfg2 -f $C2 > gpurun_out/r2p_f.log 2>&1
M="dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sectors.sum,smsp__inst_executed.sum"
$C2 > gpurun_out/r2p_plain3.log 2>&1 &&
ncu --metrics $M --clock-control none --cache-control all -k regex:"k_apply_chunk|k_expand" -s 12 -c 8 --csv --log-file gpurun_out/r2_cold_cfg2.csv $C2 > gpurun_out/r2p_c2.log 2>&1
$C3 > gpurun_out/r2p_plain4.log 2>&1 &&
ncu --metrics $M --clock-control none --cache-control all -k regex:"k_apply_chunk|k_expand" -s 12 -c 8 --csv --log-file gpurun_out/r2_cold_cfg3.csv $C3 > gpurun_out/r2p_c3.log 2>&1
tail -2 gpurun_out/r2p_f.log; ls -la gpurun_out/r2_*
./tools/bin/microbench_rmw > gpurun_out/r2_microbench_rmw.txt 2>&1
