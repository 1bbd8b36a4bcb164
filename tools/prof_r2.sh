#!/bin/bash
# Round-2 profile pass (run on the GPU box): bench, ncu launch list of the same command, full
# captures of the chain kernels.  Outputs under gpurun_out/.
set -x
python bench.py --steps 8 --warmup 3 > gpurun_out/bench_r2c.json 2> gpurun_out/bench_r2c.err
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 600 -c 500 \
    --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/bench_under_ncu_r2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_me_ctu|k_sao_ctu|k_inter_recon|k_intra_frame' -c 12 \
    -o gpurun_out/prof_r2_chain python tools/prof_one.py > gpurun_out/prof_r2_chain.log 2>&1
ls -la gpurun_out/*.ncu-rep
