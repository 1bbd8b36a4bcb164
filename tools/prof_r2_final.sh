#!/bin/bash
# Final round-2 measurement pass (run on the GPU box): GPU test-suite, bench lines (veryfast headline, medium
# for comparison), conversion micro-benchmark, ncu launch list of the bench command, full captures of the
# chain kernels + the decoder kernels.  Outputs under gpurun_out/.
set -x
[ -n "$SKIP_TESTS" ] || python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/r02f_gputests.log
python bench.py > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err
B200_BENCH_PRESET=medium python bench.py --steps 8 > gpurun_out/r02f_bench_medium.json 2>> gpurun_out/r02f_bench.err
python tools/bench_convert.py > gpurun_out/r02f_bench_convert.jsonl 2>> gpurun_out/r02f_bench.err
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 600 -c 500 \
    --csv --log-file gpurun_out/r02f_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r02f_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_me_ctu|k_sao_ctu|k_inter_recon|k_intra_frame|k_intra_modes|k_deblock' -c 12 \
    -o gpurun_out/r02f_prof_chain python tools/prof_one.py > gpurun_out/r02f_prof_chain.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_parse_rows|k_intra_decode|k_mjpg_idct|k_rgb24|k_rgb_to_i420' -c 8 \
    -o gpurun_out/r02f_prof_dec python tools/prof_dec.py > gpurun_out/r02f_prof_dec.log 2>&1
ls -la gpurun_out/*.ncu-rep
