#!/bin/bash
# Round-2 final evidence call on one B200: new-code tests first, then the whole GPU suite, the default bench line, the ncu
# launch list + full ReID capture of the same command, config-3 clocks / bench, ECC timing.  Outputs under gpurun_out/.
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_zzgpu_cmc.py -q -s 2>&1 | tail -25 > gpurun_out/r2k_cmc_tests.log; tail -4 gpurun_out/r2k_cmc_tests.log
timeout 120 python -m pytest tests/test_gpu_kernels.py -q -k dense_jv 2>&1 | tail -5 > gpurun_out/r2k_jv_kernels.log; tail -2 gpurun_out/r2k_jv_kernels.log
timeout 60 python scripts/docs_config3_clocks.py 9 > gpurun_out/r2k_clocks3_mode3.log 2>&1; tail -1 gpurun_out/r2k_clocks3_mode3.log
(time timeout 900 python -m pytest tests -q -m gpu --deselect tests/test_zzgpu_cmc.py -k "not dense_jv" 2>&1 | tail -15) > gpurun_out/r2k_gpu_suite.log 2>&1; tail -6 gpurun_out/r2k_gpu_suite.log
timeout 300 python bench.py > gpurun_out/r2k_bench_default.json 2> gpurun_out/r2k_bench_default.err; tail -1 gpurun_out/r2k_bench_default.err | cut -c1-200; cut -c1-300 gpurun_out/r2k_bench_default.json
BOXMOT_B200_REID_SPLIT=1 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 140 --csv \
    --log-file gpurun_out/r2k_launches.csv python bench.py --steps 6 --warmup 4 --skip-cpu --no-extra > gpurun_out/r2k_launches.log 2>&1
BOXMOT_B200_REID_SPLIT=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_front_tc|k_chain_tc|k_gemm_tc|k_gates_tc|k_head" \
    -s 60 -c 20 -f -o gpurun_out/r2k_reid_full python bench.py --steps 3 --warmup 3 --skip-cpu --no-extra > gpurun_out/r2k_reid_full.log 2>&1
ls -la gpurun_out | grep r2k
timeout 200 python bench.py --config 3 --steps 30 --warmup 5 --skip-cpu > gpurun_out/r2k_bench_c3.json 2> gpurun_out/r2k_bench_c3.err; cut -c1-260 gpurun_out/r2k_bench_c3.json
timeout 120 python scripts/measure_cmc.py > gpurun_out/r2k_cmc_timing.log 2>&1; cat gpurun_out/r2k_cmc_timing.log
