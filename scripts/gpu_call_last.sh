#!/bin/bash
# last short check of the round: fixed ECC tests, dense-JV variants after the batched row loads, config-3 clocks / ids / bench
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_zzgpu_cmc.py -q 2>&1 | tail -4 > gpurun_out/r2l_cmc_tests.log; tail -2 gpurun_out/r2l_cmc_tests.log
timeout 100 python -m pytest tests/test_gpu_kernels.py -q -k dense_jv 2>&1 | tail -3 > gpurun_out/r2l_jv_kernels.log; tail -1 gpurun_out/r2l_jv_kernels.log
timeout 60 python scripts/docs_config3_clocks.py 9 > gpurun_out/r2l_clocks3_mode3.log 2>&1; tail -1 gpurun_out/r2l_clocks3_mode3.log
timeout 150 python -m pytest tests/test_gpu_deepocsort_scale.py tests/test_gpu_trackers.py tests/test_zgpu_late_goldens.py -q -k "deepocsort or ocsort or config3" 2>&1 | tail -3 > gpurun_out/r2l_jv_trackers.log; tail -1 gpurun_out/r2l_jv_trackers.log
timeout 100 python bench.py --config 3 --steps 30 --warmup 5 --skip-cpu > gpurun_out/r2l_bench_c3.json 2> gpurun_out/r2l_bench_c3.err; cut -c1-260 gpurun_out/r2l_bench_c3.json
