#!/bin/bash
# GPU check of the dense-JV shortcuts (mode 3): standalone solver cases in every variant, DeepOCSORT / OC-SORT tracker
# tests, config-3 phase clocks before (mode 2) and after (mode 3), config-3 bench line.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -k dense_jv 2>&1 | tail -15 > gpurun_out/jv_kernels.log; tail -3 gpurun_out/jv_kernels.log
timeout 400 python -m pytest tests/test_gpu_deepocsort_scale.py tests/test_gpu_trackers.py tests/test_zgpu_late_goldens.py tests/test_gpu_baseline_configs.py -q -k "deepocsort or ocsort or config3 or docs or C3 or c3" 2>&1 | tail -15 > gpurun_out/jv_trackers.log; tail -3 gpurun_out/jv_trackers.log
BOXMOT_B200_JV_WIDE=2 timeout 120 python scripts/docs_config3_clocks.py 9 > gpurun_out/clocks3_mode2.log 2>&1; tail -2 gpurun_out/clocks3_mode2.log
for m in 3 7 11 19 35; do
  BOXMOT_B200_JV_WIDE=$m timeout 120 python scripts/docs_config3_clocks.py 9 > gpurun_out/clocks3_mode$m.log 2>&1; tail -1 gpurun_out/clocks3_mode$m.log
done
timeout 300 python bench.py --config 3 --steps 30 --warmup 5 --skip-cpu > gpurun_out/r2j_bench_c3.json 2> gpurun_out/r2j_bench_c3.err; tail -1 gpurun_out/r2j_bench_c3.err | cut -c1-300; cut -c1-400 gpurun_out/r2j_bench_c3.json
