# Round-end evidence run on one B200: GPU tests, smoke, the bench line, the ncu launch list of the same command and a
# full ncu capture of the LightConv launches of one step.  Outputs under gpurun_out/ (copied into profiles/ by hand).
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -3 > gpurun_out/rp_pytest.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/rp_smoke.txt 2>&1
python bench.py > gpurun_out/rp_bench_n1.json 2> gpurun_out/rp_bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 760 --csv --log-file gpurun_out/rp_launches.csv \
    python bench.py --steps 3 --warmup 3 --skip-cpu > gpurun_out/rp_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_lightconv2|k_lightchain" -s 12 -c 12 -f \
    -o gpurun_out/rp_light_full python bench.py --steps 2 --warmup 3 --skip-cpu > gpurun_out/rp_light_full.log 2>&1
tail -2 gpurun_out/rp_pytest.txt; tail -1 gpurun_out/rp_smoke.txt; cut -c1-400 gpurun_out/rp_bench_n1.json
# DeepOCSORT config-3 shape: phase clocks of both CTA-wide dense-JV variants, and the tracker-level parity of the
# column-owned one (mode 2 is opt-in until this passes: then flip the default in tracker_engine.cu::jv_wide_flag)
for m in 1 2; do
  BOXMOT_B200_JV_WIDE=$m python scripts/docs_config3_clocks.py 9 > gpurun_out/rp_clocks3_mode$m.log 2>&1
  tail -1 gpurun_out/rp_clocks3_mode$m.log
done
BOXMOT_B200_JV_WIDE=2 python -m pytest -x -q tests/test_gpu_deepocsort_scale.py tests/test_gpu_trackers.py \
    tests/test_gpu_reid.py -k "config3 or deepocsort" 2>&1 | tail -2 > gpurun_out/rp_jv_mode2_tracker.txt
cat gpurun_out/rp_jv_mode2_tracker.txt
[ -x scripts/microbench/fp64_smem_probe ] && scripts/microbench/fp64_smem_probe > gpurun_out/rp_fp64_probe.json
