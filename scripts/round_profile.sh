# Round-2 evidence run on one B200 (outputs under gpurun_out/, summaries copied into profiles/ by scripts/summarise_profile.py):
#   1. ncu launch list of `bench.py --steps 6` (per-launch device time, one frame's kernel sequence)
#   2. ncu --set full of every ReID kernel of one frame (tensor pipe %, DRAM bytes, issue activity) -> traffic.json
set -x
mkdir -p gpurun_out
BOXMOT_B200_REID_SPLIT=1 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 140 --csv \
    --log-file gpurun_out/r2h_launches.csv python bench.py --steps 6 --warmup 4 --skip-cpu --no-extra > gpurun_out/r2h_launches.log 2>&1
BOXMOT_B200_REID_SPLIT=1 ncu --set full --clock-control none --import-source on -k regex:"k_front_tc|k_chain_tc|k_gemm_tc|k_gates_tc|k_head" \
    -s 84 -c 24 -f -o gpurun_out/r2h_reid_full python bench.py --steps 3 --warmup 3 --skip-cpu --no-extra > gpurun_out/r2h_reid_full.log 2>&1
ls -la gpurun_out | tail -4
