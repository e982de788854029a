#!/bin/bash
# end-of-round evidence on one B200: launch list + ncu full capture of one frame's ReID kernels, the default bench line
# (CPU leg + parity), the other BASELINE configurations
bash scripts/round_profile.sh 2>&1 | tail -3
scripts/run_bench_configs.sh 3 4 5
