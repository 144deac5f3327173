#!/bin/bash
# ncu captures for profiles/ (run under gpurun, one GPU).  Outputs land in gpurun_out/.
set -x
mkdir -p gpurun_out
R=${1:-r01}
P="python scripts/profile_step.py --steps 1 --warmup 1 --no-graphs"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${R}.csv python scripts/profile_step.py --steps 1 --warmup 1 > gpurun_out/ncu_list.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 131 -c 6 -o gpurun_out/${R}_gemm -f $P > gpurun_out/ncu_gemm.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemv_kernel -s 3600 -c 8 -o gpurun_out/${R}_gemv -f $P > gpurun_out/ncu_gemv.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"enc_attn_kernel|dec_cross_attn|dec_self_attn" -s 640 -c 3 -o gpurun_out/${R}_attn -f $P > gpurun_out/ncu_attn.log 2>&1
ls -la gpurun_out/
