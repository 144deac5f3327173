#!/bin/bash
# Round-2 (second half) ncu captures for profiles/: the warp-MMA persistent decoder pass is the default path now.
# Run under gpurun (one GPU).  Outputs land in gpurun_out/; summarise here with
#   python scripts/summarize_ncu.py launches gpurun_out/launches_r02b.csv profiles/r02_launches_default.csv
#   python scripts/summarize_ncu.py full gpurun_out/r02b_<name>.ncu-rep profiles/r02_<name>_full.csv
set -x
mkdir -p gpurun_out
P1="python scripts/profile_step.py --steps 1 --warmup 1"
NCU="ncu --set full --clock-control none --import-source on -f"
# (1) every launch of one headline step (B = 1, beam 5): shares of the step
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02b.csv $P1 > gpurun_out/ncu_list_r02b.log 2>&1
# (2) the persistent decoder pass: two launches of the measured step
timeout 900 $NCU -k regex:dec_pass_mma_kernel -s 18 -c 2 -o gpurun_out/r02b_dec_pass_mma_kernel $P1 > gpurun_out/ncu_r02b_a.log 2>&1
# (3) the encoder kernels (programmatic dependent launch hooks added) and the fused search tail
timeout 900 $NCU -k regex:"gemm_tc_kernel|enc_attn_kernel|layernorm_f32_to_f16_kernel" -s 236 -c 14 -o gpurun_out/r02b_encoder $P1 > gpurun_out/ncu_r02b_b.log 2>&1
timeout 600 $NCU -k regex:"search_tail_kernel|topk_partial_kernel" -s 30 -c 4 -o gpurun_out/r02b_search $P1 > gpurun_out/ncu_r02b_c.log 2>&1
ls -la gpurun_out/ | tail -12
