#!/bin/bash
# Round-2 ncu captures for profiles/ (run under gpurun, one GPU).  Outputs land in gpurun_out/; summarise here with
#   python scripts/summarize_ncu.py launches gpurun_out/launches_r02.csv profiles/r02_launches_default.csv
#   python scripts/summarize_ncu.py full gpurun_out/r02_<name>.ncu-rep profiles/r02_<name>_full.csv
set -x
mkdir -p gpurun_out
P1="python scripts/profile_step.py --steps 1 --warmup 1"
PB="python scripts/bench_batch.py large-v2 16 1"
NCU="ncu --set full --clock-control none --import-source on -f"
# (1) every launch of one headline step (B = 1, beam 5): shares of the step
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02.csv $P1 > gpurun_out/ncu_list_r02.log 2>&1
# (2) the kernels of the headline step, one full capture each
timeout 600 $NCU -k regex:dec_pass_kernel -s 17 -c 2 -o gpurun_out/r02_dec_pass_kernel $P1 > gpurun_out/ncu_r02_a.log 2>&1
timeout 600 $NCU -k regex:gemm_tc_kernel -s 132 -c 6 -o gpurun_out/r02_gemm_tc $P1 > gpurun_out/ncu_r02_b.log 2>&1
timeout 600 $NCU -k regex:enc_attn_kernel -s 33 -c 2 -o gpurun_out/r02_enc_attn $P1 > gpurun_out/ncu_r02_c.log 2>&1
timeout 600 $NCU -k regex:"logmel_power_kernel|logmel_finalize_kernel|layernorm_f32_to_f16_kernel|conv1_gelu_kernel|topk_partial_kernel|topk_merge_kernel" -s 75 -c 8 -o gpurun_out/r02_small_kernels $P1 > gpurun_out/ncu_r02_d.log 2>&1
# (3) the batched decoder pass (16 utterances x beam 5 = 80 rows)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 800 -c 800 --csv --log-file gpurun_out/launches_batch_r02.csv $PB > gpurun_out/ncu_list_batch_r02.log 2>&1
timeout 900 $NCU -k regex:"bd_cross_attn_tc_kernel" -s 40 -c 2 -o gpurun_out/r02_bd_cross_attn_tc $PB > gpurun_out/ncu_r02_e.log 2>&1
timeout 900 $NCU -k regex:"gemm_tc_kernel<64|gemm_tc_kernel<128|gemm_tc_kernel<256" -s 560 -c 8 -o gpurun_out/r02_gemm_tc_decode $PB > gpurun_out/ncu_r02_f.log 2>&1
timeout 900 $NCU -k regex:"bd_self_attn_kernel|bd_resid_ln_kernel|bd_embed_ln_kernel" -s 300 -c 5 -o gpurun_out/r02_bd_small $PB > gpurun_out/ncu_r02_g.log 2>&1
ls -la gpurun_out/ | tail -20
