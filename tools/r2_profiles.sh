#!/bin/bash
# Round-2 evidence capture (run under gpurun, 1 GPU): launch lists + `ncu --set full` of every kernel of the sampling step and of
# the training step, in-graph class costs, kernel microbenchmarks.  Outputs under gpurun_out/r02_*; summarise with
# tools/ncu_summary.py / tools/launch_summary.py and copy into profiles/.
set -u
O=gpurun_out
NCU="ncu --clock-control none"
T="timeout -s KILL"
# 1. sampling step: launch list (3 eager steps; the last 128 launches are one step)
$T 200 $NCU --metrics gpu__time_duration.sum --csv --log-file $O/r02_launches_cold.csv python tools/profile_step.py 3 > /dev/null 2>&1
# 2. sampling step: --set full of the last step's kernels, one capture per kernel family
$T 300 $NCU --set full --import-source on -k regex:"gemm_wres_kernel|gemm_stream_kernel" -s 151 -c 9 -o $O/r02_ncu_gemm python tools/profile_step.py 3 > /dev/null 2>&1
$T 200 $NCU --set full --import-source on -k regex:"attn_apply_tc_kernel|attn_kv_kernel" -s 88 -c 4 -o $O/r02_ncu_attn python tools/profile_step.py 3 > /dev/null 2>&1
$T 200 $NCU --set full --import-source on -k regex:"ln_film_silu_kernel|time_table_silu|tile_rows|timestep_embed|gemm_bf16_tcgen05|pack_motion" -s 20 -c 6 -o $O/r02_ncu_small python tools/profile_step.py 3 > /dev/null 2>&1
$T 200 $NCU --set full --import-source on -k regex:"ddpm_step_kernel|recover_joints|q_sample|advance_t" -c 6 -o $O/r02_ncu_diffusion python tools/r2_diffusion_ops.py > /dev/null 2>&1
# 3. training step: launch list + --set full of the backward / optimizer kernels (one graph-replayed iteration)
$T 200 $NCU --metrics gpu__time_duration.sum --profile-from-start off --csv --log-file $O/r02_train_launches.csv python tools/train_step.py --denoiser-only --iters 1 --profile > /dev/null 2>&1
$T 300 $NCU --set full --import-source on --profile-from-start off -k regex:"eff_attn_bwd_tc|ln_film_silu_bwd|colsum_vec|cast_colsum|act_bwd|act_fwd|eff_attn_bf16|adam_flat|sumsq|mse_" -c 20 -o $O/r02_ncu_train python tools/train_step.py --denoiser-only --iters 1 --profile > /dev/null 2>&1
$T 300 $NCU --set full --import-source on --profile-from-start off -k regex:"gemm_bf16_2cta" -s 100 -c 10 -o $O/r02_ncu_train_gemm python tools/train_step.py --denoiser-only --iters 1 --profile > /dev/null 2>&1
# 4. in-graph class costs and A/B of the tcgen05 apply kernel
$T 300 python tools/step_breakdown.py 100 > $O/r02_step_breakdown.txt 2>&1
$T 300 python tools/step_ab.py 200 4 HIG_APPLY_TC=0,1 2>&1 | tail -10 > $O/r02_step_ab_apply_tc.txt
# 5. reports -> CSV tables (the .ncu-rep files are too large to bring back: gpurun merges at most 64 MiB)
for f in $O/r02_ncu_*.ncu-rep; do
  ncu -i $f --page raw --csv > ${f%.ncu-rep}.csv 2>/dev/null
done
python tools/ncu_summary.py $O/r02_ncu_*.ncu-rep > $O/r02_ncu_full_summary.txt 2>&1
python tools/ncu_stalls.py $O/r02_ncu_*.ncu-rep > $O/r02_ncu_stalls.txt 2>&1
rm -f $O/r02_ncu_train.ncu-rep $O/r02_ncu_train_gemm.ncu-rep $O/r02_ncu_gemm.ncu-rep $O/r02_ncu_small.ncu-rep $O/r02_ncu_diffusion.ncu-rep
python tools/launch_summary.py $O/r02_launches_cold.csv 3 > $O/r02_launch_summary_cold.txt 2>&1
python tools/launch_summary.py $O/r02_train_launches.csv > $O/r02_train_launch_summary.txt 2>&1
ls -la $O/r02_* | awk '{print $5, $9}'
