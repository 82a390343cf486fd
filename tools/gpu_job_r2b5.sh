#!/bin/bash
# evidence refresh after the epilogue change: ncu launch list of one steady-state forward (our kernels only)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python tools/profile_forward.py --forwards 2 > gpurun_out/r02b_profile_forward_plain.log 2>&1; echo "plain rc=$?"; tail -2 gpurun_out/r02b_profile_forward_plain.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"attn_fwd_kernel|gemm_bf16_kernel|ln_modulate_kernel|modulation_table_kernel|patchify_kernel|rmsnorm|silu_kernel|sinusoidal_kernel" -s 1670 -c 402 --csv --log-file gpurun_out/r02b_launches_one_forward.csv python tools/profile_forward.py --forwards 2 > gpurun_out/r02b_profile_forward_ncu.log 2>&1; echo "ncu rc=$?"
