#!/bin/bash
# ncu captures of the bench workload (1 GPU).  Usage: gpurun -- bash scripts/gpu_profile.sh
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
NCU="ncu --clock-control none"
# launch list of ONE step (after 2 warm-up steps = 2*~261 launches + weight packing)
timeout 600 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches.csv \
    python scripts/profile_step.py --warm 1 --steps 1 > gpurun_out/launches.log 2>&1
cap() {  # name, kernel regex, skip, count
  timeout 600 $NCU --set full --import-source on -k "regex:$2" -s "$3" -c "$4" -f \
      -o "gpurun_out/prof_$1" python scripts/profile_step.py --warm 1 --steps 1 \
      > "gpurun_out/prof_$1.log" 2>&1
  echo "capture $1 rc=$?"
}
# gemm launches per step: 20 resnet + 10 extra_conv + 1 cost volume + 4*(1+24+1) = 135
cap mixer_dw mixer_dw_kernel 48 2
cap gemm_extra gemm_tc $((135 + 20)) 2
cap gemm_mixer gemm_tc $((135 + 32)) 2
cap local_corr local_corr_kernel 4 1
cap head cost_volume_head 1 1
cap instnorm_apply instnorm_relu_split_kernel 16 1
cap cost_volume_gemm gemm_tc $((135 + 30)) 1
cap stem stem_conv_kernel 1 1
ls -la gpurun_out/*.ncu-rep
