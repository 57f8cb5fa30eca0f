#!/bin/bash
# ncu captures of the bench workload (1 GPU).  Usage: gpurun -- bash scripts/gpu_profile.sh
# Launch order of one step (256x256x48, 256 queries; python scripts/tile_model.py lists the GEMMs):
#   backbone ... cost volume (split_planes x2, gemm, head) ... 4 x (local_corr, mixer, update)
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
NCU="ncu --clock-control none"
cap() {  # name, kernel regex, skip, count, [script args]
  local name=$1 re=$2 skip=$3 cnt=$4; shift 4
  timeout 600 $NCU --set full --import-source on -k "regex:$re" -s "$skip" -c "$cnt" -f \
      -o "gpurun_out/prof_$name" python scripts/profile_step.py --warm 1 --steps 1 "$@" \
      > "gpurun_out/prof_$name.log" 2>&1
  echo "capture $name rc=$?"
}
# mixer_dw: 48 launches per step; skip the warm-up step's 48, take 2 of the timed step
cap mixer_dw mixer_dw_kernel 48 2
# the same kernel at the 4096-query regime (98304 rows per launch: HBM resident)
cap mixer_dw_c4 mixer_dw_kernel 48 1 --frames 96 --queries 1024
cap local_corr local_corr_kernel 4 1
cap head cost_volume_head 1 1
# gemm launches per step: 20 resnet + 10 extra_conv + 1 cost volume + 4*(1+24+1) = 135
cap gemm_extra gemm_tc $((135 + 20)) 2
cap gemm_mixer gemm_tc $((135 + 32)) 2
cap cost_volume_gemm gemm_tc $((135 + 30)) 1
cap instnorm_apply instnorm_relu_split_kernel 16 1
cap stem stem_conv_kernel 1 1
ls -la gpurun_out/*.ncu-rep
