#!/bin/bash
# usage (under gpurun): bash tools/gpu_batch.sh TAG "v0_base v1_flush ..." "clk variants" -- slice-phase A/B of library variants
TAG=$1; VARS=$2; CLKS=$3
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.log 2>&1
for v in $VARS; do
  LQMC_B200_LIB=$PWD/latticeqmc_b200/variants/$v.so timeout 300 python tools/phase_times.py cfg4 296 > gpurun_out/${TAG}_phase_$v.log 2>&1
done
for v in $CLKS; do
  for c in 1 148 296; do
    LQMC_B200_LIB=$PWD/latticeqmc_b200/variants/$v.so timeout 300 python tools/phase_clocks.py cfg4 $c >> gpurun_out/${TAG}_clk_$v.log 2>&1
  done
done
tail -n 6 gpurun_out/${TAG}_phase_*.log
tail -n 12 gpurun_out/${TAG}_clk_*.log
