#!/bin/bash
# usage (under gpurun): bash tools/ab_sweep.sh "variantA variantB ..." [extra bench args]  -- full-sweep A/B of library variants in one call, interleaved twice
for rep in 1 2; do
for v in $1; do
  lib=$PWD/latticeqmc_b200/variants/$v.so; [ "$v" = "main" ] && lib=$PWD/latticeqmc_b200/liblqmc_b200.so
  LQMC_B200_LIB=$lib python bench.py --no-cpu --no-configs --steps 5 --warmup 3 $2 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', round(d['ms_per_step'],2), 'ms', d['clocks']['sm_mhz'], d['clocks']['power_w_max'])"
done; done
