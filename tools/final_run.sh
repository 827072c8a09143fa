TAG=r02m
python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_cfg4_full.json 2> gpurun_out/${TAG}_bench_full.err
bash tools/ncu_cfg4.sh $TAG
python bench.py --no-cpu --no-configs --steps 3 --warmup 3 --workload cfg5 > gpurun_out/${TAG}_bench_cfg5.json 2>/dev/null
python bench.py --no-cpu --no-configs --steps 5 --warmup 3 --arith fma > gpurun_out/${TAG}_bench_cfg4_fma.json 2>/dev/null
python bench.py --no-cpu --no-configs --steps 3 --warmup 3 --mode physics --stab 8 > gpurun_out/${TAG}_bench_phys_cfg4.json 2>/dev/null
python bench.py --no-cpu --no-configs --steps 5 --warmup 3 --workload cfg2 > gpurun_out/${TAG}_bench_cfg2.json 2>/dev/null
python bench.py --no-cpu --no-configs --steps 5 --warmup 3 --workload cfg3 > gpurun_out/${TAG}_bench_cfg3.json 2>/dev/null
for f in gpurun_out/${TAG}_bench_*.json; do python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1])
print('$f', round(d['ms_per_step'],2), 'ms', '%.3e'%d['value'], 'frac', round(d['roofline']['frac'],3), d['clocks']['sm_mhz'], d['clocks']['reasons'])
"; done
