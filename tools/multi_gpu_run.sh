#!/bin/bash
# usage (under gpurun --gpus N): bash tools/multi_gpu_run.sh N TAG  -- weak and strong bench lines on N GPUs of one box
N=$1; TAG=$2
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu > gpurun_out/${TAG}_bench_${N}gpu_weak.json 2> gpurun_out/${TAG}_${N}gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu --no-configs --scaling strong > gpurun_out/${TAG}_bench_${N}gpu_strong.json 2>> gpurun_out/${TAG}_${N}gpu.err
for f in weak strong; do python -c "
import json
d=json.loads(open('gpurun_out/${TAG}_bench_${N}gpu_$f.json').read().strip().splitlines()[-1])
print('$f', d['n_gpus'], round(d['ms_per_step'],2), 'ms', '%.4e'%d['value'], d['config'].get('chains_per_gpu'), d['config'].get('ctas_per_chain'), d.get('observable_reduction',{}).get('allreduce_ms'))
"; done
