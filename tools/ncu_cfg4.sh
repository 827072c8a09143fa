#!/bin/bash
# usage (under gpurun): bash tools/ncu_cfg4.sh TAG   -- launch list + one full capture of the cfg4 sweep kernel, next to an un-profiled bench line
TAG=${1:-r02}
python bench.py --no-cpu --no-configs --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_cfg4.json 2> gpurun_out/${TAG}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/${TAG}_launches_cfg4.csv python bench.py --no-cpu --no-configs --steps 2 --warmup 3 > gpurun_out/${TAG}_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sweep_l2 -s 3 -c 1 -o gpurun_out/${TAG}_cfg4 python bench.py --no-cpu --no-configs --steps 2 --warmup 3 > gpurun_out/${TAG}_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_ncu.log; grep -c sweep_l2 gpurun_out/${TAG}_launches_cfg4.csv
