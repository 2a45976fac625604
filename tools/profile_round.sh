#!/bin/bash
# Developer aid (GPU box): the ncu evidence of a round, only after the same command has exited 0 without ncu.
#   tools/profile_round.sh TAG [WORKLOAD]
TAG=$1; WL=${2:-C3}
CMD="python bench.py --workload $WL --steps 16 --warmup 3 --no-cpu-baseline --no-extras"
$CMD > gpurun_out/${TAG}_plain.log 2> gpurun_out/${TAG}_plain.err || { echo "plain run failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_launch.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:lattice_kernel -s 6 -c 1 -o gpurun_out/${TAG}_lattice -f $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:softmax_rows -s 6 -c 1 -o gpurun_out/${TAG}_softmax -f $CMD > gpurun_out/${TAG}_ncu2.log 2>&1
ncu --set full --clock-control none -k regex:plan_kernel -s 6 -c 1 -o gpurun_out/${TAG}_plan -f $CMD > gpurun_out/${TAG}_ncu3.log 2>&1
tail -2 gpurun_out/${TAG}_ncu1.log
