#!/bin/bash
# End-of-round ncu evidence (run on the GPU box: gpurun -- bash tools/capture_evidence.sh TAG).  Every profiled command is first
# run plain and must exit 0; numbers printed under ncu are never bench values.
TAG=${1:-r2}
set -x
python tools/profile_step.py > gpurun_out/${TAG}_plain_step.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_step.csv python tools/profile_step.py > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -s 12 -c 12 -f -o gpurun_out/${TAG}_full_step python tools/profile_step.py > gpurun_out/${TAG}_ncu_full.log 2>&1
python bench.py --steps 2 --warmup 1 --no-cpu --no-matching --configs tum > gpurun_out/${TAG}_plain_bench.json 2> /dev/null || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-matching --configs tum > /dev/null 2>&1
python tools/profile_knn.py 5 > gpurun_out/${TAG}_plain_knn.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:knn2_pair_kernel -s 1 -c 1 -f -o gpurun_out/${TAG}_full_knn python tools/profile_knn.py 5 > gpurun_out/${TAG}_ncu_knn.log 2>&1
ls -la gpurun_out | tail -12
