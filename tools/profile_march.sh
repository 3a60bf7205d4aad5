#!/bin/bash
# ncu --set full of the ray-march kernel only (1 launch of 512 candidates).  usage: tools/profile_march.sh <tag>
TAG=${1:-prof}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_march_tc -s 2 -c 1 -f -o gpurun_out/${TAG}_march \
    python bench.py --poses 512 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_march.log 2>&1
ls -la gpurun_out/${TAG}_march*
