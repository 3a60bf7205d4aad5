#!/bin/bash
# ncu evidence for profiles/: run under gpurun (1 GPU).  usage: tools/profile.sh <tag>
# Writes gpurun_out/<tag>_launches.csv (every launch of a reduced bench, device time), <tag>_march.ncu-rep
# (--set full, the ray-march kernel), <tag>_gemm.ncu-rep (--set full, the five ViT GEMM shapes).
TAG=${1:-prof}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --poses 1024 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_march_tc -s 2 -c 1 -f -o gpurun_out/${TAG}_march \
    python bench.py --poses 512 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_march.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_gemm_f16 -c 10 -f -o gpurun_out/${TAG}_gemm \
    python tools/stage_bench.py --gemm-only --iters 1 --warm 1 > gpurun_out/${TAG}_gemm.log 2>&1
ls -la gpurun_out/
