#!/bin/bash
# ncu evidence for profiles/: run under gpurun (1 GPU).  usage: tools/profile.sh <tag> [launches] [march] [gemm]
#   launches: gpurun_out/<tag>_launches.csv  device time of every launch of a reduced bench (1024 poses = 2 launches/step)
#   march   : gpurun_out/<tag>_march.ncu-rep --set full, one launch of the ray-march kernel (512 candidates, 800x800)
#   gemm    : gpurun_out/<tag>_gemm.ncu-rep  --set full, the five ViT-B/32 GEMM shapes at batch 512
TAG=${1:-prof}; shift
PARTS=${@:-launches march gemm}
mkdir -p gpurun_out
for p in $PARTS; do
  case $p in
    launches) ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/${TAG}_launches.csv \
                  python bench.py --poses 1024 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_launches.log 2>&1 ;;
    march)    ncu --set full --clock-control none --import-source on -k regex:k_march_tc -s 2 -c 1 -f -o gpurun_out/${TAG}_march \
                  python bench.py --poses 512 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_march.log 2>&1 ;;
    gemm)     ncu --set full --clock-control none --import-source on -k regex:k_gemm_f16 -c 10 -f -o gpurun_out/${TAG}_gemm \
                  python tools/stage_bench.py --gemm-only --iters 1 --warm 1 > gpurun_out/${TAG}_gemm.log 2>&1 ;;
  esac
done
ls -la gpurun_out/ | grep ${TAG}
