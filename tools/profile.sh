#!/bin/bash
# ncu evidence for profiles/: run under gpurun (1 GPU).  usage: tools/profile.sh <tag> [launches] [march] [traffic] [gemm]
#   launches: gpurun_out/<tag>_launches.csv   device time of every launch of a reduced bench (1024 poses = 2 chunks/step)
#   march   : gpurun_out/<tag>_gather.ncu-rep, <tag>_mlp.ncu-rep  --set full, round 0 of a 512-candidate chunk at 800x800
#   traffic : gpurun_out/<tag>_traffic.csv    dram bytes of every march kernel of one bench run (sum per chunk = roofline.traffic)
#   gemm    : gpurun_out/<tag>_gemm.ncu-rep   --set full, the five ViT-B/32 GEMM shapes at batch 512
TAG=${1:-prof}; shift
PARTS=${@:-launches march traffic gemm}
mkdir -p gpurun_out
B="python bench.py --poses 512 --steps 1 --warmup 3 --no-cpu-baseline"
for p in $PARTS; do
  case $p in
    launches) ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/${TAG}_launches.csv \
                  python bench.py --poses 1024 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_launches.log 2>&1 ;;
    march)    # the background render (K = 1) comes first and takes ~40 rounds: skip its launches, take round 0 of the first chunk
              ncu --set full --clock-control none --import-source on -k regex:k_gather_round -s 40 -c 1 -f -o gpurun_out/${TAG}_gather $B > gpurun_out/${TAG}_gather.log 2>&1
              ncu --set full --clock-control none --import-source on -k regex:k_mlp_round -s 40 -c 1 -f -o gpurun_out/${TAG}_mlp $B > gpurun_out/${TAG}_mlp.log 2>&1 ;;
    traffic)  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
                  -k regex:"k_gather_round|k_mlp_round|k_classify|k_finish" -c 2000 --csv --log-file gpurun_out/${TAG}_traffic.csv $B > gpurun_out/${TAG}_traffic.log 2>&1 ;;
    gemm)     ncu --set full --clock-control none --import-source on -k regex:k_gemm_f16 -c 10 -f -o gpurun_out/${TAG}_gemm \
                  python tools/stage_bench.py --gemm-only --iters 1 --warm 1 > gpurun_out/${TAG}_gemm.log 2>&1 ;;
  esac
done
ls -la gpurun_out/ | grep ${TAG}
