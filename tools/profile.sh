#!/bin/bash
# ncu evidence for profiles/: run under gpurun (1 GPU).  usage: tools/profile.sh <tag> [launches] [march] [traffic] [gemm]
#   launches: gpurun_out/<tag>_launches.csv   device time of every launch of a reduced bench (1024 poses = 1 chunk/step)
#   march   : gpurun_out/<tag>_{gather,mlp,ws}.ncu-rep  --set full of one k_gather_round / k_mlp_round (round 3) / k_march_ws launch
#   traffic : gpurun_out/<tag>_traffic.csv    dram + lts bytes and time of every march kernel of ONE 512-candidate launch at 800x800
#             (tools/march_traffic.py turns it into profiles/march_ncu_summary.json, stamped with the kernel sources' hash)
#   gemm    : gpurun_out/<tag>_gemm.ncu-rep   --set full, the five ViT-B/32 GEMM shapes at batch 512
TAG=${1:-prof}; shift
PARTS=${@:-launches march traffic gemm}
mkdir -p gpurun_out
S="python tools/ws_stats.py --poses 1024 --reps 1"
for p in $PARTS; do
  case $p in
    launches) ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/${TAG}_launches.csv \
                  python bench.py --poses 1024 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_launches.log 2>&1 ;;
    march)    ncu --set full --clock-control none --import-source on -k regex:k_gather_round -s 3 -c 1 -f -o gpurun_out/${TAG}_gather $S > gpurun_out/${TAG}_gather.log 2>&1
              ncu --set full --clock-control none --import-source on -k regex:k_mlp_round -s 3 -c 1 -f -o gpurun_out/${TAG}_mlp $S > gpurun_out/${TAG}_mlp.log 2>&1
              D2R_MARCH=ws ncu --set full --clock-control none --import-source on -k regex:k_march_ws -s 1 -c 1 -f -o gpurun_out/${TAG}_ws python tools/ws_stats.py --poses 256 --chunk 256 --reps 1 > gpurun_out/${TAG}_ws.log 2>&1 ;;
    traffic)  # the background render (K = 1) and the warm-up chunk come first: tools/march_traffic.py keeps the LAST chunk's launches
              ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,gpu__time_duration.sum --clock-control none \
                  -k regex:"k_gather_round|k_mlp_round|k_march_ws|k_classify|k_finish" -c 4000 --csv --log-file gpurun_out/${TAG}_traffic.csv \
                  python tools/ws_stats.py --poses 512 --chunk 512 --reps 1 > gpurun_out/${TAG}_traffic.log 2>&1 ;;
    gemm)     ncu --set full --clock-control none --import-source on -k regex:k_gemm_f16 -c 10 -f -o gpurun_out/${TAG}_gemm \
                  python tools/stage_bench.py --gemm-only --iters 1 --warm 1 > gpurun_out/${TAG}_gemm.log 2>&1 ;;
  esac
done
ls -la gpurun_out/ | grep ${TAG}
