#!/bin/bash
# Round-2 evidence (1 GPU): launch list of the bench command, per-launch DRAM traffic of one UNet forward, ncu --set full
# of the top kernels, context timings (768x2496 forward, RGB VAE encoder). Every ncu pass runs only after the same
# command has exited 0 without ncu; numbers printed under ncu are never bench values.
set -u
TAG=${1:-r02f}
mkdir -p gpurun_out
BENCH="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-baseline"
timeout 400 $BENCH > gpurun_out/${TAG}_bench_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${TAG}_launches_bench.csv $BENCH > gpurun_out/ncu_bench.log 2>&1
echo "launch list rc=$?"
FW="python tools/profile_unet_forward.py"
timeout 300 $FW > gpurun_out/fw_plain.log 2>&1 && \
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_unet_forward_traffic.csv $FW > gpurun_out/ncu_fw.log 2>&1
echo "traffic rc=$?"
PK="python tools/profile_kernels.py --iters 1 --only attn_L0,gemm_geglu_L0,gemm_qkv_L0,conv3x3_L0,conv3x3_L2,groupnorm_L0"
timeout 300 $PK > gpurun_out/pk_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'flash_attn|gemm_tc|gn_' -o gpurun_out/${TAG}_top_kernels -f $PK > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
timeout 300 python tools/profile_kernels.py --iters 5 --json gpurun_out/${TAG}_kernels.json > gpurun_out/pk_all.log 2>&1; echo "kernels rc=$?"
timeout 600 python tools/time_unet_k2.py 2 > gpurun_out/${TAG}_unet_k2.json 2> gpurun_out/k2.err; echo "k2 rc=$?"
timeout 300 python tools/time_vae_image.py > gpurun_out/${TAG}_vae_image.json 2> gpurun_out/vae.err; echo "vae rc=$?"
tail -2 gpurun_out/ncu_bench.log gpurun_out/ncu_fw.log gpurun_out/ncu_full.log
