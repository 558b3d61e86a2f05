#!/bin/bash
# Round-end evidence: launch list of the bench command, per-launch DRAM traffic of one UNet forward, ncu --set full of the top kernels.
set -u
TAG=${1:-r01h}
mkdir -p gpurun_out
BENCH="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
$BENCH > gpurun_out/${TAG}_bench_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${TAG}_launches_bench.csv $BENCH > gpurun_out/ncu_bench.log 2>&1
echo "launch list rc=$?"
FW="python tools/profile_unet_forward.py"
$FW > gpurun_out/fw_plain.log 2>&1 && \
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_unet_forward_traffic.csv $FW > gpurun_out/ncu_fw.log 2>&1
echo "traffic rc=$?"
PK="python tools/profile_kernels.py --iters 1 --only attn_L0,gemm1x1_res_L0,gemm_geglu_L0,conv3x3_L0,conv3x3_L2,groupnorm_L0,layernorm_L0"
$PK > gpurun_out/pk_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'flash_attn|gemm_tc|gn_|layernorm' -o gpurun_out/${TAG}_top_kernels $PK > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
# context numbers outside the metric: the configs[4] frame size (768x2496) and the RGB VAE encoder in front of the sampler
timeout 600 python tools/time_unet_k2.py 2 > gpurun_out/${TAG}_unet_k2.json 2> gpurun_out/k2.err; echo "k2 rc=$?"
timeout 300 python tools/time_vae_image.py > gpurun_out/${TAG}_vae_image.json 2> gpurun_out/vae.err; echo "vae rc=$?"
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_full.log 2> gpurun_out/${TAG}_bench_full.err; echo "bench rc=$?"
