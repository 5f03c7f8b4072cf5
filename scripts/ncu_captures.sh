#!/bin/bash
# ncu --set full captures of the three dominant kernels (run on the GPU box; reports land in gpurun_out/).
#   conv:  conv_umma_kernel<2, 0, 6, 0>  launch #43 of a forward = yolo_blocks.1.body.1 (3x3 512->1024 @26^2, batch 32)
#   wgrad: wgrad_umma_kernel of a 3x3 256->512 @26^2 layer at batch 16
#   decode: decode_kernel<0> / <1> on cold heads
set -u
mkdir -p gpurun_out
COMMON="--set full --clock-control none --import-source on"
timeout 500 ncu $COMMON --kernel-name-base demangled -k 'regex:conv_umma_kernel<\(int\)2, \(bool\)0, \(int\)6, \(bool\)0>' -s 43 -c 1 \
    -o gpurun_out/r2_conv_k6_full -f python bench.py --steps 1 --warmup 3 --no-train --no-secondary --no-parity --no-cpu-baseline > gpurun_out/ncu_conv.log 2>&1
tail -2 gpurun_out/ncu_conv.log
if [ "${1:-}" = "all" ]; then
  WARM=1 STEPS=1 timeout 400 ncu $COMMON -k regex:wgrad_umma_kernel -s 100 -c 1 -o gpurun_out/r2_wgrad_full -f python scripts/train_bench.py > gpurun_out/ncu_wgrad.log 2>&1
  timeout 300 ncu $COMMON -k regex:decode_kernel -s 20 -c 3 -o gpurun_out/r2_decode_full -f python scripts/decode_probe.py > gpurun_out/ncu_decode.log 2>&1
fi
ls -la gpurun_out/*.ncu-rep
