#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_unet.py -q -x -k "conv3x3 or convt" > gpurun_out/conv_bo1.log 2>&1; echo "base_offset=dx exit $?"; tail -5 gpurun_out/conv_bo1.log
ADN_HALO_BASE_OFFSET=0 timeout 600 python -m pytest tests/test_gpu_unet.py -q -x -k "conv3x3" > gpurun_out/conv_bo0.log 2>&1; echo "base_offset=0 exit $?"; tail -5 gpurun_out/conv_bo0.log
