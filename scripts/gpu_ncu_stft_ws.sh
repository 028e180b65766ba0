#!/bin/bash
# ncu --set full with source-level sampling of the warp-specialised STFT at config 2 (launch 2 of prof_spectral.py)
set -u
mkdir -p gpurun_out
python scripts/prof_spectral.py > gpurun_out/prof_spectral_plain.log 2>&1 || { echo plain failed; tail -5 gpurun_out/prof_spectral_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:"stft_ws_kernel" -s 2 -c 1 -o gpurun_out/prof_stft_ws -f python scripts/prof_spectral.py > gpurun_out/ncu_stft.log 2>&1; echo "stft ncu exit $?"
ls -la gpurun_out/*.ncu-rep
