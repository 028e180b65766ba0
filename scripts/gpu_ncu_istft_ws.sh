#!/bin/bash
# ncu --set full with source-level sampling of the warp-specialised iSTFT: launch 2 = seeded phase, launch 5 = explicit phasor
set -u
mkdir -p gpurun_out
python scripts/prof_spectral.py > gpurun_out/prof_spectral_plain.log 2>&1 || { echo plain failed; tail -5 gpurun_out/prof_spectral_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:"istft" -s 2 -c 1 -o gpurun_out/prof_istft_seeded -f python scripts/prof_spectral.py > gpurun_out/ncu_istft.log 2>&1; echo "istft seeded ncu exit $?"
ncu --set full --clock-control none --import-source on -k regex:"istft" -s 5 -c 1 -o gpurun_out/prof_istft_phasor -f python scripts/prof_spectral.py > gpurun_out/ncu_istft2.log 2>&1; echo "istft phasor ncu exit $?"
ls -la gpurun_out/*.ncu-rep
