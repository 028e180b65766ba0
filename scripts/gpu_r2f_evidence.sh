#!/bin/bash
# Round-2 final evidence: tests of the changed kernels, the ncu launch list of the bench command (variant B, the default workload),
# ncu --set full of the 17 tensor-core conv launches of one forward (halo / dx / upm) and of the final spectral kernels.
# Every ncu command runs only after the same command has exited 0 without ncu.
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --train-steps 0 --c2-clips 0"
timeout 300 $CMD > gpurun_out/ncu_plain_B.json 2> gpurun_out/ncu_plain_B.err || { echo "plain run failed"; tail -5 gpurun_out/ncu_plain_B.err; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_B.csv $CMD > gpurun_out/ncu_launches_B.log 2>&1
echo "launch list exit $?"; wc -l gpurun_out/launches_B.csv
timeout 900 ncu --set full --clock-control none -k "regex:conv3x3_(halo|dx|upm)" -s 34 -c 17 -o gpurun_out/prof_conv_B -f $CMD > gpurun_out/ncu_conv_B.log 2>&1
echo "conv full exit $?"
python scripts/prof_spectral.py > gpurun_out/prof_spectral_plain.log 2>&1 || { echo plain failed; tail -5 gpurun_out/prof_spectral_plain.log; exit 1; }
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"stft|istft" -s 2 -c 1 -o gpurun_out/prof_stft_final -f python scripts/prof_spectral.py > gpurun_out/ncu_stft.log 2>&1; echo "stft ncu exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"istft" -s 2 -c 1 -o gpurun_out/prof_istft_seeded_final -f python scripts/prof_spectral.py > gpurun_out/ncu_istft.log 2>&1; echo "istft seeded ncu exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"istft" -s 5 -c 1 -o gpurun_out/prof_istft_phasor_final -f python scripts/prof_spectral.py > gpurun_out/ncu_istft2.log 2>&1; echo "istft phasor ncu exit $?"
# summaries are made on the box: the reports together exceed what gpurun copies back (64 MiB)
python scripts/ncu_summary.py gpurun_out/prof_conv_B.ncu-rep gpurun_out/r2f_conv_full_variantB.csv > /dev/null
for k in stft_final istft_seeded_final istft_phasor_final; do
  python scripts/ncu_stalls.py gpurun_out/prof_$k.ncu-rep > gpurun_out/r2f_${k}_stalls.txt 2>&1
  ncu -i gpurun_out/prof_$k.ncu-rep --page source --csv --print-source sass > gpurun_out/src_$k.csv 2> /dev/null && python scripts/sass_hist.py gpurun_out/src_$k.csv > gpurun_out/r2f_${k}_sass_hist.txt 2>&1
  rm -f gpurun_out/src_$k.csv
done
ls -la gpurun_out/*.ncu-rep
rm -f gpurun_out/prof_conv_B.ncu-rep gpurun_out/prof_istft_phasor_final.ncu-rep gpurun_out/prof_istft_seeded_final.ncu-rep
ls -la gpurun_out/
