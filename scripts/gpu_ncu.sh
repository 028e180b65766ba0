#!/bin/bash
# ncu evidence for bench.py (variant R keeps replays short): launch list + full sections of the conv and STFT/iSTFT kernels.
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --variant R --no-cpu-baseline"
$CMD > gpurun_out/ncu_plain.json 2> gpurun_out/ncu_plain.err || { echo "plain run failed"; tail -5 gpurun_out/ncu_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
ncu --set full --clock-control none --import-source on -k regex:conv_gemm -s 63 -c 21 -o gpurun_out/prof_conv -f $CMD > gpurun_out/ncu_conv.log 2>&1
echo "conv full exit $?"
ncu --set full --clock-control none --import-source on -k regex:"stft_kernel|istft_kernel|conv3x3_c1|maxpool" -s 6 -c 8 -o gpurun_out/prof_misc -f $CMD > gpurun_out/ncu_misc.log 2>&1
echo "misc full exit $?"
ls -la gpurun_out
