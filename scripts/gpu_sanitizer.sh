#!/bin/bash
# One compute-sanitizer tool (racecheck | synccheck | memcheck) over small-shape tests of the mbarrier / TMA / tcgen05 kernels.
# usage: bash scripts/gpu_sanitizer.sh <tool>     (one tool per gpurun call, see /opt/skills/guides/B200_PROFILING.md)
set -u
TOOL=${1:-synccheck}
mkdir -p gpurun_out
SEL='test_stft_mag_matches_oracle or test_istft_matches_oracle or test_round_trip_full_size or test_conv3x3_bn_relu or test_conv3x3_64_output_channels_all_kernels or test_convt2x2 or test_conv3x3_wgrad_matches_autograd or test_convt2x2_backward'
timeout 900 python -m pytest tests/test_gpu_spectral.py tests/test_gpu_unet.py tests/test_gpu_train_ops.py -m gpu -x -q -k "$SEL" > gpurun_out/sanitizer_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/sanitizer_plain.log; exit 1; }
tail -1 gpurun_out/sanitizer_plain.log
timeout 1500 compute-sanitizer --tool $TOOL --print-limit 20 python -m pytest tests/test_gpu_spectral.py tests/test_gpu_unet.py tests/test_gpu_train_ops.py -m gpu -x -q -k "$SEL" > gpurun_out/sanitizer_$TOOL.log 2>&1; echo "$TOOL exit $?"
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|hazard|Barrier error|Invalid" gpurun_out/sanitizer_$TOOL.log | sort | uniq -c | sort -rn | head -20
