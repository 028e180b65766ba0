#!/bin/bash
# ncu --set full with source-level sampling of ONE conv launch of a forward: $1 = kernel regex, $2 = launches to skip
set -u
mkdir -p gpurun_out
python scripts/prof_forward_once.py > gpurun_out/fwd_plain.log 2>&1 || { echo plain failed; tail -5 gpurun_out/fwd_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k "regex:$1" -s $2 -c 1 -o gpurun_out/prof_layer -f python scripts/prof_forward_once.py > gpurun_out/ncu_layer.log 2>&1; echo "ncu exit $?"
python scripts/ncu_stalls.py gpurun_out/prof_layer.ncu-rep > gpurun_out/layer_stalls.txt 2>&1; cat gpurun_out/layer_stalls.txt
ncu -i gpurun_out/prof_layer.ncu-rep --page source --csv --print-source cuda > gpurun_out/layer_src.csv 2>/dev/null
python scripts/src_lines.py gpurun_out/layer_src.csv 30 > gpurun_out/layer_src_top.txt 2>&1; cat gpurun_out/layer_src_top.txt
ncu -i gpurun_out/prof_layer.ncu-rep --page source --csv --print-source sass > gpurun_out/layer_sass.csv 2>/dev/null
python scripts/src_lines.py gpurun_out/layer_sass.csv 25 3 > gpurun_out/layer_sass_top.txt 2>&1; cat gpurun_out/layer_sass_top.txt
head -c 3000 gpurun_out/layer_src.csv > gpurun_out/layer_src_head.txt
rm -f gpurun_out/layer_src.csv gpurun_out/layer_sass.csv gpurun_out/prof_layer.ncu-rep
