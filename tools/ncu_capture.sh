#!/bin/bash
# ncu --set full over tools/ncu_ops.py; exports the raw page (all kernels) and the source page (SASS + stall samples) of the
# kernels named in $2 as CSV, then drops the .ncu-rep (gpurun copies back at most 64 MiB).   usage: ncu_capture.sh <tag> "<kernel regex for source pages>" [groups...]
tag=$1; src=$2; shift 2
K='^(ball_|bn_|csr_|edgeconv|fps_|gemm|grid_|group_|interp|knn_|segsum|select_|sumsq|pool_|maxpool|split_|absmax)'
ncu --set full --clock-control none --import-source on -k regex:"$K" -o /tmp/$tag python tools/ncu_ops.py "$@" > gpurun_out/${tag}_ncu.log 2>&1 || exit 1
ncu -i /tmp/$tag.ncu-rep --page raw --csv > gpurun_out/${tag}_raw.csv 2>/dev/null
ncu -i /tmp/$tag.ncu-rep --page source --csv -k regex:"$src" 2>/dev/null | gzip > gpurun_out/${tag}_source.csv.gz
ls -la gpurun_out/${tag}_*
