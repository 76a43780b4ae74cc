# End-of-round evidence (one gpurun call): GPU tests, the default bench line, the other workloads, the per-GEMM tables and the
# ncu launch list.  `bash tools/final_run.sh [ncu]` -- with "ncu" also the --set full capture of the hot kernels.
set -x
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/r5_gputest_summary.txt
timeout 900 python bench.py > gpurun_out/r5_bench_line.json 2> gpurun_out/r5_bench.err
timeout 300 python bench.py --model pointnetpp_msg --no-extras > gpurun_out/r5_pointnetpp_msg_bench_line.json 2>> gpurun_out/r5_bench.err
timeout 300 python bench.py --model pointnext --points 24000 --batch 8 --no-extras > gpurun_out/r5_pointnext_24k_bench_line.json 2>> gpurun_out/r5_bench.err
timeout 300 python tools/gemm_shapes.py > gpurun_out/r5_gemm_shapes.md 2>&1
timeout 300 python tools/gemm_shapes.py --no-planes > gpurun_out/r5_gemm_shapes_noplanes.md 2>&1
timeout 300 python tools/gemm3x_shapes.py --trace > gpurun_out/r5_gemm3x_wait_trace.md 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r5_dgcnn_launches.csv python bench.py --no-graph --no-extras --no-cpu-baseline --steps 2 --warmup 3 > gpurun_out/r5_ncu_launch.log 2>&1
if [ "$1" = "ncu" ]; then timeout 900 bash tools/ncu_capture.sh r5 'gemm2h|knn_tc' dgcnn gemm; fi
tail -2 gpurun_out/r5_gputest_summary.txt
