#!/bin/bash
# Round-2 evidence captures (run on the GPU box, one GPU): ncu --set full of the kernels that changed late in the
# round + the launch list of the bench command.  Summaries: tools/ncu_summary.py, tools/launch_list_summary.py.
set -u
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled"
$NCU -k regex:'tile_sort_pack|preprocess_kernel|tile_scan|scatter_kernel|render_fwd|render_bwd|preprocess_bwd' -s 33 -c 7 \
  -o gpurun_out/r2_raster -f python tools/prof_raster.py 6 > gpurun_out/r2_ev_raster.log 2>&1
$NCU -k regex:layer_gemm -s 23 -c 6 -o gpurun_out/r2_mlp_layer_res -f python tools/mlp_bench.py --profile > gpurun_out/r2_ev_mlp.log 2>&1
$NCU -k regex:'mc_count|mc_emit|mc_resolve|mc_backward|nearest_kernel' -s 5 -c 5 -o gpurun_out/r2_mc_bricks -f \
  python tools/aux_profile2.py > gpurun_out/r2_ev_mc.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_bench_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-train-step --no-cpu-baseline > gpurun_out/r2_ev_bench.log 2>&1
ls -la gpurun_out/r2_raster.ncu-rep gpurun_out/r2_mlp_layer_res.ncu-rep gpurun_out/r2_mc_bricks.ncu-rep gpurun_out/r2_bench_launches.csv
