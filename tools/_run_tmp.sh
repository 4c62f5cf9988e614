set -e
python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err
python tools/sweep.py --cases cfg2:I:0 --iters 3 > gpurun_out/r2j_sweep.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:fwd_tile_kernel|bwd_warp_kernel" -s 4 -c 2 -o gpurun_out/r2_final_fwd_bwd -f python tools/sweep.py --cases cfg2:I:0 --iters 3 > gpurun_out/r2j_ncu.log 2>&1
