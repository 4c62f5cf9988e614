set -e
python bench.py --steps 2 --warmup 3 --no-cpu --no-ref-cuda --no-configs --no-train > gpurun_out/r2k_plain.json 2> gpurun_out/r2k_plain.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_final.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-ref-cuda --no-configs --no-train > gpurun_out/r2k_ncu.log 2>&1
