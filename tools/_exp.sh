timeout 900 ncu --set full --import-source on --clock-control none -k regex:"noise_kernel|ysweep_tma|zsweep_epilogue" -s 9 -c 3 -o gpurun_out/full_r01c -f python bench.py --steps 3 --warmup 3 --no-cpu --no-sweep > gpurun_out/ncu_f.log 2>&1
tail -1 gpurun_out/ncu_f.log | cut -c1-120
