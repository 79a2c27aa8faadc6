set -e
python tools/quick_gpu.py 1024x2048_saturated_N128 > gpurun_out/q0.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:zsweep_epilogue -s 3 -c 1 -o gpurun_out/z_sat -f python tools/quick_gpu.py 1024x2048_saturated_N128 > gpurun_out/ncu_z.log 2>&1
tail -3 gpurun_out/ncu_z.log
