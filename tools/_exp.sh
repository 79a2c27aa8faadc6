timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/stepbench.py
DFB_Y_MODE=1 python tools/stepbench.py 1024x2048_profile_N128 4096x8192_profile_N128
python tools/stepbench.py 4096x8192_profile_N128 4096x8192_saturated_N128
python tools/quick_gpu.py 1024x2048_saturated_N128 2>&1 | grep "variant 0" | cut -c1-330
