timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python tools/stepbench.py
python tools/stepbench.py 4096x8192_profile_N128 4096x8192_saturated_N128
DFB_Y_MODE=1 python tools/stepbench.py 4096x8192_profile_N128
DFB_Y_MODE=0 python tools/stepbench.py 4096x8192_saturated_N128
