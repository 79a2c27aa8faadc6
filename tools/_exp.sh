timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/stepbench.py
DFB_DEBUG_Z=16 python tools/zprof.py 1024x2048_profile_N128
DFB_DEBUG_Z=16 python tools/zprof.py 1024x2048_saturated_N128
