timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/stepbench.py
DFB_Z_BLOCKS_PER_SM=3 python tools/stepbench.py
DFB_DEBUG_Z=16 python tools/zprof.py 1024x2048_profile_N128
python tools/timeline.py | tail -2
