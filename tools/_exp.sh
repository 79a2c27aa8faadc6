timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python tools/stepbench.py 512x512_N32 1024x2048_profile_N128
python tools/timeline.py default | tail -1
