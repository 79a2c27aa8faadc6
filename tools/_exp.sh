python tools/stepbench.py
DFB_Y_PERSIST=444 python tools/stepbench.py
DFB_Y_PERSIST=296 python tools/stepbench.py
DFB_Y_PERSIST=444 python tools/timeline.py | tail -2
DFB_Y_PERSIST=444 timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
