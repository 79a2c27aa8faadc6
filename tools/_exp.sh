for i in 1 2; do
python tools/timeline.py default | tail -1
DFB_Y_TK=64 python tools/timeline.py default | tail -1
DFB_Y_TK=128 python tools/timeline.py default | tail -1
done
