DFB_LIB=$PWD/digital-filtering_b200/lib/prev.so python tools/stepbench.py
python tools/stepbench.py
python tools/quick_gpu.py 1024x2048_profile_N128 2>&1 | grep "variant 0" | cut -c1-330
DFB_LIB=$PWD/digital-filtering_b200/lib/prev.so python tools/quick_gpu.py 1024x2048_profile_N128 2>&1 | grep "variant 0" | cut -c1-330
