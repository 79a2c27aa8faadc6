for v in a168 b168 b144 b128; do DFB_LIB=$PWD/digital-filtering_b200/lib/$v.so python tools/stepbench.py; done
for v in a168 b168; do DFB_Z_BLOCKS_PER_SM=2 DFB_LIB=$PWD/digital-filtering_b200/lib/$v.so python tools/stepbench.py; done
