for v in n4_1_128 n8_1_128 n16_1_64 n8_1_64 n4_1_256 n8_2_128 n16_1_128 n4_1_128; do DFB_LIB=$PWD/digital-filtering_b200/lib/$v.so python tools/stepbench.py; done
DFB_LIB=$PWD/digital-filtering_b200/lib/n8_1_128.so python tools/timeline.py | tail -2
