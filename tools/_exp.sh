for v in e010_168 e110_184 e111_184 e110_200 e111_200 e010_184 e010_168; do DFB_LIB=$PWD/digital-filtering_b200/lib/$v.so python tools/stepbench.py; done
