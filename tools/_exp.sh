python bench.py --workload 1024x2048_saturated_N128 --steps 1000 --warmup 20 --no-cpu --no-sweep > gpurun_out/bench_r01c_saturated.json
python bench.py --workload 4096x8192_saturated_N128 --steps 100 --warmup 5 --no-cpu --no-sweep > gpurun_out/bench_r01c_4096x8192_saturated.json
python bench.py --workload 4096x8192_profile_N128 --steps 100 --warmup 5 --no-cpu --no-sweep > gpurun_out/bench_r01c_4096x8192_profile.json
python bench.py --workload 512x512_N32 --steps 2000 --warmup 20 --no-cpu --no-sweep > gpurun_out/bench_r01c_512x512.json
for f in saturated 4096x8192_saturated 4096x8192_profile 512x512; do python -c "
import json,sys; d=json.load(open('gpurun_out/bench_r01c_$f.json')); k=d['roofline']['kernels']; print('$f', round(d['ms_per_step'],4), '%.3g'%d['value'], 'e2e %.3g'%d['e2e']['value'], d['roofline']['kernel'], round(d['roofline']['frac'],3), 'y', round(k['ysweep_tma_kernel']['ms'],4), round(k['ysweep_tma_kernel']['frac_fp64'],3), 'z', round(k['zsweep_epilogue_kernel']['ms'],4), round(k['zsweep_epilogue_kernel']['frac_hbm'],3), 'noise', round(k['noise_kernel']['ms'],4), 'eqTF', round(d['roofline']['step']['equivalent_tflops'],1))"; done
DFB_DEBUG_Z=16 python tools/zprof.py 1024x2048_profile_N128
DFB_DEBUG_Z=16 python tools/zprof.py 1024x2048_saturated_N128
python tools/timeline.py | tail -2
python tools/timeline.py 1024x2048_saturated_N128 | tail -2
