set -e
python bench.py --steps 1000 --warmup 20 > gpurun_out/bench_r01c_default.json
python bench.py --workload 1024x2048_saturated_N128 --steps 1000 --warmup 20 --no-cpu --no-sweep > gpurun_out/bench_r01c_saturated.json
python bench.py --workload 4096x8192_saturated_N128 --steps 100 --warmup 5 --no-cpu --no-sweep > gpurun_out/bench_r01c_4096x8192_saturated.json
for f in default saturated 4096x8192_saturated; do python -c "
import json,sys; d=json.load(open('gpurun_out/bench_r01c_$f.json')); k=d['roofline']['kernels']; print('$f', round(d['ms_per_step'],4), '%.3g'%d['value'], d['roofline']['kernel'], d['roofline']['bound'], round(d['roofline']['frac'],3), json.dumps(k['ysweep_tma_kernel'])[:300])"; done
