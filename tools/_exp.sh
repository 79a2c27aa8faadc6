N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29501 bench.py --gpus $N --steps 1000 --warmup 20 > gpurun_out/bench_gpus$N.json 2> gpurun_out/bench_gpus$N.err
python -c "
import json; d=json.load(open('gpurun_out/bench_gpus$N.json')); print('bench gpus', d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['value'])"
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/slab_bench.py --steps 20 > gpurun_out/slab_$N.log 2>&1
tail -3 gpurun_out/slab_$N.log | cut -c1-900
