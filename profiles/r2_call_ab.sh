# 2 GPUs: the default predict bench with two batches in flight per rank, and the training step in TRAIN mode (Philox draws + NCCL inside the graph)
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2ab_bench_n2.json 2> gpurun_out/r2ab_bench_n2.err; echo predict rc=$?; tail -3 gpurun_out/r2ab_bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload train --steps 30 --warmup 3 > gpurun_out/r2ab_bench_train_n2.json 2> gpurun_out/r2ab_bench_train_n2.err; echo train rc=$?; tail -3 gpurun_out/r2ab_bench_train_n2.err
python - <<'PY'
import json
for f in ['bench_n2','bench_train_n2']:
    try:
        d=json.loads(open(f'gpurun_out/r2ab_{f}.json').read().strip().splitlines()[-1])
        print(f, d['n_gpus'], round(d['ms_per_step'],3), round(d['value'],1), d['e2e'], d.get('ranks_in_sync'), d['clocks'])
    except Exception as e: print(f, 'ERR', e)
PY
