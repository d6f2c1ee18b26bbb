set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in 1 2 4 8; do
  timeout 120 $TR --nproc-per-node $n --master-port $((29600+n)) profiles/d2h_probe.py > gpurun_out/r2_d2h_n$n.txt 2> gpurun_out/r2_d2h_n$n.err
  tail -1 gpurun_out/r2_d2h_n$n.txt
done
timeout 200 python bench.py --workload train --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_train_n1.json 2> gpurun_out/r2_train_n1.err; tail -c 600 gpurun_out/r2_train_n1.json
for n in 2 8; do
  timeout 300 $TR --nproc-per-node $n --master-port $((29700+n)) bench.py --workload train --gpus $n --steps 20 --warmup 5 > gpurun_out/r2_train_n$n.json 2> gpurun_out/r2_train_n$n.err
  tail -c 900 gpurun_out/r2_train_n$n.json; tail -3 gpurun_out/r2_train_n$n.err
done
timeout 300 $TR --nproc-per-node 8 --master-port 29790 bench.py --workload train --gpus 8 --steps 20 --warmup 5 --no-graph > gpurun_out/r2_train_n8_eager.json 2> gpurun_out/r2_train_n8_eager.err; tail -c 400 gpurun_out/r2_train_n8_eager.json
for n in 2 4; do
  timeout 300 $TR --nproc-per-node $n --master-port $((29800+n)) bench.py --gpus $n --clips-total 512 --steps 5 --warmup 3 > gpurun_out/r2_strong_n$n.json 2> gpurun_out/r2_strong_n$n.err
  python -c "
import json,sys
d=json.loads(open('gpurun_out/r2_strong_n$n.json').read().strip().splitlines()[-1]); print($n, d['scaling'], d['value'], d['ms_per_step'], d['e2e'])"; tail -2 gpurun_out/r2_strong_n$n.err
done
