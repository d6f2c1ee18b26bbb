# train-mode regularisers: GPU parity tests of the training step + the configs[4] bench in both modes
timeout 600 python -m pytest tests/test_gpu_train.py -q -s > gpurun_out/r2u_train_tests.txt 2>&1; echo tests rc=$?; grep -E "passed|failed|Error|error|assert|TRAIN-mode" gpurun_out/r2u_train_tests.txt | tail -25
timeout 200 python bench.py --workload train --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/r2u_bench_train_draw.json 2> gpurun_out/r2u_bench_train_draw.err; echo rc=$?; tail -3 gpurun_out/r2u_bench_train_draw.err
timeout 200 python bench.py --workload train --steps 30 --warmup 3 --no-cpu-baseline --regularisers off > gpurun_out/r2u_bench_train_off.json 2> gpurun_out/r2u_bench_train_off.err; echo rc=$?; tail -3 gpurun_out/r2u_bench_train_off.err
python - <<'PY'
import json
for f in ['draw','off']:
    try:
        d=json.loads(open(f'gpurun_out/r2u_bench_train_{f}.json').read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step'],3), round(d['value'],1), d['e2e'], d['loss'], d['clocks'])
    except Exception as e: print(f, 'ERR', e)
PY
