# device-side draws (Philox) + graph-stable LayerDrop: GPU parity tests of the training step + the configs[4] bench in train mode
timeout 600 python -m pytest tests/test_gpu_train.py -q -s > gpurun_out/r2v_train_tests.txt 2>&1; echo tests rc=$?; grep -E "passed|failed|Error|error|assert|TRAIN-mode" gpurun_out/r2v_train_tests.txt | tail -25
timeout 200 python bench.py --workload train --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/r2v_bench_train_draw.json 2> gpurun_out/r2v_bench_train_draw.err; echo rc=$?; tail -3 gpurun_out/r2v_bench_train_draw.err
timeout 200 python bench.py --workload train --steps 30 --warmup 3 --no-cpu-baseline --clips 16 > gpurun_out/r2v_bench_train_draw_c16.json 2> gpurun_out/r2v_bench_train_draw_c16.err; echo rc=$?; tail -3 gpurun_out/r2v_bench_train_draw_c16.err
timeout 200 python bench.py --workload train --steps 30 --warmup 3 --no-cpu-baseline --clips 16 --regularisers off > gpurun_out/r2v_bench_train_off_c16.json 2> gpurun_out/r2v_bench_train_off_c16.err; echo rc=$?
python - <<'PY'
import json
for f in ['draw','draw_c16','off_c16']:
    try:
        d=json.loads(open(f'gpurun_out/r2v_bench_train_{f}.json').read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step'],3), round(d['value'],1), d['e2e'], d['loss'], d['gpu_launches'], d['clocks'])
    except Exception as e: print(f, 'ERR', e)
PY
