timeout 800 python -m pytest tests/test_gpu_prior_train.py -q -m gpu -s > gpurun_out/r2p_prior_train_tests.txt 2>&1; tail -4 gpurun_out/r2p_prior_train_tests.txt
timeout 300 python profiles/prior_train_bench.py > gpurun_out/r2p_prior_train_bench.txt 2> gpurun_out/r2p_prior_train_bench.err; cat gpurun_out/r2p_prior_train_bench.txt; tail -3 gpurun_out/r2p_prior_train_bench.err
