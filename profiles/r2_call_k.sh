set -x
timeout 400 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "gemm" > gpurun_out/r2k_gemm_tests.txt 2>&1; tail -3 gpurun_out/r2k_gemm_tests.txt
timeout 150 python profiles/gemm_shapes.py > gpurun_out/r2k_shapes_mc.txt 2>&1; cat gpurun_out/r2k_shapes_mc.txt
AVI_GEMM_NO_MULTICAST=1 timeout 150 python profiles/gemm_shapes.py > gpurun_out/r2k_shapes_nomc.txt 2>&1; cat gpurun_out/r2k_shapes_nomc.txt
timeout 250 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err; tail -c 1500 gpurun_out/r2k_bench.json; tail -5 gpurun_out/r2k_bench.err
