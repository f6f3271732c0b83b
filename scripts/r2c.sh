mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_checkerboard.py tests/test_gpu_baseline_sizes.py -x -q 2>&1 | tail -15 > gpurun_out/r2c_pytest.txt
tail -6 gpurun_out/r2c_pytest.txt
TUNE_SPARSE=0 TUNE_BETAS=1.0,1.5,0.75 TUNE_NWS=1,2,4 TUNE_VARIANTS=0,4096 timeout 600 python scripts/tune_poisson.py > gpurun_out/r2c_tune.txt 2>&1
cat gpurun_out/r2c_tune.txt
