mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_gpu.txt
timeout 1500 python -m pytest tests/test_gpu_checkerboard.py tests/test_gpu_interface.py tests/test_gpu_baseline_sizes.py -x -q 2>&1 | tail -25 > gpurun_out/r2a_pytest.txt
tail -5 gpurun_out/r2a_pytest.txt
TUNE_SPARSE=0 TUNE_BETAS=1.0,0.75,0.5,1.5,2.0 TUNE_NWS=1,2,4,6 TUNE_VARIANTS=0,1,2048 timeout 600 python scripts/tune_poisson.py > gpurun_out/r2a_tune.txt 2>&1
cat gpurun_out/r2a_tune.txt
timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k regex:k_checkerboard_tma -s 20 -c 2 -o gpurun_out/r2a_tma_warm python scripts/prof_cb.py 1.0 poisson 20 > gpurun_out/r2a_ncu.log 2>&1
tail -3 gpurun_out/r2a_ncu.log
