mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_checkerboard.py -x -q -k "tma or full_size" 2>&1 | tail -5
TUNE_SPARSE=0 TUNE_BETAS=${TB:-1.0,1.5,0.75} TUNE_NWS=${TN:-1,2} TUNE_VARIANTS=${TV:-0,4096} timeout 600 python scripts/tune_poisson.py > gpurun_out/r2d_tune.txt 2>&1
cat gpurun_out/r2d_tune.txt
