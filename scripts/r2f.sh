mkdir -p gpurun_out
TUNE_SPARSE=0 TUNE_BETAS=${TB:-1.0} TUNE_NWS=${TN:-2} TUNE_VARIANTS=${TV:-8192} timeout 600 python scripts/tune_poisson.py > gpurun_out/r2f_prof.txt 2>&1
tail -2 gpurun_out/r2f_prof.txt
