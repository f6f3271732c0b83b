TUNE_SPARSE=0 TUNE_BETAS=${TB:-1.0} TUNE_NWS=${TN:-2} TUNE_VARIANTS=${TV:-8192} timeout 600 python scripts/tune_poisson.py 2>&1 | tail -40
