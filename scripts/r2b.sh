mkdir -p gpurun_out
for cfg in "64 1024" "64 2048" "32 1024" "32 2048" "32 8192" "48 1024" "16 8192"; do
  set -- $cfg
  echo "== L=$1 R=$2"
  TUNE_L=$1 TUNE_R=$2 TUNE_SPARSE=0 TUNE_BETAS=1.0 TUNE_NWS=2 TUNE_VARIANTS=0 timeout 300 python scripts/tune_poisson.py 2>&1 | tail -2
done > gpurun_out/r2b_scaling.txt 2>&1
cat gpurun_out/r2b_scaling.txt
