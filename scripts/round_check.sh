#!/bin/bash
# One GPU-box pass over everything a round records: the -m gpu suite, smoke(), the bench (both arms), the β sweep of
# BASELINE configs[1], the extremal_opt and config-5 parallel-tempering benches. Outputs under gpurun_out/<tag>_*.
tag=${1:-check}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${tag}_pytest.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.txt 2>&1
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python bench.py --impl reference > gpurun_out/${tag}_bench_reference.json 2>> gpurun_out/${tag}_bench.err
: > gpurun_out/${tag}_beta_sweep.jsonl
for b in 0.5 0.75 1.25 1.5 2.0; do
  python bench.py --beta $b --steps 10 --warmup 3 --no-cpu-baseline >> gpurun_out/${tag}_beta_sweep.jsonl 2>> gpurun_out/${tag}_bench.err
done
timeout 250 python scripts/bench_configs.py eo > gpurun_out/${tag}_configs_eo.jsonl 2> gpurun_out/${tag}_eo.err
RRG_R=4096 timeout 250 python scripts/bench_configs.py eo --quick >> gpurun_out/${tag}_configs_eo.jsonl 2>> gpurun_out/${tag}_eo.err
timeout 300 python scripts/bench_c5_pt.py > gpurun_out/${tag}_c5_pt_n1.json 2> gpurun_out/${tag}_c5_pt.err
tail -4 gpurun_out/${tag}_pytest.txt; cat gpurun_out/${tag}_smoke.txt | tail -2
cut -c1-260 gpurun_out/${tag}_bench.json; cut -c1-200 gpurun_out/${tag}_beta_sweep.jsonl
cat gpurun_out/${tag}_configs_eo.jsonl gpurun_out/${tag}_c5_pt_n1.json; for f in eo c5_pt bench; do tail -n 3 gpurun_out/${tag}_$f.err; done; true
