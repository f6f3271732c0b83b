#!/bin/bash
# round-2 evidence pass: full GPU suite, smoke, bench both arms, ncu launch list + full capture of the bench launch
tag=${1:-r2h}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${tag}_gpu.txt
( time timeout 1700 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/${tag}_pytest.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.txt 2>&1
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2>> gpurun_out/${tag}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --cache-control none --import-source on -k regex:k_checkerboard_flow -s 3 -c 1 -o gpurun_out/${tag}_bench_flow python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_full.log 2>&1
tail -4 gpurun_out/${tag}_pytest.txt; tail -2 gpurun_out/${tag}_smoke.txt
cut -c1-400 gpurun_out/${tag}_bench.json; cut -c1-300 gpurun_out/${tag}_bench_reference.json; tail -3 gpurun_out/${tag}_bench.err; tail -2 gpurun_out/${tag}_ncu_full.log
