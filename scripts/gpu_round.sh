#!/bin/bash
mkdir -p gpurun_out
{
echo "=== full gpu test suite"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "=== bench N=1"; timeout 900 python bench.py 2>&1 | tail -1 > gpurun_out/bench_line_N1.json; cut -c1-600 gpurun_out/bench_line_N1.json
echo "=== bench reference arm"; timeout 900 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | cut -c1-500
echo "=== ncu launch list of the bench"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1; tail -3 gpurun_out/launches_bench.csv | cut -c1-200
echo "=== ncu full of the bench launch"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:rp_solve_kernel -s 6 -c 2 -o gpurun_out/solver_bench_launch python bench.py --steps 2 --warmup 3 --no-cpu-baseline 2>&1 | tail -2
} > gpurun_out/round_final1.log 2>&1
tail -40 gpurun_out/round_final1.log
