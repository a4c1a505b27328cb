#!/bin/bash
# Evidence run for profiles/: bench line, ncu of the solver bench launch, SCNet launch list, halo per-role table.
TAG=${1:-r2}
mkdir -p gpurun_out
{
echo "=== all gpu tests"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "=== bench"; timeout 1500 python bench.py 2>&1 | tail -1 > gpurun_out/${TAG}_bench_line_N1.json; python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench_line_N1.json')); print(d['value'], d['e2e']['value'], d['per_pair_p50_ms'], d['clocks'])"
echo "=== ncu solver"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:rp_solve_kernel -s 6 -c 1 -o gpurun_out/${TAG}_solver_bench_launch python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline > /dev/null 2>&1; ls -la gpurun_out/${TAG}_solver_bench_launch.ncu-rep
echo "=== launch list of the bench"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${TAG}_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-extra --no-cpu-baseline > /dev/null 2>&1; python scripts/ncu_launch_table.py gpurun_out/${TAG}_launches_bench.csv | head -6
echo "=== scnet launch list P=32"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${TAG}_scnet_launches_P32.csv python scripts/prof_scnet.py 32 > /dev/null 2>&1; python scripts/ncu_launch_table.py gpurun_out/${TAG}_scnet_launches_P32.csv | head -14
echo "=== halo roles"; RP_SCNET_HALO_FLAGS=34 timeout 300 python scripts/prof_halo_layers.py 32 3 2>&1 | tail -16
echo "=== ncu halo layers"; timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_halo_tc -o gpurun_out/${TAG}_halo_layers python scripts/prof_halo_layers.py 32 1 > /dev/null 2>&1; ls -la gpurun_out/${TAG}_halo_layers.ncu-rep
echo "=== ncu wide solver kernel, one pair"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:rp_solve_kernel -s 8 -c 1 -o gpurun_out/${TAG}_solver_wide_one_pair python scripts/phase_clk.py > /dev/null 2>&1; ls -la gpurun_out/${TAG}_solver_wide_one_pair.ncu-rep
echo "=== phase clocks"; timeout 300 python scripts/phase_clk.py 2>&1 | tail -4
echo "=== wide vs narrow"; timeout 600 python scripts/time_wide.py 2>&1 | tail -16
echo "=== launch list of one alternation call (no graph)"; RP_SCNET_GRAPH=0 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_alternation_launches.csv python scripts/prof_alternation.py > /dev/null 2>&1; python scripts/ncu_launch_table.py gpurun_out/${TAG}_alternation_launches.csv | head -24
} > gpurun_out/round_final_$TAG.log 2>&1
tail -60 gpurun_out/round_final_$TAG.log
