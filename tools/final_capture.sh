#!/bin/bash
# Round-2 final single-GPU capture: launch list, ncu --set full of the two pair kernels, bench lines, tests, smoke.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== launch list (ncu, cold-cache, serialised)"
SPH_BENCH_SKIP_EXTRA=1 SPH_BENCH_CPU_BUDGET_S=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 300 --csv --log-file gpurun_out/r02_final_launches.csv python bench.py --steps 6 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1; tail -2 gpurun_out/bench_under_ncu.log | cut -c1-200
echo "== ncu --set full, pair kernels"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_density_tma|k_force_packed' -s 200 -c 2 -f -o gpurun_out/prof_r02_final python tools/profile_step.py --warm 100 --steps 2 2>&1 | tail -2
echo "== bench --impl reference"; timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_final_bench_reference_n1.json 2> gpurun_out/r02_final_bench_reference_n1.err; cut -c1-300 gpurun_out/r02_final_bench_reference_n1.json
echo "== bench N=1"; timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_final_bench_n1.json 2> gpurun_out/r02_final_bench_n1.err; tail -c 300 gpurun_out/r02_final_bench_n1.err; cut -c1-400 gpurun_out/r02_final_bench_n1.json
echo "== bench N=1, 200 steps / 50 warm-up"; SPH_BENCH_SKIP_EXTRA=1 SPH_BENCH_CPU_BUDGET_S=2 timeout 900 python bench.py --steps 200 --warmup 50 > gpurun_out/r02_final_bench_n1_200.json 2>/dev/null; cut -c1-400 gpurun_out/r02_final_bench_n1_200.json
echo "== developed flow"; timeout 300 python tools/sweep_variants.py --pairs 1:1 --scene dragon_bath --warm 400 2>&1 | grep -v Warning
} > gpurun_out/final_capture.log 2>&1
tail -40 gpurun_out/final_capture.log
