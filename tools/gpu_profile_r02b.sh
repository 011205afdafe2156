#!/bin/sh
# Second profiling pass of round 2 on one B200 (run under gpurun from the repository root); everything lands in gpurun_out/.
# A number printed by a run under ncu is never a bench value: the timing runs come first, without ncu.
set -x
O=gpurun_out
mkdir -p $O
# 1. phase / step traces (in-kernel %globaltimer stamps), no profiler
ALLL_TRACE=1 python tools/prof_sweep.py --workload cfg4 --reps 3 --solves 2 > $O/r02b_trace_cfg4.json 2> $O/r02b_trace_cfg4.txt
ALLL_TRACE=1 python tools/prof_sweep.py --workload cfg2 --reps 3 --solves 2 > $O/r02b_trace_cfg2.json 2> $O/r02b_trace_cfg2.txt
ALLL_TRACE=1 python tools/prof_sweep.py --workload cfg3 --reps 3 --solves 2 --max-rounds 40 > $O/r02b_trace_cfg3.json 2> $O/r02b_trace_cfg3.txt
# 2. launch list of the bench command (cold-cache, serialised: shares, not absolutes)
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > $O/r02b_bench_short.json 2> $O/r02b_bench_short.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r02b_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > $O/r02b_ncu_bench.log 2>&1
# 3. full captures of the top kernels
ncu --set full --clock-control none -k regex:solve_persistent_kernel -s 1 -c 1 -f -o $O/r02b_cfg4_persistent \
    python tools/prof_sweep.py --workload cfg4 --reps 1 --solves 2 > $O/r02b_ncu_cfg4.log 2>&1
ncu --set full --clock-control none -k regex:solve_persistent_kernel -s 1 -c 1 -f -o $O/r02b_cfg2_persistent \
    python tools/prof_sweep.py --workload cfg2 --reps 1 --solves 2 > $O/r02b_ncu_cfg2.log 2>&1
ncu --set full --clock-control none -k regex:solve_persistent_kernel -s 1 -c 1 -f -o $O/r02b_cfg3_persistent \
    python tools/prof_sweep.py --workload cfg3 --reps 1 --solves 2 --max-rounds 40 > $O/r02b_ncu_cfg3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sweep_planes_kernel -s 3 -c 1 -f -o $O/r02b_sweep \
    python tools/prof_sweep.py --workload cfg4 --reps 6 > $O/r02b_ncu_sweep.log 2>&1
# the reports stay on the box (gpurun_out is capped at 64 MiB): keep their text / csv pages
for r in r02b_cfg4_persistent r02b_cfg2_persistent r02b_cfg3_persistent r02b_sweep; do
    ncu -i $O/$r.ncu-rep --page details > $O/${r}_details.txt 2>&1
    ncu -i $O/$r.ncu-rep --page raw --csv > $O/${r}_raw.csv 2>&1
done
rm -f $O/*.ncu-rep
ls -la $O | grep r02b
