#!/bin/sh
# Final profiling pass of round 2 on one B200 (run under gpurun from the repository root); everything lands in gpurun_out/.
# A number printed by a run under ncu is never a bench value: the timing runs come first, without ncu.
set -x
O=gpurun_out
mkdir -p $O
# 1. the default bench line (the command the driver runs at N=1)
S=$(date +%s)
timeout 500 python bench.py > $O/r02c_bench_n1.json 2> $O/r02c_bench_n1.err
echo "default bench wall seconds: $(( $(date +%s) - S ))" | tee $O/r02c_bench_n1_wall.txt
# 2. launch list of the bench command (cold-cache, serialised: shares, not absolutes)
timeout 200 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > $O/r02c_bench_short.json 2> $O/r02c_bench_short.err && \
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r02c_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > $O/r02c_ncu_bench.log 2>&1
# 3. full capture of the dominant kernel (one whole cfg4 solve) as it is at the end of the round
timeout 300 ncu --set full --clock-control none -k regex:solve_persistent_kernel -s 1 -c 1 -f -o $O/r02c_cfg4_persistent \
    python tools/prof_sweep.py --workload cfg4 --reps 1 --solves 2 > $O/r02c_ncu_cfg4.log 2>&1
# the report stays on the box (gpurun_out is capped at 64 MiB): keep its text / csv pages
ncu -i $O/r02c_cfg4_persistent.ncu-rep --page details > $O/r02c_cfg4_persistent_details.txt 2>&1
ncu -i $O/r02c_cfg4_persistent.ncu-rep --page raw --csv > $O/r02c_cfg4_persistent_raw.csv 2>&1
rm -f $O/*.ncu-rep
ls -la $O | grep r02c
