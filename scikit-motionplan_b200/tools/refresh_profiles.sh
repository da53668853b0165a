#!/bin/bash
# Regenerates the raw material of profiles/ on a GPU box (run through gpurun from the repo root); everything lands in
# gpurun_out/prof_*.  Post-processing (ncu_summary.py / ncu_sections.py) happens where the .ncu-rep files are read.
#   bash scikit-motionplan_b200/tools/refresh_profiles.sh
set -u
export SKB_S2_VERBOSE=1
O=gpurun_out
BK="python scikit-motionplan_b200/tools/bench_kernels.py"
# 1. contract bench line and the reference arm, no profiler
python bench.py --steps 10 --warmup 3 > $O/prof_bench.json 2> $O/prof_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/prof_bench_reference_arm.json 2> $O/prof_bench_reference_arm.err
# the named 10 M configurations on ONE GPU (92 GB of outputs), and the other named robots with their CPU baselines
python bench.py --configs-per-gpu 10000000 --steps 5 --warmup 3 --no-cpu-baseline > $O/prof_bench_10m.json 2> $O/prof_bench_10m.err
for w in cfg2 cfg4 cfg5; do python bench.py --workload $w --steps 10 --warmup 3 > $O/prof_bench_$w.json 2> $O/prof_bench_$w.err; done
$BK --reps 8 > $O/prof_kernels.jsonl 2> $O/prof_kernels.err
# 2. launch list of the bench command
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/prof_launches.csv \
    python bench.py --steps 2 --warmup 3 --configs-per-gpu 262144 > $O/prof_launches.log 2>&1
# 3. full captures of the first launch inside the NVTX range "timed" (after warm-up and autotuning): bench.py's timed
#    step, then the other kernels
# (gpurun merges at most 64 MiB back and every report embeds the 23 MB cubin: export the two pages that are read, drop the report)
export_pages() { ncu -i $O/$1.ncu-rep --page raw --csv > $O/$1_raw.csv 2>/dev/null; ncu -i $O/$1.ncu-rep --page source --csv > $O/$1_src.csv 2>/dev/null; rm -f $O/$1.ncu-rep; }
ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "timed/" -k regex:s2_kernel -c 1 -o $O/prof_cfg3 -f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/prof_cfg3.log 2>&1
export_pages prof_cfg3
for w in cfg2_all:s2_kernel cfg5_env_all:s2_kernel cfg5_self_valid:s2_kernel cfg4_409600:s2_kernel cfg5_pose:pose_kernel cfg5_com:com_kernel; do
    name=${w%%:*}; kern=${w##*:}
    ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "timed/" -k regex:$kern -c 1 -o $O/prof_$name -f \
        $BK --only $name --reps 2 > $O/prof_$name.log 2>&1
    export_pages prof_$name
done
ls -la $O/prof_* | awk '{print $5, $9}'
