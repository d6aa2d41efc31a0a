#!/bin/bash
# ncu evidence of this round's kernels (run under gpurun; B200_PROFILING.md recipe).  The .ncu-rep files stay in /tmp on the GPU box
# (gpurun_out/ is capped at 64 MiB): what comes back is the text summary of each capture (tools/ncu_summary.py), the per-instruction
# source page as gzip'ed CSV, and the launch list.
set -x
B="python bench.py --steps 4 --warmup 3 --no-cpu --no-extras"
cap() {   # name, kernel regex, launch-skip, command...
  name=$1; rx=$2; skip=$3; shift 3
  "$@" > gpurun_out/round2_plain_$name.log 2>&1 && \
  ncu --set full --import-source on --clock-control none -k regex:$rx --launch-skip $skip --launch-count 1 -f -o /tmp/prof_$name "$@" > gpurun_out/round2_ncu_$name.log 2>&1
  python tools/ncu_summary.py /tmp/prof_$name.ncu-rep > gpurun_out/round2_ncu_$name.txt 2>&1
  ncu -i /tmp/prof_$name.ncu-rep --page source --csv 2>/dev/null | gzip -9 > gpurun_out/round2_src_$name.csv.gz
}
$B > gpurun_out/round2_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/round2_launches.csv $B > gpurun_out/round2_ncu_l.log 2>&1
cap sc3 k_pass 8 $B
cap sc4 k_pass 8 $B --model sc4
cap fp64 k_pass 8 $B --precision fp64
cap cfg1_ring k_ring 1 python tools/bench_configs.py --only cfg1 --short
cap cfg4 k_pass 60 python tools/bench_configs.py --only cfg4 --short
cap cfg3_chain k_chain 1 python tools/bench_configs.py --only cfg3
ls -la gpurun_out/ | head -40
