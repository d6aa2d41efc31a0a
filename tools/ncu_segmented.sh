set -x
C="python tools/bench_segmented_bank.py 8192 2"
$C > gpurun_out/round2_plain_seg.log 2>&1 && \
ncu --set full --import-source on --clock-control none -k regex:k_pass_seg --launch-skip 6 --launch-count 1 -f -o /tmp/prof_seg $C > gpurun_out/round2_ncu_seg.log 2>&1
python tools/ncu_summary.py /tmp/prof_seg.ncu-rep > gpurun_out/round2_ncu_seg.txt 2>&1
tail -40 gpurun_out/round2_ncu_seg.txt
