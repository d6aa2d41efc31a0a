"""Summarise an .ncu-rep (read on the CPU box): duration, pipe utilisation, issue rate, stall reasons, DRAM bytes.
Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [launch_index]"""
import csv
import io
import re
import subprocess
import sys


def main(path, idx=0):
    raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2 + idx]
    get = {h: (data[i], units[i]) for i, h in enumerate(hdr)}
    print('kernel', get['Kernel Name'][0], 'grid', get['Grid Size'][0], 'block', get['Block Size'][0])
    keys = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
            'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps', 'sm__warps_active.avg.pct_of_peak_sustained_active',
            'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
            'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
            'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
            'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
            'smsp__inst_executed.sum', 'sm__cycles_elapsed.avg', 'smsp__cycles_active.avg', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
            'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
            'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'sm__cycles_active.avg']
    for k in keys:
        if k in get:
            print('%-70s %s %s' % (k, get[k][0], get[k][1]))
    st = [(float(v[0]), h) for h, v in get.items() if re.match(r'smsp__pcsamp_warps_issue_stalled_\w+$', h) and not h.endswith('not_issued')]
    tot = sum(x for x, _ in st) or 1.
    print('stall samples (pc sampling):')
    for x, h in sorted(st, reverse=True)[:10]:
        print('   %-60s %8.0f  %5.1f%%' % (h.replace('smsp__pcsamp_warps_issue_stalled_', ''), x, 100 * x / tot))


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
