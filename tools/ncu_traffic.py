"""profiles/round2_ncu_traffic.json from the ncu summaries of the three cfg2 kernels (tools/ncu_round2.sh, tools/ncu_summary.py):
DRAM bytes per launch = dram__bytes_read.sum + dram__bytes_write.sum of the one captured launch.  bench.py reads the file for
`roofline.traffic`.  Usage: python tools/ncu_traffic.py [directory with round2_ncu_{sc3,sc4,fp64}.txt, default profiles/]"""
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {'byte': 1., 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}


def main(src):
    out = {}
    for tag, key in (('sc3', 'lcf::k_pass<3,float,5,true,4>'), ('sc4', 'lcf::k_pass<4,float,5,true,2>'), ('fp64', 'lcf::k_pass<3,double,5,true,2>')):
        path = os.path.join(src, 'round2_ncu_%s.txt' % tag)
        if not os.path.exists(path):
            continue
        t = open(path).read()
        val = {}
        for name in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
            m = re.search(re.escape(name) + r'\s+([0-9.]+)\s+(\w+)', t)
            val[name] = float(m.group(1)) * UNIT[m.group(2)]
        dur = re.search(r'gpu__time_duration.sum\s+([0-9.]+ \w+)', t).group(1)
        out[key + '@cfg2'] = {
            'dram_bytes_per_launch': int(val['dram__bytes_read.sum'] + val['dram__bytes_write.sum']),
            'dram_bytes_read': int(val['dram__bytes_read.sum']), 'dram_bytes_write': int(val['dram__bytes_write.sum']), 'ncu_duration': dur,
            'source': 'ncu --set full --clock-control none, one launch of bench.py --steps 4 --warmup 3 (tools/ncu_round2.sh); profiles/round2_ncu_%s.txt' % tag,
            'note': 'algorithmic walker-state traffic of a half-step: 50 000 walkers x (2 position rows read + 1 written + log-probabilities) ~ 9 MB; '
                    'the chain step (3.6 MB) was still in L2 when the capture ended'}
    json.dump(out, open(os.path.join(ROOT, 'profiles', 'round2_ncu_traffic.json'), 'w'), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == '__main__':
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'profiles'))
