"""List the FP64 inner loop (planck_quad_f64, ShockCooling3 32-walker plain kernel) of the built library with its FP64-pipe
instruction count per Planck sample.  Usage: python tools/sass_fp64_loop.py > profiles/round2_sass_fp64_loop.txt"""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    lib = os.path.join(ROOT, 'lightcurve_fitting_b200', 'liblcf_b200.so')
    t = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
    for part in re.split(r'\n\s*Function : ', t)[1:]:
        name = part.split('\n', 1)[0]
        if not name.startswith('_ZN3lcf6k_passILi3EdLi5ELb1ELi2E'):
            continue
        ins = []
        for line in part.split('\n'):
            m = re.search(r'/\*([0-9a-f]{4,5})\*/\s+(\S.*?);', line)
            if m:
                ins.append((m.group(1), m.group(2).strip()))
        addr = {a: k for k, (a, _) in enumerate(ins)}
        best = None
        for k, (a, i) in enumerate(ins):
            m = re.search(r'BRA\s+0x([0-9a-f]+)', i)
            if not m:
                continue
            tgt = m.group(1).rjust(len(a), '0')
            if tgt in addr and addr[tgt] < k:
                body = ins[addr[tgt]:k + 1]
                n64 = sum(1 for _, x in body if re.match(r'(@!?P\d\s+)?D(FMA|MUL|ADD)', x))
                if len(body) < 400 and any('RCP64H' in x for _, x in body) and (best is None or n64 > best[1]):   # innermost loops only
                    best = (body, n64)
        body, n64 = best
        print('# FP64 inner loop of lcf::k_pass<3, double, 5, true> (planck_quad_f64<TAB>, two pair records = 8 Planck samples per iteration)')
        print('# cuobjdump -sass of lightcurve_fitting_b200/liblcf_b200.so (tools/sass_fp64_loop.py)')
        print('# instructions per iteration: %d; FP64-pipe (DFMA/DMUL/DADD): %d = %.2f per Planck sample; MUFU.RCP64H: %d; LDS: %d; integer/other: %d'
              % (len(body), n64, n64 / 8., sum(1 for _, x in body if 'RCP64H' in x), sum(1 for _, x in body if 'LDS' in x),
                 len(body) - n64 - sum(1 for _, x in body if 'RCP64H' in x or 'LDS' in x)))
        for a, x in body:
            print('/*%s*/  %s' % (a, x))
        return


if __name__ == '__main__':
    main()
