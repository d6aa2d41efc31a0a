"""Pack the reference's bundled *data* (not source) into compact npz files.

The reference ships 60-odd two-column transmission-curve text files
(lightcurve_fitting/filters/*), the SiFTO template (models/sifto.dat) and the
example light curve (example/SN2016bkv.txt).  SURVEY.md section 2 marks them
DATA, reused verbatim as read-only input.  /root/reference does not exist on
the GPU box, so the numbers travel inside the package as npz archives
(packed filter-bank format instead of per-file ASCII
parsing).  Run once in the authoring container:

    python tools/pack_reference_data.py [/root/reference]

The raw (wavelength, transmission) rows are stored exactly as they appear in
the files; unit conversion, sorting and normalisation happen at load time in
lightcurve_fitting_b200/filters.py (mirrors filters.py:181-214).
"""
import os
import sys
import numpy as np


def read_two_columns(path):
    """Two numeric columns, whitespace or comma separated; a first line that is
    not numeric is a header (astropy's ascii guesser behaves like this for
    these files)."""
    rows = []
    with open(path) as fh:
        for line in fh:
            line = line.strip()
            if not line or line.startswith('#'):
                continue
            parts = line.replace(',', ' ').split()
            try:
                rows.append((float(parts[0]), float(parts[1])))
            except (ValueError, IndexError):
                if rows:
                    raise
                continue  # header
    return np.array(rows, dtype=np.float64)


def main(ref='/root/reference'):
    pkg = os.path.join(ref, 'lightcurve_fitting')
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'lightcurve_fitting_b200', 'data')
    os.makedirs(out, exist_ok=True)

    curves = {}
    fdir = os.path.join(pkg, 'filters')
    for name in sorted(os.listdir(fdir)):
        if name.endswith(('.fits', '.py')):
            continue
        try:
            curves[name] = read_two_columns(os.path.join(fdir, name))
        except Exception as exc:  # not a curve file
            print('skip', name, exc)
    np.savez_compressed(os.path.join(out, 'filter_curves.npz'), **curves)
    print('packed', len(curves), 'curves,', sum(len(v) for v in curves.values()), 'rows')

    # SiFTO template: "#  Epoch U B V g r i" header, 106 rows
    sifto = np.loadtxt(os.path.join(pkg, 'models', 'sifto.dat'))
    np.savez_compressed(os.path.join(out, 'sifto.npz'), table=sifto,
                        columns=np.array(['Epoch', 'U', 'B', 'V', 'g', 'r', 'i']))
    print('sifto', sifto.shape)

    # Example light curve (fixed-width table with a dashed rule on line 2)
    mjd, mag, dmag, filt, source, nondet = [], [], [], [], [], []
    with open(os.path.join(pkg, 'example', 'SN2016bkv.txt')) as fh:
        lines = fh.read().splitlines()
    header, rule = lines[0], lines[1]
    # column extents from the dashed rule
    spans, start = [], None
    for i, ch in enumerate(rule + ' '):
        if ch == '-' and start is None:
            start = i
        elif ch != '-' and start is not None:
            spans.append((start, i))
            start = None
    names = [header[a:b].strip() for a, b in spans]
    assert names == ['MJD', 'mag', 'dmag', 'filter', 'source', 'nondet'], names
    for line in lines[2:]:
        if not line.strip():
            continue
        cells = [line[a:b].strip() for a, b in spans]
        mjd.append(float(cells[0]))
        mag.append(float(cells[1]))
        dmag.append(float(cells[2]) if cells[2] not in ('', '--') else np.nan)
        filt.append(cells[3])
        source.append(cells[4])
        nondet.append(cells[5] == 'True')
    np.savez_compressed(os.path.join(out, 'SN2016bkv.npz'), MJD=np.array(mjd), mag=np.array(mag),
                        dmag=np.array(dmag), filter=np.array(filt), source=np.array(source),
                        nondet=np.array(nondet))
    print('SN2016bkv', len(mjd), 'rows', sorted(set(filt)))


if __name__ == '__main__':
    main(*sys.argv[1:])
