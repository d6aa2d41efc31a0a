"""Dict-of-columns stand-in for astropy.table.Table (the calls made by reference filters.py / models.py)."""
import numpy as np
from .units import Quantity, dimensionless_unscaled


class Column(Quantity):
    pass


class MaskedColumn(Column):
    pass


class Table:
    def __init__(self, cols=None, meta=None):
        self._cols = {}
        self.meta = dict(meta or {})
        for k, v in (cols or {}).items():
            self[k] = v

    @classmethod
    def read(cls, filename, format='ascii', names=None, **kw):
        rows, header = [], None
        with open(filename) as fh:
            for line in fh:
                line = line.strip()
                if not line:
                    continue
                if line.startswith('#'):
                    if header is None:
                        header = line.lstrip('#').split()
                    continue
                parts = line.replace(',', ' ').split()
                try:
                    rows.append([float(p) for p in parts])
                except ValueError:
                    if rows:
                        raise
                    header = parts
        arr = np.array(rows, float)
        if names is None:
            names = header
        return cls({n: arr[:, i] for i, n in enumerate(names)})

    @property
    def colnames(self):
        return list(self._cols)

    def __len__(self):
        return len(next(iter(self._cols.values()))) if self._cols else 0

    def __setitem__(self, key, value):
        unit = getattr(value, 'unit', dimensionless_unscaled)
        self._cols[key] = Column(np.array(value, float), unit)

    def __getitem__(self, key):
        if isinstance(key, str):
            return self._cols[key]
        t = Table(meta=self.meta)
        for k, v in self._cols.items():
            c = Column(np.asarray(v)[key], v.unit)
            t._cols[k] = c
        return t

    def sort(self, key):
        idx = np.argsort(np.asarray(self._cols[key]), kind='stable')
        for k, v in list(self._cols.items()):
            self._cols[k] = Column(np.asarray(v)[idx], v.unit)


def vstack(tables, **kw):
    raise NotImplementedError
