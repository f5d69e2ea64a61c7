"""Coefficient tables of the supported ground-motion models.

The numbers are the published regression coefficients (data/coeffs.json,
extracted from the reference's CoeffsTable objects by tools/extract_coeffs.py).
``CoeffsTable[imt]`` mirrors openquake/hazardlib/gsim/coeffs_table.py:231-286:
exact rows for tabulated IMTs, linear interpolation in log-period between the
neighbouring SA rows otherwise, KeyError outside the tabulated range.
"""
import os
import json
import math
from .imt import from_string

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'data',
                     'coeffs.json')
_RAW = None


def _raw():
    global _RAW
    if _RAW is None:
        with open(_DATA) as f:
            _RAW = json.load(f)
    return _RAW


class CoeffsTable(object):
    """A table IMT -> {column: value}"""
    _cache = {}

    def __init__(self, name):
        raw = _raw()[name]
        self.name = name
        self.columns = raw['columns']
        self.non_sa = raw['non_sa']
        self.sa = raw['sa']   # rows [period, v0, v1, ...] sorted by period
        self._rows = {}

    @classmethod
    def get(cls, name):
        if name not in cls._cache:
            cls._cache[name] = cls(name)
        return cls._cache[name]

    def __getitem__(self, imt):
        imt = from_string(imt)
        try:
            return self._rows[imt.string]
        except KeyError:
            pass
        if not imt.string.startswith('SA'):
            if imt.string not in self.non_sa:
                raise KeyError(imt.string)
            row = dict(zip(self.columns, self.non_sa[imt.string]))
        else:
            row = self._sa_row(imt)
        self._rows[imt.string] = row
        return row

    def _sa_row(self, imt):
        lo = hi = None
        for r in self.sa:
            if r[0] == imt.period:
                return dict(zip(self.columns, r[1:]))
            if r[0] < imt.period:
                lo = r
            elif hi is None:
                hi = r
        if lo is None or hi is None:
            raise KeyError(imt.string)
        w = ((math.log(imt.period) - math.log(lo[0])) /
             (math.log(hi[0]) - math.log(lo[0])))
        return {c: (b - a) * w + a
                for c, a, b in zip(self.columns, lo[1:], hi[1:])}
