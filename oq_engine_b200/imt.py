"""Intensity measure types, the part of openquake/hazardlib/imt.py the hot
path needs: ``from_string`` (imt.py:89) giving objects with ``.string`` and
``.period``."""
import re
import collections

IMT = collections.namedtuple('IMT', 'string period damping')
_SA = re.compile(r'SA\(([^)]+)\)')


def from_string(imt):
    """
    >>> from_string('SA(0.2)').period
    0.2
    >>> from_string('PGA').period
    0.0
    """
    if isinstance(imt, IMT):
        return imt
    mo = _SA.fullmatch(imt)
    if mo:
        period = float(mo.group(1))
        return IMT('SA(%s)' % period, period, 5.0)
    if imt in ('PGA', 'PGV', 'CAV', 'IA'):
        return IMT(imt, 0.0, 5.0)
    raise KeyError('Unknown or unsupported IMT: %r' % (imt,))
