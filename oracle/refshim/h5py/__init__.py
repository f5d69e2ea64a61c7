"""Minimal h5py stand-in so that the reference's hazardlib imports in a
container without HDF5 (TEST INFRASTRUCTURE ONLY - used by the fixture
generators under oracle/ that import /root/reference; never by the product).

Only the names touched at import time are provided
(openquake/baselib/hdf5.py:43-49,334 and
openquake/hazardlib/source/rupture.py:61-72); any attempt at file I/O raises.
"""
import numpy

__version__ = '0.0-shim'


def special_dtype(vlen=None, **kw):
    return numpy.dtype(object)


def check_dtype(**kw):
    return None


class File:
    def __init__(self, *a, **k):
        raise RuntimeError('h5py shim: file I/O is not available')


class Group:
    pass


class Dataset:
    pass
