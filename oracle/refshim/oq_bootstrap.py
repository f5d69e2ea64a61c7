"""Make `import openquake.hazardlib` work in this container (TEST
INFRASTRUCTURE ONLY).

The reference is pure Python + numba but imports h5py/shapely/pyproj/... at
module level.  None of the planar-rupture hot path uses them, so they are
replaced by stubs that *raise on use* (a permissive stub silently corrupts
polygon geometry, so every attribute access on a stubbed object is an error).

Usage (only from scripts under oracle/ that generate fixtures):

    import oq_bootstrap; oq_bootstrap.install()
    from openquake.hazardlib import contexts
"""
import os
import sys
import types
import importlib.abc
import importlib.machinery

REFERENCE = os.environ.get('OQ_REFERENCE', '/root/reference')
STUBBED = ('shapely', 'pyproj', 'h3', 'alpha_shapes', 'matplotlib', 'osgeo',
           'fiona', 'geopandas', 'django', 'docutils', 'onnxruntime',
           'setproctitle')


# shapely.LineString(...).is_simple is only asked by Polygon.__init__ as a
# sanity check of module-level constants (e.g. gsim/morikawa_fujiwara_2013.py:94)
HARMLESS = {'is_simple': True}


class _Instance:
    """Result of calling a stubbed class/function: inert until touched"""
    def __init__(self, name):
        object.__setattr__(self, '_name', name)

    def __getattr__(self, n):
        if n in HARMLESS:  # validation-only queries made at import time
            return HARMLESS[n]
        raise RuntimeError('stubbed dependency used: %s().%s' % (
            object.__getattribute__(self, '_name'), n))

    def __bool__(self):
        raise RuntimeError('stubbed dependency used: bool(%s())' %
                           object.__getattribute__(self, '_name'))

    __iter__ = __len__ = __call__ = __bool__


class _Poison:
    """Placeholder name: importing/constructing is fine (some modules build
    constants at import time), using the result is an error"""
    def __init__(self, name):
        object.__setattr__(self, '_name', name)

    def __call__(self, *a, **k):
        return _Instance(self._name)

    def __getattr__(self, n):
        if n.startswith('__'):
            raise AttributeError(n)
        return _Poison(self._name + '.' + n)

    def __mro_entries__(self, bases):
        return (object,)


class _StubModule(types.ModuleType):
    __path__ = []

    def __getattr__(self, n):
        if n.startswith('__'):
            raise AttributeError(n)
        return _Poison(self.__name__ + '.' + n)


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, name, path, target=None):
        if name.split('.')[0] in STUBBED:
            return importlib.machinery.ModuleSpec(name, self, is_package=True)

    def create_module(self, spec):
        return _StubModule(spec.name)

    def exec_module(self, module):
        pass


def available():
    return os.path.isdir(os.path.join(REFERENCE, 'openquake', 'hazardlib'))


def install():
    if not available():
        raise RuntimeError('reference tree not found at %s' % REFERENCE)
    here = os.path.dirname(os.path.abspath(__file__))
    os.environ.setdefault('NUMBA_CACHE_DIR', '/tmp/oq_numba_cache')
    for p in (here, REFERENCE):
        if p not in sys.path:
            sys.path.insert(0, p)
    if not any(isinstance(f, _Finder) for f in sys.meta_path):
        sys.meta_path.insert(0, _Finder())
