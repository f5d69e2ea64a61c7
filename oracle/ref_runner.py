"""
TEST / BENCH INFRASTRUCTURE (never imported by the product): drive the REAL
reference -- the unmodified openquake.hazardlib installed under
baseline/_ref (pip install --no-deps --target baseline/_ref /root/reference;
git-ignored, it travels to the GPU box with the snapshot) -- through its own
hot path on the synthetic workloads of bench.py:

    planar.build_planar -> contexts._get_ctx_planar (distances, filter)
    -> ContextMaker.update (get_mean_stds -> gen_poes -> MapArray.update_indep)

i.e. what RmapMaker._make_src_indep does for the planar ruptures of point
sources (contexts.py:443-474, 1232-1247, 1548-1579).  h5py / shapely are
absent from the image: oracle/refshim stubs them (stubs raise on use; this
path uses neither).  `available()` says whether the install is there.
"""
import os
import sys
import time
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.path.join(ROOT, 'baseline', '_ref')
_mods = {}


def available():
    return os.path.isdir(os.path.join(REF, 'openquake', 'hazardlib'))


def load():
    """import the reference from baseline/_ref (once per process)"""
    if _mods:
        return _mods
    os.environ['OQ_REFERENCE'] = REF
    os.environ.setdefault('NUMBA_CACHE_DIR', '/tmp/numba_cache_ref')
    sys.path.insert(0, os.path.join(HERE, 'refshim'))
    import oq_bootstrap
    oq_bootstrap.REFERENCE = REF
    oq_bootstrap.install()
    from openquake.hazardlib import valid
    from openquake.hazardlib.contexts import ContextMaker, _get_ctx_planar
    from openquake.baselib.general import RecordBuilder
    from openquake.hazardlib.calc.filters import IntegrationDistance
    from openquake.hazardlib.geo.surface import planar as P
    from openquake.hazardlib.site import SiteCollection
    from openquake.hazardlib.map_array import MapArray
    from openquake.hazardlib.tom import PoissonTOM
    import openquake.hazardlib as hl
    # (inside the test-suite the reference may have been imported from
    # /root/reference already: the same files)
    _mods['path'] = os.path.dirname(os.path.realpath(hl.__file__))
    _mods.update(valid=valid, ContextMaker=ContextMaker,
                 _get_ctx_planar=_get_ctx_planar, RecordBuilder=RecordBuilder,
                 IntegrationDistance=IntegrationDistance, P=P,
                 SiteCollection=SiteCollection, MapArray=MapArray,
                 PoissonTOM=PoissonTOM)
    return _mods


def build_planar(r, usd=0., lsd=25., rar=1.5):
    """the reference's build_planar, one rupture at a time -> (U, 3)"""
    P = load()['P']
    out = []
    for u in range(len(r['mag'])):
        planin = np.zeros((1, 1), P.planin_dt).view(np.recarray)
        for k in ('mag', 'strike', 'dip', 'rake', 'rate', 'area'):
            planin[k][:] = r[k][u]
        hdd = np.array([[1.0, r['hypo_dep'][u]]])
        pla = P.build_planar(planin, hdd, r['hypo_lon'][u], r['hypo_lat'][u],
                             usd, lsd, rar)
        out.append(pla.reshape(-1, 3))
    return np.concatenate(out).view(np.recarray)


def sitecol(sites):
    m = load()

    class SM:
        pass
    sm = SM()
    for k in ('vs30', 'vs30measured', 'z1pt0', 'z2pt5'):
        setattr(sm, k, sites[k])
    return m['SiteCollection'].from_points(
        sites['lon'], sites['lat'], sites['depth'], sm,
        req_site_params=('vs30', 'vs30measured', 'z1pt0', 'z2pt5'))


def hazard_pass(names, imtls, planar, mags, sites, tl=3., itime=50.,
                maxdist='1000'):
    """
    One pass of the reference over planar ruptures x sites: returns (rows,
    seconds of get_ctx + update, the float32 pne map).
    """
    m = load()
    gsims = [m['valid'].gsim('[%s]' % n) for n in names]
    cm = m['ContextMaker']('*', gsims, dict(
        imtls=imtls, truncation_level=tl, investigation_time=itime,
        maximum_distance=m['IntegrationDistance'].new(maxdist)))
    tom = m['PoissonTOM'](itime)
    sc = sitecol(sites)
    dd = cm.defaultdict.copy()
    dd['probs_occur'] = np.zeros(0)
    builder = m['RecordBuilder'](**dd)
    cm.siteparams = [par for par in sc.array.dtype.names if par in dd]
    cm.ruptparams = cm.REQUIRES_RUPTURE_PARAMETERS | {'occurrence_rate'}
    cm.fewsites = False
    cm.tom = tom
    L = sum(len(v) for v in imtls.values())
    pmap = m['MapArray'](sc.sids, L, len(gsims)).fill(1)
    t0 = time.time()
    rows = 0
    for magi, mag in enumerate(np.unique(mags)):
        pla = planar[mags == mag]
        ctx = m['_get_ctx_planar'](cm, builder, mag, 1.0, magi, pla, sc, 0, 0,
                                   tom)
        ctx = ctx[ctx.rrup < cm.maximum_distance(mag)]
        if len(ctx):
            rows += len(ctx)
            cm.update(pmap, ctx)
    return rows, time.time() - t0, pmap.array
