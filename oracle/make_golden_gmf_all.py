"""
Golden fixture for the sampling side of the ground motion fields (SURVEY.md
section 8f row 4): the reference's GmfComputer.compute_all
(openquake/hazardlib/calc/gmf.py:394-439) run here on one planar rupture,
two gsims with their realizations, three IMTs (one MMI-free set), with

  'plain'    no correlation model, truncation 3
  'jb2009'   JB2009CorrelationModel(vs30_clustering=False) on the within-event
             residuals, truncation 3 (apply_correlation, Cholesky factor)
  'minmax'   minimum_intensity / a maximum level (set_max_min)

The fixture keeps the inputs the device path needs (the context row of the
rupture on the sites as the reference built it, the site coordinates, the
realizations, the seed and the PCG64 state numpy derives from it) and the
reference's outputs: the data frame columns, sig / eps per event and the first
uniforms of the generator.

    python oracle/make_golden_gmf_all.py   -> tests/golden/gmf_all.npz
"""
import os
import sys
from unittest.mock import Mock
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, 'refshim'))
sys.path.insert(0, HERE)
import oq_bootstrap  # noqa
oq_bootstrap.install()

from openquake.hazardlib import valid, contexts, site, geo  # noqa
from openquake.hazardlib.source.rupture import EBRupture, build_planar  # noqa
from openquake.hazardlib.calc.gmf import GmfComputer  # noqa
from openquake.hazardlib.correlation import JB2009CorrelationModel  # noqa

GOLDEN = os.path.join(HERE, '..', 'tests', 'golden')
IMTS = ['PGA', 'SA(0.3)', 'SA(1.0)']


def run(tag, out, corr=None, min_iml=None, max_iml=None, n_occ=12, seed=42):
    rng = np.random.default_rng(5)
    N = 40
    lons = rng.uniform(-0.3, 0.3, N)
    lats = rng.uniform(0.2, 0.8, N)
    siteparams = Mock(reference_vs30_value=760.)
    sitecol = site.SiteCollection.from_points(lons, lats,
                                              sitemodel=siteparams)
    rlzs_by_gsim = {valid.gsim('BooreAtkinson2008'): np.array([0, 2], 'u4'),
                    valid.gsim('AkkarEtAlRjb2014'): np.array([1], 'u4')}
    hypo = geo.point.Point(0, .5, 10)
    rup = build_planar(hypo, mag=6.5, rake=0.)
    cmaker = contexts.simple_cmaker(rlzs_by_gsim, IMTS, truncation_level=3.)
    cmaker.oq.calculation_mode = 'scenario'   # rlzs spread evenly
    if min_iml is not None:
        cmaker.min_iml = np.array(min_iml)
    if max_iml is not None:
        cmaker.oq.get_max_iml = lambda: np.array(max_iml)
    ebr = EBRupture(rup, 0, 0, n_occ=n_occ, id=1, seed=seed)
    gc = GmfComputer(ebr, sitecol, cmaker, correlation_model=corr)
    df = gc.compute_all()
    ctx = gc.ctx
    for name in ctx.dtype.names:
        out['%s_ctx_%s' % (tag, name)] = np.asarray(ctx[name])
    out[tag + '_lons'], out[tag + '_lats'] = lons, lats
    for g, gs in enumerate(cmaker.gsims):
        out['%s_ms_%d' % (tag, g)] = cmaker.get_4MN([ctx], gs).astype(
            np.float32)
        out['%s_rlzs_%d' % (tag, g)] = cmaker.gsims[gs]
    out[tag + '_eid'], out[tag + '_rlz'] = gc.eid, gc.rlz
    out[tag + '_sig'], out[tag + '_eps'] = gc.sig, gc.between_eps
    for col in df.columns:
        out['%s_df_%s' % (tag, col)] = df[col].to_numpy()
    out[tag + '_min_iml'] = cmaker.min_iml
    out[tag + '_max_iml'] = cmaker.oq.get_max_iml()
    st = np.random.PCG64(seed).state['state']
    out[tag + '_seed'] = seed
    out[tag + '_uniform'] = np.random.default_rng(seed).uniform(size=64)
    out[tag + '_state'] = np.array(
        [st['state'] >> 64, st['state'] & (2 ** 64 - 1), st['inc'] >> 64,
         st['inc'] & (2 ** 64 - 1)], np.uint64)
    print(tag, len(df), list(df.columns), float(df['PGA'].mean())
          if 'PGA' in df else list(df.columns))


def main():
    out = dict(imts=np.array(IMTS))
    run('plain', out)
    run('jb2009', out, corr=JB2009CorrelationModel(vs30_clustering=False))
    run('minmax', out, min_iml=[0.05, 0.1, 0.03], max_iml=[0.4, 1.0, 0.3],
        seed=7)
    np.savez_compressed(os.path.join(GOLDEN, 'gmf_all.npz'), **out)


if __name__ == '__main__':
    main()
