"""Conditional spectrum: oqg_cs_out through
cond_spectra / get_cs_out against the reference's cond_spectra
(oracle/make_golden_cs.py), and the Baker & Jayaram (2008) correlation
against the reference's values."""
import os
import numpy as np
import pytest

from helpers import GOLDEN

D = np.load(os.path.join(GOLDEN, 'cond_spectra.npz'))
IMTS = [str(x) for x in D['imts']]
SRC = [(0.10, 0.10, [0.02, 0.008, 0.003, 0.001], 30., 60., 90., 6.),
       (0.45, -0.20, [0.05, 0.01, 0.002], 120., 90., 0., 8.),
       (-0.30, 0.35, [0.01, 0.006, 0.004, 0.002, 0.001], 250., 45., -90., 10.)]


def test_baker_jayaram_matches_reference():
    import oq_engine_b200 as oq
    model = oq.BakerJayaram2008()
    rho = np.array([[model.rho(a, b) for b in IMTS] for a in IMTS])
    np.testing.assert_allclose(rho, D['rho'], rtol=1e-14)
    assert (np.diag(rho) == 1).all() and np.allclose(rho, rho.T)


@pytest.mark.gpu
def test_cond_spectra_golden():
    import oq_engine_b200 as oq
    oq._lib.lib()
    srcs = [oq.PointSource(
        str(i), lon, lat, oq.EvenlyDiscretizedMFD(5.05, 0.3, rates), 'WC1994',
        1.5, 0., 20., [(1., strike, dip, rake)], [(1., dep)])
        for i, (lon, lat, rates, strike, dip, rake, dep) in enumerate(SRC)]
    sc = oq.SiteCollection.from_points([0.05, 0.30], [0.0, 0.25], [0., 0.],
                                       dict(vs30=500.))
    cm = oq.ContextMaker(
        'Active Shallow Crust',
        [oq.get_gsim('BooreEtAl2014'), oq.get_gsim('BooreAtkinson2008')],
        dict(imtls={imt: [0.05, 0.2] for imt in IMTS}, truncation_level=3.,
             investigation_time=50.,
             maximum_distance=oq.IntegrationDistance.new('100')))
    ctxs = []
    for i, src in enumerate(srcs):
        ctxs.extend(cm.get_ctxs(src, sc, src_id=i))
    ctx = cm.recarray(ctxs)
    spectra, s_sigma = oq.cond_spectra(
        cm, ctx, sc, 'SA(0.2)', D['imls'], oq.PoissonTOM(50.), poes=D['poes'],
        correlation_model=oq.BakerJayaram2008())
    assert spectra.shape == D['spectra'].shape == (2, len(IMTS), 2, 2)
    np.testing.assert_allclose(spectra, D['spectra'], rtol=1e-9)
    np.testing.assert_allclose(s_sigma, D['s_sigma'], rtol=1e-8, atol=1e-12)
    # the spectrum at the conditioning period is the conditioning level
    k = IMTS.index('SA(0.2)')
    np.testing.assert_allclose(spectra[:, k], np.broadcast_to(
        D['imls'], (2, 2, 2)), rtol=1e-9)
    with pytest.raises(ValueError):
        oq.get_cs_out(cm, ctx, k, D['imls'], oq.PoissonTOM(50.),
                      poes=D['poes'])
