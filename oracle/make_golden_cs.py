"""
Golden fixture for the conditional spectrum:
the reference's cond_spectra / get_cs_out / _cs_out
(openquake/hazardlib/calc/cond_spectra.py:35-169) with the BakerJayaram2008
cross-IMT correlation, on three point sources and two sites.

    python oracle/make_golden_cs.py   -> tests/golden/cond_spectra.npz
"""
import os
import sys
import types
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, 'refshim'))
sys.path.insert(0, HERE)
import oq_bootstrap  # noqa
oq_bootstrap.install()

from openquake.hazardlib.contexts import ContextMaker  # noqa
from openquake.hazardlib import valid  # noqa
from openquake.hazardlib.imt import from_string  # noqa
from openquake.hazardlib.geo import Point, NodalPlane  # noqa
from openquake.hazardlib.pmf import PMF  # noqa
from openquake.hazardlib.mfd import EvenlyDiscretizedMFD  # noqa
from openquake.hazardlib.scalerel.wc1994 import WC1994  # noqa
from openquake.hazardlib.source import PointSource  # noqa
from openquake.hazardlib.tom import PoissonTOM  # noqa
from openquake.hazardlib.site import SiteCollection  # noqa
from openquake.hazardlib.calc.filters import IntegrationDistance  # noqa
from openquake.hazardlib.calc.cond_spectra import cond_spectra, get_cs_out  # noqa
from openquake.hazardlib.correlation_models.cross_imt.baker_jayaram_2008 \
    import BakerJayaram2008  # noqa
from openquake.baselib.general import DictArray  # noqa

GOLDEN = os.path.join(HERE, '..', 'tests', 'golden')
SRC = [(0.10, 0.10, [0.02, 0.008, 0.003, 0.001], 30., 60., 90., 6.),
       (0.45, -0.20, [0.05, 0.01, 0.002], 120., 90., 0., 8.),
       (-0.30, 0.35, [0.01, 0.006, 0.004, 0.002, 0.001], 250., 45., -90., 10.)]


def main():
    tom = PoissonTOM(50.)
    srcs = []
    for i, (lon, lat, rates, strike, dip, rake, dep) in enumerate(SRC):
        src = PointSource(
            str(i), 'src%d' % i, 'Active Shallow Crust',
            EvenlyDiscretizedMFD(5.05, 0.3, rates), 2., WC1994(), 1.5, tom,
            0., 20., Point(lon, lat),
            PMF([(1., NodalPlane(strike, dip, rake))]), PMF([(1., dep)]))
        src.id, src.grp_id, src.weight, src.offset = i, 0, 1., 0
        srcs.append(src)
    sitecol = SiteCollection.from_points(
        [0.05, 0.30], [0.0, 0.25], sitemodel=types.SimpleNamespace(
            reference_vs30_value=500., reference_vs30_type='measured',
            reference_depth_to_1pt0km_per_sec=100.,
            reference_depth_to_2pt5km_per_sec=5.))
    imts = ['PGA', 'SA(0.05)', 'SA(0.15)', 'SA(0.2)', 'SA(0.5)', 'SA(1.0)',
            'SA(2.0)']
    imtls = DictArray({imt: [0.05, 0.2] for imt in imts})
    gsims = [valid.gsim('[BooreEtAl2014]'), valid.gsim('[BooreAtkinson2008]')]
    model = BakerJayaram2008()
    cm = ContextMaker('Active Shallow Crust', gsims, dict(
        imtls=imtls, truncation_level=3., investigation_time=50.,
        maximum_distance=IntegrationDistance.new('100'),
        poes=[0.1, 0.02], cross_correl=model))
    cm.poes = [0.1, 0.02]
    imls = np.array([[0.12, 0.25], [0.10, 0.22]])       # (N, P) for SA(0.2)
    spectra, s_sigma = cond_spectra(cm, srcs, sitecol, 'SA(0.2)', imls)
    rho = np.array([[model.rho(from_string(a), from_string(b)) for b in imts]
                    for a in imts])
    np.savez_compressed(
        os.path.join(GOLDEN, 'cond_spectra.npz'), imts=np.array(imts),
        imls=imls, poes=np.array([0.1, 0.02]), spectra=spectra,
        s_sigma=s_sigma, rho=rho)
    print('cond_spectra.npz', spectra.shape, spectra[0, :, 0, 0])


if __name__ == '__main__':
    main()
