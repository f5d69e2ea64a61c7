"""Dump the published GMPE coefficient tables used by the hot path into
oq_engine_b200/data/coeffs.json.

The numbers are the published regression coefficients of the ground-motion
models (NGA-West2 etc.); they are read here from the reference's CoeffsTable
objects (openquake/hazardlib/gsim/<module>.py, attribute COEFFS*) so that the
product carries data only, no reference code.  Run in the build container:

    python tools/extract_coeffs.py

Layout of the JSON: {table: {"columns": [...], "non_sa": {"PGA": [...], ...},
"sa": [[period, v0, v1, ...], ...] sorted by period}}.
"""
import os
import sys
import json

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, '..', 'oracle', 'refshim'))
import oq_bootstrap  # noqa
oq_bootstrap.install()

from openquake.hazardlib.gsim.boore_2014 import BooreEtAl2014  # noqa
from openquake.hazardlib.gsim.chiou_youngs_2014 import ChiouYoungs2014  # noqa
from openquake.hazardlib.gsim.abrahamson_2014 import AbrahamsonEtAl2014  # noqa
from openquake.hazardlib.gsim.campbell_bozorgnia_2014 import (  # noqa
    CampbellBozorgnia2014)
from openquake.hazardlib.gsim.boore_atkinson_2008 import (  # noqa
    BooreAtkinson2008)
from openquake.hazardlib.gsim.atkinson_boore_2006 import (  # noqa
    AtkinsonBoore2006)
from openquake.hazardlib.gsim.akkar_2014 import (  # noqa
    AkkarEtAlRjb2014, AkkarEtAlRepi2014, AkkarEtAlRhyp2014)
from openquake.hazardlib.gsim.chiou_youngs_2008 import ChiouYoungs2008  # noqa
from openquake.hazardlib.gsim.toro_2002 import (  # noqa
    ToroEtAl2002, ToroEtAl2002SHARE)
from openquake.hazardlib.gsim.campbell_2003 import (  # noqa
    Campbell2003, Campbell2003SHARE)

from openquake.hazardlib.gsim.sadigh_1997 import SadighEtAl1997  # noqa

TABLES = {
    'BSSA14': BooreEtAl2014.COEFFS,
    'CY14': ChiouYoungs2014.COEFFS,
    'ASK14': AbrahamsonEtAl2014.COEFFS,
    'CB14': CampbellBozorgnia2014.COEFFS,
    'BA08': BooreAtkinson2008.COEFFS,
    'AB06_SOIL_RESPONSE': AtkinsonBoore2006.COEFFS_SOIL_RESPONSE,
    'AB06_HARD_ROCK': AtkinsonBoore2006.COEFFS_HARD_ROCK,
    'AB06_BC': AtkinsonBoore2006.COEFFS_BC,
    'AB06_STRESS': AtkinsonBoore2006.COEFFS_STRESS,
    'AKKAR14_RJB': AkkarEtAlRjb2014.COEFFS,
    'AKKAR14_REPI': AkkarEtAlRepi2014.COEFFS,
    'AKKAR14_RHYP': AkkarEtAlRhyp2014.COEFFS,
    'CY08': ChiouYoungs2008.COEFFS,
    'TORO02': ToroEtAl2002.COEFFS,
    'C03': Campbell2003.COEFFS,
    'TORO02_FS_ROCK': ToroEtAl2002SHARE.COEFFS_FS_ROCK,
    'C03_FS_ROCK': Campbell2003SHARE.COEFFS_FS_ROCK,
    'SADIGH97_ROCK_LOWMAG': SadighEtAl1997.COEFFS_ROCK_LOWMAG,
    'SADIGH97_ROCK_HIMAG': SadighEtAl1997.COEFFS_ROCK_HIMAG,
    'SADIGH97_ROCK_STDDERR': SadighEtAl1997.COEFFS_ROCK_STDDERR,
    'SADIGH97_SOIL': SadighEtAl1997.COEFFS_SOIL,
}


def dump(table):
    cols = list(table.rb.names)
    non_sa, sa = {}, []
    for imt, rec in table._coeffs.items():
        vals = [float(rec[c]) for c in cols]
        if imt.string.startswith('SA'):
            assert imt.damping == 5.0, imt
            sa.append([float(imt.period)] + vals)
        else:
            non_sa[imt.string] = vals
    sa.sort()
    return {'columns': cols, 'non_sa': non_sa, 'sa': sa}


if __name__ == '__main__':
    out = {name: dump(tab) for name, tab in TABLES.items()}
    dest = os.path.join(HERE, '..', 'oq_engine_b200', 'data', 'coeffs.json')
    with open(dest, 'w') as f:
        json.dump(out, f, separators=(',', ':'))
    for name, dic in out.items():
        print(name, len(dic['columns']), 'cols', len(dic['sa']), 'SA rows',
              list(dic['non_sa']))
