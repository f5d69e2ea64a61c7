"""Shared helpers for the tests: golden-fixture loaders and synthetic inputs
(the same generators as oracle/make_golden.py)."""
import os
import csv
import json
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, 'tests', 'golden')

IMTS = ['PGA', 'SA(0.1)', 'SA(0.2)', 'SA(0.5)', 'SA(1.0)', 'SA(2.0)',
        'SA(3.0)']


class Bag(object):
    """recarray-like bag of equally long arrays"""
    def __init__(self, **kw):
        vars(self).update(kw)

    def __len__(self):
        return len(next(iter(vars(self).values())))

    def __getitem__(self, idx):
        return Bag(**{k: v[idx] for k, v in vars(self).items()})


def load(name):
    return np.load(os.path.join(GOLDEN, name + '.npz'))


def gmpe_cases():
    return sorted(f[5:-4] for f in os.listdir(GOLDEN)
                  if f.startswith('gmpe_') and f.endswith('.npz'))


def pipeline_cases():
    return sorted(f[9:-4] for f in os.listdir(GOLDEN)
                  if f.startswith('pipeline_') and f.endswith('.npz'))


def ctx_of(d):
    """context rows of a fixture as a Bag"""
    return Bag(**{k[4:]: d[k] for k in d.files if k.startswith('ctx_')})


def rup_of(d):
    return {k[4:]: d[k] for k in d.files if k.startswith('rup_')}


def sites_of(d):
    return {k[5:]: d[k] for k in d.files if k.startswith('site_')}


def wc1994_area(mag, rake):
    """Wells & Coppersmith (1994) median rupture area, the published
    regression used by the reference's WC1994.get_median_area
    (openquake/hazardlib/scalerel/wc1994.py)"""
    mag, rake = np.asarray(mag, float), np.asarray(rake, float)
    ss = (np.abs(rake) <= 45) | (np.abs(rake) >= 135)
    out = np.where(ss, 10.0 ** (-3.42 + 0.90 * mag),
                   np.where(rake > 0, 10.0 ** (-3.99 + 0.98 * mag),
                            10.0 ** (-2.87 + 0.82 * mag)))
    return out


def synth_ruptures(U, rng, box=1.5):
    """random planar ruptures in a box"""
    mag = rng.choice(np.round(np.arange(5.0, 8.0, 0.1), 1), U)
    strike = rng.uniform(0, 360, U)
    dip = rng.choice([30., 45., 60., 90.], U)
    rake = rng.choice([-90., 0., 90.], U)
    lon = rng.uniform(0, box, U)
    lat = rng.uniform(0, box, U)
    dep = rng.uniform(5, 15, U)
    area = wc1994_area(mag, rake)
    rate = 10 ** (4 - mag) / U
    return dict(mag=mag, strike=strike, dip=dip, rake=rake, hypo_lon=lon,
                hypo_lat=lat, hypo_dep=dep, area=area, rate=rate)


def synth_sites(N, rng, box=1.5, sentinels=False):
    n_side = int(np.ceil(np.sqrt(N)))
    gx, gy = np.meshgrid(np.linspace(0, box, n_side),
                         np.linspace(0, box, n_side))
    sites = dict(sids=np.arange(N, dtype=np.uint32),
                 lon=gx.ravel()[:N], lat=gy.ravel()[:N], depth=np.zeros(N),
                 vs30=rng.uniform(180, 1500, N),
                 vs30measured=rng.random(N) < .5,
                 z1pt0=rng.uniform(5, 800, N), z2pt5=rng.uniform(0.3, 6, N))
    if sentinels:
        sites['z1pt0'][rng.random(N) < .1] = -999
        sites['z2pt5'][rng.random(N) < .1] = -999
    return sites


def read_kat(fname):
    """A reference GMPE verification table (CSV): returns (ctx Bag,
    result_type, {imt: values})"""
    with open(os.path.join(GOLDEN, fname)) as f:
        rows = list(csv.reader(f))
    header, body = rows[0], np.array(rows[1:])
    cols, imts = {}, {}
    rtype = None
    for j, h in enumerate(header):
        if h == 'result_type':
            rtype = body[0, j]
        elif h in ('damping', 'component_type') or \
                h.startswith('Unnamed') or not h:
            continue
        elif h.startswith(('rup_', 'site_', 'dist_')):
            name = h.split('_', 1)[1]
            if name == 'vs30measured':
                cols[name] = body[:, j].astype(float) != 0
            else:
                cols[name] = body[:, j].astype(float)
        else:
            key = (h.upper() if h.lower() in ('pga', 'pgv')
                   else 'SA(%s)' % float(h))
            imts[key] = body[:, j].astype(float)
    return Bag(**cols), rtype, imts


def json_kwargs(d):
    return json.loads(str(d['kwargs']))
