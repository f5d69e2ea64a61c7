"""Planar ruptures on the host: the struct-of-arrays the hot path consumes.

``planar_array_dt`` and the (U, 3) record-array layout are those of the
reference (openquake/hazardlib/geo/surface/planar.py:45-53) so that a planar
array built by the reference uploads unchanged; ``build_planar`` constructs U
independent ruptures (own hypocentre each, one hypocentral depth) with the
algorithm of _build_corners / build_planar_array (planar.py:107-185,
261-320), vectorised over the ruptures.
"""
import numpy

EARTH_RADIUS = 6371.0

planar_array_dt = numpy.dtype([
    ('corners', (float, 4)),
    ('xyz', (float, 4)),
    ('normal', float),
    ('uv1', float),
    ('uv2', float),
    ('wlr', float),
    ('sdr', float),
    ('hypo', float)])


def wc1994_median_area(mag, rake):
    """Wells & Coppersmith (1994) median rupture area in km^2 (the
    regression the reference's WC1994.get_median_area evaluates)"""
    mag, rake = numpy.asarray(mag, float), numpy.asarray(rake, float)
    ss = (numpy.abs(rake) <= 45) | (numpy.abs(rake) >= 135)
    return numpy.where(ss, 10.0 ** (-3.42 + 0.90 * mag),
                       numpy.where(rake > 0, 10.0 ** (-3.99 + 0.98 * mag),
                                   10.0 ** (-2.87 + 0.82 * mag)))


def _point_at(lon, lat, azimuth, distance):
    # geodetic.fast_point_at (geodetic.py:483-517), arrays
    lon, lat = numpy.radians(lon), numpy.radians(lat)
    tc = - numpy.radians(azimuth)
    sd, cd = numpy.sin(distance / EARTH_RADIUS), numpy.cos(
        distance / EARTH_RADIUS)
    sl, cl = numpy.sin(lat), numpy.cos(lat)
    sin_lats = sl * cd + cl * sd * numpy.cos(tc)
    dlon = numpy.arctan2(numpy.sin(tc) * sd * cl, cd - sl * sin_lats)
    lons = numpy.mod(lon - dlon + numpy.pi, 2 * numpy.pi) - numpy.pi
    return numpy.degrees(lons), numpy.degrees(numpy.arcsin(sin_lats))


def _xyz(lons, lats, deps):
    phi, theta = numpy.radians(lons), numpy.radians(lats)
    rr = EARTH_RADIUS - deps
    ctr = rr * numpy.cos(theta)
    return numpy.stack([ctr * numpy.cos(phi), ctr * numpy.sin(phi),
                        rr * numpy.sin(theta)], axis=-1)


def build_planar(mag, strike, dip, rake, rate, area, hypo_lon, hypo_lat,
                 hypo_dep, usd, lsd, rar, dip_frac=0.5):
    """
    U planar ruptures as a (U, 3) record array of planar_array_dt.

    :param mag, strike, dip, rake, rate, area: (U,) rupture parameters
    :param hypo_lon, hypo_lat, hypo_dep: (U,) hypocentres
    :param usd, lsd, rar: upper / lower seismogenic depth, aspect ratio
    """
    f = lambda x: numpy.asarray(x, float)  # noqa
    strike, dip, rake, rate, area = map(f, (strike, dip, rake, rate, area))
    clon, clat, cdep = f(hypo_lon).copy(), f(hypo_lat).copy(), f(hypo_dep)
    U = len(strike)
    # get_rupdims, planar.py:68-95
    rdip = numpy.radians(dip)
    sd, cd = numpy.sin(rdip), numpy.cos(rdip)
    max_width = (lsd - usd) / sd
    length = numpy.sqrt(area * rar)
    width = area / length
    over = width > max_width
    width = numpy.where(over, max_width, width)
    length = numpy.where(over, area / width, length)
    hl, hw, hh = length / 2., width * cd / 2., width * sd / 2.
    theta = numpy.degrees(numpy.arctan(hw / hl))
    hor = numpy.sqrt(hl ** 2 + hw ** 2)
    shift = (1. - 2. * dip_frac) * hh
    top, bot = cdep + shift - hh, cdep + shift + hh
    v = usd - top
    v2 = numpy.where(lsd - bot > 0, 0., lsd - bot)
    v = shift + numpy.where(v < 0, v2, v)
    moved = v != 0
    if moved.any():
        azi = numpy.where(v < 0, strike + 270., strike + 90.)
        with numpy.errstate(all='ignore'):
            mlon, mlat = _point_at(clon, clat, azi, numpy.abs(v / hh * hw))
        clon = numpy.where(moved, mlon, clon)
        clat = numpy.where(moved, mlat, clat)
    cdep2 = cdep + v
    pla = numpy.zeros((U, 3), planar_array_dt).view(numpy.recarray)
    corners = numpy.zeros((U, 4, 3))
    for c, azi in enumerate((strike + 180 + theta, strike - theta,
                             strike + 180 - theta, strike + theta)):
        corners[:, c, 0], corners[:, c, 1] = _point_at(clon, clat, azi, hor)
    corners[:, 0:2, 2] = (cdep2 - hh)[:, None]
    corners[:, 2:4, 2] = (cdep2 + hh)[:, None]
    # build_planar_array, planar.py:261-320
    xyz = _xyz(corners[..., 0], corners[..., 1], corners[..., 2])  # (U,4,3)
    tl, tr, bl = xyz[:, 0], xyz[:, 1], xyz[:, 2]

    def unit(x):
        return x / numpy.sqrt((x * x).sum(axis=-1))[:, None]

    def dot(a, b):
        return a[..., 0] * b[..., 0] + a[..., 1] * b[..., 1] + \
            a[..., 2] * b[..., 2]
    n = unit(numpy.cross(tl - tr, tl - bl))
    uv1 = unit(tr - tl)
    uv2 = numpy.cross(n, uv1)
    delta = xyz - xyz[:, :1]
    xx = dot(delta, uv1[:, None, :])   # (U, 4)
    yy = dot(delta, uv2[:, None, :])
    xx[:, 0] = yy[:, 0] = 0.
    plen = ((xx[:, 1] - xx[:, 0]) + (xx[:, 3] - xx[:, 2])) / 2.0
    pwid = ((yy[:, 2] - yy[:, 0]) + (yy[:, 3] - yy[:, 1])) / 2.0
    pla['corners'] = numpy.transpose(corners, (0, 2, 1))
    pla['xyz'] = numpy.transpose(xyz, (0, 2, 1))
    pla['normal'], pla['uv1'], pla['uv2'] = n, uv1, uv2
    pla['wlr'][:, 0], pla['wlr'][:, 1], pla['wlr'][:, 2] = pwid, plen, rate
    pla['sdr'][:, 0], pla['sdr'][:, 1], pla['sdr'][:, 2] = strike, dip, rake
    pla['hypo'][:, 0], pla['hypo'][:, 1], pla['hypo'][:, 2] = (
        f(hypo_lon), f(hypo_lat), cdep)
    return pla


def as_flat(planar):
    """(U, 42) float64 view/copy of a (U, 3) planar record array"""
    arr = numpy.ascontiguousarray(planar)
    assert arr.dtype == planar_array_dt and arr.shape[-1] == 3, arr.shape
    return arr.view(float).reshape(-1, 42)
