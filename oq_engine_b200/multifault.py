"""Multi-fault ruptures on the device path: the
`dparam` cache of the reference's ContextMaker.

A multi-fault source is a set of SECTIONS (surfaces) and ruptures made of
several sections.  The reference computes, once per section, rrup, rjb and
the GC2 coordinates `tuw` of the sites for both orientations of the
section's top-of-rupture line (`_build_dparam`, contexts.py:1039-1049;
get_dparam, calc/filters.py:51-71) and composes them per rupture in genctxs
(contexts.py:959-1029): rrup / rjb = minimum over the sections, rx / ry0
from the weighted GC2 coordinates of the sections ordered, flipped and
shifted by the MultiLine of the rupture (`_get_tu` :85-98, set_distances
:101-141; geo/multiline.py).

Here the section planes come from the mesh kernels and `oqg_line_tuw`, and
`oqg_multi_compose` builds the (U, N) planes of the ruptures that
`oqg_hazard_dist` consumes.  The host part below is the per-line and
per-rupture bookkeeping of the reference (a few numbers per section), in
numpy: `TorLine` (geo/line.py: projection centre, length, azimuth, the
segment unit vectors `sut_hat`), `multiline` (geo/multiline.py `_build3`,
`get_origin`: flipped, soidx, shift) and `u_max`.  The top-of-rupture lines
themselves (`surface.tor`, i.e. `Line.keep_corners` of the top row) and the
section areas are inputs: they are source-model geometry.
"""
import math
import numpy
from .fault import MeshRuptures, _middle_of_two, EARTH_RADIUS
from .area import get_spherical_bounding_box

TOLERANCE = 0.1     # geo/line.py:26
F32 = numpy.float32


def project_direct(lam0, phi0, lons, lats):
    """geo/utils.py:599-607"""
    lambdas, phis = numpy.radians(lons), numpy.radians(lats)
    cos_phis = numpy.cos(phis)
    lambdas = lambdas - lam0
    xx = numpy.cos(phis) * numpy.sin(lambdas) * EARTH_RADIUS
    yy = (math.cos(phi0) * numpy.sin(phis) - math.sin(phi0) * cos_phis
          * numpy.cos(lambdas)) * EARTH_RADIUS
    return xx, yy


def project_reverse(lam0, phi0, xs, ys):
    """geo/utils.py:577-594"""
    sin_phi0, cos_phi0 = math.sin(phi0), math.cos(phi0)
    xx, yy = xs / EARTH_RADIUS, ys / EARTH_RADIUS
    cos_c = numpy.sqrt(1. - (xx ** 2 + yy ** 2))
    phis = numpy.arcsin(cos_c * sin_phi0 + yy * cos_phi0)
    lambdas = numpy.arctan2(xx, cos_phi0 * cos_c - yy * sin_phi0)
    xx = numpy.degrees(lam0 + lambdas)
    yy = numpy.degrees(phis)
    xx = numpy.where(xx >= 180., xx - 360., xx)
    xx = numpy.where(xx <= -180., xx + 360., xx)
    return xx, yy


def projection_centre(lons, lats):
    """OrthographicProjection.from_ (geo/utils.py:652-668): the middle of
    the spherical bounding box, radians"""
    west, east, north, south = get_spherical_bounding_box(
        numpy.asarray(lons, float), numpy.asarray(lats, float))
    lon0, lat0 = _middle_of_two(west, north, east, south)
    return math.radians(float(lon0)), math.radians(float(lat0))


def _fast_distance(lon, lat, lons, lats):
    lon, lat = math.radians(lon), math.radians(lat)
    lons, lats = numpy.radians(lons), numpy.radians(lats)
    return 2 * EARTH_RADIUS * numpy.arcsin(numpy.sqrt(
        numpy.sin((lat - lats) / 2.0) ** 2.0 + numpy.cos(lat) *
        numpy.cos(lats) * numpy.sin((lon - lons) / 2.0) ** 2.0))


def _azimuth(lons1, lats1, lons2, lats2):
    """geodetic.azimuth, degrees in [0, 360)"""
    lons1, lats1, lons2, lats2 = (numpy.radians(x) for x in (
        lons1, lats1, lons2, lats2))
    cos_lat2 = numpy.cos(lats2)
    true_course = numpy.degrees(numpy.arctan2(
        numpy.sin(lons1 - lons2) * cos_lat2,
        numpy.cos(lats1) * numpy.sin(lats2) -
        numpy.sin(lats1) * cos_lat2 * numpy.cos(lons1 - lons2)))
    return (360 - true_course) % 360


def angular_mean_weighted(degrees, weights):
    """geo/utils.py:126-135"""
    mean_sin = mean_cos = 0.
    for d, w in zip(degrees, weights):
        r = math.radians(d)
        mean_sin += math.sin(r) * w
        mean_cos += math.cos(r) * w
    return numpy.degrees(numpy.arctan2(mean_sin, mean_cos))


class TorLine(object):
    """A top-of-rupture polyline (geo/line.py Line.init, sut_hat): `coo`
    (K, 2) longitudes and latitudes at constant depth"""
    def __init__(self, coo):
        self.coo = coo = numpy.asarray(coo, float)[:, :2]
        assert len(coo) >= 2, 'a line needs two points'
        self.lam0, self.phi0 = projection_centre(coo[:, 0], coo[:, 1])
        lo, la = coo[:, 0], coo[:, 1]
        if len(coo) == 2:
            self.length = float(_fast_distance(lo[0], la[0], lo[1], la[1]))
            self.azimuth = float(_azimuth(lo[0], la[0], lo[1], la[1]))
        else:
            # Point.distance of consecutive points (same depth)
            self.length = float(sum(
                _fast_distance(lo[i], la[i], lo[i + 1], la[i + 1])
                for i in range(len(coo) - 1)))
            azimuths = _azimuth(lo[:-1], la[:-1], lo[1:], la[1:])
            # geodetic_distance of arrays is float32 in the reference
            f = numpy.float32
            a1, b1, a2, b2 = (f(numpy.radians(x)) for x in (
                lo[:-1], la[:-1], lo[1:], la[1:]))
            arr = numpy.sin((b1 - b2) / 2.0) ** 2.0
            arr += numpy.cos(b1) * numpy.cos(b2) * numpy.sin(
                (a1 - a2) / 2.0) ** 2.0
            dist = 2 * EARTH_RADIUS * numpy.arcsin(numpy.sqrt(arr))
            self.azimuth = float(angular_mean_weighted(
                numpy.float64(azimuths), numpy.float64(dist)) % 360)
        # sut_hat (geo/line.py:543-565)
        sx, sy = project_direct(self.lam0, self.phi0, lo, la)
        dx, dy = sx[1:] - sx[:-1], sy[1:] - sy[:-1]
        self.slen = (dx ** 2 + dy ** 2) ** 0.5
        assert (self.slen > 0).all(), self.slen
        self.uhat = numpy.stack([dx / self.slen, dy / self.slen], axis=1)
        # versor of sg x (0, 0, 1) = (dy, -dx, 0)
        self.that = numpy.stack([dy / self.slen, -dx / self.slen], axis=1)
        self.tx, self.ty = sx, sy

    def flip(self):
        return TorLine(self.coo[::-1])

    def segments(self):
        """(nseg, 7): projected start of the segment, its length, uhat, that
        -- the record `oqg_line_tuw` reads"""
        return numpy.column_stack([self.tx[:-1], self.ty[:-1], self.slen,
                                   self.uhat, self.that])

    def get_tuw(self, lons, lats):
        """geo/line.py:178-266 in numpy (used on the host for u_max only;
        the sites go through oqg_line_tuw): (N, 3) float32"""
        sx, sy = project_direct(self.lam0, self.phi0, numpy.asarray(
            lons, float), numpy.asarray(lats, float))
        S, N = len(self.slen), len(sx)
        ui, ti = numpy.zeros((S, N)), numpy.zeros((S, N))
        for i in range(S):
            dx, dy = sx - self.tx[i], sy - self.ty[i]
            ui[i] = dx * self.uhat[i, 0] + dy * self.uhat[i, 1]
            ti[i] = dx * self.that[i, 0] + dy * self.that[i, 1]
        ws = numpy.zeros((S, N))
        iot = numpy.zeros(N, bool)
        with numpy.errstate(all='ignore'):
            for i in range(S):
                t, u, sl = ti[i], ui[i], self.slen[i]
                c0 = numpy.abs(t) >= TOLERANCE
                w = numpy.where(c0, (numpy.arctan((sl - u) / t) -
                                     numpy.arctan(-u / t)) / t, 0.)
                c1 = ~c0
                c2 = (u < 0. - TOLERANCE) | (u > sl + TOLERANCE)
                w = numpy.where(c1 & c2, 1. / (u - sl) - 1. / u, w)
                c3 = (u >= -TOLERANCE) & (u <= sl + TOLERANCE)
                w = numpy.where(c1 & c3, 1 / (-0.01 - sl) + 1 / 0.01, w)
                iot |= c1 & c3
                ws[i] = w
        wsum = ws.sum(axis=0)
        t_upp = (ti * ws / wsum).sum(axis=0)
        u_upp = ui[0] * ws[0]
        for i in range(1, S):
            u_upp += (ui[i] + self.slen[:i].sum()) * ws[i]
        u_upp = u_upp / wsum
        t_upp[iot] = 0.
        return numpy.stack([t_upp, u_upp, wsum], axis=1).astype(F32)


def multiline(lines):
    """
    MultiLine.__init__ / _build3 / get_origin (geo/multiline.py:64-176):
    which lines are flipped, their order along the average strike and the
    shift of the u coordinate of each -> (flipped bool (L,), soidx int (L,),
    shift float32 (L,))
    """
    L = len(lines)
    llen = numpy.array([ln.length for ln in lines])
    azimuths = numpy.array([ln.azimuth for ln in lines])
    ends = numpy.array([ln.coo[[0, -1]] for ln in lines])       # (L, 2, 2)
    lam0, phi0 = projection_centre(ends[:, :, 0], ends[:, :, 1])
    ave = angular_mean_weighted(azimuths, llen) % 360
    flipped = numpy.zeros(L, bool)
    if 90 <= ave <= 270:
        idx = (azimuths >= (ave - 90) % 360) & (azimuths < (ave + 90) % 360)
    else:
        idx = (azimuths >= (ave - 90) % 360) | (azimuths < (ave + 90) % 360)
    delta = numpy.abs(azimuths - ave)
    scale = numpy.abs(numpy.cos(numpy.radians(delta)))
    ratio = numpy.sum(llen[idx] * scale[idx]) / numpy.sum(llen * scale)
    if ratio > 0.5:
        flipped[~idx] = True
    else:
        flipped[idx] = True
    azimuths = numpy.where(flipped, (azimuths + 180) % 360, azimuths)
    avg_azim = angular_mean_weighted(azimuths, llen) % 360
    # get_origin
    flatlons, flatlats = ends[:, :, 0].flatten(), ends[:, :, 1].flatten()
    px, py = project_direct(lam0, phi0, flatlons, flatlats)
    DELTA = 0.1
    to_east = (avg_azim > 0) & (avg_azim <= 180)
    if to_east or abs(avg_azim) < DELTA:
        i0 = numpy.argmin(px)
    else:
        i0 = numpy.argmax(px)
    eps = numpy.minimum(px[0::2], px[1::2])
    soidx = numpy.argsort(eps)
    if not (to_east or abs(avg_azim) < DELTA):
        soidx = numpy.flipud(soidx)
    olon, olat = project_reverse(lam0, phi0, numpy.array([px[i0]]),
                                 numpy.array([py[i0]]))
    olons = numpy.array([ends[i, int(flipped[i]), 0] for i in soidx])
    olats = numpy.array([ends[i, int(flipped[i]), 1] for i in soidx])
    dist = _fast_distance(float(olon[0]), float(olat[0]), olons, olats)
    az = _azimuth(float(olon[0]), float(olat[0]), olons, olats)
    # fast_azimuth returns -degrees(true_course): the same angle mod 360
    shift = numpy.cos(numpy.radians(avg_azim - az)) * dist
    return flipped, soidx, shift.astype(F32)


def get_tu(shift, tuws):
    """geo/multiline.py:246-261, float32 accumulation as the reference"""
    N = len(tuws[0])
    ts, us, ws = (numpy.zeros(N, F32) for _ in range(3))
    for i, tuw in enumerate(tuws):
        t, u, w = tuw[:, 0], tuw[:, 1], tuw[:, 2]
        ts += t * w
        us += (u + shift[i]) * w
        ws += w
    return ts / ws, us / ws


def u_max(lines, flipped, soidx, shift):
    """MultiLine.get_u_max (geo/multiline.py:187-192, 264-278): the largest
    |u| of the end points of the lines"""
    ends = numpy.array([ln.coo[[0, -1]] for ln in lines])
    lons, lats = ends[:, :, 0].flatten(), ends[:, :, 1].flatten()
    tuws = []
    for idx in soidx:
        ln = lines[idx].flip() if flipped[idx] else lines[idx]
        tuws.append(ln.get_tuw(lons, lats))
    _t, u = get_tu(shift, tuws)
    return F32(numpy.abs(u).max())


class MultiFaultRuptures(object):
    """
    U ruptures made of sections (the ruptures of a multi-fault source).

    :param sections: a MeshRuptures-like container of the S section meshes
        (only lons / lats / deps / shape / kind are used), or a list of
        (lons, lats, depths) 2D arrays
    :param tors: for each section the (K, 2) coordinates of its
        top-of-rupture line (`surface.tor.coo` in the reference)
    :param areas: (S,) areas of the sections in km2 (`surface.get_area()`)
    :param rup_sections: for each rupture the list of its section indices
    :param mag, rake, rate: (U,)
    :param hypo: (U, 3) hypocentres lon, lat, depth
    """
    def __init__(self, sections, tors, areas, rup_sections, mag, rake, rate,
                 hypo, kind='simple'):
        if not isinstance(sections, MeshRuptures):
            S = len(sections)
            sections = MeshRuptures.from_meshes(
                sections, numpy.zeros(S), 0., numpy.zeros(S),
                numpy.zeros((S, 3)), kind=kind)
        self.sections = sections
        S = len(sections)
        self.lines = [TorLine(c) for c in tors]
        assert len(self.lines) == S, (len(self.lines), S)
        self.areas = numpy.asarray(areas, F32)
        self.rup_sections = [numpy.asarray(r, numpy.int64)
                             for r in rup_sections]
        if any(len(r) == 0 for r in self.rup_sections):
            raise ValueError('a rupture without sections')
        if max(int(r.max()) for r in self.rup_sections) >= S or min(
                int(r.min()) for r in self.rup_sections) < 0:
            raise ValueError('section index out of range')
        self.mag = numpy.ascontiguousarray(mag, float)
        U = len(self.mag)
        assert len(self.rup_sections) == U
        self.rake = numpy.broadcast_to(numpy.asarray(rake, float),
                                       (U,)).copy()
        self.rate = numpy.ascontiguousarray(rate, float)
        self.hypo = numpy.ascontiguousarray(hypo, float).reshape(U, 3)
        self.rup_id = numpy.arange(U, dtype=numpy.uint32)
        self.probs_occur = None
        # per rupture: order, flips, shifts and u_max of its MultiLine
        self.order, self.flip, self.shift, self.umax = [], [], [], []
        for secs in self.rup_sections:
            lines = [self.lines[i] for i in secs]
            flipped, soidx, shift = multiline(lines)
            self.order.append(secs[soidx])
            self.flip.append(flipped[soidx])
            self.shift.append(shift)
            # build_msparams takes the straight lines between the ends of
            # the top edges for u_max (geo/surface/multi.py:88-105)
            straight = [TorLine(ln.coo[[0, -1]]) for ln in lines]
            f2, s2, sh2 = multiline(straight)
            self.umax.append(u_max(straight, f2, s2, sh2))
        self.umax = numpy.array(self.umax, F32)

    def __len__(self):
        return len(self.mag)

    def msparams(self, secpar):
        """
        The parameters of the multi-surfaces from those of the sections
        (build_msparams, geo/surface/multi.py:107-117: weights = areas; the
        reference keeps everything in float32 records).

        :param secpar: dict of (S,) arrays dip, strike, width, ztor
        :returns: dict of (U,) float64 arrays dip, strike, width, ztor
        """
        out = {k: numpy.zeros(len(self)) for k in ('dip', 'strike', 'width',
                                                   'ztor')}
        sp = {k: numpy.asarray(v, F32) for k, v in secpar.items()}
        for u, secs in enumerate(self.rup_sections):
            areas = self.areas[secs]
            ws = areas / F32(areas.sum())
            out['dip'][u] = F32(ws @ sp['dip'][secs])
            out['strike'][u] = F32(angular_mean_weighted(
                sp['strike'][secs], ws) % 360)
            out['width'][u] = F32(ws @ sp['width'][secs])
            out['ztor'][u] = F32(ws @ sp['ztor'][secs])
        return out


# ------------------------------------------------------------------------ #
#                         planes on the device                              #
# ------------------------------------------------------------------------ #

def section_planes(cmaker, rups, sitecol, names):
    """
    The cache of the reference (`dparam`): per section the rrup and rjb
    planes (S, N) float64 and the GC2 coordinates tuw (S, 2, N, 3) float32
    for both orientations of its top-of-rupture line, CUDA tensors; and the
    section parameters dip / strike / width / ztor (numpy)
    """
    import ctypes
    import torch
    from .contexts import _dev, _ptr
    from . import fault
    h = cmaker.handle()
    dev = h.device
    cuda = 'cuda:%d' % dev
    sec = rups.sections
    S, N = len(sec), len(sitecol)
    want = {'rrup'} | ({'rjb'} & set(names))
    planes, par = fault.rupture_planes(cmaker, sec, sitecol, want)
    out = dict(rrup=planes['rrup'], rjb=planes.get('rjb'), par=par, tuw=None)
    if {'rx', 'ry0'} & set(names):
        lines = []
        for ln in rups.lines:
            lines += [ln, ln.flip()]
        seg = numpy.concatenate([ln.segments() for ln in lines])
        loff = numpy.concatenate([[0], numpy.cumsum(
            [len(ln.slen) for ln in lines])]).astype(numpy.int64)
        proj = numpy.array([[ln.lam0, ln.phi0] for ln in lines])
        st, keep = cmaker._sites_struct(sitecol, dev)
        t_seg, t_off, t_proj = _dev(seg, dev), _dev(loff, dev), _dev(proj, dev)
        tuw = torch.empty((S, 2, N, 3), dtype=torch.float32, device=cuda)
        h.call('oqg_line_tuw', _ptr(t_seg), _ptr(t_off), _ptr(t_proj), 2 * S,
               ctypes.byref(st), N, _ptr(tuw))
        torch.cuda.current_stream(dev).synchronize()
        out['tuw'] = tuw
        del keep
    return out


def rupture_planes(cmaker, rups, sitecol, names, cache=None):
    """
    (planes {name: CUDA tensor (U, N) float64}, par {dip, strike, width,
    ztor, hypo_depth}) of multi-fault ruptures, composed from the section
    cache (computed here unless passed in)
    """
    import torch
    from .contexts import _dev, _ptr
    h = cmaker.handle()
    dev = h.device
    cuda = 'cuda:%d' % dev
    names = set(names) | {'rrup'}
    bad = names - {'rrup', 'rjb', 'rx', 'ry0'}
    if bad:
        raise NotImplementedError('distances %s of multi-fault ruptures'
                                  % sorted(bad))
    if not cache:
        with cmaker.sec_mon:      # 'building dparam', contexts.py:1073
            cache = section_planes(cmaker, rups, sitecol, names)
    U, N, S = len(rups), len(sitecol), len(rups.sections)
    roff = numpy.concatenate([[0], numpy.cumsum(
        [len(o) for o in rups.order])]).astype(numpy.int64)
    rsec = numpy.concatenate(rups.order).astype(numpy.int32)
    rflip = numpy.concatenate(rups.flip).astype(numpy.uint8)
    rshift = numpy.concatenate(rups.shift).astype(numpy.float32)
    t_roff, t_rsec, t_rflip, t_rshift, t_umax = (
        _dev(x, dev) for x in (roff, rsec, rflip, rshift, rups.umax))
    planes = {n: torch.empty((U, N), dtype=torch.float64, device=cuda)
              for n in sorted(names)}
    h.call('oqg_multi_compose', _ptr(cache['rrup']), _ptr(cache['rjb']),
           _ptr(cache['tuw']), S, N, _ptr(t_roff), _ptr(t_rsec),
           _ptr(t_rflip), _ptr(t_rshift), _ptr(t_umax), U,
           _ptr(planes['rrup']), _ptr(planes.get('rjb')),
           _ptr(planes.get('rx')), _ptr(planes.get('ry0')))
    torch.cuda.current_stream(dev).synchronize()
    par = rups.msparams(cache['par'])
    par['hypo_depth'] = rups.hypo[:, 2].copy()
    return planes, par
