"""Ruptures with rectangular-mesh surfaces (the ruptures of fault sources) on
the device path: SURVEY.md section 8f row 2.

The reference builds, for every rupture of a fault source, a surface object
holding a RectangularMesh and calls, per rupture and per distance,
``get_distances`` (calc/filters.py:75-116) from ``ContextMaker.genctxs``
(contexts.py:959-1029).  Here the U ruptures of a source are ONE struct of
arrays -- the points of all the meshes concatenated, their offsets and shapes,
magnitude / rake / rate / hypocentre per rupture -- so that the mesh kernels
(oqg_mesh_min_distance, oqg_mesh_rjb, oqg_mesh_rx_ry0,
oqg_mesh_dip_strike_width) produce the (U, N) distance planes and the rupture
parameters in a few launches, and ``oqg_hazard_dist`` folds them into the map
with the same fused GMPE -> PoE -> accumulate stages as the planar ruptures.

`MeshRuptures` is the container; `SimpleFaultSource` floats ruptures over an
already built fault mesh exactly as SimpleFaultSource.iter_ruptures
(source/simple_fault.py:279-351) does and returns them as windows of that
mesh.  Building the fault mesh from the trace (Line.resample, geo/line.py) is
source-model geometry and stays outside this package: the mesh comes from the
caller (adapters/openquake_gpu.py takes it from the reference's
SimpleFaultSurface.from_fault_data).
"""
import math
import numpy
from .planar import wc1994_median_area  # noqa: re-exported
from .source import parse_msr, median_area

EARTH_RADIUS = 6371.0


def _middle_of_two(lon1, lat1, lon2, lat2):
    """geo/utils.py:671-685 get_middle_point for arrays (float64): the point
    half way on the great circle arc; identical points stay as they are"""
    lon1, lat1, lon2, lat2 = (numpy.asarray(x, float)
                              for x in (lon1, lat1, lon2, lat2))
    rl1, rp1, rl2, rp2 = (numpy.radians(x) for x in (lon1, lat1, lon2, lat2))
    # geodetic_distance of float arguments (geodetic.py fast_distance)
    dist = 2 * EARTH_RADIUS * numpy.arcsin(numpy.sqrt(
        numpy.sin((rp1 - rp2) / 2.0) ** 2.0 + numpy.cos(rp1) *
        numpy.cos(rp2) * numpy.sin((rl1 - rl2) / 2.0) ** 2.0))
    cos_lat2 = numpy.cos(rp2)
    true_course = numpy.degrees(numpy.arctan2(
        numpy.sin(rl1 - rl2) * cos_lat2,
        numpy.cos(rp1) * numpy.sin(rp2) -
        numpy.sin(rp1) * cos_lat2 * numpy.cos(rl1 - rl2)))
    azim = (360 - true_course) % 360
    # geodetic.point_at (geodetic.py:520-555)
    tc = numpy.radians(360 - azim)
    half = dist / 2.0
    sin_d, cos_d = numpy.sin(half / EARTH_RADIUS), numpy.cos(
        half / EARTH_RADIUS)
    sin_lat, cos_lat = numpy.sin(rp1), numpy.cos(rp1)
    sin_lats = sin_lat * cos_d + cos_lat * sin_d * numpy.cos(tc)
    lats = numpy.degrees(numpy.arcsin(sin_lats))
    dlon = numpy.arctan2(numpy.sin(tc) * sin_d * cos_lat,
                         cos_d - sin_lat * sin_lats)
    lons = numpy.degrees(numpy.mod(rl1 - dlon + numpy.pi, 2 * numpy.pi)
                         - numpy.pi)
    same = (lon1 == lon2) & (lat1 == lat2)
    return numpy.where(same, lon1, lons), numpy.where(same, lat1, lats)


def middle_points(lons, lats, deps):
    """RectangularMesh.get_middle_point (geo/mesh.py:683-726) of K meshes of
    the same shape at once: arrays (K, rows, cols) -> (K, 3) lon, lat, depth"""
    K, R, C = lons.shape

    def row_middle(r):
        mc = C // 2
        if C & 1:
            return lons[:, r, mc], lats[:, r, mc], deps[:, r, mc]
        lo, la = _middle_of_two(lons[:, r, mc - 1], lats[:, r, mc - 1],
                                lons[:, r, mc], lats[:, r, mc])
        return lo, la, (deps[:, r, mc - 1] + deps[:, r, mc]) / 2.0

    mr = R // 2
    if R & 1:
        lo, la, de = row_middle(mr)
    else:
        lo1, la1, d1 = row_middle(mr - 1)
        lo2, la2, d2 = row_middle(mr)
        lo, la = _middle_of_two(lo1, la1, lo2, la2)
        de = (d1 + d2) / 2.0
    return numpy.stack([lo, la, de], axis=1)


class MeshRuptures(object):
    """
    U ruptures whose surfaces are rectangular meshes, as a struct of arrays.

    :param mag, rake, rate: (U,) magnitude, rake (degrees), annual occurrence
        rate
    :param hypo: (U, 3) hypocentre lon, lat (degrees), depth (km)
    :param lons, lats, deps: (P,) the points of all the meshes, row-major,
        rupture after rupture
    :param shape: (U, 2) rows and columns of every mesh
    :param kind: 'simple' -- dip and strike from the top two rows, width from
        the first two columns of each mesh, as SimpleFaultSurface
        (geo/surface/simple_fault.py:82-114, 367-379); 'generic' -- from the
        whole mesh (BaseSurface, geo/surface/base.py)
    :param probs_occur: optional (U, K) for non-parametric ruptures
    """
    def __init__(self, mag, rake, rate, hypo, lons, lats, deps, shape,
                 kind='simple', probs_occur=None, rup_id=None):
        self.mag = numpy.ascontiguousarray(mag, float)
        U = len(self.mag)
        self.rake = numpy.broadcast_to(numpy.asarray(rake, float), (U,)).copy()
        self.rate = numpy.ascontiguousarray(rate, float)
        self.hypo = numpy.ascontiguousarray(hypo, float).reshape(U, 3)
        self.lons = numpy.ascontiguousarray(lons, float)
        self.lats = numpy.ascontiguousarray(lats, float)
        self.deps = numpy.ascontiguousarray(deps, float)
        self.shape = numpy.ascontiguousarray(shape, numpy.int32).reshape(U, 2)
        npts = self.shape[:, 0].astype(numpy.int64) * self.shape[:, 1]
        self.off = numpy.concatenate([[0], numpy.cumsum(npts)]).astype(
            numpy.int64)
        if self.off[-1] != len(self.lons) or len(self.lats) != len(
                self.lons) or len(self.deps) != len(self.lons):
            raise ValueError('the meshes have %d points, the shapes say %d'
                             % (len(self.lons), self.off[-1]))
        if kind not in ('simple', 'generic'):
            raise ValueError('kind=%r' % kind)
        if (self.shape < 2).any():
            # dip, strike and width need at least one mesh cell
            # (geo/mesh.py:728-745 asserts it)
            raise ValueError('a rupture mesh needs at least 2 x 2 points')
        self.kind = kind
        self.probs_occur = None if probs_occur is None else \
            numpy.ascontiguousarray(probs_occur, float)
        self.rup_id = numpy.arange(U, dtype=numpy.uint32) if rup_id is None \
            else numpy.asarray(rup_id, numpy.uint32)

    def __len__(self):
        return len(self.mag)

    @classmethod
    def from_meshes(cls, meshes, mag, rake, rate, hypo=None, **kw):
        """:param meshes: a list of U (lons, lats, depths) triples of 2D
        arrays; hypo None = the middle points of the meshes"""
        lo = [numpy.asarray(m[0], float) for m in meshes]
        la = [numpy.asarray(m[1], float) for m in meshes]
        de = [numpy.asarray(m[2], float) for m in meshes]
        if any(x.ndim != 2 for x in lo):
            raise ValueError('a rupture mesh must be 2D')
        if hypo is None:
            hypo = numpy.concatenate([
                middle_points(a[None], b[None], c[None])
                for a, b, c in zip(lo, la, de)])
        return cls(mag, rake, rate, hypo,
                   numpy.concatenate([x.ravel() for x in lo]),
                   numpy.concatenate([x.ravel() for x in la]),
                   numpy.concatenate([x.ravel() for x in de]),
                   [x.shape for x in lo], **kw)

    @classmethod
    def from_windows(cls, mesh, windows, mag, rake, rate, hypo=None, **kw):
        """
        Ruptures that are windows of one mesh (floating ruptures).

        :param mesh: (lons, lats, depths) 2D arrays of the whole fault
        :param windows: (U, 4) int: first row, first column, rows, columns
        """
        mlon, mlat, mdep = (numpy.asarray(x, float) for x in mesh)
        R, C = mlon.shape
        win = numpy.asarray(windows, numpy.int64).reshape(-1, 4)
        U = len(win)
        if ((win[:, 0] < 0) | (win[:, 1] < 0) | (win[:, 0] + win[:, 2] > R) |
                (win[:, 1] + win[:, 3] > C)).any():
            raise ValueError('a window is outside the %d x %d mesh' % (R, C))
        npts = win[:, 2] * win[:, 3]
        off = numpy.concatenate([[0], numpy.cumsum(npts)])
        idx = numpy.empty(off[-1], numpy.int64)
        hyp = numpy.empty((U, 3))
        # one vectorised gather per window shape
        shapes, inv = numpy.unique(win[:, 2:], axis=0, return_inverse=True)
        inv = inv.ravel()
        for k, (nr, nc) in enumerate(shapes):
            sel = numpy.nonzero(inv == k)[0]
            w = win[sel]
            ii = ((w[:, 0, None, None] + numpy.arange(nr)[None, :, None]) * C
                  + w[:, 1, None, None] + numpy.arange(nc)[None, None, :])
            pos = off[sel, None] + numpy.arange(nr * nc)[None, :]
            idx[pos] = ii.reshape(len(sel), -1)
            if hypo is None:
                hyp[sel] = middle_points(mlon.ravel()[ii], mlat.ravel()[ii],
                                         mdep.ravel()[ii])
        return cls(mag, rake, rate, hyp if hypo is None else hypo,
                   mlon.ravel()[idx], mlat.ravel()[idx], mdep.ravel()[idx],
                   win[:, 2:], **kw)

    def mesh(self, u):
        """(lons, lats, depths) 2D arrays of rupture u"""
        a, b = self.off[u], self.off[u + 1]
        sh = tuple(self.shape[u])
        return (self.lons[a:b].reshape(sh), self.lats[a:b].reshape(sh),
                self.deps[a:b].reshape(sh))

    def _gather(self, rows=None, cols=None):
        # the first `rows` rows / `cols` columns of every mesh, concatenated
        # in the same layout: (lons, lats, deps, off, shape)
        R, C = self.shape[:, 0].astype(numpy.int64), self.shape[:, 1].astype(
            numpy.int64)
        nr = R if rows is None else numpy.minimum(R, rows)
        nc = C if cols is None else numpy.minimum(C, cols)
        npts = nr * nc
        off = numpy.concatenate([[0], numpy.cumsum(npts)])
        u = numpy.repeat(numpy.arange(len(R)), npts)
        j = numpy.arange(off[-1]) - off[u]
        src = self.off[u] + (j // nc[u]) * C[u] + j % nc[u]
        return (self.lons[src], self.lats[src], self.deps[src],
                off.astype(numpy.int64),
                numpy.stack([nr, nc], axis=1).astype(numpy.int32))

    def top_edges(self):
        """the first row of every mesh: (lons, lats, off)"""
        lo, la, _de, off, _sh = self._gather(rows=1)
        return lo, la, off

    def ztor(self):
        """BaseSurface.get_top_edge_depth (geo/surface/base.py:386-399): the
        minimum depth of the first row"""
        _lo, _la, de, off, _sh = self._gather(rows=1)
        return numpy.minimum.reduceat(de, off[:-1])


class SimpleFaultSource(object):
    """
    The floating ruptures of a simple fault source over its mesh
    (source/simple_fault.py:279-351, _get_rupture_dimensions :410-450).

    :param mesh: (lons, lats, depths) 2D arrays (rows along dip, columns
        along strike) of the whole fault surface, spaced by
        `rupture_mesh_spacing` km -- SimpleFaultSurface.from_fault_data(...)
        .mesh in the reference
    :param mfd: an object with get_annual_occurrence_rates()
    :param msr: 'WC1994', 'PointMSR', 'PeerMSR' or 'CScalingMSR.C = <C>'
    """
    def __init__(self, source_id, mesh, mfd, rupture_mesh_spacing,
                 msr='WC1994', rupture_aspect_ratio=1.0, rake=0.,
                 tectonic_region_type='Active Shallow Crust'):
        self.source_id = source_id
        self.mesh = tuple(numpy.asarray(x, float) for x in mesh)
        if self.mesh[0].ndim != 2 or 1 in self.mesh[0].shape:
            raise ValueError('Mesh must have at least 2 nodes along both '
                             'length and width.')
        self.mfd = mfd
        self.rupture_mesh_spacing = float(rupture_mesh_spacing)
        parse_msr(msr)      # NotImplementedError if unsupported
        self.msr = msr
        self.rupture_aspect_ratio = float(rupture_aspect_ratio)
        self.rake = float(rake)
        self.tectonic_region_type = tectonic_region_type
        self.id = 0
        self.offset = 0
        mags = [m for m, _ in mfd.get_annual_occurrence_rates()]
        cols_rows = self._get_rupture_dimensions(float('inf'), float('inf'),
                                                 min(mags))
        if 1 in cols_rows:
            raise ValueError('mesh spacing %s is too high to represent '
                             'ruptures of magnitude %s' %
                             (rupture_mesh_spacing, min(mags)))

    def get_annual_occurrence_rates(self):
        return self.mfd.get_annual_occurrence_rates()

    def _median_area(self, mag):
        return float(median_area(self.msr, mag, self.rake))

    def _get_rupture_dimensions(self, fault_length, fault_width, mag):
        area = self._median_area(mag)
        rup_length = math.sqrt(area * self.rupture_aspect_ratio)
        rup_width = area / rup_length
        if area >= fault_length * fault_width:
            rup_length, rup_width = fault_length, fault_width
        elif rup_width > fault_width:
            rup_length = rup_length * (rup_width / fault_width)
            rup_width = fault_width
        elif rup_length > fault_length:
            rup_width = rup_width * (rup_length / fault_length)
            rup_length = fault_length
        rup_cols = int(round(rup_length / self.rupture_mesh_spacing) + 1)
        rup_rows = int(round(rup_width / self.rupture_mesh_spacing) + 1)
        return rup_cols, rup_rows

    def windows(self):
        """(mag (U,), rate (U,), windows (U, 4)) in the order of
        iter_ruptures: magnitude, first row, first column"""
        mesh_rows, mesh_cols = self.mesh[0].shape
        fault_length = float((mesh_cols - 1) * self.rupture_mesh_spacing)
        fault_width = float((mesh_rows - 1) * self.rupture_mesh_spacing)
        mags, rates, wins = [], [], []
        for mag, mag_occ_rate in self.get_annual_occurrence_rates():
            rup_cols, rup_rows = self._get_rupture_dimensions(
                fault_length, fault_width, mag)
            nl = mesh_cols - rup_cols + 1
            nw = mesh_rows - rup_rows + 1
            num_rup = nl * nw
            r0, c0 = numpy.divmod(numpy.arange(num_rup), nl)
            wins.append(numpy.stack([
                r0, c0, numpy.full(num_rup, rup_rows),
                numpy.full(num_rup, rup_cols)], axis=1))
            mags.append(numpy.full(num_rup, mag))
            rates.append(numpy.full(num_rup, mag_occ_rate / float(num_rup)))
        return (numpy.concatenate(mags), numpy.concatenate(rates),
                numpy.concatenate(wins))

    def count_ruptures(self):
        return len(self.windows()[0])

    def ruptures(self):
        """All the ruptures as one MeshRuptures (hypocentre = the middle
        point of each rupture mesh, simple_fault.py:347-351)"""
        mag, rate, win = self.windows()
        rups = MeshRuptures.from_windows(self.mesh, win, mag, self.rake, rate,
                                         kind='simple')
        rups.rup_id = (self.offset + numpy.arange(len(mag))).astype(
            numpy.uint32)
        return rups


# ------------------------------------------------------------------------ #
#                 distances and rupture parameters on the device            #
# ------------------------------------------------------------------------ #

def _slice(rups, u0, u1):
    """ruptures u0:u1 as a MeshRuptures sharing no memory with `rups`"""
    a, b = rups.off[u0], rups.off[u1]
    return MeshRuptures(
        rups.mag[u0:u1], rups.rake[u0:u1], rups.rate[u0:u1],
        rups.hypo[u0:u1], rups.lons[a:b], rups.lats[a:b], rups.deps[a:b],
        rups.shape[u0:u1], rups.kind,
        None if rups.probs_occur is None else rups.probs_occur[u0:u1],
        rups.rup_id[u0:u1])


def projection_frames(lons, lats, off):
    """
    The frames of the orthographic projections of S meshes (geo/utils.py:
    610-668: centred on the middle of the spherical bounding box of each
    mesh) -> (S, 9) float64 = the east, north and radial unit vectors at the
    centre; with them the projection of a unit vector p is (R p.e, R p.n).
    """
    from .area import get_spherical_bounding_box
    a = off[:-1]
    west, east = numpy.minimum.reduceat(lons, a), numpy.maximum.reduceat(
        lons, a)
    north, south = numpy.maximum.reduceat(lats, a), numpy.minimum.reduceat(
        lats, a)
    # meshes across the date line: the scalar rule of the reference
    for s in numpy.nonzero((east - west + 180) % 360 - 180 < 0)[0]:
        west[s], east[s], north[s], south[s] = get_spherical_bounding_box(
            lons[off[s]:off[s + 1]], lats[off[s]:off[s + 1]])
    lon0, lat0 = _middle_of_two(west, north, east, south)
    lam, phi = numpy.radians(lon0), numpy.radians(lat0)
    sl, cl, sp, cp = numpy.sin(lam), numpy.cos(lam), numpy.sin(phi), \
        numpy.cos(phi)
    zero = numpy.zeros_like(sl)
    return numpy.stack([-sl, cl, zero, -sp * cl, -sp * sl, cp,
                        cp * cl, cp * sl, sp], axis=1)


def rupture_planes(cmaker, rups, sitecol, names):
    """
    Distances of the mesh ruptures to the sites and the rupture parameters
    the models read, computed on the device.

    :param names: the distances wanted ('rrup', 'rjb', 'rx', 'ry0', 'repi',
        'rhypo'); rrup always comes
    :returns: (planes, par): planes = {name: CUDA tensor (U, N) float64},
        par = {'dip', 'strike', 'width', 'ztor', 'hypo_depth'} numpy (U,)
    """
    import ctypes
    import torch
    from .contexts import _dev, _ptr
    h = cmaker.handle()
    dev = h.device
    cuda = 'cuda:%d' % dev
    U, N = len(rups), len(sitecol)
    if U > 65535:
        raise ValueError('at most 65535 ruptures per call (chunk them)')
    names = set(names) | {'rrup'}
    st, keep = cmaker._sites_struct(sitecol, dev)
    t_lon, t_lat, t_dep = (_dev(x, dev) for x in (rups.lons, rups.lats,
                                                  rups.deps))
    t_off, t_shape = _dev(rups.off, dev), _dev(rups.shape, dev)
    P = int(rups.off[-1])
    planes = {}

    def plane():
        return torch.empty((U, N), dtype=torch.float64, device=cuda)

    # dip, strike, width: the top two rows / first two columns of a simple
    # fault rupture, the whole mesh otherwise
    def dsw(lo, la, de, off, shape):
        out = torch.empty((U, 3), dtype=torch.float64, device=cuda)
        ts = [_dev(x, dev) for x in (lo, la, de, off, shape)]
        h.call('oqg_mesh_dip_strike_width', *[_ptr(t) for t in ts], U,
               _ptr(out))
        return out.cpu().numpy()
    if rups.kind == 'simple':
        top = dsw(*rups._gather(rows=2))
        left = dsw(*rups._gather(cols=2))
        dip, strike, width = top[:, 0], top[:, 1], left[:, 2]
    else:
        whole = dsw(rups.lons, rups.lats, rups.deps, rups.off, rups.shape)
        dip, strike, width = whole.T
    par = dict(dip=dip, strike=strike, width=width, ztor=rups.ztor(),
               hypo_depth=rups.hypo[:, 2].copy())

    planes['rrup'] = plane()
    h.call('oqg_mesh_min_distance', _ptr(t_lon), _ptr(t_lat), _ptr(t_dep), P,
           _ptr(t_off), U, ctypes.byref(st), N, _ptr(planes['rrup']))
    if 'rjb' in names:
        owner = numpy.repeat(numpy.arange(U, dtype=numpy.int32),
                             numpy.diff(rups.off))
        frames = projection_frames(rups.lons, rups.lats, rups.off)
        t_owner, t_frame = _dev(owner, dev), _dev(frames, dev)
        planes['rjb'] = plane()
        h.call('oqg_mesh_rjb', _ptr(t_lon), _ptr(t_lat), P, _ptr(t_off),
               _ptr(t_shape), _ptr(t_owner), _ptr(t_frame), U,
               ctypes.byref(st), N, _ptr(planes['rjb']))
    if 'rx' in names or 'ry0' in names:
        elo, ela, eoff = rups.top_edges()
        t_elo, t_ela, t_eoff = _dev(elo, dev), _dev(ela, dev), _dev(eoff, dev)
        t_strike = _dev(strike, dev)
        if 'rx' in names:
            planes['rx'] = plane()
        if 'ry0' in names:
            planes['ry0'] = plane()
        h.call('oqg_mesh_rx_ry0', _ptr(t_elo), _ptr(t_ela), _ptr(t_eoff),
               _ptr(t_strike) if 'ry0' in names else None, U,
               ctypes.byref(st), N, _ptr(planes.get('rx')),
               _ptr(planes.get('ry0')))
    if 'repi' in names or 'rhypo' in names:
        # Point.distance_to_mesh (geo/point.py:170-197): the geodetic
        # distance of the hypocentre, combined with the depth difference
        # for rhypo (geodetic.distance, geodetic.py:138-151)
        hl, hp = (torch.deg2rad(_dev(rups.hypo[:, k], dev))[:, None]
                  for k in (0, 1))
        sl, sp = (torch.deg2rad(_dev(numpy.asarray(sitecol.array[k], float),
                                     dev))[None, :] for k in ('lon', 'lat'))
        arr = torch.sin((hp - sp) / 2.0) ** 2.0 + torch.cos(hp) * torch.cos(
            sp) * torch.sin((hl - sl) / 2.0) ** 2.0
        repi = 2 * EARTH_RADIUS * torch.arcsin(torch.sqrt(arr))
        if 'repi' in names:
            planes['repi'] = repi
        if 'rhypo' in names:
            dz = _dev(rups.hypo[:, 2], dev)[:, None] - _dev(numpy.asarray(
                sitecol.array['depth'], float), dev)[None, :]
            planes['rhypo'] = torch.sqrt(repi ** 2 + dz ** 2)
    torch.cuda.current_stream(dev).synchronize()
    del keep
    return planes, par


def _planes_of(cmaker, rups, sitecol, names):
    if isinstance(rups, MeshRuptures):
        return rupture_planes(cmaker, rups, sitecol, names)
    from . import multifault          # ruptures made of sections
    return multifault.rupture_planes(cmaker, rups, sitecol, names)


# bytes of distance planes per chunk of ruptures (HBM is 180 GB; the planes
# are scratch)
PLANE_BYTES = 2 << 30


def _chunks(cmaker, U, N, nplanes):
    per = max(1, min(65535, PLANE_BYTES // max(1, nplanes * N * 8)))
    return [(u0, min(U, u0 + per)) for u0 in range(0, U, per)]


def hazard_mesh(cmaker, pmap, rups, sitecol, count=False):
    """
    Fold the mesh ruptures into `pmap`: the equivalent of
    ``for ctx in cmaker.get_ctxs(rups, sitecol): cmaker.update(pmap, ctx)``
    (genctxs + update, contexts.py:959-1029, 1232-1247) with the distance
    planes in HBM and no context rows materialised.  Returns the number of
    rupture-site pairs kept if `count`.
    """
    import ctypes
    import torch
    from .contexts import _dev, _ptr
    from .tom import PoissonTOM
    h = cmaker.handle()
    dev = h.device
    tom = cmaker.tom or PoissonTOM(cmaker.investigation_time)
    cmaker.tom = tom
    U, N = len(rups), len(sitecol)
    if U == 0 or N == 0:
        return 0
    names = sorted(cmaker.REQUIRES_DISTANCES | {'rrup'})
    order = ['rrup', 'rjb', 'rx', 'ry0', 'repi', 'rhypo']
    bad = set(names) - set(order)
    if bad:
        raise NotImplementedError('distances %s of mesh ruptures' % bad)
    st, keep = cmaker._sites_struct(sitecol, dev)
    t_sidx = _dev(pmap.sidx[sitecol.sids].astype(numpy.uint32), dev)
    on_host = pmap.device is None
    t_map = _dev(pmap.array, dev) if on_host else pmap.array
    t_cnt = torch.zeros(2, dtype=torch.int64, device='cuda:%d' % dev) \
        if count else None
    multi = not isinstance(rups, MeshRuptures)     # MultiFaultRuptures
    for u0, u1 in ([(0, U)] if multi else _chunks(cmaker, U, N, len(names))):
        r = _slice(rups, u0, u1) if (u0, u1) != (0, U) else rups
        planes, par = _planes_of(cmaker, r, sitecol, names)
        K, t_probs = 0, None
        if cmaker.cluster:
            if pmap.rates:
                raise ValueError('a cluster needs a map of probabilities')
            t_probs, K = cmaker._fated(len(r), dev)
        elif r.probs_occur is not None:
            K = r.probs_occur.shape[1]
            t_probs = _dev(r.probs_occur, dev)
        elif hasattr(tom, 'get_pmf'):
            probs = numpy.ascontiguousarray(tom.get_pmf(r.rate))
            K = probs.shape[1]
            t_probs = _dev(probs, dev)
        rp = numpy.stack([r.mag, r.rake, par['dip'], par['ztor'],
                          par['width'], par['hypo_depth'], r.rate,
                          cmaker.magdist(r.mag)], axis=1)
        t_rp = _dev(rp, dev)
        ptrs = (ctypes.c_void_p * 6)(*[
            planes[n].data_ptr() if n in planes else None for n in order])
        with cmaker.fused_mon:
            h.call('oqg_hazard_dist', _ptr(t_rp), ptrs, _ptr(t_probs), K,
                   len(r), ctypes.byref(st), _ptr(t_sidx), N,
                   float(cmaker.minimum_distance),
                   0. if cmaker.cluster else float(tom.time_span),
                   int(pmap.rates), _ptr(t_map), _ptr(t_cnt))
        torch.cuda.current_stream(dev).synchronize()
    if on_host:
        pmap.array[:] = t_map.cpu().numpy()
    if count:
        return int(t_cnt[0].item())


def get_ctxs_mesh(cmaker, rups, sitecol, src_id=0, step=1):
    """
    Context rows of mesh ruptures: one recarray per rupture with at least
    one site within the integration distance, in the order of the
    reference's get_ctxs for a non-point source (contexts.py:1071-1099:
    ruptures sorted by magnitude; genctxs :959-1029: rrup <= magdist,
    sites filtered, one context per rupture).
    """
    U, N = len(rups), len(sitecol)
    if U == 0 or N == 0:
        return []
    cmaker.fewsites = N <= cmaker.max_sites_disagg
    names = sorted(cmaker.REQUIRES_DISTANCES | {'rrup'})
    md = cmaker.maximum_distance
    minmag, maxmag = (md.mags[0], md.mags[-1]) if hasattr(md, 'mags') else (
        -numpy.inf, numpy.inf)
    ok = numpy.nonzero((rups.mag >= minmag) & (rups.mag <= maxmag))[0]
    order = ok[numpy.argsort(rups.mag[ok], kind='stable')][::step]
    ctxs = []
    tom = cmaker.tom
    multi = not isinstance(rups, MeshRuptures)     # MultiFaultRuptures
    for u0, u1 in ([(0, U)] if multi else _chunks(cmaker, U, N, len(names))):
        r = _slice(rups, u0, u1) if (u0, u1) != (0, U) else rups
        planes, par = _planes_of(cmaker, r, sitecol, names)
        host = {n: t.cpu().numpy() for n, t in planes.items()}
        probs = r.probs_occur
        if probs is None and tom is not None and hasattr(tom, 'get_pmf'):
            probs = tom.get_pmf(r.rate)
        K = 0 if probs is None else probs.shape[1]
        chunk = {}
        for u in range(len(r)):
            mask = host['rrup'][u] <= cmaker.maximum_distance(r.mag[u])
            if not mask.any():
                continue
            ctx = cmaker.new_ctx(int(mask.sum()), K)
            for name in names:
                arr = host[name][u][mask]
                if cmaker.minimum_distance and \
                        name in cmaker.REQUIRES_DISTANCES:
                    arr = numpy.where(arr < cmaker.minimum_distance,
                                      cmaker.minimum_distance, arr)
                ctx[name] = arr
            vals = dict(mag=r.mag[u], rake=r.rake[u], dip=par['dip'][u],
                        strike=par['strike'][u], width=par['width'][u],
                        ztor=par['ztor'][u], hypo_depth=par['hypo_depth'][u])
            for name in cmaker.REQUIRES_RUPTURE_PARAMETERS:
                ctx[name] = vals[name]
            for name in cmaker.REQUIRES_SITES_PARAMETERS:
                ctx[name] = sitecol.array[name][mask]
            ctx['occurrence_rate'] = r.rate[u]
            ctx['rup_id'] = r.rup_id[u]
            ctx['src_id'] = src_id
            ctx['sids'] = sitecol.sids[mask]
            if K:
                ctx['probs_occur'] = probs[u]
            chunk[u0 + u] = ctx.view(numpy.recarray)
        ctxs.append(chunk)
    alld = {}
    for c in ctxs:
        alld.update(c)
    return [alld[u] for u in order if u in alld]
