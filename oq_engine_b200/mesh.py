"""Mesh surfaces, first step: the rupture
distance.  `Mesh.get_min_distance` mirrors
openquake/hazardlib/geo/mesh.py:365-383; `mesh_min_distance` evaluates it
for many meshes (the sections of multi-fault sources, or the ruptures of
fault sources) in one call and `section_min` composes the per-section
distances of multi-fault ruptures (contexts.py:969-971).  `mesh_rx_ry0` is
BaseSurface.get_rx_distance / get_ry0_distance
(geo/surface/base.py:319-384, :193-239) for many top edges in one call."""
import ctypes
import numpy


class Mesh(object):
    """Points on or below the Earth surface: lons, lats (degrees), depths
    (km, optional)"""
    def __init__(self, lons, lats, depths=None):
        self.lons = numpy.asarray(lons, float).ravel()
        self.lats = numpy.asarray(lats, float).ravel()
        self.depths = None if depths is None else \
            numpy.asarray(depths, float).ravel()
        assert len(self.lons) == len(self.lats)
        assert self.depths is None or len(self.depths) == len(self.lons)

    def __len__(self):
        return len(self.lons)

    def get_min_distance(self, mesh, device=None):
        """minimum distance from this mesh to each point of `mesh` (a Mesh
        or a SiteCollection), km, shape (len(mesh),)"""
        return mesh_min_distance([self], mesh, device)[0]


def _sites_struct(points, dev):
    from . import _lib
    from .contexts import _dev
    st = _lib.OqgSites()
    keep = []
    if hasattr(points, 'array'):        # SiteCollection
        lon, lat, dep = (points.array[n] for n in ('lon', 'lat', 'depth'))
    else:
        lon, lat = points.lons, points.lats
        dep = points.depths if points.depths is not None else \
            numpy.zeros(len(lon))
    for name, arr in (('lon', lon), ('lat', lat), ('depth', dep)):
        t = _dev(numpy.asarray(arr, float), dev)
        keep.append(t)
        setattr(st, name, t.data_ptr())
    return st, keep, len(lon)


def mesh_min_distance(meshes, points, device=None, as_tensor=False):
    """
    :param meshes: a list of S Mesh objects
    :param points: a Mesh or a SiteCollection of N points
    :returns: (S, N) float64 minimum distances in km
    """
    import torch
    from .contexts import get_handle, _device_index, _dev, _ptr
    h = get_handle(_device_index(device))
    dev = h.device
    S = len(meshes)
    off = numpy.concatenate([[0], numpy.cumsum([len(m) for m in meshes])]
                            ).astype(numpy.int64)
    if S == 0 or off[-1] == 0:
        raise ValueError('empty mesh')
    has_dep = any(m.depths is not None for m in meshes)
    t_lon = _dev(numpy.concatenate([m.lons for m in meshes]), dev)
    t_lat = _dev(numpy.concatenate([m.lats for m in meshes]), dev)
    t_dep = _dev(numpy.concatenate([
        m.depths if m.depths is not None else numpy.zeros(len(m))
        for m in meshes]), dev) if has_dep else None
    t_off = _dev(off, dev)
    st, keep, N = _sites_struct(points, dev)
    out = torch.empty((S, N), dtype=torch.float64, device='cuda:%d' % dev)
    for s0 in range(0, S, 65535):       # grid.y limit
        s1 = min(S, s0 + 65535)
        h.call('oqg_mesh_min_distance', _ptr(t_lon), _ptr(t_lat), _ptr(t_dep),
               int(off[-1]), ctypes.c_void_p(t_off.data_ptr() + 8 * s0),
               s1 - s0, ctypes.byref(st), N,
               ctypes.c_void_p(out.data_ptr() + 8 * N * s0))
    torch.cuda.current_stream(dev).synchronize()
    return out if as_tensor else out.cpu().numpy()


def section_min(sec_dist, rup_sections, device=None):
    """
    :param sec_dist: (S, N) distances of the sections (numpy or CUDA tensor)
    :param rup_sections: for each rupture the list of its section indices
    :returns: (U, N) float64, the minimum over the sections of each rupture
    """
    import torch
    from .contexts import get_handle, _device_index, _dev, _ptr
    h = get_handle(_device_index(device))
    dev = h.device
    t_sec = _dev(sec_dist, dev, torch.float64)
    S, N = t_sec.shape
    U = len(rup_sections)
    roff = numpy.concatenate([[0], numpy.cumsum([len(r) for r in rup_sections])
                              ]).astype(numpy.int64)
    rsec = numpy.concatenate([numpy.asarray(r, numpy.int32)
                              for r in rup_sections]).astype(numpy.int32)
    if len(rsec) and (rsec.min() < 0 or rsec.max() >= S):
        raise ValueError('section index out of range')
    if (numpy.diff(roff) == 0).any():
        raise ValueError('a rupture without sections')
    t_roff, t_rsec = _dev(roff, dev), _dev(rsec, dev)
    out = torch.empty((U, N), dtype=torch.float64, device='cuda:%d' % dev)
    for u0 in range(0, U, 65535):
        u1 = min(U, u0 + 65535)
        h.call('oqg_section_min', _ptr(t_sec), S, N,
               ctypes.c_void_p(t_roff.data_ptr() + 8 * u0), _ptr(t_rsec),
               u1 - u0, ctypes.c_void_p(out.data_ptr() + 8 * N * u0))
    torch.cuda.current_stream(dev).synchronize()
    return out.cpu().numpy()


def mesh_rx_ry0(top_edges, strikes, points, device=None, rx=True, ry0=True):
    """
    :param top_edges: a list of S top edges (the first row of the surface
        meshes): Mesh objects or (lons, lats) pairs of at least two vertices
    :param strikes: the S mean strikes (surface.get_strike()), degrees; may
        be None when ry0 is False
    :param points: a Mesh or a SiteCollection of N points
    :returns: (rx, ry0), each (S, N) float64 km or None if not asked
    """
    import torch
    from .contexts import get_handle, _device_index, _dev, _ptr
    h = get_handle(_device_index(device))
    dev = h.device
    edges = [e if isinstance(e, Mesh) else Mesh(*e) for e in top_edges]
    S = len(edges)
    if S == 0:
        raise ValueError('no surface')
    if any(len(e) < 2 for e in edges):
        # the reference fails in _get_p1_p2 (base.py:126-138)
        raise ValueError('a top edge needs at least two points')
    if ry0:
        if strikes is None or len(strikes) != S:
            raise ValueError('ry0 needs one mean strike per surface')
        t_strike = _dev(numpy.asarray(strikes, float), dev)
    else:
        t_strike = None
    off = numpy.concatenate([[0], numpy.cumsum([len(e) for e in edges])]
                            ).astype(numpy.int64)
    t_lon = _dev(numpy.concatenate([e.lons for e in edges]), dev)
    t_lat = _dev(numpy.concatenate([e.lats for e in edges]), dev)
    t_off = _dev(off, dev)
    st, keep, N = _sites_struct(points, dev)
    cuda = 'cuda:%d' % dev
    t_rx = torch.empty((S, N), dtype=torch.float64, device=cuda) if rx \
        else None
    t_ry0 = torch.empty((S, N), dtype=torch.float64, device=cuda) if ry0 \
        else None
    for s0 in range(0, S, 65535):       # grid.y limit
        s1 = min(S, s0 + 65535)
        h.call('oqg_mesh_rx_ry0', _ptr(t_lon), _ptr(t_lat),
               ctypes.c_void_p(t_off.data_ptr() + 8 * s0),
               ctypes.c_void_p(t_strike.data_ptr() + 8 * s0) if ry0 else None,
               s1 - s0, ctypes.byref(st), N,
               ctypes.c_void_p(t_rx.data_ptr() + 8 * N * s0) if rx else None,
               ctypes.c_void_p(t_ry0.data_ptr() + 8 * N * s0) if ry0 else None)
    torch.cuda.current_stream(dev).synchronize()
    out = []
    for t, name in ((t_rx, 'Rx'), (t_ry0, 'Ry0')):
        a = None if t is None else t.cpu().numpy()
        if a is not None and numpy.isnan(a).any():
            raise ValueError('NaN in %s' % name)      # base.py:236, :382
        out.append(a)
    return tuple(out)


def mesh_dip_strike_width(meshes, device=None):
    """
    Mean inclination, azimuth and width of rectangular meshes
    (RectangularMesh.get_mean_inclination_and_azimuth, geo/mesh.py:728-783,
    and get_mean_width, :861-885).

    :param meshes: a list of S (lons, lats, depths) triples of 2D arrays
        (rows, cols), rows and cols >= 2, rows not shallower than the
        previous one
    :returns: (S, 3) float64: dip and strike in degrees, width in km
    """
    import torch
    from .contexts import get_handle, _device_index, _dev, _ptr
    h = get_handle(_device_index(device))
    dev = h.device
    S = len(meshes)
    if S == 0:
        raise ValueError('no mesh')
    lon, lat, dep, shape = [], [], [], []
    for lons, lats, deps in meshes:
        lons = numpy.asarray(lons, float)
        deps = numpy.zeros_like(lons) if deps is None else \
            numpy.asarray(deps, float)
        if lons.ndim != 2 or 1 in lons.shape:
            raise ValueError(
                'inclination and azimuth are only defined for mesh of more '
                'than one row and more than one column of points')
        if not ((deps[1:] - deps[:-1]) >= 0).all():
            raise ValueError('a mesh row is shallower than the previous one')
        lon.append(lons.ravel())
        lat.append(numpy.asarray(lats, float).ravel())
        dep.append(deps.ravel())
        shape.append(lons.shape)
    off = numpy.concatenate([[0], numpy.cumsum([len(x) for x in lon])]
                            ).astype(numpy.int64)
    t_lon, t_lat, t_dep = (_dev(numpy.concatenate(x), dev)
                           for x in (lon, lat, dep))
    t_off = _dev(off, dev)
    t_shape = _dev(numpy.array(shape, numpy.int32), dev)
    out = torch.empty((S, 3), dtype=torch.float64, device='cuda:%d' % dev)
    h.call('oqg_mesh_dip_strike_width', _ptr(t_lon), _ptr(t_lat), _ptr(t_dep),
           _ptr(t_off), _ptr(t_shape), S, _ptr(out))
    torch.cuda.current_stream(dev).synchronize()
    return out.cpu().numpy()


def projection_frames(meshes):
    """
    The frames of the orthographic projections of S meshes
    (geo/utils.py:610-668: centred on the middle point of the spherical
    bounding box of each mesh): (S, 9) float64 = the east, north and radial
    unit vectors at the centre.  With them the projection of a point p (unit
    vector) is (R p.e, R p.n).
    """
    from .area import OrthographicProjection, get_spherical_bounding_box
    out = numpy.zeros((len(meshes), 9))
    for s, (lons, lats) in enumerate(meshes):
        proj = OrthographicProjection(*get_spherical_bounding_box(
            numpy.asarray(lons, float).ravel(),
            numpy.asarray(lats, float).ravel()))
        sl, cl = numpy.sin(proj.lam0), numpy.cos(proj.lam0)
        sp, cp = numpy.sin(proj.phi0), numpy.cos(proj.phi0)
        out[s, 0:3] = -sl, cl, 0.
        out[s, 3:6] = -sp * cl, -sp * sl, cp
        out[s, 6:9] = cp * cl, cp * sl, sp
    return out


def mesh_rjb(meshes, points, device=None, as_tensor=False):
    """
    RectangularMesh.get_joyner_boore_distance (geo/mesh.py:453-540) for S
    rectangular meshes at once.

    :param meshes: a list of S (lons, lats) pairs of 2D arrays (rows, cols);
        a 1D array is a mesh of one row
    :param points: a Mesh or a SiteCollection of N points
    :returns: (S, N) float64 Joyner-Boore distances in km (zero inside the
        surface projection of a mesh)
    """
    import torch
    from .contexts import get_handle, _device_index, _dev, _ptr
    h = get_handle(_device_index(device))
    dev = h.device
    S = len(meshes)
    if S == 0:
        raise ValueError('no mesh')
    lon, lat, shape, owner = [], [], [], []
    for s, (lons, lats) in enumerate(meshes):
        lons = numpy.atleast_2d(numpy.asarray(lons, float))
        lats = numpy.atleast_2d(numpy.asarray(lats, float))
        assert lons.shape == lats.shape and lons.ndim == 2, lons.shape
        lon.append(lons.ravel())
        lat.append(lats.ravel())
        shape.append(lons.shape)
        owner.append(numpy.full(lons.size, s, numpy.int32))
    off = numpy.concatenate([[0], numpy.cumsum([len(x) for x in lon])]
                            ).astype(numpy.int64)
    frames = projection_frames(list(zip(lon, lat)))
    t_lon, t_lat = _dev(numpy.concatenate(lon), dev), _dev(
        numpy.concatenate(lat), dev)
    t_off, t_shape = _dev(off, dev), _dev(numpy.array(shape, numpy.int32), dev)
    t_owner, t_frame = _dev(numpy.concatenate(owner), dev), _dev(frames, dev)
    st, keep, N = _sites_struct(points, dev)
    out = torch.empty((S, N), dtype=torch.float64, device='cuda:%d' % dev)
    h.call('oqg_mesh_rjb', _ptr(t_lon), _ptr(t_lat), int(off[-1]), _ptr(t_off),
           _ptr(t_shape), _ptr(t_owner), _ptr(t_frame), S, ctypes.byref(st),
           N, _ptr(out))
    torch.cuda.current_stream(dev).synchronize()
    return out if as_tensor else out.cpu().numpy()
