"""Area sources on the host: the discretisation of the polygon into point
sources, mirroring openquake/hazardlib/source/area.py:232-290 and
geo/polygon.py:163-184, 232-264 (with geo/utils.py:538-607, 671-685 and
geo/geodetic.py:64-103, 133-140, 426-475, 520-555).

The only step of the reference that needs shapely is the point-in-polygon
test in the orthographic projection (polygon.py:252); an even-odd ray cast
replaces it here.  Everything downstream (the ruptures of the point sources,
the hazard) runs on the device (source.PointSourceBatch).
"""
import math
import numpy

from .source import (PointSource, TruncatedGRMFD, EvenlyDiscretizedMFD,
                     ArbitraryMFD)

EARTH_RADIUS = 6371.0
UPSAMPLING_STEP_KM = 100


def get_longitudinal_extent(lon1, lon2):
    return (lon2 - lon1 + 180) % 360 - 180


def fast_distance(lon, lat, lons, lats):
    lon, lat = math.radians(lon), math.radians(lat)
    lons, lats = numpy.radians(lons), numpy.radians(lats)
    return 2 * EARTH_RADIUS * numpy.arcsin(numpy.sqrt(
        numpy.sin((lat - lats) / 2.0) ** 2.0 + numpy.cos(lat) *
        numpy.cos(lats) * numpy.sin((lon - lons) / 2.0) ** 2.0))


def geodetic_distance(lons1, lats1, lons2, lats2):
    # geodetic.py:64-90: float32 for arrays "to save memory"
    if isinstance(lons1, float):
        return fast_distance(lons1, lats1, lons2, lats2)
    f32 = numpy.float32
    lons1, lats1, lons2, lats2 = (f32(numpy.radians(x))
                                  for x in (lons1, lats1, lons2, lats2))
    arr = numpy.sin((lats1 - lats2) / 2.0) ** 2.0
    arr += numpy.cos(lats1) * numpy.cos(lats2) * numpy.sin(
        (lons1 - lons2) / 2.0) ** 2.0
    return 2 * EARTH_RADIUS * numpy.arcsin(numpy.sqrt(arr))


def azimuth(lons1, lats1, lons2, lats2):
    lons1, lats1, lons2, lats2 = (numpy.radians(x)
                                  for x in (lons1, lats1, lons2, lats2))
    cos_lat2 = numpy.cos(lats2)
    true_course = numpy.degrees(numpy.arctan2(
        numpy.sin(lons1 - lons2) * cos_lat2,
        numpy.cos(lats1) * numpy.sin(lats2) -
        numpy.sin(lats1) * cos_lat2 * numpy.cos(lons1 - lons2)))
    return (360 - true_course) % 360


def npoints_towards(lon, lat, azim, hdist, npoints):
    # geodetic.py:426-475 without depths
    rlon, rlat = numpy.radians(lon), numpy.radians(lat)
    tc = numpy.radians(360 - azim)
    hdists = numpy.arange(npoints, dtype=float)
    hdists *= (hdist / EARTH_RADIUS) / (npoints - 1)
    sin_d, cos_d = numpy.sin(hdists), numpy.cos(hdists)
    sin_lat, cos_lat = numpy.sin(rlat), numpy.cos(rlat)
    sin_lats = sin_lat * cos_d + cos_lat * sin_d * numpy.cos(tc)
    lats = numpy.degrees(numpy.arcsin(sin_lats))
    dlon = numpy.arctan2(numpy.sin(tc) * sin_d * cos_lat,
                         cos_d - sin_lat * sin_lats)
    lons = numpy.degrees(numpy.mod(rlon - dlon + numpy.pi, 2 * numpy.pi)
                         - numpy.pi)
    return lons, lats


def point_at(lon, lat, azim, distance):
    lons, lats = npoints_towards(lon, lat, azim, 2 * distance, 3)
    return lons[1], lats[1]


def get_resampled_coordinates(lons, lats):
    # polygon.py:267-310
    n = len(lons)
    lons1, lats1 = numpy.array(lons), numpy.array(lats)
    lons2 = numpy.concatenate((lons1[1:], lons1[:1]))
    lats2 = numpy.concatenate((lats1[1:], lats1[:1]))
    distances = geodetic_distance(lons1, lats1, lons2, lats2)
    rlons, rlats = [lons[0]], [lats[0]]
    for i in range(n):
        j = (i + 1) % n
        num_points = int(distances[i] / UPSAMPLING_STEP_KM) + 1
        if num_points >= 2:
            hdist = geodetic_distance(float(lons[i]), float(lats[i]),
                                      float(lons[j]), float(lats[j]))
            nl, nt = npoints_towards(
                lons[i], lats[i], azimuth(lons[i], lats[i], lons[j], lats[j]),
                hdist, num_points)
            nl[-1], nt[-1] = lons[j], lats[j]
            rlons.extend(nl[1:])
            rlats.extend(nt[1:])
        else:
            rlons.append(lons[j])
            rlats.append(lats[j])
    return numpy.array(rlons[:-1]), numpy.array(rlats[:-1])


def get_spherical_bounding_box(lons, lats):
    north, south = lats.max(), lats.min()
    west, east = lons.min(), lons.max()
    if get_longitudinal_extent(west, east) < 0:
        west = lons[lons > 0].min()
        east = lons[lons < 0].max()
        if not ((get_longitudinal_extent(west, lons) >= 0) &
                (get_longitudinal_extent(lons, east) >= 0)).all():
            raise ValueError('points collection has longitudinal extent '
                             'wider than 180 degrees')
    return west, east, north, south


class OrthographicProjection(object):
    """geo/utils.py:599-668, forward direction"""
    def __init__(self, west, east, north, south):
        if west == east and north == south:
            lon0, lat0 = west, north
        else:
            dist = geodetic_distance(float(west), float(north), float(east),
                                     float(south))
            lon0, lat0 = point_at(west, north,
                                  azimuth(west, north, east, south),
                                  dist / 2.0)
        self.lam0, self.phi0 = numpy.radians(lon0), numpy.radians(lat0)

    def __call__(self, lons, lats):
        lambdas, phis = numpy.radians(lons), numpy.radians(lats)
        lambdas = lambdas - self.lam0
        xx = numpy.cos(phis) * numpy.sin(lambdas) * EARTH_RADIUS
        yy = (math.cos(self.phi0) * numpy.sin(phis) - math.sin(self.phi0) *
              numpy.cos(phis) * numpy.cos(lambdas)) * EARTH_RADIUS
        return xx, yy


def contains(xy, x, y):
    """even-odd ray cast: is (x, y) strictly inside the polygon xy (K, 2)"""
    inside = False
    n = len(xy)
    for i in range(n):
        x1, y1 = xy[i]
        x2, y2 = xy[(i + 1) % n]
        if (y1 > y) != (y2 > y) and x < (x2 - x1) * (y - y1) / (y2 - y1) + x1:
            inside = not inside
    return inside


def discretize(lons, lats, mesh_spacing):
    """
    Polygon.discretize, polygon.py:232-264: points of a mesh with
    `mesh_spacing` km between them that lie inside the polygon, from the
    highest latitude to the lowest and from west to east.
    """
    plons, plats = get_resampled_coordinates(lons, lats)
    west, east, north, south = bbox = get_spherical_bounding_box(plons, plats)
    proj = OrthographicProjection(*bbox)
    xy = numpy.column_stack(proj(plons, plats))
    out_lons, out_lats = [], []
    latitude = north
    while latitude > south:
        longitude = west
        while get_longitudinal_extent(longitude, east) > 0:
            x, y = proj(longitude, latitude)
            if contains(xy, x, y):
                out_lons.append(longitude)
                out_lats.append(latitude)
            longitude, _ = point_at(longitude, latitude, 90, mesh_spacing)
        _, latitude = point_at(west, latitude, 180, mesh_spacing)
    if not out_lons:
        raise ValueError("The area_source_discretization is too large "
                         "(%d km)" % mesh_spacing)
    return numpy.array(out_lons), numpy.array(out_lats)


class AreaSource(object):
    """
    source/area.py: uniform seismicity over a polygon.  Iterating yields the
    point sources of the discretised polygon, all sharing one rescaled MFD
    (area.py:232-290), nodal-plane and depth distribution, which
    PointSourceBatch stores once.
    """
    def __init__(self, source_id, polygon_lons, polygon_lats,
                 area_discretization, mfd, msr, rupture_aspect_ratio,
                 upper_seismogenic_depth, lower_seismogenic_depth,
                 nodal_plane_distribution, hypocenter_distribution,
                 hypo_dip_fracs=None, tectonic_region_type='*'):
        self.source_id = source_id
        self.polygon_lons = numpy.asarray(polygon_lons, float)
        self.polygon_lats = numpy.asarray(polygon_lats, float)
        if len(self.polygon_lons) < 3:
            raise ValueError('polygon must have at least 3 points')
        self.area_discretization = area_discretization
        self.mfd, self.msr = mfd, msr
        self.args = (rupture_aspect_ratio, upper_seismogenic_depth,
                     lower_seismogenic_depth, list(nodal_plane_distribution),
                     list(hypocenter_distribution), hypo_dip_fracs,
                     tectonic_region_type)
        self.tectonic_region_type = tectonic_region_type

    def rescaled_mfd(self, num_points):
        m = self.mfd
        if isinstance(m, TruncatedGRMFD):
            return TruncatedGRMFD(m.min_mag, m.max_mag, m.bin_width,
                                  m.a_val - math.log10(num_points), m.b_val)
        if isinstance(m, EvenlyDiscretizedMFD):
            return EvenlyDiscretizedMFD(
                m.min_mag, m.bin_width,
                [x / num_points for x in m.occurrence_rates])
        if isinstance(m, ArbitraryMFD):
            return ArbitraryMFD(m.magnitudes,
                                [x / num_points for x in m.occurrence_rates])
        raise NotImplementedError('MFD %s' % type(m).__name__)

    def __iter__(self):
        lons, lats = discretize(self.polygon_lons, self.polygon_lats,
                                self.area_discretization)
        mfd = self.rescaled_mfd(len(lons))
        for i, (lon, lat) in enumerate(zip(lons, lats)):
            yield PointSource('%s:%d' % (self.source_id, i), lon, lat, mfd,
                              self.msr, *self.args)

    def count_ruptures(self):
        return sum(src.count_ruptures() for src in self)
