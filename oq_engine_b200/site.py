"""A minimal site collection: the struct-of-arrays the hot path reads
(openquake/hazardlib/site.py:249-324): sids, lon, lat, depth and the site
parameters vs30, vs30measured, z1pt0, z2pt5."""
import numpy

EARTH_RADIUS = 6371.0

site_dt = numpy.dtype([
    ('sids', numpy.uint32), ('lon', float), ('lat', float), ('depth', float),
    ('vs30', float), ('vs30measured', bool), ('z1pt0', float),
    ('z2pt5', float)])


class SiteCollection(object):
    def __init__(self, array):
        self.array = array

    @classmethod
    def from_points(cls, lons, lats, depths=None, sitemodel=None,
                    req_site_params=()):
        """Same call as the reference's SiteCollection.from_points;
        `sitemodel` is any object (or dict) with the site parameters as
        scalars or arrays"""
        n = len(lons)
        arr = numpy.zeros(n, site_dt)
        arr['sids'] = numpy.arange(n, dtype=numpy.uint32)
        arr['lon'], arr['lat'] = lons, lats
        arr['depth'] = 0 if depths is None else depths
        arr['z1pt0'] = arr['z2pt5'] = -999.
        for name in ('vs30', 'vs30measured', 'z1pt0', 'z2pt5'):
            if sitemodel is None:
                continue
            val = (sitemodel.get(name) if isinstance(sitemodel, dict)
                   else getattr(sitemodel, name, None))
            if val is not None:
                arr[name] = val
        return cls(arr)

    def __len__(self):
        return len(self.array)

    def __getattr__(self, name):
        if name != 'array' and name in site_dt.names:
            return self.array[name]
        raise AttributeError(name)

    @property
    def sids(self):
        return self.array['sids']

    @property
    def xyz(self):
        """(N, 3) cartesian coordinates in km (site.py:316-324,
        geo/utils.py spherical_to_cartesian)"""
        lon, lat = numpy.radians(self.array['lon']), numpy.radians(
            self.array['lat'])
        rr = EARTH_RADIUS - self.array['depth']
        return numpy.stack([rr * numpy.cos(lat) * numpy.cos(lon),
                            rr * numpy.cos(lat) * numpy.sin(lon),
                            rr * numpy.sin(lat)], axis=1)

    def filter(self, mask):
        return SiteCollection(self.array[mask])

    def split(self, ntiles):
        """Tiles by sids % ntiles like TileGetter (site.py:133-161)"""
        return [self.filter(self.array['sids'] % ntiles == k)
                for k in range(ntiles)]
