"""Point sources: the host-side description the rupture generator on the
device consumes.

Mirrors the slice of the reference a classical calculation touches before the
hot path: the magnitude-frequency distributions
(openquake/hazardlib/mfd/truncated_gr.py:103-164, evenly_discretized.py,
arbitrary_mfd.py), `PointSource` with its nodal-plane and hypocentral-depth
distributions (source/point.py:60-196) and `get_annual_occurrence_rates`
(source/base.py:459-477).  The ruptures themselves (`get_planar`,
point.py:258-289) are built on the device by `oqg_point_ruptures`; nothing
here loops over ruptures.
"""
import math
import ctypes
import numpy

from . import _lib

MSR = {'WC1994': 0, 'PointMSR': 1, 'PeerMSR': 2, 'CScalingMSR': 3}


def parse_msr(msr):
    """(name, parameter) of a magnitude-scaling relationship given as the
    reference writes it in the source model: 'WC1994', 'PointMSR', 'PeerMSR'
    or the TOML form 'CScalingMSR.C = 3.90' (scalerel/cscaling.py:36-40).
    Anything else is an error: there is no device implementation to fall
    back from."""
    name, _, rest = str(msr).partition('.')
    name = name.strip()
    if name not in MSR or (rest and name != 'CScalingMSR'):
        raise NotImplementedError(
            'magnitude-scaling relationship %s has no device '
            'implementation' % msr)
    if name == 'CScalingMSR':
        key, _, val = rest.partition('=')
        if key.strip() != 'C' or not val.strip():
            raise ValueError('expected "CScalingMSR.C = <value>", got %r'
                             % msr)
        return name, float(val)
    return name, 4.0 if name == 'PeerMSR' else 0.


def median_area(msr, mags, rake):
    """get_median_area of the supported relationships (scalerel/wc1994.py,
    point.py:40-52, peer.py:32-36, cscaling.py:33-34)"""
    from .planar import wc1994_median_area
    name, par = parse_msr(msr)
    mags = numpy.asarray(mags, float)
    if name == 'WC1994':
        return wc1994_median_area(mags, numpy.full(mags.shape, rake))
    if name == 'PointMSR':
        return numpy.full(mags.shape, 1e-4)
    return 10. ** (mags - par)


def _round(x):
    # baselib/general.py:566-571: the Python-2 rounding the MFDs use
    return float(math.floor(x + math.copysign(0.5, x)))


class TruncatedGRMFD(object):
    """mfd/truncated_gr.py: rates of a doubly truncated Gutenberg-Richter
    distribution evenly discretised in bins of `bin_width`"""
    def __init__(self, min_mag, max_mag, bin_width, a_val, b_val):
        if not bin_width > 0:
            raise ValueError('bin width %g must be positive' % bin_width)
        if not min_mag >= 0:
            raise ValueError('minimum magnitude %g must be non-negative'
                             % min_mag)
        if not max_mag >= min_mag + bin_width:
            raise ValueError('maximum magnitude %g must be higher than '
                             'minimum magnitude %g by bin width %g at least'
                             % (max_mag, min_mag, bin_width))
        if b_val <= 0:
            raise ValueError('b-value %g must be non-negative' % b_val)
        self.min_mag, self.max_mag, self.bin_width = min_mag, max_mag, bin_width
        self.a_val, self.b_val = a_val, b_val

    def get_annual_occurrence_rates(self):
        bw = self.bin_width
        lo = _round(self.min_mag / bw) * bw
        hi = _round(self.max_mag / bw) * bw
        if lo != hi:
            lo += bw / 2.0
            hi -= bw / 2.0
        nbins = int(_round((hi - lo) / bw)) + 1
        out, mag = [], lo
        for _ in range(nbins):
            out.append((mag, 10 ** (self.a_val - self.b_val * (mag - bw / 2.0))
                        - 10 ** (self.a_val - self.b_val * (mag + bw / 2.0))))
            mag += bw
        return out


class EvenlyDiscretizedMFD(object):
    """mfd/evenly_discretized.py"""
    def __init__(self, min_mag, bin_width, occurrence_rates):
        if not bin_width > 0:
            raise ValueError('bin width must be positive')
        if not len(occurrence_rates):
            raise ValueError('at least one bin must be specified')
        if not all(r >= 0 for r in occurrence_rates):
            raise ValueError('all occurrence rates must not be negative')
        self.min_mag, self.bin_width = min_mag, bin_width
        self.occurrence_rates = list(occurrence_rates)

    def get_annual_occurrence_rates(self):
        return [(self.min_mag + i * self.bin_width, r)
                for i, r in enumerate(self.occurrence_rates)]


class ArbitraryMFD(object):
    """mfd/arbitrary_mfd.py"""
    def __init__(self, magnitudes, occurrence_rates):
        if len(magnitudes) != len(occurrence_rates):
            raise ValueError('magnitudes and rates of different length')
        self.magnitudes = list(magnitudes)
        self.occurrence_rates = list(occurrence_rates)

    def get_annual_occurrence_rates(self):
        return list(zip(self.magnitudes, self.occurrence_rates))


class PointSource(object):
    """
    source/point.py PointSource, the attributes the rupture generation reads.

    :param mfd: an object with get_annual_occurrence_rates()
    :param msr: 'WC1994', 'PointMSR', 'PeerMSR' or 'CScalingMSR.C = <C>'
        (anything else is an error: there is no device implementation to
        fall back from)
    :param nodal_plane_distribution: [(prob, strike, dip, rake), ...]
    :param hypocenter_distribution: [(prob, depth), ...]
    :param hypo_dip_fracs: optional, one down-dip fraction per depth
    """
    def __init__(self, source_id, lon, lat, mfd, msr, rupture_aspect_ratio,
                 upper_seismogenic_depth, lower_seismogenic_depth,
                 nodal_plane_distribution, hypocenter_distribution,
                 hypo_dip_fracs=None, tectonic_region_type='*'):
        parse_msr(msr)      # NotImplementedError if unsupported
        if not lower_seismogenic_depth > upper_seismogenic_depth:
            raise ValueError('lower seismogenic depth must be below '
                             'upper seismogenic depth')
        for _p, depth in hypocenter_distribution:
            if depth > lower_seismogenic_depth:
                raise ValueError('source_id=%r depth=%r is below '
                                 'lower_seismogenic_depth=%r' % (
                                     source_id, depth,
                                     lower_seismogenic_depth))
            elif depth < upper_seismogenic_depth:
                raise ValueError('source_id=%r depth=%r is over '
                                 'upper_seismogenic_depth=%r' % (
                                     source_id, depth,
                                     upper_seismogenic_depth))
        self.source_id, self.lon, self.lat = source_id, lon, lat
        self.mfd, self.msr = mfd, msr
        self.rupture_aspect_ratio = rupture_aspect_ratio
        self.upper_seismogenic_depth = upper_seismogenic_depth
        self.lower_seismogenic_depth = lower_seismogenic_depth
        self.nodal_plane_distribution = list(nodal_plane_distribution)
        self.hypocenter_distribution = list(hypocenter_distribution)
        self.hypo_dip_fracs = hypo_dip_fracs
        self.tectonic_region_type = tectonic_region_type

    def get_annual_occurrence_rates(self, min_rate=0):
        # source/base.py:459-477
        return [(m, r) for m, r in self.mfd.get_annual_occurrence_rates()
                if min_rate is None or r > min_rate]

    def count_nphc(self):
        return len(self.nodal_plane_distribution) * len(
            self.hypocenter_distribution)

    def count_ruptures(self):
        return len(self.get_annual_occurrence_rates()) * self.count_nphc()


class MultiPointSource(object):
    """
    source/multi_point.py: point sources sharing everything but location and
    MFD.  Iterating yields them (as the engine's split_source does before the
    calculation).

    :param lons, lats: the locations
    :param mfds: one MFD per point (a MultiMFD of the reference is a list of
        MFDs of one kind, mfd/multi_mfd.py), or a single MFD for all
    """
    def __init__(self, source_id, lons, lats, mfds, msr, rupture_aspect_ratio,
                 upper_seismogenic_depth, lower_seismogenic_depth,
                 nodal_plane_distribution, hypocenter_distribution,
                 hypo_dip_fracs=None, tectonic_region_type='*'):
        self.source_id = source_id
        self.lons = numpy.asarray(lons, float)
        self.lats = numpy.asarray(lats, float)
        if len(self.lons) != len(self.lats) or len(self.lons) == 0:
            raise ValueError('lons and lats must have the same, non zero '
                             'length')
        if hasattr(mfds, 'get_annual_occurrence_rates'):
            mfds = [mfds] * len(self.lons)
        self.mfds = list(mfds)
        if len(self.mfds) != len(self.lons):
            raise ValueError('expected %d MFDs, got %d' % (
                len(self.lons), len(self.mfds)))
        parse_msr(msr)
        self.msr = msr
        self.args = (rupture_aspect_ratio, upper_seismogenic_depth,
                     lower_seismogenic_depth, list(nodal_plane_distribution),
                     list(hypocenter_distribution), hypo_dip_fracs,
                     tectonic_region_type)
        self.tectonic_region_type = tectonic_region_type

    def __iter__(self):
        for i, (lon, lat, mfd) in enumerate(zip(self.lons, self.lats,
                                                self.mfds)):
            yield PointSource('%s:%d' % (self.source_id, i), lon, lat, mfd,
                              self.msr, *self.args)

    def __len__(self):
        return len(self.lons)

    def count_ruptures(self):
        return sum(src.count_ruptures() for src in self)


class OqgPointSources(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in (
        'lon', 'lat', 'usd', 'lsd', 'rar', 'mag_beg', 'mag_cnt', 'np_beg',
        'np_cnt', 'hd_beg', 'hd_cnt', 'rup_off', 'mag', 'mag_rate',
        'mag_dist', 'np_strike', 'np_dip', 'np_rake', 'np_prob', 'hd_depth',
        'hd_prob', 'hd_frac')] + [('msr', ctypes.c_int32),
                                  ('shift_hypo', ctypes.c_int32),
                                  ('pointsource_distance', ctypes.c_double),
                                  ('mode', ctypes.c_void_p),
                                  ('msr_par', ctypes.c_double)]


class PointSourceBatch(object):
    """
    S point sources as the struct of arrays of `oqg_point_sources`
    (include/oqg.h).  Sources sharing the same MFD object or equal nodal-plane /
    depth distributions (the points of an area source) share one slice of
    the lists.

    :param pointsource_distance: None / 0 / >= 1000 km: every site sees
        every rupture.  Otherwise the approximation of _quintets
        (contexts.py:321-370): the sources with more than one (nodal plane,
        depth) combination get, per magnitude, one more rupture -- the
        average one of PointSource.iruptures -- for the sites farther than
        psdist(mag) from the source.
    """
    def __init__(self, sources, pointsource_distance=None):
        sources = list(sources)
        psd = float(pointsource_distance or 0.)
        self.pointsource_distance = psd if 0. < psd < 1000. else 0.
        if not sources:
            raise ValueError('no sources')
        msrs = {s.msr for s in sources}
        if len(msrs) != 1:
            raise ValueError('one magnitude-scaling relationship per batch, '
                             'got %s' % sorted(msrs))
        self.msr_name = msrs.pop()
        name, self.msr_par = parse_msr(self.msr_name)
        self.msr = MSR[name]
        S = self.S = len(sources)
        f = lambda name: numpy.array([getattr(s, name) for s in sources],  # noqa
                                     float)
        self.lon, self.lat = f('lon'), f('lat')
        self.usd, self.lsd = (f('upper_seismogenic_depth'),
                              f('lower_seismogenic_depth'))
        self.rar = f('rupture_aspect_ratio')
        lists = {k: [] for k in ('mag', 'mag_rate', 'np', 'hd', 'frac')}
        slices = {'mag': {}, 'np': {}, 'hd': {}}
        beg = {k: numpy.zeros(S, numpy.int32) for k in slices}
        cnt = {k: numpy.zeros(S, numpy.int32) for k in slices}
        any_frac = any(s.hypo_dip_fracs is not None for s in sources)
        for i, s in enumerate(sources):
            key = id(s.mfd)
            if key not in slices['mag']:
                mr = s.get_annual_occurrence_rates()
                slices['mag'][key] = (len(lists['mag']), len(mr))
                lists['mag'] += [m for m, _ in mr]
                lists['mag_rate'] += [r for _, r in mr]
            npkey = tuple(tuple(x) for x in s.nodal_plane_distribution)
            if npkey not in slices['np']:
                slices['np'][npkey] = (len(lists['np']), len(npkey))
                lists['np'] += list(npkey)
            fr = s.hypo_dip_fracs
            fracs = tuple(0.5 if fr is None or fr[d] is None else fr[d]
                          for d in range(len(s.hypocenter_distribution)))
            hdkey = (tuple(tuple(x) for x in s.hypocenter_distribution), fracs)
            if hdkey not in slices['hd']:
                slices['hd'][hdkey] = (len(lists['hd']), len(fracs))
                lists['hd'] += list(hdkey[0])
                lists['frac'] += list(fracs)
            for k, key in (('mag', id(s.mfd)), ('np', npkey), ('hd', hdkey)):
                beg[k][i], cnt[k][i] = slices[k][key]
        self.beg, self.cnt = beg, cnt
        self.mag = numpy.array(lists['mag'], float)
        self.mag_rate = numpy.array(lists['mag_rate'], float)
        npd = numpy.array(lists['np'], float).reshape(-1, 4)
        self.np_prob, self.np_strike, self.np_dip, self.np_rake = (
            numpy.ascontiguousarray(npd[:, k]) for k in range(4))
        hdd = numpy.array(lists['hd'], float).reshape(-1, 2)
        self.hd_prob, self.hd_depth = (
            numpy.ascontiguousarray(hdd[:, k]) for k in range(2))
        self.hd_frac = numpy.array(lists['frac'], float) if any_frac else None
        nphc = cnt['np'].astype(numpy.int64) * cnt['hd']
        # ruptures per magnitude: all the planes and depths, plus the
        # average rupture where the approximation applies
        self.block = nphc + ((nphc > 1) & (self.pointsource_distance > 0))
        n = cnt['mag'].astype(numpy.int64) * self.block
        self.rup_off = numpy.concatenate([[0], numpy.cumsum(n)]).astype(
            numpy.int64)
        self.U = int(self.rup_off[-1])

    def __len__(self):
        return self.U

    def expand(self):
        """(mag, occurrence_rate) of every rupture in device order, on the
        host (needed only for non-Poissonian temporal occurrence models)"""
        mags, rates = [], []
        for s in range(self.S):
            mb, mc = self.beg['mag'][s], self.cnt['mag'][s]
            nb, nc = self.beg['np'][s], self.cnt['np'][s]
            hb, hc = self.beg['hd'][s], self.cnt['hd'][s]
            w = (self.np_prob[nb:nb + nc, None] *
                 self.hd_prob[None, hb:hb + hc]).ravel()
            if self.block[s] > nc * hc:      # the average rupture
                w = numpy.append(w, 1.)
            mags.append(numpy.repeat(self.mag[mb:mb + mc], len(w)))
            rates.append((self.mag_rate[mb:mb + mc, None] * w[None]).ravel())
        return numpy.concatenate(mags), numpy.concatenate(rates)

    def psdist(self, mag_dist):
        """
        (psdist, close-enough radius) per source and magnitude as a list of
        arrays, PointSource.get_psdist (point.py:215-256): min(radius,
        magdist / 2) + pointsource_distance with radius half the diagonal of
        the surface projection of the rupture of the LAST nodal plane
        """
        from .planar import wc1994_median_area
        out = []
        for s in range(self.S):
            mb, mc = self.beg['mag'][s], self.cnt['mag'][s]
            last = self.beg['np'][s] + self.cnt['np'][s] - 1
            mags = self.mag[mb:mb + mc]
            if self.msr == MSR['PointMSR']:
                radius = numpy.zeros(mc)
            else:
                area = median_area(self.msr_name, mags, self.np_rake[last])
                rdip = numpy.radians(self.np_dip[last])
                max_width = (self.lsd[s] - self.usd[s]) / numpy.sin(rdip)
                length = numpy.sqrt(area * self.rar[s])
                width = area / length
                over = width > max_width
                width = numpy.where(over, max_width, width)
                length = numpy.where(over, area / width, length)
                radius = numpy.sqrt(length ** 2 +
                                    (width * numpy.cos(rdip)) ** 2) / 2.
            out.append(numpy.minimum(radius, mag_dist[mb:mb + mc] / 2.) +
                       self.pointsource_distance)
        return out

    def fewsites_mode(self, mag_dist, sitecol):
        """
        The rule of the reference for few sites (contexts.py:358-362): for
        each source and magnitude, if some site is within psdist all the
        sites see all the planes and depths, otherwise all the sites see the
        average rupture.  Returns the (U,) uint8 visibility of each rupture
        (0 all sites, 3 none).
        """
        from .site import EARTH_RADIUS
        lon, lat = numpy.radians(self.lon), numpy.radians(self.lat)
        loc = EARTH_RADIUS * numpy.stack([
            numpy.cos(lat) * numpy.cos(lon), numpy.cos(lat) * numpy.sin(lon),
            numpy.sin(lat)], axis=1)                      # (S, 3)
        xyz = sitecol.xyz                                  # (N, 3)
        mode = numpy.zeros(self.U, numpy.uint8)
        psd = self.psdist(mag_dist)
        for s in range(self.S):
            B = int(self.block[s])
            nphc = int(self.cnt['np'][s]) * int(self.cnt['hd'][s])
            if B == nphc:
                continue
            cdist = numpy.sqrt(((xyz - loc[s]) ** 2).sum(axis=1))
            close = (cdist[None, :] <= psd[s][:, None]).any(axis=1)   # (M,)
            blk = numpy.zeros((len(close), B), numpy.uint8)
            blk[close, nphc:] = 3        # something is close: no average
            blk[~close, :nphc] = 3       # all far: only the average
            mode[self.rup_off[s]:self.rup_off[s + 1]] = blk.ravel()
        return mode

    def to_device(self, dev, mag_dist=None, shift_hypo=False, mode=None):
        """:returns: (OqgPointSources struct, list of tensors kept alive)"""
        from .contexts import _dev
        keep = {}

        def up(name, arr):
            if arr is None:
                return None
            keep[name] = _dev(arr, dev)
            return keep[name].data_ptr()
        st = OqgPointSources()
        for name in ('lon', 'lat', 'usd', 'lsd', 'rar', 'rup_off', 'mag',
                     'mag_rate', 'np_strike', 'np_dip', 'np_rake', 'np_prob',
                     'hd_depth', 'hd_prob', 'hd_frac'):
            setattr(st, name, up(name, getattr(self, name)))
        for k in ('mag', 'np', 'hd'):
            setattr(st, k + '_beg', up(k + '_beg', self.beg[k]))
            setattr(st, k + '_cnt', up(k + '_cnt', self.cnt[k]))
        st.mag_dist = up('mag_dist', mag_dist)
        st.msr = self.msr
        st.msr_par = self.msr_par
        st.shift_hypo = int(bool(shift_hypo))
        st.pointsource_distance = self.pointsource_distance
        st.mode = up('mode', mode)
        return st, keep


def point_ruptures_device(batch, handle, mag_dist=None, shift_hypo=False,
                          mode=None, with_aux=False):
    """
    Build the planar ruptures of a PointSourceBatch on the handle's device:
    returns torch CUDA tensors (planar (U, 42), mag (U,), magdist (U,) or
    None) and, if `with_aux` or the batch uses the pointsource_distance
    approximation, the (U, 5) visibility array for oqg_hazard_planar_aux.
    The equivalent of PointSource.get_planar (and iruptures) for every
    source.
    """
    import torch
    dev = handle.device
    want_aux = with_aux or batch.pointsource_distance > 0 or mode is not None
    if batch.pointsource_distance > 0 and mag_dist is None:
        raise ValueError('pointsource_distance needs the integration '
                         'distances (mag_dist)')
    st, keep = batch.to_device(dev, mag_dist, shift_hypo, mode)
    U = batch.U
    cuda = 'cuda:%d' % dev
    t_pla = torch.empty((U, _lib.PLANAR_LEN), dtype=torch.float64, device=cuda)
    t_mag = torch.empty(U, dtype=torch.float64, device=cuda)
    t_md = torch.empty(U, dtype=torch.float64, device=cuda) \
        if mag_dist is not None else None
    t_aux = torch.empty((U, 5), dtype=torch.float64, device=cuda) \
        if want_aux else None
    handle.call('oqg_point_ruptures', ctypes.byref(st), batch.S, U,
                ctypes.c_void_p(t_pla.data_ptr()),
                ctypes.c_void_p(t_mag.data_ptr()),
                ctypes.c_void_p(t_md.data_ptr() if t_md is not None else None),
                ctypes.c_void_p(t_aux.data_ptr() if want_aux else None))
    # the source arrays must outlive the kernel: the launch is asynchronous
    torch.cuda.current_stream(dev).synchronize()
    del keep
    if want_aux:
        return t_pla, t_mag, t_md, t_aux
    return t_pla, t_mag, t_md
