"""RmapMaker: the driver on top of the hot path, mirroring
openquake/hazardlib/contexts.py:1478-1662 for point-like sources.

`RmapMaker(cmaker, sitecol, group).make()` returns the same dictionary as
the reference ('rmap' = a MapArray of rates with the `gid` of the gsims,
'cfactor', 'rup_data', 'source_data', 'task_no').  The sources are folded
into the map by the fused kernels (ContextMaker.hazard_point_sources /
hazard_planar) instead of the chunked get_ctxs -> update loop of
_make_src_indep (:1548-1579).  Supported groups: independent sources,
mutually exclusive sources (src_interdep='mutex', _make_src_mutex
:1581-1606) and cluster groups (:1626-1638).  Mutually exclusive RUPTURES
(rup_interdep='mutex', update_pmap_m) have no device implementation: an
error.
"""
import time
import collections
import numpy

from .map_array import MapArray, F32
from .tom import PoissonTOM
from .contexts import PlanarRuptures
from .fault import MeshRuptures, SimpleFaultSource
from .multifault import MultiFaultRuptures
from .source import PointSource, PointSourceBatch


class SourceGroup(object):
    """The attributes of sourceconverter.SourceGroup the driver reads"""
    def __init__(self, trt, sources, name=None, src_interdep='indep',
                 rup_interdep='indep', grp_probability=1., cluster=False,
                 temporal_occurrence_model=None):
        if src_interdep not in ('indep', 'mutex') or \
                rup_interdep not in ('indep', 'mutex'):
            raise ValueError('interdependence must be indep or mutex')
        self.trt, self.sources, self.name = trt, list(sources), name
        self.src_interdep, self.rup_interdep = src_interdep, rup_interdep
        self.grp_probability = grp_probability
        self.cluster = cluster
        self.temporal_occurrence_model = temporal_occurrence_model

    def __iter__(self):
        return iter(self.sources)

    def __len__(self):
        return len(self.sources)

    def __getitem__(self, i):
        return self.sources[i]


class RmapMaker(object):
    def __init__(self, cmaker, sitecol, group, task_no=0):
        self.cmaker = cmaker
        self.sitecol = sitecol
        self.task_no = task_no
        self.N = len(sitecol)
        try:
            self.sources = group.sources
        except AttributeError:  # already a list of sources
            self.sources = list(group)
        if not self.sources:
            raise ValueError('no sources')
        # distributed sources are split into their point sources first, like
        # the engine does (split_source) before the calculation
        from .area import AreaSource
        from .source import MultiPointSource
        if not getattr(group, 'src_interdep', None) == 'mutex':
            self.sources = [
                s for src in self.sources for s in (
                    src if isinstance(src, (AreaSource, MultiPointSource))
                    else [src])]
        self.src_mutex = getattr(group, 'src_interdep', None) == 'mutex'
        if getattr(group, 'rup_interdep', None) == 'mutex':
            raise NotImplementedError(
                'mutually exclusive ruptures (update_pmap_m) have no device '
                'implementation')
        self.grp_probability = getattr(group, 'grp_probability', 1.)
        self.cluster = cmaker.cluster = int(bool(getattr(group, 'cluster', 0)))
        if self.cluster:
            tom = group.temporal_occurrence_model
        else:
            tom = getattr(self.sources[0], 'temporal_occurrence_model', None
                          ) or PoissonTOM(cmaker.investigation_time)
        self.cmaker.tom = self.tom = tom

    # ----------------------------------------------------------- folding
    @staticmethod
    def _tom_key(tom):
        return (type(tom).__name__, getattr(tom, 'time_span', None),
                repr(getattr(tom, 'mu', None)),
                repr(getattr(tom, 'alpha', None)))

    def _fold(self, pmap, srcs):
        """all the ruptures of `srcs` into pmap; returns the kept pairs.
        The temporal occurrence model belongs to the source (the reference
        reads src.temporal_occurrence_model per source, contexts.py:443-474,
        944-953): sources with different models (e.g. NegativeBinomialTOM
        with their own mu / alpha) are folded one model at a time."""
        if self.cluster:
            return self._fold_tom(pmap, srcs)
        by_tom = {}
        for s in srcs:
            tom = getattr(s, 'temporal_occurrence_model', None) or self.tom
            by_tom.setdefault(self._tom_key(tom), (tom, []))[1].append(s)
        kept = 0
        for tom, group in by_tom.values():
            self.cmaker.tom = tom
            kept += self._fold_tom(pmap, group)
        self.cmaker.tom = self.tom
        return kept

    def _fold_tom(self, pmap, srcs):
        cm, kept = self.cmaker, 0
        points = [s for s in srcs if isinstance(s, PointSource)]
        if points:
            for msr in sorted({s.msr for s in points}):
                # the pointsource_distance of the ContextMaker applies here
                # as in get_ctxs (_quintets, contexts.py:321-370)
                batch = PointSourceBatch([s for s in points if s.msr == msr],
                                         cm.pointsource_distance)
                kept += cm.hazard_point_sources(pmap, batch, self.sitecol,
                                                count=True)
        for s in srcs:
            if isinstance(s, PlanarRuptures):
                kept += cm.hazard_planar(pmap, s, self.sitecol, count=True)
            elif isinstance(s, (MeshRuptures, SimpleFaultSource,
                                MultiFaultRuptures)):
                kept += cm.hazard_mesh(pmap, s, self.sitecol, count=True)
            elif not isinstance(s, PointSource):
                raise NotImplementedError(
                    '%s sources have no device implementation'
                    % type(s).__name__)
        return kept

    @staticmethod
    def _nrupts(src):
        return src.count_ruptures() if hasattr(src, 'count_ruptures') \
            else len(src)

    def _make_src_indep(self):
        cm = self.cmaker
        t0 = time.time()
        pnemap = MapArray(self.sitecol.sids, cm.L, cm.G, not self.cluster
                          ).fill(self.cluster)
        kept = self._fold(pnemap, self.sources)
        dt = (time.time() - t0) / len(self.sources)
        tot = sum(self._nrupts(s) for s in self.sources)
        for src in self.sources:
            self.update_source_data(
                src, dt, kept * self._nrupts(src) / max(tot, 1))
        return pnemap

    def _make_src_mutex(self):
        cm = self.cmaker
        pmap = MapArray(self.sitecol.sids, cm.L, cm.G).fill(0)
        for src in self.sources:
            t0 = time.time()
            pm = MapArray(pmap.sids, cm.L, cm.G).fill(1)
            n = self._fold(pm, [src])
            if n == 0:
                continue
            pmap.array += (1. - pm.array) * F32(src.mutex_weight)
            self.update_source_data(src, time.time() - t0, n)
        pmap.array *= F32(self.grp_probability)
        return ~pmap

    def update_source_data(self, src, dt, nctxs):
        sd = self.source_data
        sd['src_id'].append(getattr(src, 'source_id',
                                    getattr(src, 'src_id', 0)))
        sd['grp_id'].append(getattr(src, 'grp_id', 0))
        sd['nctxs'].append(nctxs)
        sd['nrupts'].append(self._nrupts(src))
        sd['weight'].append(getattr(src, 'weight', 1.))
        sd['ctimes'].append(dt)
        sd['taskno'].append(self.task_no)

    def make(self):
        dic = {}
        self.source_data = collections.defaultdict(list)
        if not self.src_mutex:
            pnemap = self._make_src_indep()
        else:
            pnemap = self._make_src_mutex()
        if self.cluster:
            # contexts.py:1626-1638: the group happens n times with the
            # probabilities of its own temporal occurrence model
            MINFLOAT = 1.4E-45
            probs = F32(self.tom.get_probability_n_occurrences(
                self.tom.occurrence_rate, numpy.arange(20)))
            probs = probs[probs > MINFLOAT]
            array = numpy.full(pnemap.shape, probs[0], dtype=F32)
            for nocc, probn in enumerate(probs[1:], 1):
                array += pnemap.array ** nocc * probn
            pnemap.array = array
        dic['rmap'] = pnemap.to_rates()
        dic['rmap'].gid = self.cmaker.gid
        dic['cfactor'] = self.cmaker.cfactor
        dic['rup_data'] = ()
        dic['source_data'] = dict(self.source_data)
        dic['task_no'] = self.task_no
        return dic
