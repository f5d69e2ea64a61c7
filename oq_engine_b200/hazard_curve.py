"""calc_hazard_curves: the hazardlib entry point on top of the hot path,
mirroring openquake/hazardlib/calc/hazard_curve.py:40-164 for point-like
sources: the sources are grouped by tectonic region type, each group goes
through RmapMaker (fused kernels) with the GSIM of its TRT, and the groups
are composed as ``pmap = 1 - (1 - pmap) * exp(-rmap)`` (:162)."""
import numpy

from .contexts import ContextMaker
from .map_array import MapArray
from .rmap_maker import RmapMaker, SourceGroup


def classical(group, sitecol, cmaker):
    """hazard_curve.py:40-73: the rates map of a group of sources (or of a
    list of atomic groups, whose rates add)"""
    if isinstance(group[0], SourceGroup):
        dic = RmapMaker(cmaker, sitecol, group[0]).make()
        for grp in group[1:]:
            res = RmapMaker(cmaker, sitecol, grp).make()
            dic['rmap'].array += res['rmap'].array
            for k, v in res['source_data'].items():
                dic['source_data'][k] += v
        return dic
    return RmapMaker(cmaker, sitecol, group).make()


def calc_hazard_curves(groups, sitecol, imtls, gsim_by_trt,
                       truncation_level=99., **kwargs):
    """
    :param groups: a list of sources, or of SourceGroup objects
    :param sitecol: the sites
    :param imtls: dictionary imt -> levels
    :param gsim_by_trt: dictionary tectonic region type -> GMPE
    :param kwargs: investigation_time (used when the sources carry no
        temporal occurrence model), maximum_distance, shift_hypo,
        pointsource_distance
    :returns: an array of N records with a field per IMT holding the L1
        probabilities of exceedance, as MapArray.convert (map_array.py:362)
    """
    if not isinstance(groups[0], SourceGroup):      # a list of sources
        by_trt = {}
        for src in groups:
            by_trt.setdefault(src.tectonic_region_type, []).append(src)
        groups = [SourceGroup(trt, srcs, 'src_group', 'indep', 'indep')
                  for trt, srcs in by_trt.items()]
    span = kwargs.get('investigation_time')
    for grp in groups:
        for src in grp:
            tom = getattr(src, 'temporal_occurrence_model', None)
            if tom is not None:
                span = tom.time_span
    if span is None:
        raise ValueError('investigation_time is needed')
    imtls = {imt: numpy.asarray(imls, float) for imt, imls in imtls.items()}
    L = sum(len(v) for v in imtls.values())
    param = dict(imtls=imtls, truncation_level=truncation_level,
                 investigation_time=span,
                 shift_hypo=kwargs.get('shift_hypo', False))
    for k in ('maximum_distance', 'minimum_distance',
              'pointsource_distance'):
        if k in kwargs:
            param[k] = kwargs[k]
    pmap = MapArray(sitecol.sids, L, 1).fill(0)
    for group in groups:
        cmaker = ContextMaker(group.trt, [gsim_by_trt[group.trt]], param)
        rmap = classical(group, sitecol, cmaker)['rmap']
        pmap.array[:] = 1. - (1. - pmap.array) * numpy.exp(-rmap.array)
    # MapArray.convert
    L1 = len(next(iter(imtls.values())))
    dt = numpy.dtype([(str(imt), numpy.float64, (len(imls),))
                      for imt, imls in imtls.items()])
    curves = numpy.zeros(len(sitecol), dt)
    start = 0
    for imt, imls in imtls.items():
        curves[imt][pmap.sidx[sitecol.sids]] = pmap.array[
            :, start:start + len(imls), 0]
        start += len(imls)
    return curves
