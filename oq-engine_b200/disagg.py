"""Disaggregation: the hot function of openquake/hazardlib/calc/disagg.py
(`_disaggregate`, :188-243 -- "the crucial bit for performance") on the
device.

`disaggregate(ctx, mea, std, cmaker, g, iml2, bin_edges, epsstar, gp)` keeps
the argument meaning of the reference (minus the eps4 tuple, which is
derived from the epsilon edges and the truncation level as get_eps4 does,
and the monitors): ctx are the context rows of ONE site and ONE magnitude
bin with columns rrup, clon, clat, occurrence_rate; mea, std the (G, M, U)
arrays of ContextMaker.get_mean_stds; iml2 the (M, P) logarithmic
intensities; bin_edges = [mag, dist, lon, lat, eps] edges.  Returns the
6D matrix (D, Lo, La, E, M, P) of probabilities of exceedance.
Non-Poissonian ruptures (probs_occur) are an error: no device path.
"""
import ctypes
import numpy

from . import _lib


def disaggregate(ctx, mea, std, cmaker, g, iml2, bin_edges, epsstar=False,
                 gp=1., device=None):
    import torch
    from .contexts import get_handle, _device_index, _dev, _ptr
    po = getattr(ctx, 'probs_occur', None)
    if po is not None and numpy.size(po):
        raise NotImplementedError(
            'disaggregation of non-Poissonian ruptures has no device '
            'implementation')
    h = get_handle(_device_index(device if device is not None
                                 else getattr(cmaker, 'device', None)))
    dev = h.device
    iml2 = numpy.ascontiguousarray(iml2, float)
    M, P = iml2.shape
    dist_edges, lon_edges, lat_edges, eps_edges = (
        numpy.ascontiguousarray(e, float) for e in bin_edges[1:5])
    D, Lo, La, E = (len(e) - 1 for e in (dist_edges, lon_edges, lat_edges,
                                         eps_edges))
    U = len(ctx)
    mea_g = numpy.ascontiguousarray(numpy.asarray(mea)[g], float)
    std_g = numpy.ascontiguousarray(numpy.asarray(std)[g], float)
    assert mea_g.shape == (M, U), (mea_g.shape, M, U)
    t = {n: _dev(numpy.asarray(ctx[n] if not hasattr(ctx, n)
                               else getattr(ctx, n), float), dev)
         for n in ('rrup', 'clon', 'clat', 'occurrence_rate')} if U else {}
    t_mea, t_std = (_dev(mea_g, dev), _dev(std_g, dev)) if U else (None, None)
    out = torch.empty((D, Lo, La, E, M, P), dtype=torch.float64,
                      device='cuda:%d' % dev)
    dp = _lib.c_double_p
    h.call('oqg_disaggregate', _ptr(t.get('rrup')), _ptr(t.get('clon')),
           _ptr(t.get('clat')), _ptr(t.get('occurrence_rate')), U,
           _ptr(t_mea), _ptr(t_std), M, P, iml2.ctypes.data_as(dp),
           dist_edges.ctypes.data_as(dp), D, lon_edges.ctypes.data_as(dp), Lo,
           lat_edges.ctypes.data_as(dp), La, eps_edges.ctypes.data_as(dp), E,
           float(cmaker.truncation_level), float(cmaker.investigation_time),
           int(bool(epsstar)), float(gp), _ptr(out))
    return out.cpu().numpy()
