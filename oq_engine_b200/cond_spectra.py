"""Conditional spectrum: mirrors
openquake/hazardlib/calc/cond_spectra.py:35-169 (Lin et al. 2013) with the
Baker & Jayaram (2008) cross-IMT correlation
(openquake/hazardlib/correlation_models/cross_imt/baker_jayaram_2008.py).

`get_cs_out` evaluates, per site and gsim, the weighted sums of `_cs_out`
over the context rows on the device (`oqg_cs_out`, a reduction over the
ruptures with the mean and standard deviations of ContextMaker.get_mean_stds
staying on the device); `cond_spectra` composes them as the reference does.
Non-Poissonian rows (probs_occur) are an error: no device path.
"""
import math
import ctypes
import numpy

from . import _lib
from .imt import from_string


class BakerJayaram2008(object):
    """Total-residual cross-IMT correlation, PGA taken as SA(0.01)"""
    def rho(self, from_imt, to_imt):
        fp = from_string(from_imt).period or 0.01
        tp = from_string(to_imt).period or 0.01
        if abs(fp - tp) < 1E-10:
            return 1.0
        tmin, tmax = min(fp, tp), max(fp, tp)
        c1 = 1 - math.cos(math.pi / 2 - 0.366 * math.log(
            tmax / max(tmin, 0.109)))
        c2 = 0.0
        if tmax < 0.2:
            term1 = 1.0 - 1.0 / (1.0 + math.exp(100.0 * tmax - 5.0))
            term2 = (tmax - tmin) / (tmax - 0.0099)
            c2 = 1 - 0.105 * term1 * term2
        c3 = c2 if tmax < 0.109 else c1
        c4 = c1 + 0.5 * (math.sqrt(c3) - c3) * (
            1 + math.cos(math.pi * tmin / 0.109))
        if tmax < 0.109:
            return c2
        if tmin > 0.109:
            return c1
        if tmax < 0.2:
            return min(c2, c4)
        return c4


def outdict(M, N, P, start, stop):
    return {_g: numpy.zeros((M, N, 3, P)) for _g in range(start, stop)}


def get_cs_out(cmaker, ctxt, imti, imlsNP, tom, _c=None, poes=None,
               correlation_model=None):
    """
    :param ctxt: a context array
    :param imti: index of the conditioning IMT
    :param imlsNP: (N, P) levels of the conditioning IMT, one row per site,
        in correspondence with the probabilities `poes`
    :param tom: the temporal occurrence model (Poissonian)
    :param _c: the previously computed mean contribution (M, N, 3, P), for
        the standard deviations
    :returns: a dictionary g -> array (M, N, 3, P)
    """
    import torch
    from .contexts import _dev, _ptr
    imlsNP = numpy.asarray(imlsNP, float)
    N, P = imlsNP.shape
    poes = numpy.asarray(poes if poes is not None else cmaker.poes, float)
    assert P == len(poes), (len(poes), P)
    model = correlation_model or getattr(
        cmaker, 'total_residual_correlation_model', None)
    if model is None:
        raise ValueError('a cross-IMT correlation model is needed')
    imts = [im.string for im in cmaker.imts]
    M = len(imts)
    rho = numpy.array([model.rho(imts[imti], imt) for imt in imts])
    out = outdict(M, N, P, int(cmaker.gid.min()), int(cmaker.gid.max()) + 1)
    h = cmaker.handle()
    dev = h.device
    dp = _lib.c_double_p
    for sid, imls in enumerate(imlsNP):
        ctx = ctxt[ctxt.sids == sid]
        if len(ctx) == 0:
            continue
        if ctx.probs_occur.size:
            raise NotImplementedError(
                'conditional spectra of non-Poissonian ruptures have no '
                'device implementation')
        mean_stds = cmaker.get_mean_stds([ctx])       # (4, G, M, U)
        probs = 1. - numpy.exp(-ctx.occurrence_rate * tom.time_span)
        t_probs = _dev(probs, dev)
        for c, g in enumerate(cmaker.gid):
            i = c % len(cmaker.gsims)
            t_mu = _dev(numpy.ascontiguousarray(mean_stds[0, i]), dev)
            t_sig = _dev(numpy.ascontiguousarray(mean_stds[1, i]), dev)
            res = numpy.zeros((M, 3, P))
            prev = None
            if _c is not None:
                prev = numpy.ascontiguousarray(_c[:, sid, 1, :])
            imls_c = numpy.ascontiguousarray(imls, float)
            h.call('oqg_cs_out', _ptr(t_mu), _ptr(t_sig), _ptr(t_probs),
                   len(ctx), M, rho.ctypes.data_as(dp), int(imti),
                   imls_c.ctypes.data_as(dp), poes.ctypes.data_as(dp), P,
                   float(cmaker.truncation_level),
                   float(cmaker.investigation_time),
                   prev.ctypes.data_as(dp) if prev is not None else None,
                   res.ctypes.data_as(dp))
            out[int(g)][:, sid] = res
    return out


def cond_spectra(cmaker, ctxt, sitecol, imt_ref, imls, tom, poes=None,
                 correlation_model=None):
    """
    :param ctxt: the context rows of the sources on the sites (the reference
        takes the sources and calls cmaker.from_srcs)
    :param imt_ref: the conditioning IMT (a string)
    :param imls: (N, P) levels of the conditioning IMT
    :returns: conditional spectra and sigmas, arrays of shape (G, M, N, P)
    """
    imti = list(cmaker.imtls).index(imt_ref)
    kw = dict(poes=poes, correlation_model=correlation_model)
    out0 = get_cs_out(cmaker, ctxt, imti, imls, tom, **kw)
    mean = numpy.mean([out0[g] for g in out0], axis=0)       # MNOP
    out = get_cs_out(cmaker, ctxt, imti, imls, tom, mean, **kw)
    imls = numpy.asarray(imls)
    G, M, N, P = len(cmaker.gsims), len(cmaker.imtls), len(sitecol), \
        imls.shape[1]
    spectra = numpy.zeros((G, M, N, P))
    s_sigma = numpy.zeros((G, M, N, P))
    for g, arr in out.items():
        spectra[g] = numpy.exp(arr[:, :, 1] / arr[:, :, 0])
        s_sigma[g] = numpy.sqrt(arr[:, :, 2] / arr[:, :, 0])
    return spectra, s_sigma
