"""Ground motion fields given the epsilons: the arithmetic of
GmfComputer._compute (openquake/hazardlib/calc/gmf.py:476-525) on the device
(SURVEY.md section 8f row 4).  The random numbers stay with the caller, as in
the reference (numpy generators on the host): this module evaluates
``exp(mean + phi * within_eps + tau * between_eps)`` for all the IMTs, sites
and events at once, in float32 like the reference.  Spatial / cross-IMT
correlation of the residuals is applied by the caller to `within_eps` before
the call (the reference's non-spatial models do the same, gmf.py:441-457);
SpatialCorrelationModel.apply_correlation is not built.
"""
import ctypes
import numpy

F32 = numpy.float32
TRUNCATION_THRESHOLD = 1e-9


def compute_gmf(mean_stds, imts, within_eps, between_eps=None,
                total_stddev_only=False, truncation_within=3.,
                truncation_between=3., device=None):
    """
    :param mean_stds: (4, M, N) mean, sig, tau, phi of one gsim
        (ContextMaker.get_mean_stds(...)[:, g])
    :param imts: the M intensity measure types (strings); MMI is not
        exponentiated
    :param within_eps: (M, N, E) float32
    :param between_eps: (E, M) float32 (not needed without between-event
        variability)
    :param total_stddev_only: the gsim defines only the total standard
        deviation (gmf.py:497-506)
    :returns: (M, N, E) float32 ground motion values
    """
    import torch
    from .contexts import get_handle, _device_index, _dev, _ptr
    h = get_handle(_device_index(device))
    dev = h.device
    ms = numpy.ascontiguousarray(mean_stds, F32)
    _4, M, N = ms.shape
    assert len(imts) == M
    within = numpy.ascontiguousarray(within_eps, F32)
    E = within.shape[2]
    assert within.shape == (M, N, E), within.shape
    if truncation_within <= TRUNCATION_THRESHOLD and \
            truncation_between <= TRUNCATION_THRESHOLD:
        mode = 0
    elif total_stddev_only:
        mode = 1
    else:
        mode = 2
        if between_eps is None:
            raise ValueError('between_eps is needed')
    t_ms, t_w = _dev(ms, dev), _dev(within, dev)
    t_b = _dev(numpy.ascontiguousarray(between_eps, F32), dev) \
        if mode == 2 else None
    if mode == 2:
        assert t_b.shape == (E, M), t_b.shape
    not_mmi = (ctypes.c_ubyte * M)(*[int(str(imt) != 'MMI') for imt in imts])
    out = torch.empty((M, N, E), dtype=torch.float32, device='cuda:%d' % dev)
    h.call('oqg_gmf', _ptr(t_ms), _ptr(t_w), _ptr(t_b), not_mmi, M, N, E, mode,
           _ptr(out))
    return out.cpu().numpy()
