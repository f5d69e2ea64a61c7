"""
ORACLE (test infrastructure, never imported by the product): numpy
restatement of the probability part of the classical-PSHA hot path:
truncated-normal exceedance probabilities, probabilities of no exceedance
and the update of the per-site map.

Follows openquake/hazardlib/stats.py:27-73 (ndtr, truncnorm_sf),
openquake/hazardlib/contexts.py:482-489 (_set_poes), :1155-1203 (gen_poes),
openquake/hazardlib/tom.py:185-217 (get_pnes) and
openquake/hazardlib/map_array.py:238-260 (update_pmap_i / update_pmap_r).
Pinned by tests/golden/poes_*.npz and tests/golden/truncnorm.json
(reference outputs) .
"""
import math
import numpy as np
from scipy.special import erf

SQRT05 = math.sqrt(0.5)


def ndtr(z):
    # stats.py:31-33
    return 0.5 * (1.0 + erf(np.asarray(z, float) * SQRT05))


def truncnorm_sf(phi_b, values):
    # stats.py:38-73
    z = phi_b * 2. - 1.
    with np.errstate(all='ignore'):
        return ((phi_b - ndtr(values)) / z).clip(0., 1.)


def set_poes(mean_std, loglevels, truncation_level):
    """
    contexts.py:482-489 + :1173: mean_std (2, M, N) f64, loglevels (M, L1)
    -> poes (N, M*L1) float32
    """
    phi_b = float(ndtr(truncation_level))
    M, L1 = loglevels.shape
    N = mean_std.shape[2]
    out = np.empty((N, M * L1), np.float32)
    for m in range(M):
        mea, std = mean_std[0, m], mean_std[1, m]
        with np.errstate(all='ignore'):
            z = (loglevels[m][None, :] - mea[:, None]) / std[:, None]
        out[:, m * L1:(m + 1) * L1] = truncnorm_sf(phi_b, z)
    return out


def get_pnes(rate, probs, poes, time_span):
    # tom.py:185-217; poes any shape, probs 1-D (empty => Poisson)
    if time_span == 0.:
        return 1. - poes
    elif len(probs) == 0:
        return np.exp(- rate * time_span * poes)
    pnes = np.full(poes.shape, probs[0])
    for p, prob in enumerate(probs[1:], 1):
        pnes[:] += prob * (1 - poes) ** p
    return pnes.clip(0., 1.)


def _rounds(sidxs):
    """
    Split row indices into rounds such that inside a round every site index
    appears at most once and the per-site order of the rows is preserved:
    applying the rounds one after the other reproduces exactly the
    row-by-row read-modify-write sequence of map_array.py:238-260.
    """
    order = np.argsort(sidxs, kind='stable')
    s = sidxs[order]
    start = np.r_[0, np.flatnonzero(np.diff(s)) + 1]
    rank = np.arange(len(s)) - np.repeat(start, np.diff(np.r_[start, len(s)]))
    for k in range(rank.max() + 1 if len(s) else 0):
        yield order[rank == k]


def update_pmap(arr, poes, rates, probs_occur, sidxs, itime, rates_mode):
    """
    map_array.py:238-260: arr (Nsites, L, G) float32 updated in place;
    poes (n, L, G) float64, rates (n,), probs_occur (n, K), sidxs (n,)
    """
    K = probs_occur.shape[1] if probs_occur.ndim == 2 else 0
    for rows in _rounds(np.asarray(sidxs)):
        sx = sidxs[rows]
        if K == 0:
            x = rates[rows, None, None] * poes[rows] * itime
            if rates_mode:
                arr[sx] += x
            else:
                arr[sx] *= np.exp(-x)
        else:
            pnes = np.empty_like(poes[rows])
            for i, r in enumerate(rows):
                pnes[i] = get_pnes(rates[r], probs_occur[r], poes[r], itime)
            if rates_mode:
                arr[sx] += -np.log(pnes)
            else:
                arr[sx] *= pnes
    return arr
