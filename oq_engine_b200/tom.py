"""Temporal occurrence models, the part of openquake/hazardlib/tom.py the
hot path needs: PoissonTOM (tom.py:107-173) and NegativeBinomialTOM.get_pmf
(tom.py:220-319)."""
import numpy
from scipy.stats import nbinom


class PoissonTOM(object):
    def __init__(self, time_span):
        if time_span <= 0:
            raise ValueError('time_span must be positive')
        self.time_span = time_span

    def get_probability_no_exceedance(self, occurrence_rate, poes):
        return numpy.exp(- occurrence_rate * self.time_span * poes)

    def get_probability_n_occurrences(self, occurrence_rate, num):
        # tom.py:123-140 (scipy.stats.poisson.pmf)
        from scipy.stats import poisson
        return poisson(occurrence_rate * self.time_span).pmf(num)


class NegativeBinomialTOM(object):
    """
    :param time_span: the investigation time
    :param mu, alpha: mean rate and dispersion (Kagan and Jackson 2010)
    """
    x = numpy.arange(10, dtype=float)  # size of the returned PMF

    def __init__(self, time_span, mu, alpha):
        if time_span <= 0:
            raise ValueError('time_span must be positive')
        if numpy.any(numpy.asarray(mu) <= 0) or numpy.any(
                numpy.asarray(alpha) <= 0):
            raise ValueError('Mean rate and rate dispersion must be greater '
                             'than 0')
        self.time_span = time_span
        self.mu = mu
        self.alpha = alpha

    def get_pmf(self, mean_rate):
        """
        Probabilities of 0..9 occurrences, shape (10,) for a scalar rate and
        (n, 10) for an array of rates (tom.py:300-319)
        """
        tau = 1 / self.alpha
        theta = tau / (tau + mean_rate * self.time_span)
        if isinstance(theta, numpy.ndarray):
            theta = theta[:, numpy.newaxis]
        return nbinom.pmf(self.x, tau, theta)
