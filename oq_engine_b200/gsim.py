"""The ground-motion models with a device implementation, behind the
reference's GMPE operator: ``compute(ctx, imts, mean, sig, tau, phi)`` fills
four (M, N) float64 arrays in place (openquake/hazardlib/gsim/base.py:389-401).

Each class mirrors the constructor arguments and REQUIRES_* sets of the
reference class of the same name; options that have no
device implementation raise NotImplementedError at construction (there is no
fallback to a CPU implementation).
"""
import math
import numpy
from . import _lib
from .imt import from_string
from .coeffs import CoeffsTable

registry = {}


def _period(imt):
    return from_string(imt).period


class GMPE(object):
    """Base class: subclasses define family, tables and device flags"""
    REQUIRES_DISTANCES = frozenset()
    REQUIRES_RUPTURE_PARAMETERS = frozenset()
    REQUIRES_SITES_PARAMETERS = frozenset()
    # the reference's StdDev constants (const.py) as strings
    DEFINED_FOR_STANDARD_DEVIATION_TYPES = frozenset(
        {'Total', 'Inter event', 'Intra event'})
    family = None
    tables = {}   # prefix -> table name; '' is the main table

    def __init_subclass__(cls, register=True, **kw):
        super().__init_subclass__(**kw)
        if register:     # by class name, for get_gsim
            registry[cls.__name__] = cls

    def __init__(self, **kwargs):
        self.kwargs = kwargs

    def __repr__(self):
        kw = ', '.join('%s=%r' % kv for kv in sorted(self.kwargs.items()))
        return '%s(%s)' % (self.__class__.__name__, kw)

    # ---- what the device needs -------------------------------------- #
    def device_spec(self):
        """(family id, flags, [par0..par3])"""
        raise NotImplementedError

    def extra_fields(self, imt):
        """values of the non-tabulated fields of the coefficient row"""
        imt = from_string(imt)
        return {'period': imt.period}

    def coeff_row(self, imt):
        """The KMAX doubles of the device coefficient row for an IMT; raises
        KeyError if the model is not defined for it"""
        imt = from_string(imt)
        extra = self.extra_fields(imt)
        out = []
        for name in _lib.family_fields(self.family):
            if name in extra:
                out.append(float(extra[name]))
                continue
            prefix, _, col = name.rpartition('_')
            if prefix + '_' in self.tables:
                table = self.tables[prefix + '_']
            else:
                table, col = self.tables[''], name
            out.append(float(CoeffsTable.get(table)[imt][col]))
        assert len(out) <= _lib.KMAX
        return out + [0.] * (_lib.KMAX - len(out))

    def pga_row(self):
        return self.coeff_row('PGA')

    # ---- the reference operator ------------------------------------- #
    def compute(self, ctx: numpy.recarray, imts, mean, sig, tau, phi):
        """Fill mean, sig, tau, phi (M, N) in place on the GPU"""
        from .contexts import mean_stds_device
        out = mean_stds_device([self], ctx, [from_string(i) for i in imts])
        mean[:], sig[:], tau[:], phi[:] = out[:, 0]


class BooreEtAl2014(GMPE):
    """boore_2014.py:290-359"""
    REQUIRES_DISTANCES = frozenset({'rjb'})
    REQUIRES_RUPTURE_PARAMETERS = frozenset({'mag', 'rake'})
    REQUIRES_SITES_PARAMETERS = frozenset({'vs30'})
    family = 'BSSA14'
    tables = {'': 'BSSA14'}

    def __init__(self, region='nobasin', sof=True, sigma_mu_epsilon=0.0,
                 **kwargs):
        if kwargs:
            raise NotImplementedError(
                'BooreEtAl2014 options %s have no device implementation'
                % sorted(kwargs))
        super().__init__(region=region, sof=sof,
                         sigma_mu_epsilon=sigma_mu_epsilon)
        self.region, self.sof = region, bool(sof)
        self.sigma_mu_epsilon = sigma_mu_epsilon
        if region != 'nobasin':
            self.REQUIRES_SITES_PARAMETERS = frozenset({'vs30', 'z1pt0'})
        try:
            self._region = {'nobasin': 0, 'CAL': 1, 'JPN': 2}[region]
        except KeyError:
            raise ValueError('region=%s' % region)

    def device_spec(self):
        return (_lib.FAMILIES[self.family],
                self._region | (int(self.sof) << 4),
                [self.sigma_mu_epsilon, 0., 0., 0.])


class ChiouYoungs2014(GMPE):
    """chiou_youngs_2014.py:638-805"""
    REQUIRES_DISTANCES = frozenset({'rrup', 'rjb', 'rx'})
    REQUIRES_RUPTURE_PARAMETERS = frozenset({'mag', 'rake', 'dip', 'ztor'})
    REQUIRES_SITES_PARAMETERS = frozenset({'vs30', 'vs30measured', 'z1pt0'})
    family = 'CY14'
    tables = {'': 'CY14'}

    def __init__(self, region='CAL', sigma_mu_epsilon=0.0, use_hw=True,
                 alpha_nm=1.0, peer=False, **kwargs):
        if kwargs:
            raise NotImplementedError(
                'ChiouYoungs2014 options %s have no device implementation'
                % sorted(kwargs))
        super().__init__(region=region, sigma_mu_epsilon=sigma_mu_epsilon,
                         use_hw=use_hw, alpha_nm=alpha_nm, peer=peer)
        self.region, self.sigma_mu_epsilon = region, sigma_mu_epsilon
        self.use_hw, self.alpha_nm, self.peer = use_hw, alpha_nm, peer
        try:
            self._region = {'CAL': 0, 'JPN': 1, 'ITA': 2, 'WEN': 3}[region]
        except KeyError:
            raise ValueError('region=%s' % region)

    def device_spec(self):
        flags = self._region | (int(bool(self.use_hw)) << 4) | (
            int(bool(self.peer)) << 5)
        return (_lib.FAMILIES[self.family], flags,
                [self.sigma_mu_epsilon, self.alpha_nm, 0., 0.])


class AbrahamsonEtAl2014(GMPE):
    """abrahamson_2014.py:467-560"""
    REQUIRES_DISTANCES = frozenset({'rrup', 'rjb', 'rx', 'ry0'})
    REQUIRES_RUPTURE_PARAMETERS = frozenset(
        {'mag', 'rake', 'dip', 'ztor', 'width'})
    REQUIRES_SITES_PARAMETERS = frozenset({'vs30', 'vs30measured', 'z1pt0'})
    family = 'ASK14'
    tables = {'': 'ASK14'}

    def __init__(self, region=None, sigma_mu_epsilon=0.0, **kwargs):
        if kwargs:
            raise NotImplementedError(
                'AbrahamsonEtAl2014 options %s have no device implementation'
                % sorted(kwargs))
        super().__init__(region=region, sigma_mu_epsilon=sigma_mu_epsilon)
        self.region, self.sigma_mu_epsilon = region, sigma_mu_epsilon
        try:
            self._region = {None: 0, 'CAL': 0, 'JPN': 1, 'TWN': 2,
                            'CHN': 3}[region]
        except KeyError:
            raise ValueError('region=%s' % region)

    def extra_fields(self, imt):
        # _get_v1, abrahamson_2014.py:345-365
        t = imt.period
        if imt.string.startswith('SA'):
            if t <= 0.50:
                v1 = 1500.0
            elif t < 3.0:
                v1 = float(numpy.exp(-0.35 * numpy.log(t / 0.5) +
                                     numpy.log(1500.)))
            else:
                v1 = 800.0
        else:
            v1 = 1500.0
        return {'period': t, 'v1': v1}

    def device_spec(self):
        return (_lib.FAMILIES[self.family], self._region,
                [self.sigma_mu_epsilon, 0., 0., 0.])


class CampbellBozorgnia2014(GMPE):
    """campbell_bozorgnia_2014.py:466-567"""
    REQUIRES_DISTANCES = frozenset({'rrup', 'rjb', 'rx'})
    REQUIRES_RUPTURE_PARAMETERS = frozenset(
        {'mag', 'rake', 'dip', 'ztor', 'width', 'hypo_depth'})
    REQUIRES_SITES_PARAMETERS = frozenset({'vs30', 'z2pt5'})
    family = 'CB14'
    tables = {'': 'CB14'}

    def __init__(self, SJ=False, sigma_mu_epsilon=0.0, **kwargs):
        if kwargs:
            raise NotImplementedError(
                'CampbellBozorgnia2014 options %s have no device '
                'implementation' % sorted(kwargs))
        super().__init__(SJ=SJ, sigma_mu_epsilon=sigma_mu_epsilon)
        self.SJ, self.sigma_mu_epsilon = bool(SJ), sigma_mu_epsilon

    def coeff_row(self, imt):
        if from_string(imt).string in ('CAV', 'IA'):
            raise KeyError('%s has no device implementation' % (imt,))
        return super().coeff_row(imt)

    def device_spec(self):
        return (_lib.FAMILIES[self.family], int(self.SJ),
                [self.sigma_mu_epsilon, 0., 0., 0.])


class BooreAtkinson2008(GMPE):
    """boore_atkinson_2008.py:56-151"""
    REQUIRES_DISTANCES = frozenset({'rjb'})
    REQUIRES_RUPTURE_PARAMETERS = frozenset({'mag', 'rake'})
    REQUIRES_SITES_PARAMETERS = frozenset({'vs30'})
    family = 'BA08'
    tables = {'': 'BA08', 'so_': 'AB06_SOIL_RESPONSE'}

    def __init__(self, **kwargs):
        if kwargs:
            raise NotImplementedError(
                'BooreAtkinson2008 options %s have no device implementation'
                % sorted(kwargs))
        super().__init__()

    def extra_fields(self, imt):
        is_pga = imt.string in ('PGA', 'SA(0.0)', 'SA(0)')
        return {'period': imt.period, 'is_pga': float(is_pga)}

    def device_spec(self):
        return (_lib.FAMILIES[self.family], 0, [0., 0., 0., 0.])


class AtkinsonBoore2006(GMPE):
    """atkinson_boore_2006.py:363-448"""
    REQUIRES_DISTANCES = frozenset({'rrup'})
    REQUIRES_RUPTURE_PARAMETERS = frozenset({'mag'})
    REQUIRES_SITES_PARAMETERS = frozenset({'vs30'})
    DEFINED_FOR_STANDARD_DEVIATION_TYPES = frozenset({'Total'})
    family = 'AB06'
    tables = {'bc_': 'AB06_BC', 'hr_': 'AB06_HARD_ROCK',
              'st_': 'AB06_STRESS', 'so_': 'AB06_SOIL_RESPONSE'}

    def __init__(self, mag_eq='NA', scale_fac=0, **kwargs):
        if kwargs:
            raise NotImplementedError(
                'AtkinsonBoore2006 options %s have no device implementation'
                % sorted(kwargs))
        super().__init__(mag_eq=mag_eq, scale_fac=scale_fac)
        self.mag_eq, self.scale_fac = mag_eq, scale_fac
        try:
            self._mag_eq = {'NA': 0, 'Mblg87': 1, 'Mblg96': 2}[mag_eq]
        except KeyError:
            raise ValueError('mag_eq=%s' % mag_eq)

    def extra_fields(self, imt):
        return {'period': imt.period, 'is_pgv': float(imt.string == 'PGV')}

    def device_spec(self):
        return (_lib.FAMILIES[self.family], self._mag_eq,
                [float(self.scale_fac), 0., 0., 0.])


class AkkarEtAlRjb2014(GMPE):
    """akkar_2014.py:132-198"""
    REQUIRES_DISTANCES = frozenset({'rjb'})
    REQUIRES_RUPTURE_PARAMETERS = frozenset({'mag', 'rake'})
    REQUIRES_SITES_PARAMETERS = frozenset({'vs30'})
    family = 'AKKAR14'
    tables = {'': 'AKKAR14_RJB'}
    _kind = 0

    def __init__(self, adjustment_factor=1.0, **kwargs):
        if kwargs:
            raise NotImplementedError(
                '%s options %s have no device implementation'
                % (self.__class__.__name__, sorted(kwargs)))
        super().__init__(adjustment_factor=adjustment_factor)
        self.adjustment_factor = adjustment_factor

    def device_spec(self):
        return (_lib.FAMILIES[self.family], self._kind,
                [math.log(self.adjustment_factor), 0., 0., 0.])


class AkkarEtAlRepi2014(AkkarEtAlRjb2014):
    """akkar_2014.py:280"""
    REQUIRES_DISTANCES = frozenset({'repi'})
    tables = {'': 'AKKAR14_REPI'}
    _kind = 1


class AkkarEtAlRhyp2014(AkkarEtAlRjb2014):
    """akkar_2014.py:361"""
    REQUIRES_DISTANCES = frozenset({'rhypo'})
    tables = {'': 'AKKAR14_RHYP'}
    _kind = 2


class _PlainGMPE(GMPE, register=False):
    """A model without options and without a rock-PGA pass"""
    def __init__(self, **kwargs):
        if kwargs:
            raise NotImplementedError(
                '%s options %s have no device implementation'
                % (self.__class__.__name__, sorted(kwargs)))
        super().__init__()

    def device_spec(self):
        return (_lib.FAMILIES[self.family], 0, [0., 0., 0., 0.])


class ChiouYoungs2008(_PlainGMPE):
    """chiou_youngs_2008.py:168-228"""
    REQUIRES_DISTANCES = frozenset({'rrup', 'rjb', 'rx'})
    REQUIRES_RUPTURE_PARAMETERS = frozenset({'dip', 'rake', 'mag', 'ztor'})
    REQUIRES_SITES_PARAMETERS = frozenset({'vs30', 'vs30measured', 'z1pt0'})
    family = 'CY08'
    tables = {'': 'CY08'}


class ToroEtAl2002(_PlainGMPE):
    """toro_2002.py:59-146 (total standard deviation only)"""
    REQUIRES_DISTANCES = frozenset({'rjb'})
    REQUIRES_RUPTURE_PARAMETERS = frozenset({'mag'})
    REQUIRES_SITES_PARAMETERS = frozenset()
    DEFINED_FOR_STANDARD_DEVIATION_TYPES = frozenset({'Total'})
    family = 'TORO02'
    tables = {'': 'TORO02'}

    def extra_fields(self, imt):
        dic = super().extra_fields(imt)
        dic.update(adj_ss=0., adj_rv=0., adj_nm=0.)     # SHARE variant only
        return dic


def _share_terms(Frss, pR=0.81, Fnss=0.95, pN=0.01):
    """_compute_faulting_style_term (campbell_2003.py:272-281) for the three
    rake classes: strike slip / other, reverse (30, 150], normal (-120, -60];
    CONSTS_FS of both SHARE classes"""
    return (numpy.power(Frss, -pR) * numpy.power(Fnss, -pN),
            numpy.power(Frss, 1 - pR) * numpy.power(Fnss, -pN),
            numpy.power(Frss, -pR) * numpy.power(Fnss, 1 - pN))


class ToroEtAl2002SHARE(ToroEtAl2002):
    """toro_2002.py:149-183: style of faulting and rock adjustment of the
    SHARE project, mean = log(exp(mean) * (term(rake) * AFrock))"""
    REQUIRES_RUPTURE_PARAMETERS = frozenset({'mag', 'rake'})

    def device_spec(self):
        return (_lib.FAMILIES[self.family], 1, [0., 0., 0., 0.])

    def extra_fields(self, imt):
        dic = super().extra_fields(imt)
        C = CoeffsTable.get('TORO02_FS_ROCK')[from_string(imt)]
        ss, rv, nm = _share_terms(C['Frss'])
        dic.update(adj_ss=ss * C['AFrock'], adj_rv=rv * C['AFrock'],
                   adj_nm=nm * C['AFrock'])
        return dic


class Campbell2003(_PlainGMPE):
    """campbell_2003.py:138-212 (kind "base"; total standard deviation
    only)"""
    REQUIRES_DISTANCES = frozenset({'rrup'})
    REQUIRES_RUPTURE_PARAMETERS = frozenset({'mag'})
    REQUIRES_SITES_PARAMETERS = frozenset()
    DEFINED_FOR_STANDARD_DEVIATION_TYPES = frozenset({'Total'})
    family = 'C03'
    tables = {'': 'C03'}

    def extra_fields(self, imt):
        dic = super().extra_fields(imt)
        dic.update(fs_ss=0., fs_rv=0., fs_nm=0., afrock=0.)   # SHARE only
        return dic


class Campbell2003SHARE(Campbell2003):
    """campbell_2003.py:215-270: mean = log(exp(mean) * term(rake) * AFrock)"""
    REQUIRES_RUPTURE_PARAMETERS = frozenset({'mag', 'rake'})

    def device_spec(self):
        return (_lib.FAMILIES[self.family], 1, [0., 0., 0., 0.])

    def extra_fields(self, imt):
        dic = super().extra_fields(imt)
        C = CoeffsTable.get('C03_FS_ROCK')[from_string(imt)]
        ss, rv, nm = _share_terms(C['Frss'])
        dic.update(fs_ss=ss, fs_rv=rv, fs_nm=nm, afrock=C['AFrock'])
        return dic


class SadighEtAl1997(_PlainGMPE):
    """sadigh_1997.py:110-161 (rock above 750 m/s, deep soil otherwise; total
    standard deviation only): the model of the PEER verification tests"""
    REQUIRES_DISTANCES = frozenset({'rrup'})
    REQUIRES_RUPTURE_PARAMETERS = frozenset({'rake', 'mag'})
    REQUIRES_SITES_PARAMETERS = frozenset({'vs30'})
    DEFINED_FOR_STANDARD_DEVIATION_TYPES = frozenset({'Total'})
    family = 'SADIGH97'
    tables = {'lo_': 'SADIGH97_ROCK_LOWMAG', 'hi_': 'SADIGH97_ROCK_HIMAG',
              'sd_': 'SADIGH97_ROCK_STDDERR', 'so_': 'SADIGH97_SOIL'}


def get_gsim(name, **kwargs):
    """Instantiate a supported model by class name (like valid.gsim); an
    unsupported model is a hard error: there is no CPU fallback"""
    if isinstance(name, GMPE):
        return name
    try:
        cls = registry[name]
    except KeyError:
        raise NotImplementedError(
            'GMPE %r has no device implementation (supported: %s)'
            % (name, ', '.join(sorted(registry)))) from None
    return cls(**kwargs)
