"""oq-engine_b200: a B200-native (sm_100a) implementation of OpenQuake's
classical-PSHA hot path behind the reference's operator surface
(ContextMaker / GMPE.compute / MapArray).  The computation runs in
hand-written CUDA kernels (csrc/) called through the C ABI of include/oqg.h;
there is no CPU fallback."""
from . import _lib, _build  # noqa
from .imt import from_string  # noqa
from .coeffs import CoeffsTable  # noqa
from .gsim import (  # noqa
    GMPE, BooreEtAl2014, ChiouYoungs2014, AbrahamsonEtAl2014,
    CampbellBozorgnia2014, BooreAtkinson2008, AtkinsonBoore2006,
    AkkarEtAlRjb2014, AkkarEtAlRepi2014, AkkarEtAlRhyp2014, ChiouYoungs2008,
    ToroEtAl2002, Campbell2003, ToroEtAl2002SHARE, Campbell2003SHARE,
    SadighEtAl1997, get_gsim,
    registry)
from .tom import PoissonTOM, NegativeBinomialTOM  # noqa
from .map_array import MapArray  # noqa
from .planar import build_planar, planar_array_dt, wc1994_median_area  # noqa
from .site import SiteCollection  # noqa
from .contexts import ContextMaker, IntegrationDistance, PlanarRuptures  # noqa
from .source import (  # noqa
    PointSource, PointSourceBatch, TruncatedGRMFD, EvenlyDiscretizedMFD,
    ArbitraryMFD, MultiPointSource, point_ruptures_device)
from .postproc import (  # noqa
    postclassical, compute_hazard_maps, compute_hmaps, calc_mean_rates,
    rates_to_records, rates_dt)
from .rmap_maker import RmapMaker, SourceGroup  # noqa
from .mesh import (Mesh, mesh_min_distance, section_min,  # noqa
                   mesh_rx_ry0, mesh_dip_strike_width, mesh_rjb)
from .fault import MeshRuptures, SimpleFaultSource  # noqa
from .multifault import MultiFaultRuptures  # noqa
from .disagg import (disaggregate, Disaggregator, FarAwayRupture,  # noqa
                     disaggregation, build_bin_edges, lon_lat_bins)
from .cond_spectra import (  # noqa
    cond_spectra, get_cs_out, BakerJayaram2008)
from .hazard_curve import calc_hazard_curves, classical  # noqa
from .area import AreaSource  # noqa
from .gmf import (compute_gmf, GmfComputer,  # noqa
                  JB2009CorrelationModel)
from .performance import Monitor  # noqa
from . import postproc  # noqa
from . import sharding  # noqa

__version__ = '0.1.0'
