"""ctypes binding of liboqg.so (include/oqg.h).  There is no fallback: if
the library is missing or no CUDA device is present, using it raises."""
import os
import ctypes as C

HERE = os.path.dirname(os.path.abspath(__file__))
# OQG_LIB selects another build of the same library (A/B experiments)
LIBPATH = os.environ.get('OQG_LIB') or os.path.join(HERE, 'liboqg.so')

KMAX = 56
PLANAR_LEN = 42
FAMILIES = {'BSSA14': 0, 'CY14': 1, 'ASK14': 2, 'CB14': 3, 'BA08': 4,
            'AB06': 5, 'AKKAR14': 6, 'CY08': 7, 'TORO02': 8, 'C03': 9,
            'SADIGH97': 10}
DIST = ['rrup', 'rjb', 'rx', 'ry0', 'repi', 'rhypo', 'azimuth', 'clon',
        'clat']

c_double_p = C.POINTER(C.c_double)
c_void = C.c_void_p


class OqgGsim(C.Structure):
    _fields_ = [('family', C.c_int32), ('flags', C.c_int32),
                ('par', C.c_double * 4)]


class OqgCtx(C.Structure):
    _fields_ = [(n, c_void) for n in (
        'mag', 'rake', 'dip', 'ztor', 'width', 'hypo_depth', 'rrup', 'rjb',
        'rx', 'ry0', 'repi', 'rhypo', 'vs30', 'z1pt0', 'z2pt5',
        'vs30measured')]


class OqgSites(C.Structure):
    _fields_ = [(n, c_void) for n in (
        'lon', 'lat', 'depth', 'vs30', 'z1pt0', 'z2pt5', 'vs30measured')]


_lib = None


def lib():
    """The loaded library; raises if it has not been built"""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIBPATH):
        raise RuntimeError(
            '%s is missing: build it with `python -c "import __graft_entry__ '
            'as g; g.build()"` (nvcc, sm_100a); there is no CPU fallback'
            % LIBPATH)
    L = C.CDLL(LIBPATH)
    H = c_void
    sig = {
        'oqg_create': (C.c_int, [C.POINTER(H), C.c_int]),
        'oqg_destroy': (None, [H]),
        'oqg_last_error': (C.c_char_p, [H]),
        'oqg_set_stream': (C.c_int, [H, c_void]),
        'oqg_use_own_stream': (C.c_int, [H]),
        'oqg_synchronize': (C.c_int, [H]),
        'oqg_family_fields': (C.c_char_p, [C.c_int]),
        'oqg_required_distances': (C.c_int, [C.POINTER(OqgGsim)]),
        'oqg_set_gsims': (C.c_int, [H, C.c_int, C.POINTER(OqgGsim), C.c_int,
                                    c_double_p]),
        'oqg_set_imt_mask': (C.c_int, [H, C.POINTER(C.c_ubyte)]),
        'oqg_set_conversion': (C.c_int, [H, c_double_p]),
        'oqg_set_levels': (C.c_int, [H, C.c_int, C.c_int, c_double_p,
                                     C.c_double]),
        'oqg_distances': (C.c_int, [H, c_void, C.c_int, C.POINTER(OqgSites),
                                    C.c_int, C.c_int, c_void]),
        'oqg_mean_stds': (C.c_int, [H, C.POINTER(OqgCtx), C.c_int64, c_void]),
        'oqg_mean_stds_staged': (C.c_int, [H, C.POINTER(OqgCtx), C.c_int64,
                                           c_void]),
        'oqg_poes': (C.c_int, [H, c_void, C.c_int64, c_void]),
        'oqg_update': (C.c_int, [H, c_void, C.c_int, c_void, C.c_int64,
                                 c_void, c_void, c_void, c_void, C.c_int,
                                 C.c_double, C.c_int]),
        'oqg_update_mutex': (C.c_int, [H, c_void, C.c_int, c_void, C.c_int64,
                                       c_void, c_void, c_void, c_void,
                                       C.c_int, C.c_double, c_void]),
        'oqg_hazard_planar': (C.c_int, [
            H, c_void, c_void, c_void, c_void, C.c_int, C.c_int,
            C.POINTER(OqgSites), c_void, C.c_int, C.c_double, C.c_double,
            C.c_int, c_void, c_void]),
        'oqg_hazard_planar_host': (C.c_int, [
            H, c_void, c_void, c_void, c_void, C.c_int, C.c_int,
            C.POINTER(OqgSites), C.c_int, C.c_double, C.c_double, C.c_int,
            c_void, C.POINTER(C.c_ulonglong)]),
        'oqg_host_begin': (C.c_int, [H, C.POINTER(OqgSites), C.c_int,
                                     C.c_int, c_void]),
        'oqg_host_add': (C.c_int, [H, c_void, c_void, c_void, c_void,
                                   C.c_int, C.c_int, C.c_double, C.c_double]),
        'oqg_host_end': (C.c_int, [H, c_void, C.POINTER(C.c_ulonglong)]),
        'oqg_point_ruptures': (C.c_int, [H, c_void, C.c_int, C.c_int64,
                                         c_void, c_void, c_void, c_void]),
        'oqg_fused_planar': (C.c_int, [
            H, c_void, c_void, C.c_int, C.POINTER(OqgSites), C.c_int, c_void,
            C.c_double, C.c_double, C.c_int, c_void, c_void]),
        'oqg_hazard_planar_aux': (C.c_int, [
            H, c_void, c_void, c_void, c_void, c_void, C.c_int, C.c_int,
            C.POINTER(OqgSites), c_void, C.c_int, C.c_double, C.c_double,
            C.c_int, c_void, c_void]),
        'oqg_hazard_dist': (C.c_int, [
            H, c_void, C.POINTER(c_void), c_void, C.c_int, C.c_int,
            C.POINTER(OqgSites), c_void, C.c_int, C.c_double, C.c_double,
            C.c_int, c_void, c_void]),
        'oqg_postclassical': (C.c_int, [
            H, c_void, C.c_int, C.c_int, C.c_int, C.c_int, c_void, c_void,
            C.c_int, c_void, C.c_int, c_double_p, C.c_int,
            C.POINTER(C.c_int32), c_double_p, C.c_int, c_void, c_void]),
        'oqg_mean_rates': (C.c_int, [H, c_void, C.c_int64, C.c_int, C.c_int,
                                     C.c_int, c_void, C.c_int, c_void]),
        'oqg_hazard_maps': (C.c_int, [H, c_void, C.c_int64, C.c_int, C.c_int,
                                      c_void, C.c_int, c_void, c_void]),
        'oqg_rates_to_records': (C.c_int, [
            H, c_void, C.c_int, C.c_int, C.c_int, c_void, c_void, C.c_int,
            c_void, C.c_int64, c_void, C.c_int64, C.POINTER(C.c_int64)]),
        'oqg_mesh_min_distance': (C.c_int, [
            H, c_void, c_void, c_void, C.c_int64, c_void, C.c_int,
            C.POINTER(OqgSites), C.c_int, c_void]),
        'oqg_mesh_rx_ry0': (C.c_int, [
            H, c_void, c_void, c_void, c_void, C.c_int, C.POINTER(OqgSites),
            C.c_int, c_void, c_void]),
        'oqg_mesh_rjb': (C.c_int, [
            H, c_void, c_void, C.c_int64, c_void, c_void, c_void, c_void,
            C.c_int, C.POINTER(OqgSites), C.c_int, c_void]),
        'oqg_line_tuw': (C.c_int, [H, c_void, c_void, c_void, C.c_int,
                                   C.POINTER(OqgSites), C.c_int, c_void]),
        'oqg_multi_compose': (C.c_int, [
            H, c_void, c_void, c_void, C.c_int, C.c_int, c_void, c_void,
            c_void, c_void, c_void, C.c_int, c_void, c_void, c_void, c_void]),
        'oqg_mesh_dip_strike_width': (C.c_int, [
            H, c_void, c_void, c_void, c_void, c_void, C.c_int, c_void]),
        'oqg_section_min': (C.c_int, [H, c_void, C.c_int, C.c_int, c_void,
                                      c_void, C.c_int, c_void]),
        'oqg_disaggregate': (C.c_int, [
            H, c_void, c_void, c_void, c_void, C.c_int64, c_void, c_void,
            C.c_int, C.c_int, c_double_p, c_double_p, C.c_int, c_double_p,
            C.c_int, c_double_p, C.c_int, c_double_p, C.c_int, C.c_double,
            C.c_double, C.c_int, C.c_double, c_void]),
        'oqg_cs_out': (C.c_int, [
            H, c_void, c_void, c_void, C.c_int64, C.c_int, c_double_p, C.c_int,
            c_double_p, c_double_p, C.c_int, C.c_double, C.c_double,
            c_double_p, c_double_p]),
        'oqg_gmf': (C.c_int, [H, c_void, c_void, c_void,
                              C.POINTER(C.c_ubyte), C.c_int, C.c_int, C.c_int,
                              C.c_int, c_void]),
        'oqg_set_exact_poe': (C.c_int, [H, C.c_int]),
        'oqg_poe_table_info': (C.c_int, [H, C.POINTER(C.c_int), c_double_p]),
        'oqg_set_fused_config': (C.c_int, [H, C.c_int, C.c_int]),
        'oqg_last_fused_subs': (C.c_int, [H]),
        'oqg_set_site_grouping': (C.c_int, [H, C.c_int, C.c_double]),
        'oqg_set_post_grouping': (C.c_int, [H, C.c_int]),
        'oqg_truncnorm_rvs': (C.c_int, [
            H, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.c_uint64,
            C.c_int64, C.c_double, c_void]),
        'oqg_corr_matrix': (C.c_int, [H, c_void, c_void, C.c_int, C.c_double,
                                      c_void]),
        'oqg_gmf_max_min': (C.c_int, [
            H, c_void, c_void, c_double_p, c_double_p,
            C.POINTER(C.c_ubyte), C.c_int, C.c_int, C.c_int]),
        'oqg_last_post_path': (C.c_int, [H, C.POINTER(C.c_int)]),
        'oqg_set_count_evals': (C.c_int, [H, C.c_int]),
        'oqg_launch_count': (C.c_ulonglong, [H]),
        'oqg_last_fused_ms': (C.c_int, [H, C.POINTER(C.c_float), C.c_int]),
        'oqg_measure_fp64_peak': (C.c_int, [H, c_double_p]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)   # AttributeError if a symbol is missing
        fn.restype = res
        fn.argtypes = args
    L._signatures = sig
    _lib = L
    return L


def family_fields(family):
    s = lib().oqg_family_fields(FAMILIES[family] if isinstance(family, str)
                                else family)
    return [f for f in s.decode().split(',') if f]


class Handle(object):
    """RAII wrapper of an oqg_handle"""
    def __init__(self, device=0):
        self._h = c_void()
        L = lib()
        if L.oqg_create(C.byref(self._h), device):
            raise RuntimeError('oqg_create: %s' %
                               L.oqg_last_error(None).decode())
        self.device = device

    def check(self, rc):
        if rc:
            raise RuntimeError(lib().oqg_last_error(self._h).decode())

    def call(self, name, *args):
        self.check(getattr(lib(), name)(self._h, *args))

    def close(self):
        if self._h:
            lib().oqg_destroy(self._h)
            self._h = c_void()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
