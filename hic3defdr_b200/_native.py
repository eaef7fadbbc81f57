"""
ctypes binding of libh3d.so (include/h3d.h).  There is no fallback: if the
library is missing or a call fails, an exception is raised.
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# H3D_LIB selects an alternative build of the same library (A/B measurements)
LIB_PATH = os.environ.get('H3D_LIB') or os.path.join(HERE, 'libh3d.so')

c_int, c_ll, c_dbl, c_sz = ctypes.c_int, ctypes.c_longlong, ctypes.c_double, \
    ctypes.c_size_t
vp = ctypes.c_void_p

# name -> (restype, argtypes); mirrors include/h3d.h line by line
SIGNATURES = {
    'h3d_version': (c_int, []),
    'h3d_last_error': (ctypes.c_char_p, []),
    'h3d_launch_count': (ctypes.c_ulonglong, []),
    'h3d_reset_launch_count': (None, []),
    'h3d_fp64_peak': (c_int, [vp, vp, vp]),
    'h3d_bias_filter': (c_int, [vp, c_int, c_int, c_dbl, vp]),
    'h3d_union_count': (c_int, [c_int, vp, c_int, vp, vp, c_int, vp, c_int,
                                c_int, vp, vp, c_sz, vp]),
    'h3d_union_ws_bytes': (c_sz, [c_int]),
    'h3d_union_emit': (c_int, [c_int, vp, c_int, vp, vp, c_int, vp, c_int,
                               c_int, vp, vp, vp, vp, vp, vp, vp]),
    'h3d_size_factors': (c_int, [vp, vp, c_ll, c_int, c_int, c_int, c_int, vp,
                                 vp, c_sz, vp]),
    'h3d_size_factors_ws_bytes': (c_sz, [c_ll, c_int, c_int]),
    'h3d_sf_num_groups': (c_int, [c_int, c_int, c_int]),
    'h3d_sf_group_bounds': (c_int, [c_ll, c_int, c_int, c_int, vp, vp, vp]),
    'h3d_sf_values': (c_int, [vp, vp, c_ll, c_int, c_int, vp, vp]),
    'h3d_sf_group_reduce': (c_int, [vp, c_ll, vp, c_int, c_int, c_int, vp, vp,
                                    vp]),
    'h3d_sf_table': (c_int, [vp, vp, vp, c_int, c_int, c_int, c_int, c_int,
                             vp, vp, c_sz, vp]),
    'h3d_sf_table_ws_bytes': (c_sz, [c_int, c_int]),
    'h3d_scale_filter': (c_int, [vp, vp, vp, vp, c_int, vp, c_ll, c_int, c_int,
                                 c_int, c_dbl, c_int, vp, vp, vp]),
    'h3d_mask_to_index': (c_int, [vp, c_ll, vp, vp, vp, c_sz, vp]),
    'h3d_mask_to_index_ws_bytes': (c_sz, [c_ll]),
    'h3d_loop_membership': (c_int, [vp, vp, vp, c_ll, vp, c_ll, vp, vp]),
    'h3d_gather_counts_factors': (c_int, [vp, vp, vp, c_ll, vp, vp, c_int, vp,
                                          c_int, vp, c_ll, vp, vp, vp, vp]),
    'h3d_peer_alloc': (c_int, [c_sz, vp, vp]),
    'h3d_peer_free': (c_int, [vp]),
    'h3d_peer_open': (c_int, [vp, vp]),
    'h3d_peer_close': (c_int, [vp]),
    'h3d_peer_copy': (c_int, [vp, vp, vp, vp, vp, c_int, vp]),
    'h3d_pool_index': (c_int, [vp, vp, c_ll, c_int, vp, vp, vp]),
    'h3d_pool_pull': (c_int, [vp, c_ll, vp, c_int, vp, vp, vp, vp, c_int, c_int,
                              vp, c_int, c_ll, vp]),
    'h3d_stable_rank': (c_int, [vp, c_ll, c_int, vp, vp, vp, c_sz, vp]),
    'h3d_stable_rank_ws_bytes': (c_sz, [c_ll, c_int]),
    'h3d_estimate_dispersion': (c_int, [vp, vp, c_ll, vp, c_int, vp, c_int,
                                        c_int, c_int, vp, vp, vp, c_sz, vp]),
    'h3d_estimate_dispersion_ws_bytes': (c_sz, [c_ll, c_int, c_int, c_int]),
    'h3d_estimate_dispersion_runs': (c_int, [vp, vp, c_ll, vp, vp, vp, c_int,
                                             c_int, vp, c_int, c_int, c_int,
                                             vp, vp, vp, c_sz, vp]),
    'h3d_estimate_dispersion_runs_ws_bytes': (c_sz, [c_ll, c_int, c_int,
                                                     c_int, c_int]),
    'h3d_equalize': (c_int, [vp, vp, c_ll, c_ll, c_int, c_dbl, vp, vp, vp,
                             c_sz, vp]),
    'h3d_equalize_ws_bytes': (c_sz, [c_ll]),
    'h3d_cml_nll': (c_int, [vp, c_ll, c_ll, c_int, c_dbl, vp, vp, c_sz, vp]),
    'h3d_cml_nll_ws_bytes': (c_sz, [c_ll]),
    'h3d_lowess': (c_int, [vp, vp, c_int, c_dbl, c_int, c_dbl, vp, vp, c_sz,
                           vp]),
    'h3d_lowess_ws_bytes': (c_sz, [c_int]),
    'h3d_lowess_batch': (c_int, [vp, vp, vp, vp, vp, c_int, vp, c_int, vp, vp,
                                 c_sz, vp]),
    'h3d_lowess_batch_ws_bytes': (c_sz, [c_int, c_int]),
    'h3d_gather_table': (c_int, [vp, c_ll, vp, c_int, c_int, vp, vp]),
    'h3d_fit_mu_hat': (c_int, [vp, vp, vp, c_ll, c_ll, c_ll, c_int, vp, vp,
                               vp]),
    'h3d_lrt': (c_int, [vp, vp, vp, vp, c_ll, c_int, c_int, c_int, vp, vp, vp,
                        vp, vp, vp]),
    'h3d_lrt_fused': (c_int, [vp, vp, vp, c_ll, vp, vp, c_int, vp, vp, vp,
                              c_int, c_int, c_int, vp, vp, vp, vp, vp, vp]),
    'h3d_publish': (c_int, [vp, vp, c_sz, vp]),
    'h3d_bh': (c_int, [vp, c_ll, vp, vp, c_sz, vp]),
    'h3d_bh_ws_bytes': (c_sz, [c_ll]),
    'h3d_bh_ranked': (c_int, [vp, c_ll, c_ll, c_ll, vp, vp, vp, c_sz, vp]),
    'h3d_bh_apply_carry': (c_int, [vp, c_ll, c_dbl, vp]),
    'h3d_bh_apply_carry_dev': (c_int, [vp, c_ll, vp, vp]),
    'h3d_roc_sort': (c_int, [vp, vp, c_ll, vp, vp, vp, vp, c_sz, vp]),
    'h3d_roc_sort_ws_bytes': (c_sz, [c_ll]),
    'h3d_roc_points': (c_int, [vp, vp, c_ll, vp, c_ll, vp, vp, vp, vp, vp]),
    'h3d_nb_simulate': (c_int, [vp, vp, vp, c_ll, vp, c_int, vp, c_int, c_int,
                                vp, c_int, c_int, ctypes.c_ulonglong, vp, vp,
                                vp]),
    'h3d_perturb': (c_int, [vp, c_ll, vp, vp, c_ll, vp, vp]),
    'h3d_kr_balance': (c_int, [vp, vp, vp, c_int, c_dbl, vp, c_dbl, c_dbl,
                               c_int, vp, vp, c_int, vp, vp, vp, c_sz, vp]),
    'h3d_kr_balance_ws_bytes': (c_sz, [c_int]),
    'h3d_band_nnz': (c_int, [vp, vp, vp, c_int, c_int, vp, vp, vp]),
    'h3d_host_expand_by_distance': (c_int, [vp, c_int, c_int, vp, vp, vp, c_ll,
                                            vp, c_int]),
    'h3d_narrow_i64': (c_int, [vp, c_ll, vp, vp, vp]),
    'h3d_host_widen_i32': (c_int, [vp, vp, c_ll, c_int]),
    'h3d_connected_components': (c_int, [vp, vp, c_ll, vp, vp, vp, c_sz, vp]),
    'h3d_connected_components_ws_bytes': (c_sz, [c_ll]),
}


class H3DError(RuntimeError):
    pass


class _Lib(object):
    def __init__(self):
        if not os.path.exists(LIB_PATH):
            raise H3DError(
                'libh3d.so not found at %s: build it with '
                '`python -m hic3defdr_b200.build` (there is no CPU fallback)'
                % LIB_PATH)
        self.cdll = ctypes.CDLL(LIB_PATH)
        self.missing = []
        for name, (res, args) in SIGNATURES.items():
            try:
                fn = getattr(self.cdll, name)
            except AttributeError:
                self.missing.append(name)
                continue
            fn.restype = res
            fn.argtypes = args

    def call(self, name, *args):
        if name in self.missing:
            raise H3DError('libh3d.so does not export %s' % name)
        rc = getattr(self.cdll, name)(*args)
        if rc != 0:
            raise H3DError('%s failed (%d): %s' % (
                name, rc, self.cdll.h3d_last_error().decode()))

    def query(self, name, *args):
        return getattr(self.cdll, name)(*args)


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = _Lib()
    return _lib


def ptr(t):
    """Device (or host) address of a torch tensor / numpy array / None."""
    if t is None:
        return None
    if hasattr(t, 'data_ptr'):
        return t.data_ptr()
    return t.ctypes.data
