"""ctypes binding of liblcf_b200.so (C ABI declared in include/lcf.h).

The shared library is built in-tree by ``__graft_entry__.build()`` (nvcc, sm_100a).  There is no
CPU fallback: if the library is missing, or no CUDA device is usable, every compute call raises.
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# LCF_B200_LIB: developer override used by tools/microbench to time experimental builds of the SAME library
LIB_PATH = os.environ.get('LCF_B200_LIB') or os.path.join(_HERE, 'liblcf_b200.so')

LCF_ERR_ARG, LCF_ERR_CUDA, LCF_ERR_NAN, LCF_ERR_STATE, LCF_ERR_NWALKERS = -1, -2, -3, -4, -5
MODEL_IDS = {'ShockCooling': 1, 'ShockCooling2': 2, 'ShockCooling3': 3, 'ShockCooling4': 4,
             'CompanionShocking': 5, 'CompanionShocking2': 6, 'CompanionShocking3': 7, 'BlackbodySED': 8}
PRECISIONS = {'fp64': 0, 'fp32': 1}
ROLE_KASEN_RU, ROLE_SIFTO_RR, ROLE_SIFTO_RI, ROLE_DT_U, ROLE_DT_I = 1, 2, 4, 8, 16

_pd = C.POINTER(C.c_double)
_pi = C.POINTER(C.c_int32)


class ProblemDesc(C.Structure):
    _fields_ = [
        ('model_id', C.c_int32), ('precision', C.c_int32), ('ndim', C.c_int32), ('use_sigma', C.c_int32),
        ('sigma_type', C.c_int32), ('npoints', C.c_int32), ('nfilters', C.c_int32), ('reserved0', C.c_int32),
        ('sigma_unit_abs', C.c_double), ('model_consts', C.c_double * 16),
        ('bank_offsets', _pi), ('bank_alpha', _pd), ('bank_w', _pd), ('bank_kappa', _pd), ('filter_role', _pi),
        ('sifto_nknots', C.c_int32), ('reserved1', C.c_int32), ('sifto_x0', C.c_double), ('sifto_dx', C.c_double),
        ('sifto_coef', _pd),
        ('t', _pd), ('point_filter', _pi), ('y', _pd), ('dy', _pd),
        ('prior_kind', _pi), ('prior_min', _pd), ('prior_max', _pd), ('prior_mean', _pd), ('prior_std', _pd),
    ]


class LcfError(RuntimeError):
    pass


_lib = None

# every symbol include/lcf.h declares: (name, restype, argtypes)
_vp = C.c_void_p
SYMBOLS = [
    ('lcf_abi_version', C.c_int, []),
    ('lcf_last_error', C.c_char_p, []),
    ('lcf_device_count', C.c_int, []),
    ('lcf_set_device', C.c_int, [C.c_int]),
    ('lcf_set_tuning', C.c_int, [C.c_int, C.c_int]),
    ('lcf_set_tuning_ex', C.c_int, [C.c_int, C.c_int, C.c_int]),
    ('lcf_set_tuning_split', C.c_int, [C.c_int]),
    ('lcf_set_tuning_flat', C.c_int, [C.c_int]),
    ('lcf_plan_bank_segments', C.c_int, [C.POINTER(C.c_int), C.c_int, C.c_int64, C.POINTER(C.c_int), C.c_int]),
    ('lcf_problem_last_launch_ex', C.c_int, [_vp, C.POINTER(C.c_int64), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    ('lcf_blackbody_lstsq_batch', C.c_int, [C.c_int64, C.POINTER(C.c_int32), _pd, _pd, C.c_double, C.c_double, C.c_double, _pd, _pd,
                                            _pd, _pd, _pd, C.POINTER(C.c_int32)]),
    ('lcf_ensemble_diagnostics', C.c_int, [_vp, C.c_int64, C.c_double, C.c_int64, _pd, C.POINTER(C.c_int64), _pd]),
    ('lcf_ensemble_ipc_export', C.c_int, [_vp, C.c_char_p]),
    ('lcf_ensemble_peers_attach_ipc', C.c_int, [_vp, C.c_char_p]),
    ('lcf_ensemble_peers_attach_ptrs', C.c_int, [_vp, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp)]),
    ('lcf_ensemble_peers_detach', C.c_int, [_vp]),
    ('lcf_ensemble_exchange_view', C.c_int, [_vp, C.POINTER(_vp), C.POINTER(C.c_int)]),
    ('lcf_problem_create', C.c_int, [C.POINTER(ProblemDesc), C.POINTER(_vp)]),
    ('lcf_problem_destroy', None, [_vp]),
    ('lcf_problem_last_launch', C.c_int, [_vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int64),
                                          C.POINTER(C.c_int)]),
    ('lcf_model_eval', C.c_int, [_vp, C.c_int64, _pd, _pd]),
    ('lcf_log_likelihood', C.c_int, [_vp, C.c_int64, _pd, _pd]),
    ('lcf_log_posterior', C.c_int, [_vp, C.c_int64, _pd, _pd, C.POINTER(C.c_int64)]),
    ('lcf_ensemble_create', C.c_int, [_vp, C.c_int64, C.c_uint64, C.c_int, C.c_int, C.POINTER(_vp)]),
    ('lcf_ensemble_destroy', None, [_vp]),
    ('lcf_ensemble_set_state', C.c_int, [_vp, _pd, _pd]),
    ('lcf_ensemble_set_state_slice', C.c_int, [_vp, C.c_int64, C.c_int64, _pd]),
    ('lcf_ensemble_get_state', C.c_int, [_vp, _pd, _pd]),
    ('lcf_ensemble_reset', C.c_int, [_vp]),
    ('lcf_ensemble_run', C.c_int, [_vp, C.c_int64, C.c_int]),
    ('lcf_ensemble_run_to_host', C.c_int, [_vp, C.c_int64, _pd, _pd]),
    ('lcf_ensemble_run_to_host_slice', C.c_int, [_vp, C.c_int64, C.c_int64, C.c_int64, _pd, _pd]),
    ('lcf_ensemble_run_replay', C.c_int, [_vp, C.c_int64, C.c_int, _pi, _pd, _pi, _pd]),
    ('lcf_ensemble_reserve', C.c_int, [_vp, C.c_int64]),
    ('lcf_ensemble_half_step', C.c_int, [_vp, C.c_int, C.c_int]),
    ('lcf_ensemble_end_step', C.c_int, [_vp, C.c_int]),
    ('lcf_ensemble_nstored', C.c_int64, [_vp]),
    ('lcf_ensemble_get_chain', C.c_int, [_vp, _pd]),
    ('lcf_ensemble_get_log_prob', C.c_int, [_vp, _pd]),
    ('lcf_ensemble_get_chain_slice', C.c_int, [_vp, C.c_int64, C.c_int64, _pd, _pd]),
    ('lcf_ensemble_get_accepted', C.c_int, [_vp, C.POINTER(C.c_int64)]),
    ('lcf_ensemble_device_view', C.c_int, [_vp, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), C.POINTER(C.c_int64),
                                           C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    ('lcf_ensemble_set_stream', C.c_int, [_vp, _vp]),
    ('lcf_ensemble_sync', C.c_int, [_vp]),
    ('lcf_ensemble_last_timing', C.c_int, [_vp, _pd, C.POINTER(C.c_int64)]),
    ('lcf_batch_create', C.c_int, [C.c_int64, C.POINTER(_vp), C.c_int64, C.c_uint64, C.POINTER(_vp)]),
    ('lcf_batch_destroy', None, [_vp]),
    ('lcf_batch_set_state', C.c_int, [_vp, _pd]),
    ('lcf_batch_run', C.c_int, [_vp, C.c_int64, C.c_int64]),
    ('lcf_batch_get_chain', C.c_int, [_vp, _pd]),
    ('lcf_batch_get_log_prob', C.c_int, [_vp, _pd]),
    ('lcf_batch_get_accepted', C.c_int, [_vp, C.POINTER(C.c_int64)]),
    ('lcf_batch_get_status', C.c_int, [_vp, _pi]),
    ('lcf_batch_last_timing', C.c_int, [_vp, _pd, C.POINTER(C.c_int64)]),
    ('lcf_sed_batch_create', C.c_int, [C.c_int64, _pi, _pi, _pd, _pd, C.c_int32, _pi, _pd, _pd, C.c_int32, C.c_int32, C.c_int32,
                                       _pi, _pd, _pd, _pd, _pd, C.c_int32, C.c_int64, C.c_uint64, C.POINTER(_vp)]),
    ('lcf_batch_summary', C.c_int, [_vp, _vp, C.c_double, C.c_double, _pd]),
]


def lib():
    """Load liblcf_b200.so (once).  Raises if it has not been built: there is no fallback path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LcfError('CUDA extension %s is missing: run `python -c "import __graft_entry__ as g; g.build()"` '
                           '(nvcc, sm_100a).  lightcurve_fitting_b200 has no CPU fallback.' % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, res, args in SYMBOLS:
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc):
    """Map C return codes onto the exceptions the reference (via emcee) raises."""
    if rc == 0:
        return
    msg = lib().lcf_last_error().decode()
    if rc == LCF_ERR_NAN:
        raise ValueError(msg)               # emcee: ValueError("Probability function returned NaN")
    if rc == LCF_ERR_NWALKERS:
        raise RuntimeError(msg)             # emcee RedBlueMove
    if rc == LCF_ERR_ARG:
        raise ValueError(msg)
    raise LcfError('liblcf_b200: %s (code %d)' % (msg, rc))


def dptr(a):
    return a.ctypes.data_as(_pd)


def iptr(a):
    return a.ctypes.data_as(_pi)


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)
