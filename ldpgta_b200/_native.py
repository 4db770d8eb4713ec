"""
ctypes binding of libldpgta.so (include/ldpgta.h).  There is no CPU path: if the
library or a CUDA device is missing, every entry point raises.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# LDPGTA_LIB: an alternative build of the library (kernel experiments, tools/build_variant.py); default = the in-tree build
LIB_PATH = os.environ.get('LDPGTA_LIB') or os.path.join(_HERE, 'libldpgta.so')

HOMOGENEOUS, RECENT_ADMIXTURE, DISTANT_ADMIXTURE = 0, 1, 2
MAX_READS_PER_DRAW = 18
NUM_LLR_PAIRS = 5
COMM_HANDLE_BYTES = 128

# every symbol include/ldpgta.h declares: (name, restype, argtypes)
_c = ctypes
_P = _c.c_void_p
SYMBOLS = (
    ('ldpgta_last_error', _c.c_char_p, []),
    ('ldpgta_version', _c.c_int, []),
    ('ldpgta_device_count', _c.c_int, []),
    ('ldpgta_create', _c.c_int, [_c.c_int, _c.c_int, _P, _c.POINTER(_P)]),
    ('ldpgta_destroy', _c.c_int, [_P]),
    ('ldpgta_row_words', _c.c_int, [_P, _c.c_int]),
    ('ldpgta_upload_allele_rows', _c.c_int, [_P, _c.c_int, _P, _c.c_int64, _P]),
    ('ldpgta_panel_create', _c.c_int, [_c.c_int, _c.c_int, _P, _c.c_int64, _c.POINTER(_P)]),
    ('ldpgta_panel_upload', _c.c_int, [_P, _c.c_int, _c.c_int64, _P, _c.c_int64]),
    ('ldpgta_panel_destroy', _c.c_int, [_P]),
    ('ldpgta_gather_allele_rows', _c.c_int, [_P, _P, _P, _P, _c.c_int64]),
    ('ldpgta_build_read_masks', _c.c_int, [_P, _P, _P, _c.c_int64]),
    ('ldpgta_read_scores', _c.c_int, [_P, _P, _P, _c.c_int64, _P, _c.c_double, _c.c_double, _P]),
    ('ldpgta_upload_read_masks', _c.c_int, [_P, _c.c_int, _P, _c.c_int64]),
    ('ldpgta_padded_row_words', _c.c_int, [_P, _c.c_int]),
    ('ldpgta_attach_read_masks_dev', _c.c_int, [_P, _c.c_int, _P, _c.c_int64]),
    ('ldpgta_num_reads', _c.c_int64, [_P]),
    ('ldpgta_joint_counts', _c.c_int, [_P, _P, _c.c_int, _P]),
    ('ldpgta_get_likelihoods', _c.c_int, [_P, _c.c_int, _P, _P, _c.c_int, _P, _c.c_int, _P, _P]),
    ('ldpgta_likelihoods_batch', _c.c_int, [_P, _c.c_int, _P, _P, _c.c_int64, _P, _P]),
    ('ldpgta_likelihoods_uniform', _c.c_int, [_P, _c.c_int, _c.c_int, _P, _c.c_int64, _P, _P]),
    ('ldpgta_bootstrap_uniform', _c.c_int, [_P, _c.c_int, _c.c_int, _P, _P, _c.c_int64, _P, _c.c_uint64, _P, _P, _P]),
    ('ldpgta_likelihoods_uniform_dev', _c.c_int, [_P, _c.c_int, _c.c_int, _P, _c.c_int64, _P, _P]),
    ('ldpgta_synchronize', _c.c_int, [_P]),
    ('ldpgta_stream', _c.c_uint64, [_P]),
    ('ldpgta_launch_count', _c.c_int64, [_P]),
    ('ldpgta_graph_replays', _c.c_int64, [_P]),
    ('ldpgta_set_windows', _c.c_int, [_P, _P, _P, _c.c_int64, _P, _P, _P, _c.c_int, _P, _P]),
    ('ldpgta_bootstrap_stats', _c.c_int, [_P, _c.c_int, _c.c_uint64, _P, _P, _P, _P, _P, _P, _P]),
    ('ldpgta_replay_stats', _c.c_int, [_P, _c.c_int, _P, _P, _P, _P, _P, _P, _P]),
    ('ldpgta_bootstrap_stats_dev', _c.c_int, [_P, _c.c_int, _c.c_uint64, _P, _P]),
    ('ldpgta_finish_segments_dev', _c.c_int, [_P, _P, _P]),
    ('ldpgta_plan_info', _c.c_int, [_P, _P]),
    ('ldpgta_comm_create', _c.c_int, [_P, _c.c_int, _c.c_int, _c.c_int64, _P]),
    ('ldpgta_comm_connect', _c.c_int, [_P, _P]),
    ('ldpgta_allreduce_chrom_dev', _c.c_int, [_P, _P, _c.c_int64]),
    ('ldpgta_comm_destroy', _c.c_int, [_P]),
    ('ldpgta_kernel_timing', _c.c_int, [_P, _c.c_int]),
    ('ldpgta_kernel_times', _c.c_int, [_P, _P, _c.c_int]),
    ('ldpgta_window_stats', _c.c_int, [_P, _P, _P, _c.c_int64, _c.c_int, _P, _P, _P]),
    ('ldpgta_finish_chromosome', _c.c_int, [_P, _c.c_int, _c.c_int, _P]),
    ('ldpgta_bin_stats', _c.c_int, [_P, _P, _c.c_int, _P, _c.c_int64, _P, _c.c_int64, _P]),
    ('ldpgta_host_alloc', _c.c_int, [_c.POINTER(_P), _c.c_int64]),
    ('ldpgta_host_free', _c.c_int, [_P]),
)

_lib = None


class LdpgtaError(Exception):
    """ A failure reported by libldpgta.so (negative return code + message). """
    def __init__(self, code, message):
        super().__init__('libldpgta error %d: %s' % (code, message))
        self.code = code


def load():
    """ Loads the shared library once; raises if it has not been built. """
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError('%s is missing: build it with `python -c "import __graft_entry__ as g; g.build()"` '
                              '(nvcc, sm_100a). The likelihood engine has no CPU fallback.' % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, restype, argtypes in SYMBOLS:
            fn = getattr(lib, name)
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = lib
    return _lib


def check(rc):
    if rc < 0:
        raise LdpgtaError(rc, load().ldpgta_last_error().decode('utf-8', 'replace'))
    return rc


def ptr(a):
    """ Address of a C-contiguous numpy array (or None). """
    if a is None:
        return None
    assert a.flags['C_CONTIGUOUS']
    return a.ctypes.data_as(ctypes.c_void_p)


class PinnedArray:
    """ A numpy array over page-locked host memory from ldpgta_host_alloc. """
    def __init__(self, shape, dtype):
        self.lib = load()
        dtype = np.dtype(dtype)
        n = int(np.prod(shape)) * dtype.itemsize
        self._ptr = ctypes.c_void_p()
        check(self.lib.ldpgta_host_alloc(ctypes.byref(self._ptr), n))
        buf = (ctypes.c_char * max(n, 1)).from_address(self._ptr.value)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def free(self):
        if self._ptr is not None and self._ptr.value:
            self.array = None
            self.lib.ldpgta_host_free(self._ptr)
            self._ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass
