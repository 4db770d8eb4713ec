"""
oracle/coracle.py -- TEST INFRASTRUCTURE: ctypes front end of oracle/oracle_c.c
(built by oracle/Makefile or on first use with gcc).  Imported only by tests/,
__graft_entry__.smoke() and bench.py as the checker.
"""
import ctypes, os, subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    so, src = os.path.join(_HERE, 'liboracle.so'), os.path.join(_HERE, 'oracle_c.c')
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(['gcc', '-O2', '-shared', '-fPIC', '-o', so, src, '-lm'])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        L.oracle_batch.restype = ctypes.c_int
        L.oracle_llr.restype = ctypes.c_double
        L.oracle_llr.argtypes = [ctypes.c_double, ctypes.c_double]
        _LIB = L
    return _LIB


def pack_rows(ints, n_bits):
    """ Python ints (bit i = haplotype i) -> uint64[len, ceil(n_bits/64)], little-endian words. """
    W = (n_bits + 63) // 64
    buf = b''.join(int(x).to_bytes(8 * W, 'little') for x in ints)
    return np.frombuffer(buf, dtype='<u8').reshape(len(ints), W).copy()


def joint_counts(rows):
    """ rows uint64[n, W] -> uint32[2^n] (index = subset bitmask, [0] = 0). """
    rows = np.ascontiguousarray(rows, dtype=np.uint64)
    n, W = rows.shape
    out = np.zeros(1 << n, dtype=np.uint32)
    lib().oracle_joint_counts(rows.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(n), ctypes.c_int(W),
                              out.ctypes.data_as(ctypes.c_void_p))
    return out


def batch(masks_per_group, n_hap, read_idx, offsets, kind, proportions=None):
    """ masks_per_group: list of uint64[n_reads, W] (same W); returns float64[n_draws, 4]. """
    G = len(masks_per_group)
    masks = [np.ascontiguousarray(m, dtype=np.uint64) for m in masks_per_group]
    W = masks[0].shape[1]
    assert all(m.shape[1] == W for m in masks)
    ptrs = (ctypes.c_void_p * G)(*[m.ctypes.data for m in masks])
    nh = np.asarray(n_hap, dtype=np.uint32)
    idx = np.ascontiguousarray(read_idx, dtype=np.int32)
    off = np.ascontiguousarray(offsets, dtype=np.int64)
    nd = len(off) - 1
    out = np.zeros((nd, 4), dtype=np.float64)
    prop = np.ascontiguousarray(proportions if proportions is not None else np.zeros(G), dtype=np.float64)
    rc = lib().oracle_batch(ptrs, ctypes.c_int(G), nh.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(W),
                            idx.ctypes.data_as(ctypes.c_void_p), off.ctypes.data_as(ctypes.c_void_p),
                            ctypes.c_int64(nd), ctypes.c_int(kind), prop.ctypes.data_as(ctypes.c_void_p),
                            out.ctypes.data_as(ctypes.c_void_p))
    if rc:
        raise ValueError('oracle_batch: unsupported draw size')
    return out
