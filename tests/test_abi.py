"""CPU: the C-ABI library loads and exports every symbol include/stair_b200.h declares (no compute without a GPU)."""
import ctypes
import os
import re

from stair_b200 import _lib as L, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, 'include', 'stair_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(stair_[a-z0-9_]+)\s*\(', src)))


def test_library_builds_and_exports_every_declared_symbol():
    build.build()
    lib = ctypes.CDLL(L.LIB_PATH)
    names = declared_functions()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), 'missing export %s' % n
    assert lib.stair_version() == 9


def test_host_only_entry_points():
    lib = L.lib()
    il = L.StairItabLayout()
    assert lib.stair_itab_layout(L.i32(1000), L.i32(17), ctypes.byref(il)) == 0
    assert il.total == lib.stair_itab_ints(L.i32(1000), L.i32(17))
    assert il.perm == 0 and il.out_slot >= 1000 and il.arg_slot - il.aux_slot >= 1000
    assert L.OP['WORD'] == 0 and L.OP['ARRAY2'] == 18 and L.OP['COUNT'] == 19
    assert L.W['COUNT'] == len([k for k in L.W if k != 'COUNT' and not k.endswith(('_AFTER', '_BETWEEN', '_ACTIONS', '_OBJECTS', '_RELATIONS'))]) \
        or L.W['COUNT'] > 80


def test_struct_sizes_match_header_layout():
    # int32 fields first, then pointers: ctypes mirrors of the C structs must have the C sizes
    lib = L.lib()
    lib.stair_sizeof.restype = ctypes.c_longlong
    for which, struct in enumerate((L.StairModel, L.StairGroup, L.StairBatch, L.StairBuffers, L.StairItabLayout, L.StairTrain,
                                    L.StairAdamSeg)):
        assert ctypes.sizeof(struct) == lib.stair_sizeof(L.i32(which)), struct.__name__
    assert ctypes.sizeof(L.StairGroup) == 9 * 4
    assert ctypes.sizeof(L.StairModel) == 10 * 4 + 2 * 8 * L.W_COUNT


def test_dropout_mask_host_matches_oracle_restatement():
    """The counter-based dropout mask (csrc/stair_common.cuh) evaluated on the host by the library == the numpy restatement the
    oracle injects (oracle/nmn_oracle.py dropout_keep); keep rate ~ 1 - p."""
    import numpy as np
    from oracle import nmn_oracle as orc
    lib = L.lib()
    for p, seed, site, row0, rows, cols in ((0.25, 12345678901234567, 17, 0, 64, 512), (0.5, 7, 3, 1000000, 33, 172),
                                            (0.1, 2 ** 63 + 5, 99, 123456789, 8, 1), (0.0, 1, 2, 3, 4, 5)):
        out = np.zeros((rows, cols), np.uint8)
        rc = lib.stair_dropout_mask_host(ctypes.c_float(p), ctypes.c_uint64(seed), ctypes.c_int(site), ctypes.c_longlong(row0),
                                         ctypes.c_int(rows), ctypes.c_int(cols), out.ctypes.data_as(ctypes.c_void_p))
        assert rc == 0
        want = orc.dropout_keep(seed, site, row0, rows, cols, p) if p > 0 else np.ones((rows, cols), bool)
        assert np.array_equal(out.astype(bool), want)
        if rows * cols > 4000:
            assert abs(out.mean() - (1 - p)) < 0.02
    # different sites / seeds give different masks
    a = orc.dropout_keep(5, 1, 0, 32, 256, 0.25); b = orc.dropout_keep(5, 2, 0, 32, 256, 0.25); c = orc.dropout_keep(6, 1, 0, 32, 256, 0.25)
    assert (a != b).mean() > 0.2 and (a != c).mean() > 0.2
    assert lib.stair_dropout_mask_host(ctypes.c_float(1.0), ctypes.c_uint64(0), 0, ctypes.c_longlong(0), 1, 1, None) != 0
