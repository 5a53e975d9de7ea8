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
    assert lib.stair_version() == 4


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
    for which, struct in enumerate((L.StairModel, L.StairGroup, L.StairBatch, L.StairBuffers, L.StairItabLayout, L.StairTrain)):
        assert ctypes.sizeof(struct) == lib.stair_sizeof(L.i32(which)), struct.__name__
    assert ctypes.sizeof(L.StairGroup) == 9 * 4
    assert ctypes.sizeof(L.StairModel) == 10 * 4 + 2 * 8 * L.W_COUNT
