"""TEST-ONLY binding of tests/host_harness (device logic compiled for the CPU)."""
import ctypes as C
import os
import subprocess

import numpy as np

import oracle_py as o

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "host_harness")
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        subprocess.check_call(["make", "-s", "-C", _DIR], stderr=subprocess.DEVNULL)
        L = C.CDLL(os.path.join(_DIR, "libhostharness.so"))
        vp, u64 = C.c_void_p, C.c_uint64
        L.hh_compact.argtypes = [C.c_int, vp, vp, vp, vp, vp, C.c_int, u64, C.c_int64, C.c_int, u64, C.c_int, u64,
                                 C.c_char_p, u64, C.c_char_p, u64, C.c_char_p, u64, u64]
        for f in ("hh_keys", "hh_vals", "hh_koff", "hh_voff"):
            getattr(L, f).restype = vp
        L.hh_num.restype = u64
        L.hh_group_prefix_len.argtypes = [C.c_char_p, C.c_int, C.c_int]
        _LIB = L
    return _LIB


def compact_runs(runs, params):
    """Same contract as oracle_py.compact_runs but through the device logic. Returns kv list or
    raises RuntimeError(dev error code)."""
    L = lib()
    flat = [kv for r in runs for kv in r]
    starts = np.zeros(len(runs) + 1, np.uint64)
    starts[1:] = np.cumsum([len(r) for r in runs])
    kb, ko = o._flat([k for k, _ in flat])
    vb, vo = o._flat([v for _, v in flat])
    luk = params._luk
    if luk is None:
        lasts = [r[-1][0][:-8] for r in runs if r]
        luk = max(lasts) if lasts else b""
    rc = L.hh_compact(len(runs), starts.ctypes.data, kb.ctypes.data, ko.ctypes.data, vb.ctypes.data, vo.ctypes.data,
                      params.retention_enabled, params.primary_cutoff_ht, params.table_ttl_ns,
                      params.retain_delete_markers, params.other_min_ht, params.bottommost_level,
                      params.last_sequence, luk, len(luk), params._lo, len(params._lo), params._up, len(params._up),
                      params.cotables_cutoff_ht)
    if rc != 0:
        raise RuntimeError("device logic error %d" % rc)
    n = L.hh_num()

    def arr(ptr, count, ty):
        if count == 0:
            return np.zeros(0, ty)
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(np.ctypeslib.as_ctypes_type(ty))), (count,)).copy()
    koff = arr(L.hh_koff(), n + 1, np.uint64)
    voff = arr(L.hh_voff(), n + 1, np.uint64)
    keys = arr(L.hh_keys(), int(koff[-1]), np.uint8).tobytes()
    vals = arr(L.hh_vals(), int(voff[-1]), np.uint8).tobytes()
    return [(keys[int(koff[i]):int(koff[i + 1])], vals[int(voff[i]):int(voff[i + 1])]) for i in range(n)]


# ---- the Snappy kernels' source on emulated warps (tests/host_harness/warp_emu.cc) -------------------------------------
_WARP = None


def warp_lib():
    global _WARP
    if _WARP is None:
        subprocess.check_call(["make", "-s", "-C", _DIR], stderr=subprocess.DEVNULL)
        L = C.CDLL(os.path.join(_DIR, "libwarpemu.so"))
        L.we_compress_table.restype = C.c_uint64
        L.we_compress_table.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_void_p, C.c_void_p]
        L.we_uncompress_table.restype = C.c_uint64
        L.we_uncompress_table.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64, C.c_void_p]
        _WARP = L
    return _WARP


def warp_compress_table(data_file: bytes, offsets, variant=0):
    """k_snappy_compress<variant> + k_snappy_gather over an uncompressed data file (blocks + trailers back to back;
    offsets = the blocks' start offsets): (compressed data file, final block offsets incl. the end)."""
    raw = np.frombuffer(data_file, np.uint8)
    off = np.array(list(offsets) + [len(data_file)], np.uint64)
    out = np.zeros(len(data_file) + 64, np.uint8)
    foff = np.zeros(len(off), np.uint64)
    n = warp_lib().we_compress_table(raw.ctypes.data, off.ctypes.data, len(off) - 1, variant, out.ctypes.data, foff.ctypes.data)
    return out[:n].tobytes(), [int(x) for x in foff]


def warp_uncompress_table(data_file: bytes, offsets, sizes):
    """k_snappy_sizes + k_snappy_decode: (uncompressed image, its block offsets incl. the end); RuntimeError(device error code)
    when the kernels flag a block."""
    raw = np.frombuffer(data_file, np.uint8)
    off = np.array(list(offsets), np.uint64)
    sz = np.array(list(sizes), np.uint32)
    cap = 64 + sum(int(s) for s in sz) * 40 + 5 * len(sz)
    out = np.zeros(cap, np.uint8)
    ooff = np.zeros(len(off) + 1, np.uint64)
    n = warp_lib().we_uncompress_table(raw.ctypes.data, raw.size, off.ctypes.data, sz.ctypes.data, len(off), out.ctypes.data, cap, ooff.ctypes.data)
    if n > 2**63:
        raise RuntimeError(2**64 - 1 - n)
    return out[:n].tobytes(), [int(x) for x in ooff]
