"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes binding of oracle/liboracle.so (the CPU restatement of the reference compaction path) for
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.  Nothing under
yugabyte-db_b200/ imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
_LIB = None

HT_MIN = 0
HT_MAX = 2**64 - 1
HT_INVALID = 2**64 - 2
MAX_TTL_NS = 2**63 - 1
MAX_SEQ = (1 << 56) - 1
YB_EPOCH_US = 1500000000 * 1000000


def build():
    subprocess.check_call(["make", "-s", "-C", _DIR])


def varint(b):
    """Decode one varint64 from the start of b."""
    v = shift = 0
    for c in b:
        v |= (c & 0x7f) << shift
        shift += 7
        if not c & 0x80:
            break
    return v


class TableOptions(C.Structure):
    _fields_ = [("block_size", C.c_uint32), ("block_restart_interval", C.c_int32),
                ("key_encoding", C.c_int32), ("block_size_deviation", C.c_int32),
                ("index_block_size", C.c_uint32), ("min_keys_per_index_block", C.c_uint32),
                ("filter_policy", C.c_int32), ("filter_block_size", C.c_uint32), ("compression", C.c_int32)]

    def __init__(self, block_size=32768, restart=16, key_encoding=1, deviation=10,
                 index_block_size=32768, min_keys_per_index_block=100, filter_policy=0, filter_block_size=65536, compression=0):
        super().__init__(block_size, restart, key_encoding, deviation, index_block_size,
                         min_keys_per_index_block, filter_policy, filter_block_size, compression)


class CompactionParams(C.Structure):
    _fields_ = [("bottommost_level", C.c_int32), ("last_sequence", C.c_uint64),
                ("largest_user_key", C.c_char_p), ("largest_user_key_len", C.c_uint64),
                ("has_largest_user_key", C.c_int32), ("retention_enabled", C.c_int32),
                ("primary_cutoff_ht", C.c_uint64), ("cotables_cutoff_ht", C.c_uint64),
                ("table_ttl_ns", C.c_int64), ("retain_delete_markers", C.c_int32),
                ("other_min_ht", C.c_uint64),
                ("lower_bound", C.c_char_p), ("lower_len", C.c_uint64),
                ("upper_bound", C.c_char_p), ("upper_len", C.c_uint64)]

    def __init__(self, bottommost=True, last_sequence=MAX_SEQ, largest_user_key=None,
                 retention=True, cutoff_ht=HT_MIN, cotables_cutoff_ht=HT_INVALID,
                 table_ttl_ns=MAX_TTL_NS, retain_delete_markers=False, other_min_ht=HT_MAX,
                 lower=b"", upper=b""):
        super().__init__()
        self.bottommost_level = int(bottommost)
        self.last_sequence = last_sequence
        self._luk = largest_user_key
        self.largest_user_key = largest_user_key
        self.largest_user_key_len = len(largest_user_key) if largest_user_key is not None else 0
        self.has_largest_user_key = int(largest_user_key is not None)
        self.retention_enabled = int(retention)
        self.primary_cutoff_ht = cutoff_ht
        self.cotables_cutoff_ht = cotables_cutoff_ht
        self.table_ttl_ns = table_ttl_ns
        self.retain_delete_markers = int(retain_delete_markers)
        self.other_min_ht = other_min_ht
        self._lo, self._up = lower, upper
        self.lower_bound, self.lower_len = lower, len(lower)
        self.upper_bound, self.upper_len = upper, len(upper)


class Stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in
                ("num_input_records", "num_output_records", "num_dropped_hidden",
                 "num_dropped_obsolete", "num_dropped_feed", "in_key_bytes", "in_val_bytes",
                 "out_key_bytes", "out_val_bytes", "kv_hash")] + [("seconds", C.c_double)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class GenConfig(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("num_rows", C.c_uint64), ("cols", C.c_uint32),
                ("versions", C.c_uint32), ("num_files", C.c_uint32), ("value_len", C.c_uint32),
                ("base_micros", C.c_uint64), ("tombstone_per_1024", C.c_uint32),
                ("tombstone_newest", C.c_uint32), ("row_offset", C.c_uint64),
                ("hash_rows_total", C.c_uint64)]

    def __init__(self, seed=1, num_rows=1000, cols=1, versions=1, num_files=2, value_len=256,
                 base_micros=1790000000 * 1000000, tombstone_per_1024=0, tombstone_newest=0,
                 row_offset=0, hash_rows_total=0):
        super().__init__(seed, num_rows, cols, versions, num_files, value_len, base_micros,
                         tombstone_per_1024, tombstone_newest, row_offset, hash_rows_total)


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.path.join(_DIR, "liboracle.so")
    if not os.path.exists(path):
        build()
    L = C.CDLL(path)
    vp, u64, u8p = C.c_void_p, C.c_uint64, C.POINTER(C.c_uint8)
    L.orc_last_error.restype = C.c_char_p
    L.orc_encode_doc_ht.argtypes = [u64, C.c_uint32, C.c_char_p]
    L.orc_decode_doc_ht.argtypes = [C.c_char_p, u64, C.POINTER(u64), C.POINTER(C.c_uint32)]
    L.orc_signed_varint.argtypes = [C.c_int64, C.c_char_p]
    L.orc_unsigned_varint.argtypes = [u64, C.c_char_p]
    L.orc_decode_signed_varint.argtypes = [C.c_char_p, u64, C.POINTER(C.c_int64)]
    L.orc_crc32c.argtypes = [C.c_char_p, u64]
    L.orc_crc32c.restype = C.c_uint32
    L.orc_crc32c_mask.argtypes = [C.c_uint32]
    L.orc_crc32c_mask.restype = C.c_uint32
    L.orc_subdockey_ends.argtypes = [C.c_char_p, u64, C.POINTER(u64), C.c_int]
    L.orc_shortest_separator.argtypes = [C.c_char_p, u64, C.c_char_p, u64, C.c_char_p, u64]
    L.orc_sst_build.argtypes = [u64, vp, vp, vp, vp, C.POINTER(TableOptions)]
    L.orc_sst_build.restype = vp
    L.orc_sst_from_bytes.argtypes = [C.c_char_p, u64, C.c_char_p, u64]
    L.orc_sst_from_bytes.restype = vp
    L.orc_sst_free.argtypes = [vp]
    for f in ("data_size", "meta_size", "num_entries", "raw_key_bytes", "raw_val_bytes", "num_blocks"):
        getattr(L, "orc_sst_" + f).argtypes = [vp]
        getattr(L, "orc_sst_" + f).restype = u64
    L.orc_sst_data.argtypes = [vp]
    L.orc_sst_data.restype = vp
    L.orc_sst_meta.argtypes = [vp]
    L.orc_sst_meta.restype = vp
    L.orc_sst_block_handles.argtypes = [vp, vp, vp]
    L.orc_sst_key_encoding.argtypes = [vp]
    L.orc_sst_read_all.argtypes = [vp, C.c_int]
    L.orc_sst_read_all.restype = vp
    L.orc_compact.argtypes = [C.c_int, C.POINTER(vp), C.POINTER(u64), C.POINTER(CompactionParams),
                              C.POINTER(TableOptions), C.c_int, C.c_int]
    L.orc_compact.restype = vp
    L.orc_compact_runs.argtypes = [C.c_int, vp, vp, vp, vp, vp, C.POINTER(CompactionParams)]
    L.orc_compact_runs.restype = vp
    L.orc_result_free.argtypes = [vp]
    L.orc_result_error.argtypes = [vp]
    L.orc_result_error.restype = C.c_char_p
    L.orc_result_stats.argtypes = [vp]
    L.orc_result_stats.restype = C.POINTER(Stats)
    L.orc_result_sst.argtypes = [vp]
    L.orc_result_sst.restype = vp
    for f in ("num_kv", "keys_size", "vals_size"):
        getattr(L, "orc_result_" + f).argtypes = [vp]
        getattr(L, "orc_result_" + f).restype = u64
    for f in ("keys", "vals", "koff", "voff"):
        getattr(L, "orc_result_" + f).argtypes = [vp]
        getattr(L, "orc_result_" + f).restype = vp
    L.orc_gen_sst.argtypes = [C.POINTER(GenConfig), C.c_uint32, C.POINTER(TableOptions)]
    L.orc_gen_sst.restype = vp
    L.orc_gen_ssts.argtypes = [C.POINTER(GenConfig), C.POINTER(TableOptions), C.POINTER(vp), C.c_int]
    _LIB = L
    return L


# ------------------------------------------------------------------------------------------------
def encode_doc_ht(micros, logical=0, write_id=0):
    buf = C.create_string_buffer(32)
    n = lib().orc_encode_doc_ht((micros << 12) + logical, write_id, buf)
    return buf.raw[:n]


def encode_doc_ht_repr(ht_repr, write_id=0):
    buf = C.create_string_buffer(32)
    n = lib().orc_encode_doc_ht(ht_repr, write_id, buf)
    return buf.raw[:n]


def decode_doc_ht(b):
    ht, wid = C.c_uint64(), C.c_uint32()
    if lib().orc_decode_doc_ht(b, len(b), C.byref(ht), C.byref(wid)) != 0:
        raise ValueError(lib().orc_last_error().decode())
    return ht.value >> 12, ht.value & 0xfff, wid.value


def signed_varint(v):
    buf = C.create_string_buffer(16)
    n = lib().orc_signed_varint(v, buf)
    return buf.raw[:n]


def unsigned_varint(v):
    buf = C.create_string_buffer(16)
    n = lib().orc_unsigned_varint(v, buf)
    return buf.raw[:n]


def crc32c(b):
    return lib().orc_crc32c(b, len(b))


def subdockey_ends(key):
    ends = (C.c_uint64 * 32)()
    n = lib().orc_subdockey_ends(key, len(key), ends, 32)
    if n < 0:
        raise ValueError(lib().orc_last_error().decode())
    return list(ends[:n])


def ht_from_micros(micros, logical=0):
    return (micros << 12) + logical


def ikey(user_key, seq, vtype=1):
    return user_key + int((seq << 8) | vtype).to_bytes(8, "little")


def _flat(items):
    offs = np.zeros(len(items) + 1, dtype=np.uint64)
    if items:
        offs[1:] = np.cumsum([len(x) for x in items], dtype=np.uint64)
    blob = np.frombuffer(b"".join(items), dtype=np.uint8) if items else np.zeros(0, np.uint8)
    return np.ascontiguousarray(blob), offs


class Sst:
    """Owns an orc_sst handle."""

    def __init__(self, handle):
        if not handle:
            raise RuntimeError("oracle: " + lib().orc_last_error().decode())
        self.h = handle
        self._own = True

    @classmethod
    def borrowed(cls, handle):
        s = cls.__new__(cls)
        s.h = handle
        s._own = False
        return s

    def __del__(self):
        if getattr(self, "_own", False) and self.h and _LIB is not None:
            _LIB.orc_sst_free(self.h)
            self.h = None

    @classmethod
    def build(cls, kvs, opts=None):
        """kvs: list of (internal_key, value) in internal-key order."""
        opts = opts or TableOptions()
        kb, ko = _flat([k for k, _ in kvs])
        vb, vo = _flat([v for _, v in kvs])
        return cls(lib().orc_sst_build(len(kvs), kb.ctypes.data, ko.ctypes.data, vb.ctypes.data,
                                       vo.ctypes.data, C.byref(opts)))

    @classmethod
    def from_bytes(cls, meta, data):
        return cls(lib().orc_sst_from_bytes(meta, len(meta), data, len(data)))

    @classmethod
    def generate(cls, cfg, file_index, opts=None):
        opts = opts or TableOptions()
        return cls(lib().orc_gen_sst(C.byref(cfg), file_index, C.byref(opts)))

    @classmethod
    def generate_all(cls, cfg, opts=None, max_threads=None):
        opts = opts or TableOptions()
        arr = (C.c_void_p * cfg.num_files)()
        rc = lib().orc_gen_ssts(C.byref(cfg), C.byref(opts), arr, max_threads or os.cpu_count() or 1)
        if rc != 0:
            raise RuntimeError("oracle: " + lib().orc_last_error().decode())
        return [cls(arr[i]) for i in range(cfg.num_files)]

    # zero-copy numpy views (valid while self is alive)
    def data_view(self):
        n = lib().orc_sst_data_size(self.h)
        return np.ctypeslib.as_array(C.cast(lib().orc_sst_data(self.h), C.POINTER(C.c_uint8)), (n,)) if n else np.zeros(0, np.uint8)

    def meta_view(self):
        n = lib().orc_sst_meta_size(self.h)
        return np.ctypeslib.as_array(C.cast(lib().orc_sst_meta(self.h), C.POINTER(C.c_uint8)), (n,)) if n else np.zeros(0, np.uint8)

    @property
    def data(self):
        return self.data_view().tobytes()

    @property
    def meta(self):
        return self.meta_view().tobytes()

    @property
    def num_entries(self):
        return lib().orc_sst_num_entries(self.h)

    @property
    def raw_bytes(self):
        return lib().orc_sst_raw_key_bytes(self.h) + lib().orc_sst_raw_val_bytes(self.h)

    def block_handles(self):
        n = lib().orc_sst_num_blocks(self.h)
        off = np.zeros(n, np.uint64)
        sz = np.zeros(n, np.uint64)
        lib().orc_sst_block_handles(self.h, off.ctypes.data, sz.ctypes.data)
        return off, sz

    @property
    def key_encoding(self):
        return lib().orc_sst_key_encoding(self.h)

    def read_all(self, verify=True):
        r = Result(lib().orc_sst_read_all(self.h, int(verify)))
        return r.kv_list()

    def _meta_dump(self):
        L = lib()
        L.orc_sst_meta_dump.restype = C.c_uint64
        L.orc_sst_meta_dump.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64]
        n = L.orc_sst_meta_dump(self.h, None, 0)
        buf = C.create_string_buffer(int(n))
        L.orc_sst_meta_dump(self.h, buf, n)
        raw, pos, out = buf.raw, 0, []
        for _ in range(2):
            cnt = int.from_bytes(raw[pos:pos + 4], "little"); pos += 4
            items = []
            for _ in range(cnt):
                pair = []
                for _ in range(2):
                    ln = int.from_bytes(raw[pos:pos + 4], "little"); pos += 4
                    pair.append(raw[pos:pos + ln]); pos += ln
                items.append(tuple(pair))
            out.append(items)
        return out

    def properties(self):
        """Properties block of the metadata file: {name: raw value bytes}."""
        return {k.decode(): v for k, v in self._meta_dump()[0]}

    def filter_blocks(self):
        """[(filter index key, filter block contents)] in index order."""
        return self._meta_dump()[1]


class Result:
    def __init__(self, handle):
        self.h = handle
        err = lib().orc_result_error(handle)
        if err:
            msg = err.decode()
            lib().orc_result_free(handle)
            self.h = None
            raise RuntimeError("oracle: " + msg)

    def __del__(self):
        if getattr(self, "h", None) and _LIB is not None:
            _LIB.orc_result_free(self.h)
            self.h = None

    @property
    def stats(self):
        return lib().orc_result_stats(self.h).contents

    def sst(self):
        h = lib().orc_result_sst(self.h)
        if not h:
            return None
        s = Sst.borrowed(h)
        s._keepalive = self
        return s

    def user_values(self):
        """(smallest, largest): {tag: encoded key component} — FileMetaData user boundary values as
        DocDBCompactionFeed::UpdateBoundaryValues accumulates them (tag = 10 + range component index)."""
        L = lib()
        L.orc_result_num_user_values.argtypes = [C.c_void_p, C.c_int]
        L.orc_result_num_user_values.restype = C.c_uint32
        L.orc_result_user_value.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
        L.orc_result_user_value.restype = C.c_uint32
        out = []
        for which in (0, 1):
            d = {}
            for i in range(L.orc_result_num_user_values(self.h, which)):
                p, n = C.c_void_p(), C.c_uint64()
                tag = L.orc_result_user_value(self.h, which, i, C.byref(p), C.byref(n))
                d[tag] = C.string_at(p, n.value)
            out.append(d)
        return tuple(out)

    def flat(self):
        L = lib()
        n = L.orc_result_num_kv(self.h)
        ks, vs = L.orc_result_keys_size(self.h), L.orc_result_vals_size(self.h)

        def arr(ptr, count, ty):
            if count == 0:
                return np.zeros(0, ty)
            return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(np.ctypeslib.as_ctypes_type(ty))), (count,)).copy()
        keys = arr(L.orc_result_keys(self.h), ks, np.uint8)
        vals = arr(L.orc_result_vals(self.h), vs, np.uint8)
        koff = arr(L.orc_result_koff(self.h), n + 1, np.uint64)
        voff = arr(L.orc_result_voff(self.h), n + 1, np.uint64)
        return keys, koff, vals, voff

    def kv_list(self):
        keys, koff, vals, voff = self.flat()
        kb, vb = keys.tobytes(), vals.tobytes()
        return [(kb[int(koff[i]):int(koff[i + 1])], vb[int(voff[i]):int(voff[i + 1])])
                for i in range(len(koff) - 1)]


COLLECT_KV = 1
BUILD_SST = 2
NO_HASH = 4


class _CotableFilters(C.Structure):
    _fields_ = [("db_oids", C.c_void_p), ("hybrid_times", C.c_void_p), ("n", C.c_uint64)]


def compact(ssts, params=None, opts=None, mode=COLLECT_KV | BUILD_SST, verify=True, ht_filters=None, cotable_filters=None):
    """cotable_filters: per input None or (sorted database oids, hybrid times) — the per-database part of
    user_filter_data (docdb_rocksdb_util.cc:503-509)."""
    params = params or CompactionParams()
    opts = opts or TableOptions()
    arr = (C.c_void_p * len(ssts))(*[s.h for s in ssts])
    filt = None
    if ht_filters is not None:
        filt = (C.c_uint64 * len(ssts))(*ht_filters)
    if cotable_filters is None:
        return Result(lib().orc_compact(len(ssts), arr, filt, C.byref(params), C.byref(opts), mode, int(verify)))
    cf = (_CotableFilters * len(ssts))()
    keep = []
    for i, f in enumerate(cotable_filters):
        if not f:
            continue
        oids = np.ascontiguousarray(f[0], dtype=np.uint32)
        hts = np.ascontiguousarray(f[1], dtype=np.uint64)
        keep += [oids, hts]
        cf[i] = _CotableFilters(oids.ctypes.data, hts.ctypes.data, oids.size)
    L = lib()
    L.orc_compact2.restype = C.c_void_p
    L.orc_compact2.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    return Result(L.orc_compact2(len(ssts), arr, filt, cf, C.byref(params), C.byref(opts), mode, int(verify)))


EXP_NORMAL, EXP_TABLE_ONLY, EXP_TRUST_VALUE = 0, 1, 2
NO_EXPIRATION, USE_DEFAULT_TTL = HT_MAX, HT_MIN + 1      # dockv/doc_ttl_util.h:65-73


def ttl_is_expired(ttl_expiration_ht, created_ht, table_ttl_ns, now, mode=EXP_NORMAL):
    """docdb::TtlIsExpired (compaction_file_filter.cc:126-144)."""
    L = lib()
    L.orc_ttl_is_expired.argtypes = [C.c_uint64, C.c_uint64, C.c_int64, C.c_uint64, C.c_int]
    return bool(L.orc_ttl_is_expired(ttl_expiration_ht, created_ht, table_ttl_ns, now, mode))


def file_filter(frontiers, table_ttl_ns, now, primary_cutoff_ht=HT_MAX, cotables_cutoff_ht=HT_INVALID, mode=EXP_NORMAL):
    """DocDBCompactionFileFilterFactory::CreateCompactionFileFilter over the files + Filter of each
    (compaction_file_filter.cc:150-243). frontiers: per file None (no largest user frontier) or (hybrid_time,
    max_value_level_ttl_expiration_time or HT_INVALID). Returns a list of booleans, True = kDiscard."""
    n = len(frontiers)
    has = np.array([f is not None for f in frontiers], np.uint8)
    created = np.array([f[0] if f is not None else 0 for f in frontiers], np.uint64)
    vttl = np.array([f[1] if f is not None else HT_INVALID for f in frontiers], np.uint64)
    out = np.zeros(max(n, 1), np.uint8)
    L = lib()
    L.orc_file_filter.argtypes = [C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.c_void_p]
    L.orc_file_filter.restype = None
    L.orc_file_filter(n, has.ctypes.data, created.ctypes.data, vttl.ctypes.data, table_ttl_ns, primary_cutoff_ht, cotables_cutoff_ht, now, mode, out.ctypes.data)
    return [bool(x) for x in out[:n]]


def snappy_compress(raw: bytes) -> bytes:
    """The repository's Snappy-format encoder (oracle_sst.cc SnappyCompress)."""
    L = lib()
    L.orc_snappy_compress.restype = C.c_int64
    L.orc_snappy_compress.argtypes = [C.c_char_p, C.c_uint64, C.c_char_p, C.c_uint64]
    n = L.orc_snappy_compress(raw, len(raw), None, 0)
    buf = C.create_string_buffer(max(1, n))
    assert L.orc_snappy_compress(raw, len(raw), buf, n) == n
    return buf.raw[:n]


def snappy_uncompress(comp: bytes) -> bytes:
    L = lib()
    L.orc_snappy_uncompress.restype = C.c_int64
    L.orc_snappy_uncompress.argtypes = [C.c_char_p, C.c_uint64, C.c_char_p, C.c_uint64]
    n = L.orc_snappy_uncompress(comp, len(comp), None, 0)
    if n < 0:
        raise ValueError("snappy: malformed stream")
    buf = C.create_string_buffer(max(1, n))
    assert L.orc_snappy_uncompress(comp, len(comp), buf, n) == n
    return buf.raw[:n]


def compact_runs(runs, params=None):
    """runs: list of sorted lists of (internal_key, value)."""
    params = params or CompactionParams()
    flat = [kv for r in runs for kv in r]
    starts = np.zeros(len(runs) + 1, np.uint64)
    starts[1:] = np.cumsum([len(r) for r in runs])
    kb, ko = _flat([k for k, _ in flat])
    vb, vo = _flat([v for _, v in flat])
    return Result(lib().orc_compact_runs(len(runs), starts.ctypes.data, kb.ctypes.data, ko.ctypes.data,
                                         vb.ctypes.data, vo.ctypes.data, C.byref(params)))
