"""Test-side DocDB key/value builders (python restatement of the dockv ENCODERS, used only to make
test inputs; pinned against the reference's byte-exact expectations in test_oracle_codec.py).

References: dockv/doc_kv_util.cc (zero-encoded strings), dockv/primitive_value.cc:775-900
(KeyEntryValue::AppendToKey), dockv/doc_key.cc (DocKey::AppendTo), dockv/value_type.h.
"""
import oracle_py as o

INT_KEY1 = 123456


def zero_encode(b):
    return b.replace(b"\x00", b"\x00\x01") + b"\x00\x00"


def kstr(s):
    if isinstance(s, str):
        s = s.encode()
    return b"S" + zero_encode(s)


def kstr_desc(s):
    if isinstance(s, str):
        s = s.encode()
    return b"a" + bytes((~c) & 0xff for c in zero_encode(s))


def kint64(v):
    return b"I" + ((v + (1 << 63)) & (2**64 - 1)).to_bytes(8, "big")


def kint32(v):
    return b"H" + ((v + (1 << 31)) & (2**32 - 1)).to_bytes(4, "big")


def kcol(cid):
    return b"K" + o.signed_varint(cid)


def ksyscol(cid):
    return b"J" + o.signed_varint(cid)


def varint_comparable(v, reserved_bits=0):
    """util/varint.cc:102-157 VarInt::EncodeToComparable: sign bit, unary byte count, big-endian magnitude;
    negatives are complemented; `reserved_bits` leading bits are left zero for the caller."""
    if v == 0:
        return bytes([0x80 >> reserved_bits])
    mag = abs(v)
    num_bits = mag.bit_length()
    total = num_bits + 1 + reserved_bits
    num_bytes = (total + 6) // 7
    x = mag | (((1 << (num_bytes + reserved_bits)) - 1) << (num_bytes * 8 - num_bytes - reserved_bits))
    if v < 0:
        x ^= (1 << (num_bytes * 8)) - 1
    out = bytearray(x.to_bytes(num_bytes, "big"))
    if reserved_bits:
        out[0] &= (1 << (8 - reserved_bits)) - 1
    return bytes(out)


def decimal_comparable(digits, exponent, positive=True):
    """util/decimal.cc:270-310 Decimal::EncodeToComparable for value 0.d1d2.. * 10^exponent."""
    if not digits:
        return bytes([128])
    pairs = bytearray()
    n = (len(digits) + 1) // 2
    for i in range(n):
        lo = digits[2 * i + 1] if 2 * i + 1 < len(digits) else 0
        pairs.append((digits[2 * i] * 10 + lo) * 2 + (1 if i < n - 1 else 0))
    out = bytearray(varint_comparable(exponent, 2) + bytes(pairs))
    out[0] |= 0xc0
    if not positive:
        out = bytearray((~c) & 0xff for c in out)
    return bytes(out)


def kbson(b):
    return Enc(b"o" + zero_encode(b))                       # dockv/doc_bson.cc:29-31 BsonKeyToComparableBinary


def kbson_desc(b):
    return Enc(b"p" + bytes((~c) & 0xff for c in zero_encode(b)))


class Enc(bytes):
    """An already encoded key entry (kprim passes it through whatever its first byte is)."""


def kvarint(v):
    return Enc(b"B" + varint_comparable(v))


def kvarint_desc(v):
    return Enc(b"f" + varint_comparable(-v))          # primitive_value.cc:869-872: the negated number


def kdecimal(digits, exponent, positive=True):
    return Enc(b"E" + decimal_comparable(digits, exponent, positive))


def kdecimal_desc(digits, exponent, positive=True):
    return Enc(b"d" + decimal_comparable(digits, exponent, not positive))   # :861-864: the negated number


def kprim(v):
    if isinstance(v, Enc):
        return bytes(v)
    if isinstance(v, bytes) and v[:1] in (b"S", b"I", b"H", b"K", b"J", b"a", b"["):
        return v
    if isinstance(v, (str, bytes)):
        return kstr(v)
    if isinstance(v, int):
        return kint64(v)
    raise TypeError(v)


def doc_key(range_components=(), hash_code=None, hashed=(), cotable=None, colocation=None):
    out = b""
    if cotable is not None:
        out += b"y" + cotable
    elif colocation is not None:
        out += b"0" + colocation.to_bytes(4, "big")
    if hash_code is not None:
        out += b"G" + hash_code.to_bytes(2, "big") + b"".join(kprim(x) for x in hashed) + b"!"
    out += b"".join(kprim(x) for x in range_components) + b"!"
    return out


def sub_doc_key(dk, subkeys=(), micros=None, logical=0, write_id=0, ht=None):
    out = dk + b"".join(kprim(s) for s in subkeys)
    if ht is not None:
        micros, logical, write_id = ht
    if micros is not None:
        out += b"#" + o.encode_doc_ht(micros, logical, write_id)
    return out


def table_tombstone_key(cotable=None, colocation=None, micros=0):
    idp = (b"y" + cotable) if cotable is not None else (b"0" + colocation.to_bytes(4, "big"))
    return idp + b"!" + b"#" + o.encode_doc_ht(micros)


# values (dockv/primitive_value.cc AppendToValue; value_type.h)
def vstr(s):
    if isinstance(s, str):
        s = s.encode()
    return b"S" + s


TOMBSTONE = b"X"
OBJECT = b"{"


def with_ttl(value, ttl_ms):
    return b"t" + o.signed_varint(ttl_ms) + value


def ttl_merge_row(ttl_ms, value=b"$"):
    # merge flags 0x1 = kTtlFlag (dockv/value.h) then TTL then a value
    return b"k" + o.unsigned_varint(1) + b"t" + o.signed_varint(ttl_ms) + value


def with_intent_ht(value, micros, logical=0, write_id=0):
    return b"#" + o.encode_doc_ht(micros, logical, write_id) + value
