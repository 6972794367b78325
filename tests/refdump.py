"""Parser of the reference's DocDB debug dumps (the text of ASSERT_DOC_DB_DEBUG_DUMP_STR_EQ in
src/yb/docdb/docdb-test-wrapper.cc / docdb-ttl-test.cc: DocDBDebugDump -> SubDocKey::ToString,
Value::ToString) into (user key bytes, value bytes), so that the reference's expected states can be
replayed verbatim against the oracle and the device logic.

Grammar handled (what the compaction tests use):
  SubDocKey(DocKey(<id>?<hash>?[hashed], [range]), [subkeys; HT{ physical: N [logical: N] [w: N] }]) -> value[; merge flags: N][; ttl: S.MMMs][; timestamp: N]
    <id>    := CoTableId=<uuid>,  |  ColocationId=<n>,
    <hash>  := 0xHHHH,            (then two component lists; without it only the range list follows "[], ")
    entries := "string" | integer | ColumnId(n) | SystemColumnId(n)
    value   := "string" | {} | DEL | null | integer
Lines may be continued with a trailing backslash, exactly as in the reference sources.
"""
import re

import dockv_util as dk
import oracle_py as o


def _entries(text):
    out = []
    text = text.strip()
    pos = 0
    while pos < len(text):
        if text[pos] in ", ":
            pos += 1
            continue
        if text[pos] == '"':
            end = text.index('"', pos + 1)
            out.append(dk.kstr(text[pos + 1:end]))
            pos = end + 1
            continue
        m = re.match(r"(System)?ColumnId\((\d+)\)", text[pos:])
        if m:
            out.append((dk.ksyscol if m.group(1) else dk.kcol)(int(m.group(2))))
            pos += m.end()
            continue
        m = re.match(r"-?\d+", text[pos:])
        if m:
            out.append(dk.kint64(int(m.group(0))))
            pos += m.end()
            continue
        raise ValueError("cannot parse key entries: %r" % text[pos:])
    return out


_LINE = re.compile(
    r"SubDocKey\(DocKey\((?P<dk>.*?)\), \[(?P<sub>[^;\]]*?)(?:; )?HT\{ (?P<ht>[^}]*)\}\]\) -> (?P<val>.*)$")


def _doc_key(text):
    kw = {}
    m = re.match(r"CoTableId=([0-9a-fA-F-]+), ", text)
    if m:
        kw["cotable"] = bytes.fromhex(m.group(1).replace("-", ""))
        text = text[m.end():]
    m = re.match(r"ColocationId=(\d+), ", text)
    if m:
        kw["colocation"] = int(m.group(1))
        text = text[m.end():]
    m = re.match(r"0x([0-9a-fA-F]{4}), \[(.*?)\], \[(.*)\]$", text)
    if m:
        return dk.doc_key(_entries(m.group(3)), hash_code=int(m.group(1), 16), hashed=_entries(m.group(2)), **kw)
    m = re.match(r"\[\], \[(.*)\]$", text)
    if not m:
        raise ValueError("cannot parse DocKey(%s)" % text)
    return dk.doc_key(_entries(m.group(1)), **kw)


def _value(text):
    text = text.strip()
    if text.startswith('"'):                                # a string may itself contain "; "
        end = text.index('"', 1)
        body, rest = text[:end + 1], text[end + 1:]
    else:
        body, _, rest = text.partition(";")
        rest = ";" + rest if rest else ""
    body = body.strip()
    fields = [p.strip() for p in rest.split(";") if p.strip()]
    if body.startswith('"') and body.endswith('"'):
        v = dk.vstr(body[1:-1])
    elif body == "{}":
        v = dk.OBJECT
    elif body == "DEL":
        v = dk.TOMBSTONE
    elif body == "null":
        v = b"$"                                            # ValueEntryType::kNullLow
    elif re.fullmatch(r"-?\d+", body):
        v = b"I" + (int(body) & (2**64 - 1)).to_bytes(8, "big")
    else:
        raise ValueError("cannot parse value %r" % body)
    prefix = b""
    for f in fields:                                        # dockv/value.cc:114-129: merge flags, ttl, timestamp
        name, _, arg = f.partition(": ")
        if name == "merge flags":
            prefix += b"k" + o.unsigned_varint(int(arg))
        elif name == "ttl":
            prefix += b"t" + o.signed_varint(round(float(arg.rstrip("s")) * 1000))
        elif name in ("timestamp", "user_timestamp"):
            prefix += b"u" + (int(arg) & (2**64 - 1)).to_bytes(8, "big")
        else:
            raise ValueError("unknown control field %r" % f)
    return prefix + v


def parse(dump, with_files=False):
    """Returns [(user_key, value)] in dump order (= RocksDB order); with_files=True adds the number of the
    trailing "// file N" comment some reference dumps carry: [(user_key, value, file)]."""
    text = re.sub(r"\\\n\s*", "", dump)
    out = []
    for line in text.splitlines():
        line = line.strip()
        if not line:
            continue
        file_no = None
        fm = re.search(r"\s*// file (\d+)$", line)
        if fm:
            file_no = int(fm.group(1))
            line = line[:fm.start()]
        m = _LINE.match(line)
        if not m:
            raise ValueError("cannot parse dump line %r" % line)
        ht = dict((k, int(v)) for k, v in re.findall(r"(physical|logical|w): (\d+)", m.group("ht")))
        key = dk.sub_doc_key(_doc_key(m.group("dk")), _entries(m.group("sub")),
                             micros=ht["physical"], logical=ht.get("logical", 0), write_id=ht.get("w", 0))
        out.append((key, _value(m.group("val")), file_no) if with_files else (key, _value(m.group("val"))))
    return out
