"""Whole-file TTL expiration (docdb/compaction_file_filter.cc). The oracle's restatement is pinned by the
reference's own known-answer tests, replayed here value for value: docdb/compaction_file_filter-test.cc:137-484
(ExpirationFilterTest). `now` is a fixed HybridTime instead of the clock's: every expectation is relative to it."""
import oracle_py as o

M = 2**64
NOW = (1_790_000_000 * 1_000_000) << 12
NS = 1_000_000_000


def ht_add_seconds(ht, s):            # HybridTime::AddSeconds -> AddMicroseconds (hybrid_time.h:150-161), wraps like uint64
    return (ht + ((s * 1_000_000) << 12)) % M


def ht_add_ms(ht, ms):
    return (ht + ((ms * 1000) << 12)) % M


def usec_ht(us):                      # 1000_usec_ht
    return us << 12


KEEP, DISCARD = False, True
NOEXP, DEFAULT, INVALID = o.NO_EXPIRATION, o.USE_DEFAULT_TTL, o.HT_INVALID


def test_extract_expiration_time():
    # :137-158 ExtractFromNullFileOrFrontier / FileWithNoDefinedExpiration / FileWithExpiration, seen through the filter:
    # a file without a frontier never expires (created = kMax); an unset value expiration reads as kNoExpiration
    assert o.file_filter([None], 1 * NS, NOW) == [KEEP]
    assert o.file_filter([(usec_ht(1000), INVALID)], 1 * NS, NOW) == [KEEP]
    assert o.file_filter([(usec_ht(1000), usec_ht(2000))], 1 * NS, NOW) == [DISCARD]


def test_expiration_no_table_ttl():
    # :160-213 TestExpirationNoTableTTL
    cur, fut, past, ttl = NOW, ht_add_seconds(NOW, 1000), usec_ht(1000), o.MAX_TTL_NS
    E = o.ttl_is_expired
    assert E(NOEXP, o.HT_MAX, ttl, cur) is False                      # 1
    assert E(NOEXP, fut, ttl, cur) is False                           # 2
    assert E(NOEXP, past, ttl, cur) is False                          # 3, 4
    assert E(past, past, ttl, cur) is True                            # 5
    assert E(DEFAULT, past, ttl, cur) is False                        # 6
    assert E(INVALID, past, ttl, cur) is False                        # 7
    assert E(INVALID, past, ttl, cur, o.EXP_TABLE_ONLY) is False      # 8
    assert E(past, past, ttl, cur, o.EXP_TABLE_ONLY) is False         # 9
    assert E(INVALID, past, ttl, cur, o.EXP_TRUST_VALUE) is False     # 10
    assert E(DEFAULT, past, ttl, cur, o.EXP_TRUST_VALUE) is False     # 11
    assert E(past, past, ttl, cur, o.EXP_TRUST_VALUE) is True         # 12


def test_expiration_table_ttl_that_will_not_expire():
    # :215-266
    cur, key, ttl = NOW, ht_add_ms(NOW, -100), 1000 * NS
    E = o.ttl_is_expired
    assert E(NOEXP, o.HT_MAX, ttl, cur) is False
    assert E(NOEXP, key, ttl, cur) is False
    assert E(ht_add_seconds(cur, 1000), key, ttl, cur) is False
    assert E(ht_add_seconds(cur, -1000), key, ttl, cur) is False      # kept to accommodate the table TTL
    assert E(DEFAULT, key, ttl, cur) is False
    assert E(INVALID, key, ttl, cur) is False
    assert E(INVALID, key, ttl, cur, o.EXP_TABLE_ONLY) is False
    assert E(key, key, ttl, cur, o.EXP_TABLE_ONLY) is False
    assert E(INVALID, key, ttl, cur, o.EXP_TRUST_VALUE) is False
    assert E(DEFAULT, key, ttl, cur, o.EXP_TRUST_VALUE) is False
    assert E(ht_add_seconds(cur, -1000), key, ttl, cur, o.EXP_TRUST_VALUE) is True


def test_expiration_table_ttl_that_will_expire():
    # :268-319
    cur, key, ttl = NOW, ht_add_seconds(NOW, -100), 1 * NS
    E = o.ttl_is_expired
    assert E(NOEXP, o.HT_MAX, ttl, cur) is False
    assert E(NOEXP, key, ttl, cur) is False
    assert E(ht_add_seconds(cur, 1000), key, ttl, cur) is False
    assert E(ht_add_seconds(cur, -100), key, ttl, cur) is True
    assert E(DEFAULT, key, ttl, cur) is True
    assert E(INVALID, key, ttl, cur) is False
    assert E(INVALID, key, ttl, cur, o.EXP_TABLE_ONLY) is True
    assert E(ht_add_seconds(cur, 1000), key, ttl, cur, o.EXP_TABLE_ONLY) is True
    assert E(INVALID, key, ttl, cur, o.EXP_TRUST_VALUE) is False
    assert E(DEFAULT, key, ttl, cur, o.EXP_TRUST_VALUE) is True
    assert E(ht_add_seconds(cur, 1000), key, ttl, cur, o.EXP_TRUST_VALUE) is False


A = ht_add_seconds
# (reference test, table TTL, mode, frontiers [(created, value expiration)], expected decisions); the manual retention
# policy's history cutoff is kMax (:71) in all of them
FILTER_CASES = [
    ("TestFilterBasedOnTableTTLOnlyNoTableTTL :321", o.MAX_TTL_NS, o.EXP_NORMAL,
     [(A(NOW, -100), DEFAULT), (A(NOW, 100), DEFAULT), (A(NOW, -10000), DEFAULT), (A(NOW, 10000), DEFAULT)], [KEEP, KEEP, KEEP, KEEP]),
    ("TestFilterBasedOnTableTTLOnly :338", 1 * NS, o.EXP_NORMAL,
     [(A(NOW, -100), DEFAULT), (A(NOW, 100), DEFAULT), (A(NOW, -10000), DEFAULT), (A(NOW, 10000), DEFAULT)], [DISCARD, KEEP, DISCARD, KEEP]),
    ("TestFilterBasedOnTableTTLNoValueTTLData :355", 1 * NS, o.EXP_NORMAL,
     [(A(NOW, -100), INVALID), (A(NOW, 100), INVALID), (A(NOW, -10000), INVALID), (A(NOW, 10000), INVALID)], [KEEP, KEEP, KEEP, KEEP]),
    ("TestFilterBasedOnValueTTLData :372", o.MAX_TTL_NS, o.EXP_NORMAL,
     [(A(NOW, 1), A(NOW, -100)), (NOW, A(NOW, 100)), (A(NOW, -1), A(NOW, -10000)), (A(NOW, 2), A(NOW, 10000))], [KEEP, KEEP, DISCARD, KEEP]),
    ("TestFilterMixTableAndValueTTL :389", 1 * NS, o.EXP_NORMAL,
     [(A(NOW, -100), DEFAULT), (A(NOW, -50), NOEXP), (A(NOW, -20), DEFAULT), (A(NOW, -10), A(NOW, -10))], [DISCARD, KEEP, KEEP, KEEP]),
    ("TestFilterNoTableTTLWithIgnoreValueTTLFlag :407", o.MAX_TTL_NS, o.EXP_TABLE_ONLY,
     [(A(NOW, -100), DEFAULT), (A(NOW, 100), NOEXP), (A(NOW, -10000), A(NOW, -100)), (A(NOW, 10000), INVALID)], [KEEP, KEEP, KEEP, KEEP]),
    ("TestFilterMixTableAndValueTTLWithIgnoreValueTTLFlag :425", 1 * NS, o.EXP_TABLE_ONLY,
     [(A(NOW, -100), DEFAULT), (A(NOW, -50), NOEXP), (A(NOW, -20), INVALID), (A(NOW, -10), A(NOW, -10))], [DISCARD, DISCARD, DISCARD, DISCARD]),
    ("TestFilterNoTableTTLWithTrustValueTTLFlag :445", o.MAX_TTL_NS, o.EXP_TRUST_VALUE,
     [(A(NOW, -10), DEFAULT), (A(NOW, 100), A(NOW, -100)), (A(NOW, -10000), A(NOW, -100)), (A(NOW, -100), INVALID)], [KEEP, KEEP, DISCARD, KEEP]),
    ("TestFilterMixTableAndValueTTLWithTrustValueTTLFlag :463", 1 * NS, o.EXP_TRUST_VALUE,
     [(A(NOW, -100), DEFAULT), (A(NOW, 50), NOEXP), (A(NOW, 10), A(NOW, -10)), (A(NOW, 30), A(NOW, 10))], [DISCARD, KEEP, DISCARD, KEEP]),
]


def test_filter_factory_known_answers():
    for name, ttl, mode, frontiers, want in FILTER_CASES:
        assert o.file_filter(frontiers, ttl, NOW, mode=mode) == want, name


def test_history_cutoff_keeps_files_inside_the_retention_window():
    # compaction_file_filter.cc:196-233: a file whose latest key is not older than the history cutoff is never expired, and it
    # shields every file created after it; the smaller of the two cutoffs counts
    fr = [(A(NOW, -100), DEFAULT), (A(NOW, -50), DEFAULT), (A(NOW, -20), DEFAULT)]
    assert o.file_filter(fr, 1 * NS, NOW) == [DISCARD, DISCARD, DISCARD]
    assert o.file_filter(fr, 1 * NS, NOW, primary_cutoff_ht=A(NOW, -60)) == [DISCARD, KEEP, KEEP]
    assert o.file_filter(fr, 1 * NS, NOW, primary_cutoff_ht=o.HT_MAX, cotables_cutoff_ht=A(NOW, -200)) == [KEEP, KEEP, KEEP]
    assert o.file_filter(fr, 1 * NS, NOW, primary_cutoff_ht=o.HT_INVALID, cotables_cutoff_ht=o.HT_INVALID) == [DISCARD, DISCARD, DISCARD]
