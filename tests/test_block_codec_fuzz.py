"""Host-side parity of the data-block codec on arbitrary record shapes: the product's run builder / block decoder
(host/host_util.cpp) and the oracle's BlockBuilder / BlockIter (oracle/orc_lsm.cpp) are independent restatements of
RocksDB's block format (varint shared/non_shared/value_len entries, restart array, restart interval 16);
blocks written by one must decode to the same records with the other, for ragged keys and values, long shared
prefixes, empty values, tombstones and several versions of one key."""
import numpy as np
from hypothesis import HealthCheck, given, settings, strategies as st

key_piece = st.binary(min_size=0, max_size=12)


@st.composite
def record_sets(draw):
    """sorted (user key asc, seq desc) records with a fair amount of shared prefixes"""
    stems = draw(st.lists(st.binary(min_size=0, max_size=40), min_size=1, max_size=6, unique=True))
    keys = set()
    for _ in range(draw(st.integers(1, 120))):
        k = draw(st.sampled_from(stems)) + draw(key_piece)
        keys.add(b"\x00\x02" + k if draw(st.booleans()) else k)
    items = []
    seq = 10_000
    for k in sorted(keys):
        for _ in range(draw(st.sampled_from([1, 1, 1, 2, 3]))):  # versions of one key: newest first
            seq -= draw(st.integers(1, 5))
            t = draw(st.sampled_from([1, 1, 1, 0]))
            v = b"" if t == 0 else draw(st.one_of(st.binary(max_size=20), st.binary(min_size=200, max_size=5000)))
            items.append((k, seq + 1_000_000 * (len(items) % 3 == 0), t, v))
        # keep seq descending inside one key
    fixed, last_key, last_seq = [], None, None
    for k, s, t, v in items:
        if k == last_key and s >= last_seq:
            s = last_seq - 1
        fixed.append((k, s, t, v))
        last_key, last_seq = k, s
    return fixed


@settings(max_examples=150, deadline=None, derandomize=True, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])
@given(items=record_sets())
def test_product_and_oracle_codecs_agree(pgs, oracle, items):
    r = pgs.Records.from_list(items)
    br = pgs.build_run(r)                                      # product builder
    assert np.all(br.blk_off % 16 == 0)
    assert pgs.decode_blocks(br).same_as(r)                    # product decoder
    assert oracle.Run.from_blocks(br).records().same_as(r)     # oracle decoder reads the product's blocks
    ob = oracle.BlockRunCPU.from_run(oracle.Run.from_records(r))   # oracle builder
    assert ob.decode().records().same_as(r)
