"""Host-side parity of the `user_specified_compaction` env parser: the product's JSON -> binary ops table
(pgs_compaction_ops_parse, host/host_util.cpp) and the oracle's restatement of create_compaction_operations
(src/server/compaction_operation.cpp:162-186, compaction_filter_rule.cpp) must accept the same operations for any
input -- well-formed, partly invalid (the reference skips an operation with an unknown type, bad params or no
valid rule) or garbage -- and neither may crash."""
import json

import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

OP_TYPES = ["COT_DELETE", "COT_UPDATE_TTL", "COT_INVALID", "nope", ""]
RULE_TYPES = ["FRT_HASHKEY_PATTERN", "FRT_SORTKEY_PATTERN", "FRT_TTL_RANGE", "FRT_INVALID", "x"]
MATCH = ["SMT_MATCH_ANYWHERE", "SMT_MATCH_PREFIX", "SMT_MATCH_POSTFIX", "SMT_INVALID", "y"]
TTL_TYPES = ["UTOT_FROM_NOW", "UTOT_FROM_CURRENT", "UTOT_TIMESTAMP", "UTOT_INVALID", "z"]

text = st.text(alphabet=st.characters(min_codepoint=1, max_codepoint=0x7E, blacklist_characters='"\\'), max_size=8)
maybe_bad = lambda s: st.one_of(s, st.just("{"), st.just(""), st.just("[]"), st.just("null"))


@st.composite
def rule(draw):
    t = draw(st.sampled_from(RULE_TYPES))
    if t == "FRT_TTL_RANGE":
        p = {"start_ttl": draw(st.integers(0, 2**32 - 1)), "stop_ttl": draw(st.integers(0, 2**32 - 1))}
        if draw(st.booleans()) and draw(st.booleans()):
            p.pop("stop_ttl")
    else:
        p = {"pattern": draw(text), "match_type": draw(st.sampled_from(MATCH))}
        if draw(st.integers(0, 9)) == 0:
            p.pop("match_type")
    params = draw(maybe_bad(st.just(json.dumps(p))))
    r = {"type": t, "params": params}
    if draw(st.integers(0, 19)) == 0:
        r.pop("params")
    return r


@st.composite
def op(draw):
    t = draw(st.sampled_from(OP_TYPES))
    if t == "COT_UPDATE_TTL":
        p = json.dumps({"type": draw(st.sampled_from(TTL_TYPES)), "value": draw(st.integers(0, 2**32 - 1))})
    else:
        p = ""
    o = {"type": t, "params": draw(maybe_bad(st.just(p))), "rules": draw(st.lists(rule(), max_size=3))}
    if draw(st.integers(0, 19)) == 0:
        o.pop(draw(st.sampled_from(["type", "params", "rules"])))
    return o


def counts(pgs, oracle, s):
    try:
        g = int.from_bytes(pgs.parse_ops(s)[:4].tobytes(), "little")
    except pgs.PegasusError:
        g = -1
    o = len(oracle.Ops(s))
    return g, o


@settings(max_examples=300, deadline=None, derandomize=True, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(ops=st.lists(op(), max_size=4))
def test_structured_ops(pgs, oracle, ops):
    s = json.dumps({"ops": ops})
    g, o = counts(pgs, oracle, s)
    assert max(g, 0) == o, s


@settings(max_examples=300, deadline=None, derandomize=True, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(s=st.one_of(st.text(max_size=80), st.from_regex(r'\{"ops":\[[\{\}\[\]",:a-zA-Z0-9_ ]{0,60}', fullmatch=True)))
def test_garbage_never_crashes(pgs, oracle, s):
    try:
        s.encode()
    except UnicodeEncodeError:
        return
    g, o = counts(pgs, oracle, s)
    assert max(g, 0) == o, s
