#!/usr/bin/env python
"""Regenerates the golden fixtures in this directory from the reference's own tests.
Run in the build container (needs /root/reference); the JSON files are committed because the
reference tree does not exist on the GPU box.  Only known-answer TABLES are extracted, no code.

  multi_get_basic.json   <- src/test/function_test/base_api/test_basic.cpp:82-600 (forward) and :606-1110 (reverse)
The hand-transcribed tables (compaction rules / ops, range_read limits, value schema, hashkey
transform) live in tables.json with their source lines noted per entry.
"""
import json
import os
import re
import sys

REF = "/root/reference/src/test/function_test/base_api/test_basic.cpp"
HERE = os.path.dirname(os.path.abspath(__file__))

FT = {"FT_NO_FILTER": 0, "FT_MATCH_ANYWHERE": 1, "FT_MATCH_PREFIX": 2, "FT_MATCH_POSTFIX": 3}
PERR = {"PERR_OK": 0, "PERR_INCOMPLETE": 7}


def cstr(s):
    return bytes(s, "utf-8").decode("unicode_escape")


def parse_pairs(text):
    return [[cstr(a), cstr(b)] for a, b in re.findall(r'\{\s*"((?:[^"\\]|\\.)*)"\s*,\s*"((?:[^"\\]|\\.)*)"\s*\}', text)]


def extract(lines, first, last):
    src = "".join(lines[first - 1:last])
    fixture = parse_pairs(re.search(r"kvs\(\{(.*?)\}\);", src, re.S).group(1))
    cases = []
    # blocks at 4-space indentation, preceded by a comment line
    for m in re.finditer(r"\n    // ([^\n]*)\n    \{\n(.*?)\n    \}\n", src, re.S):
        title, body = m.group(1), m.group(2)
        call = re.search(r'client_->multi_get\(\s*"[^"]*",\s*"((?:[^"\\]|\\.)*)",\s*"((?:[^"\\]|\\.)*)",\s*options,\s*new_values'
                         r'(?:,\s*(-?\d+))?(?:,\s*(-?\d+))?\)', body)
        if not call or "client_->set(" in body:
            continue
        opts = {"start_inclusive": True, "stop_inclusive": False, "reverse": False, "sort_key_filter_type": 0,
                "sort_key_filter_pattern": "", "no_value": False}
        for k, v in re.findall(r"options\.(\w+) = ([^;]+);", body):
            v = v.strip()
            if k == "sort_key_filter_type":
                opts[k] = FT[v.split("::")[-1]]
            elif k == "sort_key_filter_pattern":
                opts[k] = cstr(v.strip('"'))
            else:
                opts[k] = v == "true"
        err = re.search(r"ASSERT_EQ\(\s*(PERR_\w+)\s*,", body).group(1)
        if "ASSERT_TRUE(new_values.empty())" in body:
            expect = []
        elif re.search(r"ASSERT_EQ\(kvs, new_values\)", body):
            expect = fixture
        elif "expect_kvs(kvs)" in body:
            erased = [cstr(x) for x in re.findall(r'expect_kvs\.erase\("((?:[^"\\]|\\.)*)"\)', body)]
            expect = [p for p in fixture if p[0] not in erased]
        else:
            em = re.search(r"expect_kvs\(\s*(\{.*?\})\s*\);", body, re.S)
            expect = parse_pairs(em.group(1))
        cases.append({"title": title, "start": cstr(call.group(1)), "stop": cstr(call.group(2)),
                      "max_count": int(call.group(3)) if call.group(3) else 100, "options": opts,
                      "error": PERR[err], "expect": sorted(expect)})
    return {"fixture": fixture, "cases": cases}


def main():
    if not os.path.exists(REF):
        sys.exit("reference tree not present; fixtures are already committed")
    lines = open(REF).read().splitlines(keepends=True)
    out = {"source": "src/test/function_test/base_api/test_basic.cpp",
           "forward": extract(lines, 82, 604), "reverse": extract(lines, 606, 1112)}
    with open(os.path.join(HERE, "multi_get_basic.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("forward cases:", len(out["forward"]["cases"]), "reverse cases:", len(out["reverse"]["cases"]))


if __name__ == "__main__":
    main()
