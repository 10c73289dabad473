"""Regenerates tests/golden/sst_two_blocks.hex: a hand-assembled two-block BlockBasedTable file (format_version 2, no
compression) written by the pure-Python restatement oracle/sst_py.py from the records below.  Run from the repo root:
    python tests/golden/make_sst_golden.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import sst_py  # noqa: E402


def raw_key(hk: bytes, sk: bytes) -> bytes:
    return len(hk).to_bytes(2, "big") + hk + sk


def value(expire_ts: int, user: bytes) -> bytes:  # pegasus value schema v1: BE32 expire_ts | BE64 timetag | user data
    return expire_ts.to_bytes(4, "big") + (0x0005F5E10000000A).to_bytes(8, "big") + user


BLOCKS = [
    [(raw_key(b"alice", b"a"), 9, 1, value(0, b"first")),
     (raw_key(b"alice", b"a"), 4, 1, value(0, b"older version")),
     (raw_key(b"alice", b"b"), 7, 0, b""),
     (raw_key(b"alice", b"c"), 5, 1, value(300000100, b"ttl"))],
    [(raw_key(b"bob", b""), 8, 1, value(0, b"empty sort key")),
     (raw_key(b"bob", b"x" * 20), 6, 1, value(0, bytes(range(40)))),
     (raw_key(b"carol", b"k1"), 3, 1, value(0, b"z"))],
]

if __name__ == "__main__":
    img = sst_py.write_sst(BLOCKS, restart_interval=2)
    with open(os.path.join(ROOT, "tests", "golden", "sst_two_blocks.hex"), "w") as f:
        h = img.hex()
        f.write("\n".join(h[i:i + 96] for i in range(0, len(h), 96)) + "\n")
    print(len(img), "bytes")
