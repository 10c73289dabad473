"""CPU models of the index arithmetic behind two experimental k_merge items (csrc/compact.cu, EXP instantiation):
the chunk ownership of the staged-heads writer (every entry byte written exactly once, nothing outside the entries) and
the sample-then-refine rank search (positions equal a plain lower bound).  Restatements for review, not the CUDA code."""
import random
def sim_staged(seed):
    rnd = random.Random(seed)
    nblocks = rnd.randint(1, 4)
    entries = []  # (e0, hs, vl, block)
    cuts = [0]
    off = 0
    for b in range(nblocks):
        off = (off + 15) & ~15
        n = rnd.randint(1, 20)
        for i in range(n):
            hs = rnd.randint(11, 70) if rnd.random() < 0.2 else rnd.randint(11, 20)
            vl = rnd.choice([0, 0, 1, 5, 15, 16, 17, 31, 32, 100, 268, 300, 1000])
            entries.append((off, hs, vl, b))
            off += hs + vl
        cuts.append(len(entries))
        off += 4 * ((n + 15) // 16 + 1)  # restart array
    total = (off + 15) & ~15
    img = [None] * total
    def src(q, o):
        e0, hs, vl, b = entries[q]
        return (q, 'h', o) if o < hs else (q, 'v', o - hs)
    expect = [None] * total
    for q, (e0, hs, vl, b) in enumerate(entries):
        for o in range(hs + vl):
            expect[e0 + o] = src(q, o)
    writes = [0] * total
    for p, (e0, hs, vl, b) in enumerate(entries):
        # (b1) interior value chunks
        if vl:
            dv = e0 + hs
            lead = dv & 15
            nch = (lead + vl + 15) >> 4
            for c in range(nch):
                if (c == 0 and lead != 0) or lead + vl < (c << 4) + 16:
                    continue
                a = dv - lead + (c << 4)
                for i in range(16):
                    img[a + i] = (p, 'v', a + i - dv); writes[a + i] += 1
        # (b2) owned boundary chunks
        s0 = (e0 + 15) & ~15; hend = e0 + hs; end = hend + vl
        nh = (hend - s0 + 15) >> 4 if hend > s0 else 0
        sl = (end - 1) & ~15
        extra = 1 if (sl >= s0 + 16 * nh and sl + 16 > end) else 0
        qend = cuts[b + 1]
        for bi in range(nh + extra):
            sc = s0 + 16 * bi if bi < nh else sl
            q = p
            valid = 16
            for i in range(16):
                if i >= valid: break
                pos = sc + i
                while pos >= entries[q][0] + entries[q][1] + entries[q][2]:
                    q += 1
                    if q >= qend: break
                if q >= qend:
                    valid = i
                else:
                    img[pos] = src(q, pos - entries[q][0]); writes[pos] += 1
    for x in range(total):
        if expect[x] is not None:
            assert img[x] == expect[x], (seed, x, img[x], expect[x])
            assert writes[x] == 1, (seed, x, writes[x])
        else:
            assert writes[x] == 0, (seed, x, 'wrote outside entries')
for s in range(3000):
    sim_staged(s)

import bisect
SS = 8
def sim_sample(seed):
    rnd = random.Random(seed)
    nv = rnd.randint(0, 70); no = rnd.randint(0, 70)
    pool = rnd.sample(range(1000), nv + no)
    a = sorted(pool[:nv]); o = sorted(pool[nv:])
    want = [bisect.bisect_left(o, x) for x in a]
    pos = [None] * nv
    nsamp = (nv - 1 + SS - 1) // SS + 1 if nv else 0
    samples = set()
    for si in range(nsamp):
        rel = min(si * SS, nv - 1)
        if si > 0 and rel == (si - 1) * SS: continue
        pos[rel] = bisect.bisect_left(o, a[rel]); samples.add(rel)
    for rel in range(nv):
        if rel % SS == 0 or rel + 1 == nv:
            assert rel in samples, (seed, rel, nv)
            continue
        s0 = rel - rel % SS; s1 = min(s0 + SS, nv - 1)
        assert pos[s0] is not None and pos[s1] is not None
        lo, hi = pos[s0], pos[s1]
        pos[rel] = bisect.bisect_left(o, a[rel], lo, hi)
    assert pos == want, (seed, pos, want)
for s in range(5000):
    sim_sample(s)

print("ok")
