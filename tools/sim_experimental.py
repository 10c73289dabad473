"""CPU models of the index arithmetic behind two experimental k_merge items (csrc/compact.cu, EXP instantiation):
the chunk ownership of the staged-heads writer (every entry byte written exactly once, nothing outside the entries) and
the sample-then-refine rank search (positions equal a plain lower bound) and the shuffle-scan key rebuild (keys,
zero padding and straddling trailers equal a plain decode).  Restatements for review, not the CUDA code."""
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

M32 = 0xffffffff
def sim_d2(seed):
    rnd = random.Random(seed)
    nrec = rnd.randint(1, 45)
    KS = 64
    # internal keys = user key + 8-byte trailer; build sorted-ish keys with shared prefixes
    keys = []
    base = bytes(rnd.randrange(256) for _ in range(rnd.randint(0, 30)))
    prev = None
    for i in range(nrec):
        if prev is not None and rnd.random() < 0.15:
            uk = prev[:-8]  # same user key, different trailer (another version)
        else:
            uk = base + bytes(rnd.randrange(256) for _ in range(rnd.randint(0, 20)))
        tr_same = prev is not None and uk == prev[:-8]
        tr = (prev[-8:-3] if tr_same and rnd.random() < 0.7 else bytes(rnd.randrange(256) for _ in range(5))) + bytes(rnd.randrange(256) for _ in range(3))
        tr = tr[:8]
        keys.append(uk + tr)
        prev = keys[-1]
    # encode: restart every 16 (shared = 0), else LCP with previous
    IN = bytearray(b"\xAA" * 3)
    meta = []
    for i, k in enumerate(keys):
        sh = 0
        if i % 16 != 0:
            p = keys[i - 1]
            while sh < min(len(p), len(k)) and p[sh] == k[sh]: sh += 1
            sh = rnd.randint(0, sh)  # any shorter prefix is a legal encoding too
        ns = len(k) - sh
        ko = len(IN)
        IN += k[sh:] + bytes(rnd.randrange(256) for _ in range(rnd.randint(0, 9)))  # value bytes follow
        meta.append((sh, ns, len(k) - 8, ko, 1 if ns < 8 else 0))
    IN += b"\x55" * 16
    def ld32(off):
        a = off & ~3
        w0 = int.from_bytes(IN[a:a+4], 'little'); w1 = int.from_bytes(IN[a+4:a+8], 'little')
        s = (off & 3) * 8
        return ((w0 | (w1 << 32)) >> s) & M32
    arena = [bytearray(KS) for _ in range(nrec)]
    trailer = [0] * nrec
    for i in range(nrec):
        if meta[i][4] == 0:
            trailer[i] = int.from_bytes(keys[i][-8:], 'little')
    maxk = max(m[2] + 8 for m in meta)
    for p0 in range(0, maxk, 4):
        carry = 0
        for seg in range(0, nrec, 16):
            lanes = []
            for hl in range(16):
                i = seg + hl
                if i < nrec:
                    sh, ns, ulen, ko, fl = meta[i]
                    a = max(sh, p0); b = min(sh + ns, p0 + 4)
                    msk = d = 0
                    if a < b:
                        x = ld32(ko + (a - sh))
                        s0 = 8 * (a - p0); s1 = 8 * (p0 + 4 - b)
                        msk = ((M32 << s0) & M32) & (M32 >> s1)
                        d = ((x << s0) & M32) & msk
                    lanes.append([msk, d])
                else:
                    lanes.append([0, 0])
            dl = 1
            while dl < 16:
                new = [l[:] for l in lanes]
                for hl in range(16):
                    if hl >= dl:
                        pm, pd = lanes[hl - dl]
                        msk, d = lanes[hl]
                        new[hl] = [msk | pm, (pd & ~msk & M32) | d]
                lanes = new
                dl <<= 1
            words = [((carry & ~l[0]) & M32) | l[1] for l in lanes]
            carry = words[15]
            for hl in range(16):
                i = seg + hl
                if i >= nrec: continue
                sh, ns, ulen, ko, fl = meta[i]
                word = words[hl]
                pad = (ulen + 7) & ~7
                if p0 < pad:
                    keep = ulen - p0 if ulen > p0 else 0
                    v = word if keep >= 4 else word & ((1 << (8 * keep)) - 1)
                    arena[i][p0:p0+4] = v.to_bytes(4, 'little')
                if fl and p0 < ulen + 8 and p0 + 4 > ulen:
                    c = (word << (8 * (p0 - ulen))) & 0xffffffffffffffff if p0 >= ulen else word >> (8 * (ulen - p0))
                    trailer[i] |= c
    for i, k in enumerate(keys):
        ulen = len(k) - 8
        pad = (ulen + 7) & ~7
        assert bytes(arena[i][:ulen]) == k[:ulen], (seed, i, 'key')
        assert all(x == 0 for x in arena[i][ulen:pad]), (seed, i, 'pad')
        assert trailer[i] == int.from_bytes(k[-8:], 'little'), (seed, i, 'trailer', hex(trailer[i]), k[-8:].hex(), meta[i])
for s in range(2000):
    sim_d2(s)
print("ok")
