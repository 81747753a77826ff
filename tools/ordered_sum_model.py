#!/usr/bin/env python
"""Model (not a test) of the round-2 idea recorded in DESIGN.md 5.2: a sequential round-to-nearest-even sum of
non-negative doubles, evaluated block-wise with an integer prefix sum per binade of the accumulator.

While the accumulator stays inside one binade its ulp u is fixed and acc = A*u, A in [2^52, 2^53):
    fl(A*u + t) = (A + T + r)*u,  T = floor(t/u),  r = 0/1 as the remainder is below/above u/2, exact halves make the
    result even -- so the only state a block of terms depends on is the parity of the incoming A, and blocks compose
    (scan_sum below folds every block into a (parity -> increment) function).  A binade crossing ends the scan; the
    crossing block is added with real float64 adds and the scan restarts with the new ulp.
Checked here against numpy float64 sequential addition on random and adversarial data.  A device version
(warp-level, 32 blocks per scan) was built and is bit-exact, but per 1024-term tile it does not beat seven
sequential DADD chains (restarts after every crossing, bank conflicts of block-contiguous reads); it needs the
whole ordered cluster in one scan to pay."""
import numpy as np, struct, random
def bits(x): return struct.unpack('<Q', struct.pack('<d', x))[0]
def frombits(b): return struct.unpack('<d', struct.pack('<Q', b))[0]
def decode(x):
    b = bits(x); e = (b >> 52) & 0x7FF; m = b & ((1 << 52) - 1)
    if e == 0: return m, -1074            # denormal / zero: value = m * 2^-1074
    return m | (1 << 52), e - 1075        # value = M * 2^(e-1075)
def seq_sum(acc, terms):
    acc = np.float64(acc)
    for t in terms: acc = np.float64(acc + np.float64(t))
    return float(acc)
def term_class(t, k):
    """relative to ulp u = 2^k: returns (T, cls) cls 0 down 1 up 2 tie"""
    if t == 0.0: return 0, 0
    mt, et = decode(t)
    shift = k - et
    if shift <= 0:
        if -shift > 10: return 1 << 53, 0
        return mt << (-shift), 0
    if shift >= 54: return 0, 0
    T = mt >> shift; rem = mt & ((1 << shift) - 1); half = 1 << (shift - 1)
    return T, (1 if rem > half else 2 if rem == half else 0)
def block_fn(cls_list):
    x = [0, 1]
    for T, c in cls_list:
        for p in (0, 1):
            v = x[p] + T
            if c == 1: v += 1
            elif c == 2: v += (v & 1)
            x[p] = v
    return x[0], x[1] - 1
def scan_sum(acc, terms, L=4):
    i = 0; n = len(terms)
    while i < n and acc == 0.0:
        acc = float(np.float64(acc) + np.float64(terms[i])); i += 1
    while i < n:
        A, k = decode(acc)
        assert A >= (1 << 52), "acc must be normal here"
        # blocks of L terms
        blocks = [terms[j:j + L] for j in range(i, n, L)]
        fns = [block_fn([term_class(t, k) for t in blk]) for blk in blocks]
        cur = A; ok_blocks = 0
        for f in fns:
            nxt = cur + f[cur & 1]
            if nxt >= (1 << 53): break
            cur = nxt; ok_blocks += 1
        acc = frombits(((k + 1075) << 52) | (cur & ((1 << 52) - 1)))
        i += ok_blocks * L
        if ok_blocks < len(blocks):
            blk = terms[i:i + L]
            acc = seq_sum(acc, blk); i += len(blk)
    return acc
random.seed(1); bad = 0
for trial in range(3000):
    n = random.randint(1, 400)
    mode = trial % 5
    if mode == 0: w = [1.0 / random.randint(1, 10**7)] * n
    elif mode == 1: w = [random.randint(1, 1000) / float(random.randint(1000, 10**7)) for _ in range(n)]
    elif mode == 2: w = [2.0 ** -random.randint(1, 40) for _ in range(n)]
    elif mode == 3: w = [random.randint(1, 4) * 0.25 for _ in range(n)]
    else: w = [random.random() * 10 ** random.randint(-12, 3) for _ in range(n)]
    terms = [float(np.float64(wi) * np.float64(random.randint(0, 255) ** random.choice([1, 2]))) for wi in w]
    a0 = random.choice([0.0, 0.0, terms[0], 1.0, 0.3])
    s1 = seq_sum(a0, terms); s2 = scan_sum(a0, terms, L=random.choice([1, 3, 4, 32]))
    if bits(s1) != bits(s2):
        bad += 1
        if bad < 5: print("MISMATCH", trial, mode, n, s1, s2)
print("bad", bad)
