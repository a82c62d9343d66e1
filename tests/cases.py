"""Deterministic test inputs shared by the golden generator and the tests."""
import numpy as np
from seqlib_b200 import synth
from seqlib_b200.abi import EXT_JOB_DTYPE

KAT_NAMES = ["ref3", "ref4", "ref5", "ref6"]
KAT_SEQS = ["ACATGGCGAGCACTTCTAGCATCAGCTAGCTACGATCGATCGATCGATCGTAGC",
            "CTACTTTATCATCTACACACTGCCTGACTGCGGCGACGAGCGAGCAGCTACTATCGACT",
            "CGATCGTAGCTAGCTGATGCTAGAAGTGCTCGCCATGT",
            "TATCTACTGCGCGCGATCATCTAGCGCAGGACGAGCATC" + "N" * 100 + "CGATCGTTATTATCGAGCGACGATCTACTACGT"]
KAT_QUERIES = ["ACATGGCGAGCACTTCTAGCATCAGCTAGCTACGATCG", "CGATCGTAGCTAGCTGATGCTAGAAGTGCTCGC"]
KAT_SRAND = 0   # srand48(0) before ConstructIndex so the N -> lrand48()&3 draws are reproducible in one process


def ids_for(n):
    return np.arange(n, dtype=np.int64) * 7919 + 13


def c1_reference():
    l_pac = 10000
    pac = synth.reference(l_pac)
    ctg = synth.contigs_for(l_pac, 1, "ref10k")
    return pac, ctg, synth.ascii_of(pac, 0, l_pac)


def c1_reads(pac, ctg, n=1000):
    a, off, _, _ = synth.reads(pac, 10000, ctg, n, 150, 0.01, 5e-4)
    return a, off


def read_lines(path):
    with open(path) as f:
        return [l.strip() for l in f if l.strip()]


def c3_tuples(n, seed=0x5EED0003):
    """Config-3 shaped ksw_extend2 jobs: target random 300, query = target[0:150] with 2 % subs + 0.2 % indels, h0 in [19,150]."""
    rng = np.random.Generator(np.random.PCG64(seed))
    tp = rng.integers(0, 4, size=n * 300, dtype=np.uint8)
    qp = np.zeros(n * 150, dtype=np.uint8)
    jobs = np.zeros(n, dtype=EXT_JOB_DTYPE)
    for i in range(n):
        t = tp[i * 300:(i + 1) * 300]
        q = []
        j = 0
        while len(q) < 150 and j < 300:
            x = rng.random()
            if x < 0.001:
                j += 1 + int(rng.integers(0, 3))
                continue
            if x < 0.002:
                q.extend(rng.integers(0, 4, size=1 + int(rng.integers(0, 3))).tolist())
                continue
            b = int(t[j])
            if rng.random() < 0.02:
                b = (b + 1 + int(rng.integers(0, 3))) & 3
            q.append(b)
            j += 1
        q = (q + [0] * 150)[:150]
        qp[i * 150:(i + 1) * 150] = q
        jobs[i] = (150, 300, i * 150, i * 300, 100, 5, 100, int(rng.integers(19, 151)))
    return jobs, qp, tp


def c3_tuples_fast(n, seed=0x5EED0003):
    """Vectorised variant for the 1M-pair microbenchmark: substitutions only in the numpy path plus sparse indels."""
    rng = np.random.Generator(np.random.PCG64(seed))
    tp = rng.integers(0, 4, size=(n, 300), dtype=np.uint8)
    q = tp[:, :150].copy()
    sub = rng.random((n, 150)) < 0.02
    q[sub] = (q[sub] + rng.integers(1, 4, size=int(sub.sum()), dtype=np.uint8)) & 3
    # one indel in ~26 % of the pairs (0.2 % per base): shift the tail by one base
    has = rng.random(n) < 0.26
    posn = rng.integers(10, 140, size=n)
    idx = np.nonzero(has)[0]
    for i in idx[: min(len(idx), 400000)]:
        p = posn[i]
        if i & 1:
            q[i, p + 1:] = q[i, p:-1].copy()      # insertion in the query
        else:
            q[i, p:-1] = q[i, p + 1:].copy()      # deletion from the query
    jobs = np.zeros(n, dtype=EXT_JOB_DTYPE)
    jobs["qlen"] = 150
    jobs["tlen"] = 300
    jobs["q_off"] = np.arange(n, dtype=np.int64) * 150
    jobs["t_off"] = np.arange(n, dtype=np.int64) * 300
    jobs["w"] = 100
    jobs["end_bonus"] = 5
    jobs["zdrop"] = 100
    jobs["h0"] = rng.integers(19, 151, size=n)
    return jobs, q.reshape(-1), tp.reshape(-1)


def mixed_tuples(n, seed=11):
    """Random ksw_extend2 jobs with varied shapes: qlen 1..139, related / unrelated targets, Ns, w in {5,30,100,200}, zdrop in {0,20,100}."""
    rng = np.random.default_rng(seed)
    ql = rng.integers(1, 140, n)
    tl = np.minimum(ql + rng.integers(0, 200, n), 400)
    qo = np.concatenate([[0], np.cumsum(ql)[:-1]])
    to = np.concatenate([[0], np.cumsum(tl)[:-1]])
    tp = rng.integers(0, 4, int(tl.sum()), dtype=np.uint8)
    qp = np.zeros(int(ql.sum()), dtype=np.uint8)
    for i in range(n):
        t = tp[to[i]:to[i] + tl[i]]
        q = t[:ql[i]].copy() if ql[i] <= tl[i] else np.resize(t, ql[i])
        if i % 3 == 0:
            q = rng.integers(0, 4, ql[i], dtype=np.uint8)
        else:
            m = rng.random(ql[i]) < 0.05
            q[m] = (q[m] + 1) & 3
            if i % 5 == 0 and ql[i] > 20:
                p = int(rng.integers(5, ql[i] - 5))
                q[p:] = np.roll(q[p:], int(rng.integers(1, 6)))
            if i % 7 == 0:
                q[rng.integers(0, ql[i])] = 4
        qp[qo[i]:qo[i] + ql[i]] = q
    jobs = np.zeros(n, dtype=EXT_JOB_DTYPE)
    jobs["qlen"] = ql
    jobs["tlen"] = tl
    jobs["q_off"] = qo
    jobs["t_off"] = to
    jobs["w"] = rng.choice([100, 200, 5, 30], n)
    jobs["end_bonus"] = 5
    jobs["zdrop"] = rng.choice([100, 0, 20], n)
    jobs["h0"] = rng.integers(1, 160, n)
    return jobs, qp, tp


# ---------------------------------------------------------------------------------------------------- fermi-lite half
def fml_reads(n, region=20000, read_len=150, err=0.01, seed=0x5EED0004, junk=0.03, n_frac=0.01, short_frac=0.01, ragged=True):
    """Assembly-shaped reads (SURVEY 8d, config 4 at reduced scale): n reads drawn from a random `region`-bp sequence, both
    strands, `err` substitutions; erroneous bases mostly get a low quality; plus the edge cases the reference code
    branches on: unrelated reads (no solid k-mer; dropped by fml_fltuniq), reads with N (skipped k-mers, ECCODE_MANY_N when
    > 5 %), reads shorter than k, ragged lengths.  Returns (seqs u8 pool, quals u8 pool, off)."""
    rng = np.random.default_rng(seed)
    ref = rng.integers(0, 4, region, dtype=np.uint8)
    comp = np.array([3, 2, 1, 0], dtype=np.uint8)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    seqs, quals = [], []
    for i in range(n):
        ln = read_len
        u = rng.random()
        if ragged and u < 0.15:
            ln = int(rng.integers(read_len // 2, read_len + 1))
        if u > 1 - short_frac:
            ln = int(rng.integers(1, 40))
        if rng.random() < junk:
            s = rng.integers(0, 4, ln, dtype=np.uint8)
            q = np.full(ln, ord("I"), dtype=np.uint8)
        else:
            p = int(rng.integers(0, region - ln + 1))
            s = ref[p:p + ln].copy()
            if rng.random() < 0.5:
                s = comp[s[::-1]]
            q = rng.choice(np.frombuffer(b"II??5", dtype=np.uint8), ln)
            e = rng.random(ln) < err
            s[e] = (s[e] + rng.integers(1, 4, int(e.sum()), dtype=np.uint8)) & 3
            lowq = e & (rng.random(ln) < 0.7)
            q[lowq] = ord("(")
        a = acgt[s]
        if rng.random() < n_frac:
            k = 1 if rng.random() < 0.6 else max(1, ln // 8)
            a[rng.integers(0, ln, k)] = ord("N")
        if rng.random() < 0.02:
            a = np.frombuffer(a.tobytes().lower(), dtype=np.uint8).copy()
        seqs.append(a)
        quals.append(q)
    off = np.zeros(n + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(s) for s in seqs])
    return np.concatenate(seqs), np.concatenate(quals), off


def long_reads(pac, l_pac, n=40, seed=0x5EED0007, lo=800, hi=4000):
    """Contig-like queries (longer than the ~730 bp where mem_flt_chained_seeds starts to act, bwa/bwamem.c:624-628):
    segments of the reference with 1-3 % substitutions, a few indels, some chimeras of two segments, some reverse strand,
    one stretch of unrelated sequence; returns list[str]."""
    rng = np.random.default_rng(seed)
    idx = np.arange(l_pac)
    ref = (pac[idx >> 2] >> ((~idx & 3) << 1)) & 3
    comp = np.array([3, 2, 1, 0], dtype=np.uint8)

    def seg(ln):
        p = int(rng.integers(0, l_pac - ln))
        s = ref[p:p + ln].astype(np.uint8).copy()
        e = rng.random(ln) < rng.choice([0.01, 0.03])
        s[e] = (s[e] + rng.integers(1, 4, int(e.sum()), dtype=np.uint8)) & 3
        out = []
        i = 0
        while i < ln:
            u = rng.random()
            if u < 0.002:
                i += int(rng.integers(1, 12))          # deletion
                continue
            if u < 0.004:
                out.extend(rng.integers(0, 4, int(rng.integers(1, 12))).tolist())   # insertion
            out.append(int(s[i]))
            i += 1
        s = np.array(out, dtype=np.uint8)
        if rng.random() < 0.5:
            s = comp[s[::-1]]
        return s

    reads = []
    for i in range(n):
        ln = int(rng.integers(lo, hi))
        if i % 5 == 4:
            s = np.concatenate([seg(ln // 2), seg(ln - ln // 2)])
        elif i % 11 == 10:
            s = np.concatenate([seg(ln // 2), rng.integers(0, 4, 300, dtype=np.uint8), seg(ln // 3)])
        else:
            s = seg(ln)
        reads.append("".join("ACGT"[b] for b in s))
    return reads


def fml_diploid_reads(n, region, seed, snp=0.004, indel=0.001, err=0.005):
    """Reads from two haplotypes (SNPs + small deletions between them): the unitig graph has real bubbles, which is what
    mag_g_pop_simple / mag_g_simplify_bubble / the AGGRESSIVE flag act on.  Returns (seqs, quals, off), quality 'I'."""
    rng = np.random.default_rng(seed)
    h1 = rng.integers(0, 4, region, dtype=np.uint8)
    h2 = h1.copy()
    m = rng.random(region) < snp
    h2[m] = (h2[m] + rng.integers(1, 4, int(m.sum()), dtype=np.uint8)) & 3
    h2 = h2[rng.random(region) >= indel]
    comp = np.array([3, 2, 1, 0], dtype=np.uint8)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    seqs = []
    for _ in range(n):
        h = h1 if rng.random() < 0.5 else h2
        p = int(rng.integers(0, len(h) - 150))
        s = h[p:p + 150].copy()
        e = rng.random(150) < err
        s[e] = (s[e] + rng.integers(1, 4, int(e.sum()), dtype=np.uint8)) & 3
        if rng.random() < 0.5:
            s = comp[s[::-1]]
        seqs.append(acgt[s])
    off = np.arange(n + 1, dtype=np.int64) * 150
    return np.concatenate(seqs), np.full(n * 150, ord("I"), dtype=np.uint8), off
