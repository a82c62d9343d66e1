"""Field-by-field comparison helpers shared by the parity tests."""
import numpy as np

REG_FIELDS = ["rb", "re", "qb", "qe", "rid", "score", "truesc", "sub", "alt_sc", "csub", "sub_n", "w", "seedcov",
              "secondary", "secondary_all", "seedlen0", "n_comp", "is_alt", "frac_rep", "hash"]
ALN_FIELDS = ["pos", "flag", "is_rev", "mapq", "NM", "aln_sub", "n_cigar", "md_len"]


def compare_results(got, exp, max_report=5):
    """Returns a list of human readable mismatch strings (empty = bit-identical)."""
    bad = []
    if not np.array_equal(got.hit_off, exp.hit_off):
        d = np.nonzero(np.diff(got.hit_off) != np.diff(exp.hit_off))[0]
        bad.append("hit counts differ for %d reads, first %s: got %s exp %s" % (
            len(d), d[:5], np.diff(got.hit_off)[d[:5]], np.diff(exp.hit_off)[d[:5]]))
        return bad
    for f in REG_FIELDS + ALN_FIELDS:
        ne = np.nonzero(got.hits[f] != exp.hits[f])[0]
        if len(ne):
            rd = np.searchsorted(got.hit_off, ne[:max_report], side="right") - 1
            bad.append("field %s differs in %d hits; reads %s got %s exp %s" % (
                f, len(ne), rd, got.hits[f][ne[:max_report]], exp.hits[f][ne[:max_report]]))
    if bad:
        return bad
    # cigars and MDs (offsets may differ, contents may not)
    gc = np.concatenate([got.cigar_of(h) for h in got.hits]) if len(got.hits) else np.zeros(0, np.uint32)
    ec = np.concatenate([exp.cigar_of(h) for h in exp.hits]) if len(exp.hits) else np.zeros(0, np.uint32)
    if not np.array_equal(gc, ec):
        for i in range(len(got.hits)):
            g, e = got.hits[i], exp.hits[i]
            if not np.array_equal(got.cigar_of(g), exp.cigar_of(e)):
                bad.append("cigar differs at hit %d: %s vs %s" % (i, got.cigar_str(g), exp.cigar_str(e)))
                if len(bad) >= max_report:
                    break
    for i in range(len(got.hits)):
        g, e = got.hits[i], exp.hits[i]
        if got.md_of(g) != exp.md_of(e):
            bad.append("MD differs at hit %d: %s vs %s" % (i, got.md_of(g), exp.md_of(e)))
            if len(bad) >= max_report:
                break
    return bad


def _gather(pool, offs, lens):
    """Concatenation of pool[offs[i]:offs[i]+lens[i]] without a Python loop."""
    lens = lens.astype(np.int64)
    tot = int(lens.sum())
    if tot == 0:
        return pool[:0]
    starts = np.repeat(offs.astype(np.int64) - np.concatenate([[0], np.cumsum(lens)[:-1]]), lens)
    return pool[starts + np.arange(tot, dtype=np.int64)]


def compare_prefix(got, exp, n_reads, max_report=5):
    """Vectorised bit-exactness check of the first n_reads reads of `got` against `exp` (which holds exactly n_reads reads):
    hit counts, every region / alignment field, CIGAR words and MD bytes.  Returns (n_mismatching_reads, messages)."""
    bad = []
    g_off = got.hit_off[:n_reads + 1] - got.hit_off[0]
    if not np.array_equal(g_off, exp.hit_off):
        d = np.nonzero(np.diff(g_off) != np.diff(exp.hit_off))[0]
        bad.append("hit counts differ for %d reads, first %s" % (len(d), d[:max_report]))
        return len(d), bad
    nh = int(exp.hit_off[-1])
    gh = got.hits[int(got.hit_off[0]):int(got.hit_off[0]) + nh]
    eh = exp.hits
    wrong = np.zeros(nh, dtype=bool)
    for f in REG_FIELDS + ALN_FIELDS:
        ne = gh[f] != eh[f]
        if ne.any():
            k = np.nonzero(ne)[0]
            rd = np.searchsorted(exp.hit_off, k[:max_report], side="right") - 1
            bad.append("field %s differs in %d hits; reads %s got %s exp %s" % (f, len(k), rd, gh[f][k[:max_report]], eh[f][k[:max_report]]))
            wrong |= ne
    if not wrong.any():
        gc = _gather(got.cigar, gh["cigar_off"], gh["n_cigar"])
        ec = _gather(exp.cigar, eh["cigar_off"], eh["n_cigar"])
        if not np.array_equal(gc, ec):
            per = np.repeat(np.arange(nh), eh["n_cigar"].astype(np.int64))
            k = np.unique(per[gc != ec])
            wrong[k] = True
            bad.append("cigar differs in %d hits, first %s" % (len(k), k[:max_report]))
        gm = _gather(np.frombuffer(got.md, dtype=np.uint8), gh["md_off"], gh["md_len"])
        em = _gather(np.frombuffer(exp.md, dtype=np.uint8), eh["md_off"], eh["md_len"])
        if not np.array_equal(gm, em):
            per = np.repeat(np.arange(nh), eh["md_len"].astype(np.int64))
            k = np.unique(per[gm != em])
            wrong[k] = True
            bad.append("MD differs in %d hits, first %s" % (len(k), k[:max_report]))
    rd = np.unique(np.searchsorted(exp.hit_off, np.nonzero(wrong)[0], side="right") - 1)
    return len(rd), bad


def truth_recovery(res, pos, strand, contigs, tol=8):
    """Fraction of reads whose first (primary) hit lies within tol bp of the simulated origin on the right strand."""
    has = np.diff(res.hit_off) > 0
    first = res.hits[res.hit_off[:-1][has]]
    coff = np.array([c[1] for c in contigs], dtype=np.int64)
    gpos = coff[np.clip(first["rid"], 0, len(coff) - 1)] + first["pos"]
    ok = (np.abs(gpos - pos[has]) <= tol) & (first["is_rev"] == strand[has]) & (first["rid"] >= 0)
    return float(ok.sum()) / max(1, len(pos)), float(has.mean())


def read_fastq(path, n=None):
    names, seqs, quals = [], [], []
    with open(path) as f:
        while True:
            h = f.readline()
            if not h:
                break
            s = f.readline().strip()
            f.readline()
            q = f.readline().strip()
            names.append(h[1:].strip())
            seqs.append(s)
            quals.append(q)
            if n is not None and len(seqs) >= n:
                break
    return names, seqs, quals
