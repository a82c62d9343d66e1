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
