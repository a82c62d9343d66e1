"""Brute-force model of the BWT fml_seq2fmi builds (fermi-lite/misc.c:65-128 + mrope.c in MR_SO_RCLO order): the collection is
{s, revcomp(s)} for every read without N (an even-length reverse palindrome first loses its last base); row (string, offset i)
sorts by s[i:] $ revcomp(s[:i]) $ with $ < A < C < G < T and carries s[i-1] ($ for i = 0).  Small inputs only."""
import numpy as np

_COMP = {"A": "T", "C": "G", "G": "C", "T": "A"}
_CODE = {"$": 0, "A": 1, "C": 2, "G": 3, "T": 4}


def bwt(reads):
    strs = []
    for r in reads:
        r = (r.decode() if isinstance(r, bytes) else r).upper()
        if not r or any(c not in "ACGT" for c in r):
            continue
        n = len(r)
        if n % 2 == 0 and all(_COMP[r[i]] == r[n - 1 - i] for i in range(n // 2)):
            r = r[:-1]
        strs.append(r)
        strs.append("".join(_COMP[c] for c in reversed(r)))
    rows = []
    for j, s in enumerate(strs):
        t = strs[j ^ 1]
        n = len(s)
        for i in range(n + 1):
            key = s[i:] + "$" + t[n - i:] + "$"
            rows.append((tuple(_CODE[c] for c in key), _CODE[s[i - 1]] if i > 0 else 0))
    rows.sort(key=lambda x: x[0])
    return np.array([r[1] for r in rows], dtype=np.uint8)
