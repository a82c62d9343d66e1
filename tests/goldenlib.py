"""Load committed golden fixtures (tests/golden) as Results-like objects."""
import os
import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class GoldResults:
    def __init__(self, z):
        self.hit_off = z["hit_off"]
        self.hits = z["hits"]
        self.cigar = z["cigar"]
        self.md = z["md"].tobytes()

    def read_hits(self, i):
        return self.hits[self.hit_off[i]:self.hit_off[i + 1]]

    def cigar_of(self, h):
        return self.cigar[h["cigar_off"]:h["cigar_off"] + h["n_cigar"]]

    def cigar_str(self, h):
        return "".join("%d%s" % (c >> 4, "MIDSH"[c & 0xf]) for c in self.cigar_of(h))

    def md_of(self, h):
        return self.md[h["md_off"]:h["md_off"] + h["md_len"]].decode()


def load(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    return GoldResults(z), z


def path(*p):
    return os.path.join(GOLD, *p)
