"""The BFC stages of fml_assemble (fermi-lite/misc.c:280-290) as a driver over any backend exposing
default_opt() / correct_flat(opt, seqs, quals, off, flt_uniq): fml_opt_adjust, fml_correct, then fml_fltuniq on the corrected
reads.  Used with the reference binding (golden generation), the host emulation and the CUDA library."""
import os
import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FML_SETS = ["fml_bcr_2k", "fml_mt_2k", "fml_mixed", "fml_mixed_noqual"]


def adjusted_ec_k(tot_len, ec_k=0):
    """fml_opt_adjust (fermi-lite/misc.c:43-54)."""
    log_len = 10
    while log_len < 32 and (1 << log_len) <= tot_len:
        log_len += 1
    if ec_k == 0:
        ec_k = (log_len + 12) // 2
    if ec_k % 2 == 0:
        ec_k += 1
    return ec_k


def compact(seqs, quals, off, lens):
    """Drop the filtered reads / tails the way fml_fltuniq leaves the fseq1_t array (reads keep their slots; l_seq = 0)."""
    n = len(off) - 1
    ps = [seqs[int(off[i]):int(off[i]) + int(lens[i])] for i in range(n)]
    noff = np.zeros(n + 1, dtype=np.int64)
    noff[1:] = np.cumsum(lens)
    s = np.concatenate(ps) if n else np.zeros(0, np.uint8)
    q = None
    if quals is not None:
        q = np.concatenate([quals[int(off[i]):int(off[i]) + int(lens[i])] for i in range(n)]) if n else np.zeros(0, np.uint8)
    return s, q, noff


def pipeline(correct_flat, opt, seqs, quals, off):
    opt.ec_k = adjusted_ec_k(int(off[-1]), 0)
    r1 = correct_flat(opt, seqs, quals, off, flt_uniq=False)
    es, eq, _, ek = r1[0], r1[1], r1[2], r1[3]
    r2 = correct_flat(opt, es, eq, off, flt_uniq=True)
    fs, fq, fl, fk = r2[0], r2[1], r2[2], r2[3]
    cs, cq, coff = compact(fs, fq, off, fl)
    out = dict(ec_k=np.int32(opt.ec_k), ec_seqs=es, ec_kcov=np.float32(ek), flt_lens=fl, flt_seqs=cs, flt_kcov=np.float32(fk))
    if quals is not None:
        out["ec_quals"] = eq
        out["flt_quals"] = cq
    return out


def reference_pipeline(pyref_fml, seqs, quals, off):
    return pipeline(pyref_fml.correct_flat, pyref_fml.default_opt(), seqs, quals, off)


def load(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    quals = z["quals"] if len(z["quals"]) else None
    return z["seqs"], quals, z["off"], z


def compare(got, z):
    bad = []
    for k, v in got.items():
        e = z[k]
        if np.issubdtype(np.asarray(v).dtype, np.floating):
            if not (np.float32(v) == np.float32(e) or (np.isnan(v) and np.isnan(e))):
                bad.append("%s: %r != %r" % (k, float(v), float(e)))
        elif not np.array_equal(v, e):
            bad.append("%s differs (%d entries)" % (k, int((np.asarray(v) != np.asarray(e)).sum()) if np.shape(v) == np.shape(e) else -1))
    return bad


# ---------------------------------------------------------------------------------------------------- assembly half
def utg_text(utgs):
    """Unitigs in fml_utg_print's layout (fermi-lite/misc.c:215-246), one canonical text for comparisons."""
    out = []
    for i, u in enumerate(utgs):
        l0 = "".join("%d,%d;" % (o[2] << 1 | o[3], o[0]) for o in u["ovlp"][:u["n_ovlp"][0]]) or "."
        l1 = "".join("%d,%d;" % (o[2] << 1 | o[3], o[0]) for o in u["ovlp"][u["n_ovlp"][0]:]) or "."
        out.append("@%d:%d\t%d\t%s\t%s\n%s\n+\n%s\n" % (i << 1, i << 1 | 1, u["nsr"], l0, l1, u["seq"].decode(), u["cov"].decode()))
    return "".join(out)


def filtered_reads(z, off):
    """The reads as fml_fltuniq leaves them (input of fml_seq2fmi), from a BFC fixture."""
    fl = z["flt_lens"]
    foff = np.zeros(len(fl) + 1, dtype=np.int64)
    foff[1:] = np.cumsum(fl)
    return z["flt_seqs"], foff


def asm_opt_for(opt, tot_len, n, kcov, clean=True):
    """Options as fml_assemble hands them to fml_fmi2mag / fml_mag_clean (fermi-lite/misc.c:286-298)."""
    opt.ec_k = adjusted_ec_k(tot_len, 0)
    opt.mag_opt.min_elen = int(float(tot_len) / n * 2.5 + .499)
    if clean:
        me = opt.mag_opt.min_ensr if opt.mag_opt.min_ensr > kcov * .1 else int(kcov * .1 + .499)
        me = me if me < opt.max_cnt else opt.max_cnt
        me = me if me > opt.min_cnt else opt.min_cnt
        opt.mag_opt.min_ensr = me
        opt.mag_opt.min_insr = me - 1
    return opt


def load_asm(name):
    z = np.load(os.path.join(GOLD, name + "_asm.npz"))
    return dict(bwt_md5=str(z["bwt_md5"]), cnt=z["cnt"], mag0=z["mag0"].tobytes().decode(), mag1=z["mag1"].tobytes().decode(),
                utg=z["utg"].tobytes().decode(), rdist=float(z["rdist"]), rank_q=z["rank_q"], rank_r=z["rank_r"], rank_s=z["rank_s"])
