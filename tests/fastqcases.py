"""FASTA/FASTQ texts for the ingest parity tests: every behaviour of kseq_read (bwa/kseq.h:176-226) the reference's reader
exposes -- name/comment split on any isspace() character, multi-line sequences and qualities, blank lines, CR LF, quality
lines starting with '@' or '+', FASTA and FASTQ mixed, garbage before the first header, empty records, a missing final
newline, a truncated quality string (kseq_read = -2) and a missing quality line."""
import numpy as np


def strict_fastq(n=2000, seed=7, read_len=150, crlf=False, final_newline=True, comments=True):
    rng = np.random.default_rng(seed)
    nl = b"\r\n" if crlf else b"\n"
    out = []
    for i in range(n):
        ln = read_len if read_len > 0 else int(rng.integers(1, 200))
        seq = bytes(rng.choice(np.frombuffer(b"ACGTN", dtype=np.uint8), size=ln, p=[.245, .245, .245, .245, .02]))
        qual = bytes(rng.integers(33, 74, size=ln, dtype=np.uint8))          # includes '@' (64) and '+' (43) as first characters
        name = b"r%d" % i
        if comments and i % 3 == 0:
            name += (b" ", b"\t")[i % 2] + b"len=%d extra words" % ln
        out.append(b"@" + name + nl + seq + nl + b"+" + (name if i % 5 == 0 else b"") + nl + qual + nl)
    text = b"".join(out)
    if not final_newline:
        text = text[:-len(nl)]
    return text


CASES = {
    "strict": strict_fastq(300),
    "strict_ragged": strict_fastq(300, seed=8, read_len=0),
    "strict_crlf": strict_fastq(100, seed=9, crlf=True),
    "strict_no_final_newline": strict_fastq(50, seed=10, final_newline=False),
    "fasta_multiline": b">c1 first contig\nACGTACGT\nGGCC\n\nTTAA\n>c2\nNNNN\n>c3\tcomment\twith\ttabs\nAC\n",
    "fastq_multiline": b"@m1\nACGT\nACGT\n+\nIIII\nIIII\n@m2 x\nAC\nGT\n+m2\nII\nI\nI\n",
    "mixed": b">f1\nACGT\n@q1\nAC\n+\nII\n>f2\nGG\n",
    "garbage_before": b"# a header line\n\n@r1\nACGT\n+\nIIII\n",
    "blank_lines": b"@r1\n\nACGT\n\n+\nIIII\n\n\n@r2\nGG\n+\nII\n\n",
    "qual_starts_with_at": b"@r1\nACGT\n+\n@III\n@r2\nAC\n+\n+I\n",
    "empty_name": b"@\nACGT\n+\nIIII\n@ c only\nAC\n+\nII\n",
    "empty_seq": b"@r1\n\n+\n\n@r2\nAC\n+\nII\n",
    "cr_only_comment": b"@r1 \r\nAC\r\n+\r\nII\r\n",
    "single_char_cr": b"@r1\r\nA\r\n+\r\nI\r\n",
    "no_final_newline_fasta": b">c1\nACGT",
    "truncated_qual": b"@r1\nACGT\n+\nIIII\n@r2\nACGT\n+\nII\n",
    "long_qual": b"@r1\nACGT\n+\nIIIIII\n@r2\nAC\n+\nII\n",
    "missing_qual": b"@r1\nACGT\n+\nIIII\n@r2\nACGT\n+",
    "header_only": b"@r1",
    "empty": b"",
    "only_garbage": b"no records here\n",
    "vt_ff_delims": b"@n1\x0bcomment a\nAC\n+\nII\n@n2\x0ccomment b\nGT\n+\nII\n",
    "seq_line_starts_plus": b">c1\nAC\n+GT\nII\n",
}
