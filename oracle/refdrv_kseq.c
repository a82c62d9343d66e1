/* refdrv_kseq.c -- TEST INFRASTRUCTURE ONLY.
 * Driver around the reference's own FASTA/FASTQ parser: bwa/kseq.h from the mount, instantiated over gzread exactly as
 * SeqLib does (SeqLib/FastqReader.h:11-14 KSEQ_DECLARE(gzFile); bwa/bwa.c KSEQ_INIT2(, gzFile, err_gzread)), driven the way
 * FastqReader::GetNextSequence drives it (src/FastqReader.cpp:37-59): records until kseq_read() < 0.
 * Output: the four fields of every record as newline-free flat buffers + offsets, and kseq_read's final return value. */
#include <zlib.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include "kseq.h"
KSEQ_INIT(gzFile, gzread)

typedef struct { char *p; int64_t n, m; } buf_t;
static void put(buf_t *b, const char *s, int64_t l)
{
    if (b->n + l + 1 > b->m) { b->m = (b->n + l + 1) * 2 + 64; b->p = (char *)realloc(b->p, b->m); }
    if (l) memcpy(b->p + b->n, s, l);
    b->n += l;
}

/* returns the number of records; *last = the kseq_read value that ended the loop (-1 end, -2 truncated quality);
 * has[r] bit 0: comment.s != NULL, bit 1: qual.s != NULL after record r (what FastqReader.cpp:49-56 tests) */
int64_t refdrv_kseq_parse(const char *path, int64_t max_rec, char **fields /* 4: name, comment, seq, qual */, int64_t **offs /* 4 */,
                          int32_t **has, int *last)
{
    gzFile fp = gzopen(path, "r");
    if (!fp) return -1;
    kseq_t *ks = kseq_init(fp);
    buf_t b[4]; memset(b, 0, sizeof(b));
    int64_t n = 0, cap = 1024, *o[4];
    int32_t *h = (int32_t *)malloc(cap * sizeof(int32_t));
    for (int f = 0; f < 4; ++f) { o[f] = (int64_t *)malloc((cap + 1) * sizeof(int64_t)); o[f][0] = 0; }
    int r;
    while (n < max_rec && (r = kseq_read(ks)) >= 0) {
        if (n + 1 >= cap) {
            cap *= 2;
            h = (int32_t *)realloc(h, cap * sizeof(int32_t));
            for (int f = 0; f < 4; ++f) o[f] = (int64_t *)realloc(o[f], (cap + 1) * sizeof(int64_t));
        }
        put(&b[0], ks->name.s, ks->name.l);
        put(&b[1], ks->comment.s, ks->comment.l);
        put(&b[2], ks->seq.s, ks->seq.l);
        put(&b[3], ks->qual.s, ks->qual.l);
        for (int f = 0; f < 4; ++f) o[f][n + 1] = b[f].n;
        h[n] = (ks->comment.s ? 1 : 0) | (ks->qual.s ? 2 : 0);
        ++n;
    }
    *last = n < max_rec ? r : 0;
    for (int f = 0; f < 4; ++f) { fields[f] = b[f].p ? b[f].p : (char *)calloc(1, 1); offs[f] = o[f]; }
    *has = h;
    kseq_destroy(ks);
    gzclose(fp);
    return n;
}
void refdrv_kseq_free(void *p) { free(p); }
