/* synth.c -- deterministic synthetic references and reads:
 * splitmix64 streams, uniform i.i.d. bases, 150-bp reads with uniform start,
 * 50 % reverse strand, i.i.d. substitutions and optional short indels.
 * Shared by the tests, bench.py and the CPU baseline so every arm sees the
 * same inputs.  Plain C, multi-threaded with pthreads (pure data generation). */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

static inline uint64_t splitmix64(uint64_t *s)
{
	uint64_t z = (*s += 0x9e3779b97f4a7c15ull);
	z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
	z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
	return z ^ (z >> 31);
}

/* forward pac, 4 bases per byte, base i in bits ((~i)&3)*2 of byte i>>2 (bwa order); 32 bases per draw */
typedef struct { uint64_t seed; int64_t l_pac; uint8_t *pac; int64_t b, e; } refjob_t;

static void *ref_worker(void *d)
{
	refjob_t *j = d;
	int64_t w;
	for (w = j->b; w < j->e; ++w) {               /* word w covers bases [32w, 32w+32) */
		uint64_t s = j->seed + (uint64_t)w * 0x9e3779b97f4a7c15ull, r;
		int k;
		r = splitmix64(&s);
		for (k = 0; k < 8; ++k) {
			int64_t byte = w * 8 + k;
			if (byte * 4 >= j->l_pac) break;
			uint8_t v = 0; int t;
			for (t = 0; t < 4; ++t) {
				int64_t i = byte * 4 + t;
				if (i >= j->l_pac) break;
				v |= (uint8_t)(((r >> (2 * (k * 4 + t))) & 3) << ((~i & 3) << 1));
			}
			j->pac[byte] = v;
		}
	}
	return 0;
}

void synth_reference(uint64_t seed, int64_t l_pac, uint8_t *pac, int n_threads)
{
	int64_t n_words = (l_pac + 31) / 32, per;
	int t;
	if (n_threads < 1) n_threads = 1;
	per = (n_words + n_threads - 1) / n_threads;
	pthread_t *th = calloc(n_threads, sizeof(pthread_t));
	refjob_t *jobs = calloc(n_threads, sizeof(refjob_t));
	memset(pac, 0, l_pac / 4 + 1);
	for (t = 0; t < n_threads; ++t) {
		jobs[t].seed = seed; jobs[t].l_pac = l_pac; jobs[t].pac = pac;
		jobs[t].b = t * per; jobs[t].e = (t + 1) * per < n_words ? (t + 1) * per : n_words;
		pthread_create(&th[t], 0, ref_worker, &jobs[t]);
	}
	for (t = 0; t < n_threads; ++t) pthread_join(th[t], 0);
	free(th); free(jobs);
}

void synth_pac_to_ascii(const uint8_t *pac, int64_t beg, int64_t end, char *out)
{
	int64_t i;
	for (i = beg; i < end; ++i) out[i - beg] = "ACGT"[pac[i >> 2] >> ((~i & 3) << 1) & 3];
}

typedef struct {
	uint64_t seed; const uint8_t *pac; int64_t l_pac; const int64_t *coff; int n_contigs;
	int64_t n, b, e; int len; double sub, indel; char *out; int64_t *pos; int8_t *strand; int32_t *lens;
} readjob_t;

static inline double u01(uint64_t *s) { return (splitmix64(s) >> 11) * (1.0 / 9007199254740992.0); }

static void *read_worker(void *d)
{
	readjob_t *j = d;
	int64_t r;
	char *tmp = malloc(j->len * 2 + 64);
	for (r = j->b; r < j->e; ++r) {
		uint64_t s = j->seed ^ ((uint64_t)r * 0xd1342543de82ef95ull + 0x632be59bd9b4e019ull);
		int64_t p; int c, i, l = 0, src_len = j->len + 16;
		for (;;) {                               /* uniform start inside one contig */
			p = (int64_t)(u01(&s) * (double)(j->l_pac - src_len));
			if (p < 0) p = 0;
			for (c = 0; c < j->n_contigs; ++c) if (p >= j->coff[c] && p < j->coff[c + 1]) break;
			if (c < j->n_contigs && p + src_len <= j->coff[c + 1]) break;
		}
		int rev = u01(&s) < 0.5;
		int64_t q = p;
		while (l < j->len) {
			int base = j->pac[q >> 2] >> ((~q & 3) << 1) & 3;
			if (j->indel > 0) {
				double x = u01(&s);
				if (x < j->indel) {              /* deletion from the read: skip reference bases (geometric, p = .5) */
					do { ++q; } while (u01(&s) < 0.5 && q < p + src_len - 1);
					continue;
				} else if (x < 2 * j->indel) {   /* insertion into the read */
					do { tmp[l++] = (char)(splitmix64(&s) & 3); } while (u01(&s) < 0.5 && l < j->len);
					continue;
				}
			}
			if (u01(&s) < j->sub) base = (base + 1 + (int)(splitmix64(&s) % 3)) & 3;
			tmp[l++] = (char)base; ++q;
			if (q >= p + src_len) break;
		}
		while (l < j->len) tmp[l++] = 0;
		char *o = j->out + r * (int64_t)j->len;
		if (!rev) for (i = 0; i < j->len; ++i) o[i] = "ACGT"[(int)tmp[i]];
		else for (i = 0; i < j->len; ++i) o[i] = "TGCA"[(int)tmp[j->len - 1 - i]];
		if (j->pos) j->pos[r] = p;
		if (j->strand) j->strand[r] = (int8_t)rev;
	}
	free(tmp);
	return 0;
}

/* out: n * len ASCII bytes (reads back to back); coff: n_contigs+1 contig offsets */
void synth_reads(uint64_t seed, const uint8_t *pac, int64_t l_pac, const int64_t *coff, int n_contigs, int64_t n, int len,
                 double sub, double indel, char *out, int64_t *pos, int8_t *strand, int n_threads)
{
	int t;
	if (n_threads < 1) n_threads = 1;
	int64_t per = (n + n_threads - 1) / n_threads;
	pthread_t *th = calloc(n_threads, sizeof(pthread_t));
	readjob_t *jobs = calloc(n_threads, sizeof(readjob_t));
	for (t = 0; t < n_threads; ++t) {
		readjob_t *j = &jobs[t];
		j->seed = seed; j->pac = pac; j->l_pac = l_pac; j->coff = coff; j->n_contigs = n_contigs; j->n = n; j->len = len;
		j->sub = sub; j->indel = indel; j->out = out; j->pos = pos; j->strand = strand;
		j->b = t * per < n ? t * per : n; j->e = (t + 1) * per < n ? (t + 1) * per : n;
		pthread_create(&th[t], 0, read_worker, j);
	}
	for (t = 0; t < n_threads; ++t) pthread_join(th[t], 0);
	free(th); free(jobs);
}
