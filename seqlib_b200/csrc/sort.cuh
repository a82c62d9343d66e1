// sort.cuh -- an introsort whose element order is identical, comparison by
// comparison, to klib's ks_introsort (bwa/ksort.h:159-226) including the
// median-of-3 pivot, the "leave runs <= 16 for one final insertion sort" rule
// and the comb-sort escape (bwa/ksort.h:137-158).  BWA-MEM's tie-breaks between
// equal keys (chain weights, region ends) depend on this exact permutation, so
// the kernels sort with the same algorithm instead of a stable or bitonic sort.
#pragma once
#include "common.cuh"

namespace b200 {

template <typename T, typename Less>
HD void insertion_sort_(T *s, T *t, Less lt)   // [s, t)
{
    for (T *i = s + 1; i < t; ++i)
        for (T *j = i; j > s && lt(*j, *(j - 1)); --j) swap_(*j, *(j - 1));
}

template <typename T, typename Less>
HD void comb_sort_(size_t n, T *a, Less lt)
{
    const double shrink = 1.2473309501039786540366528676643;
    bool swapped;
    size_t gap = n;
    do {
        if (gap > 2) {
            gap = (size_t)(gap / shrink);
            if (gap == 9 || gap == 10) gap = 11;
        }
        swapped = false;
        for (T *i = a; i < a + n - gap; ++i) {
            T *j = i + gap;
            if (lt(*j, *i)) { swap_(*i, *j); swapped = true; }
        }
    } while (swapped || gap > 2);
    if (gap != 1) insertion_sort_(a, a + n, lt);
}

template <typename T, typename Less>
HD void introsort(size_t n, T *a, Less lt)
{
    struct Frame { T *left, *right; int depth; };
    if (n < 1) return;
    if (n == 2) { if (lt(a[1], a[0])) swap_(a[0], a[1]); return; }
    int d;
    for (d = 2; (1ul << d) < n; ++d) {}
    Frame stack[48];   // the larger side is pushed, the smaller continued: depth <= log2(n)
    Frame *top = stack;
    T *s = a, *t = a + (n - 1);
    d <<= 1;
    for (;;) {
        if (s < t) {
            if (--d == 0) { comb_sort_((size_t)(t - s + 1), s, lt); t = s; continue; }
            T *i = s, *j = t, *k = i + ((j - i) >> 1) + 1;
            if (lt(*k, *i)) { if (lt(*k, *j)) k = j; }
            else k = lt(*j, *i) ? i : j;
            T rp = *k;
            if (k != t) swap_(*k, *t);
            for (;;) {
                do ++i; while (lt(*i, rp));
                do --j; while (i <= j && lt(rp, *j));
                if (j <= i) break;
                swap_(*i, *j);
            }
            swap_(*i, *t);
            if (i - s > t - i) {
                if (i - s > 16) { top->left = s; top->right = i - 1; top->depth = d; ++top; }
                s = t - i > 16 ? i + 1 : t;
            } else {
                if (t - i > 16) { top->left = i + 1; top->right = t; top->depth = d; ++top; }
                t = i - s > 16 ? i - 1 : s;
            }
        } else {
            if (top == stack) { insertion_sort_(a, a + n, lt); return; }
            --top; s = top->left; t = top->right; d = top->depth;
        }
    }
}

} // namespace b200
