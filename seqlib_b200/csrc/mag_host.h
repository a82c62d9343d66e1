// mag_host.h -- the unitig graph ("mag") and its cleaning passes, host C++.
//
//   Mag::build_hash / amend / cal_rdist          <- mag_g_build_hash, mag_g_amend, mag_cal_rdist (fermi-lite/mag.c:89-145,492-533)
//   Mag::merge / rm_vext / rm_vint / rm_edge     <- mag_g_merge, mag_g_rm_vext, mag_g_rm_vint, mag_g_rm_edge (mag.c:310-483)
//   Mag::pop_simple / pop_open                   <- mag_g_pop_simple, mag_g_pop_open (fermi-lite/bubble.c:179-366)
//   Mag::clean / trim_open / to_utg              <- mag_g_clean, mag_g_trim_open (mag.c:559-620), fml_mag2utg (misc.c:153-213)
//
// These passes are sequential and vertex-order dependent in the reference (each deletion / merge changes what the next
// vertex sees) and touch 10^2..10^5 vertices, < 3 % of the assembly time: they stay on the host, written
// against std containers; the order of every loop, the unstable klib introsort (sort.cuh reproduces it comparison for
// comparison) and the float/double mix of every threshold follow the reference so the surviving unitigs are identical.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <climits>
#include <string>
#include <vector>
#include <unordered_map>
#include <stdexcept>
#include "common.cuh"
#include "sort.cuh"
#include "../../include/seqlib_b200.h"

namespace b200 {

struct MagEdge { u64 x, y; };                     // ku128_t: x = neighbour's end id, y = overlap length

inline void edge_mark_del(MagEdge &e) { e.x = (u64)-2; e.y = 0; }
inline bool edge_is_del(const MagEdge &e) { return e.x == (u64)-2 || e.y == 0; }

struct MagV {                                     // magv_t (fermi-lite/mag.h:16-24); seq holds symbols 1..4
    int len = -1, nsr = 0;
    u64 k[2] = {0, 0};
    std::vector<MagEdge> nei[2];
    std::string seq, cov;
};

struct MagUtgOvlp { u32 len; int from; u32 id; int to; };
struct MagUtg { std::string seq, cov; int nsr; std::vector<MagUtgOvlp> ovlp; int n_ovlp[2]; };

struct Mag {
    std::vector<MagV> v;
    float rdist = 0;
    int min_ovlp = 0;
    std::unordered_map<u64, u64> h;               // end id -> vertex << 1 | side   ((u64)-1: duplicated end)

    // ------------------------------------------------------------------ small vector helpers (mag.c:37-83)
    static void v128_clean(std::vector<MagEdge> &r)
    {
        size_t j = 0;
        for (size_t i = 0; i < r.size(); ++i) if (!edge_is_del(r[i])) r[j++] = r[i];
        r.resize(j);
    }
    struct XLess { bool operator()(const MagEdge &a, const MagEdge &b) const { return a.x < b.x || (a.x == b.x && a.y > b.y); } };
    static void v128_rmdup(std::vector<MagEdge> &r)
    {
        size_t l; int cnt = 0;
        if (r.size() > 1) introsort(r.size(), r.data(), XLess());
        for (l = 0; l < r.size(); ++l) { if (edge_is_del(r[l])) ++cnt; else break; }
        if (l == r.size()) { r.clear(); return; }
        u64 x = r[l].x;
        for (++l; l < r.size(); ++l) {
            if (edge_is_del(r[l]) || r[l].x == x) { edge_mark_del(r[l]); ++cnt; }
            else x = r[l].x;
        }
        if (cnt) v128_clean(r);
    }

    // ------------------------------------------------------------------ hash (mag.c:89-119)
    void build_hash()
    {
        h.clear();
        h.reserve(v.size() * 2 + 16);
        for (size_t i = 0; i < v.size(); ++i)
            for (int j = 0; j < 2; ++j) {
                auto it = h.find(v[i].k[j]);
                if (it != h.end()) it->second = (u64)-1;
                else h.emplace(v[i].k[j], (u64)i << 1 | (u64)j);
            }
    }
    u64 tid2idd(u64 tid) const
    {
        auto it = h.find(tid);
        if (it == h.end()) throw std::logic_error("mag: dangling end id");       // the reference asserts
        return it->second;
    }

    void amend()                                   // mag_g_amend
    {
        for (size_t i = 0; i < v.size(); ++i) {
            MagV &p = v[i];
            for (int j = 0; j < 2; ++j) {
                for (size_t l = 0; l < p.nei[j].size(); ++l) {
                    u64 x = p.nei[j][l].x;
                    auto it = h.find(x);
                    if (it == h.end()) { edge_mark_del(p.nei[j][l]); continue; }
                    u64 z = it->second;
                    if (z == (u64)-1) throw std::logic_error("mag: duplicated end id");   // the reference would index out of range
                    const std::vector<MagEdge> &r = v[z >> 1].nei[z & 1];
                    size_t ll; u64 me = p.k[j];
                    for (ll = 0; ll < r.size(); ++ll) if (r[ll].x == me) break;
                    if (ll == r.size()) edge_mark_del(p.nei[j][l]);
                }
                v128_rmdup(p.nei[j]);
            }
        }
    }

    double cal_rdist() const                       // mag_cal_rdist
    {
        const double A_THRES = 20.;
        std::vector<u64> srt(v.size());
        i64 sum_n_all = 0, sum_n = 0, sum_l = 0;
        double rd = -1.;
        for (size_t i = 0; i < v.size(); ++i) { srt[i] = (u64)(uint32_t)v[i].nsr << 32 | i; sum_n_all += v[i].nsr; }
        struct U64Less { bool operator()(u64 a, u64 b) const { return a < b; } };
        introsort(srt.size(), srt.data(), U64Less());
        for (int j = 0; j < 2; ++j) {
            sum_n = sum_l = 0;
            for (i64 i = (i64)v.size() - 1; i >= 0; --i) {
                const MagV &p = v[srt[i] << 32 >> 32];
                int tmp1 = 0, tmp2 = 0;
                if (!p.nei[0].empty()) { ++tmp1; tmp2 += (int)p.nei[0][0].y; }
                if (!p.nei[1].empty()) { ++tmp1; tmp2 += (int)p.nei[1][0].y; }
                if (tmp1) tmp2 /= tmp1;
                if (rd > 0.) {
                    double A = (p.len - tmp1) / rd - p.nsr * M_LN2;
                    if (A < A_THRES) continue;
                }
                sum_n += p.nsr;
                sum_l += p.len - tmp1;
                if (sum_n >= sum_n_all * 0.5) break;
            }
            rd = (double)sum_l / sum_n;
        }
        return rd;
    }

    // ------------------------------------------------------------------ basic operations (mag.c:225-308)
    void eh_add(u64 u, u64 w, int ovlp)
    {
        if ((i64)u < 0) return;
        u64 idd = tid2idd(u);
        std::vector<MagEdge> &r = v[idd >> 1].nei[idd & 1];
        for (size_t i = 0; i < r.size(); ++i) if (r[i].x == w) return;
        r.push_back(MagEdge{w, (u64)(i64)ovlp});
    }
    void eh_markdel(u64 u, u64 w)
    {
        if ((i64)u < 0) return;
        u64 idd = tid2idd(u);
        std::vector<MagEdge> &r = v[idd >> 1].nei[idd & 1];
        for (size_t i = 0; i < r.size(); ++i) if (r[i].x == w) edge_mark_del(r[i]);
    }
    static void v_destroy(MagV &p) { p.nei[0].clear(); p.nei[1].clear(); p.seq.clear(); p.cov.clear(); p.nsr = 0; p.k[0] = p.k[1] = 0; p.len = -1; }
    void v_del(MagV &p)
    {
        if (p.len < 0) return;
        for (int i = 0; i < 2; ++i) {
            std::vector<MagEdge> &r = p.nei[i];
            for (size_t j = 0; j < r.size(); ++j)
                if (!edge_is_del(r[j]) && r[j].x != p.k[0] && r[j].x != p.k[1]) eh_markdel(r[j].x, p.k[i]);
        }
        for (int i = 0; i < 2; ++i) h.erase(p.k[i]);
        v_destroy(p);
    }
    void v_transdel(MagV &p, int min_ovlp_)
    {
        if (!p.nei[0].empty() && !p.nei[1].empty()) {
            for (size_t i = 0; i < p.nei[0].size(); ++i) {
                if (edge_is_del(p.nei[0][i]) || p.nei[0][i].x == p.k[0] || p.nei[0][i].x == p.k[1]) continue;
                for (size_t j = 0; j < p.nei[1].size(); ++j) {
                    if (edge_is_del(p.nei[1][j]) || p.nei[1][j].x == p.k[0] || p.nei[1][j].x == p.k[1]) continue;
                    int ovlp = (int)(p.nei[0][i].y + p.nei[1][j].y) - p.len;
                    if (ovlp >= min_ovlp_) {
                        // eh_add may grow a neighbour list of ANOTHER vertex only (self loops are skipped above)
                        u64 a = p.nei[0][i].x, b = p.nei[1][j].x;
                        eh_add(a, b, ovlp);
                        eh_add(b, a, ovlp);
                    }
                }
            }
        }
        v_del(p);
    }
    static void revcomp6(std::string &s)
    {
        size_t l = s.size();
        for (size_t i = 0; i < l >> 1; ++i) {
            int t = s[l - 1 - i];
            t = (t >= 1 && t <= 4) ? 5 - t : t;
            int u = s[i];
            s[l - 1 - i] = (char)((u >= 1 && u <= 4) ? 5 - u : u);
            s[i] = (char)t;
        }
        if (l & 1) { int u = s[l >> 1]; s[l >> 1] = (char)((u >= 1 && u <= 4) ? 5 - u : u); }
    }
    static void reverse(std::string &s) { for (size_t i = 0, l = s.size(); i < l >> 1; ++i) std::swap(s[i], s[l - 1 - i]); }
    void v_flip(MagV &p)
    {
        revcomp6(p.seq); reverse(p.cov);
        std::swap(p.k[0], p.k[1]);
        p.nei[0].swap(p.nei[1]);
        h.at(p.k[0]) ^= 1;
        h.at(p.k[1]) ^= 1;
    }

    // ------------------------------------------------------------------ unambiguous merge (mag.c:310-399)
    int merge_try(size_t pi, int min_merge_len)
    {
        MagV &p = v[pi];
        if (p.nei[1].size() != 1) return -1;
        if ((i64)p.nei[1][0].x < 0) return -2;
        if ((int)p.nei[1][0].y < min_merge_len) return -5;
        u64 qx = p.nei[1][0].x;
        u64 kq = tid2idd(qx);
        MagV &q = v[kq >> 1];
        if (&p == &q) return -3;
        if (q.nei[kq & 1].size() != 1) return -4;
        if (kq & 1) v_flip(q);
        h.erase(p.k[1]); h.erase(qx);
        int ov = (int)p.nei[1][0].y;
        if (!(p.k[1] == q.nei[0][0].x && q.k[0] == qx) || p.nei[1][0].y != q.nei[0][0].y || p.len < ov || q.len < ov)
            throw std::logic_error("mag: inconsistent topology in merge");          // the reference asserts
        p.nsr += q.nsr;
        int new_l = p.len + q.len - ov;
        p.seq.resize(new_l); p.cov.resize(new_l);
        for (int i = p.len - ov, j = 0; j < q.len; ++i, ++j) {
            p.seq[i] = q.seq[j];
            if (i < p.len) {
                if ((int)p.cov[i] + (q.cov[j] - 33) > 126) p.cov[i] = 126;
                else p.cov[i] = (char)(p.cov[i] + (q.cov[j] - 33));
            } else p.cov[i] = q.cov[j];
        }
        p.len = new_l;
        p.nei[1].swap(q.nei[1]); p.k[1] = q.k[1];
        q.nei[1].clear();
        h.at(p.k[1]) = (u64)pi << 1 | 1;
        v_destroy(q);
        return 0;
    }
    void merge(int rmdup, int min_merge_len)
    {
        for (size_t i = 0; i < v.size(); ++i) {
            if (rmdup) { v128_rmdup(v[i].nei[0]); v128_rmdup(v[i].nei[1]); }
            else { v128_clean(v[i].nei[0]); v128_clean(v[i].nei[1]); }
        }
        for (size_t i = 0; i < v.size(); ++i) {
            if (v[i].len < 0) continue;
            while (merge_try(i, min_merge_len) == 0) {}
            v_flip(v[i]);
            while (merge_try(i, min_merge_len) == 0) {}
        }
    }

    // ------------------------------------------------------------------ easy simplification (mag.c:401-483)
    struct VLess1 { bool operator()(const MagV *a, const MagV *b) const { return a->nsr < b->nsr || (a->nsr == b->nsr && a->len < b->len); } };
    int rm_vext(int min_len, int min_nsr)
    {
        std::vector<MagV *> a;
        for (size_t i = 0; i < v.size(); ++i) {
            MagV *p = &v[i];
            if (p->len < 0 || (!p->nei[0].empty() && !p->nei[1].empty())) continue;
            if (p->len >= min_len || p->nsr >= min_nsr) continue;
            a.push_back(p);
        }
        introsort(a.size(), a.data(), VLess1());
        for (size_t i = 0; i < a.size(); ++i) v_del(*a[i]);
        return (int)a.size();
    }
    int rm_vint(int min_len, int min_nsr, int min_ovlp_)
    {
        std::vector<MagV *> a;
        for (size_t i = 0; i < v.size(); ++i) {
            MagV *p = &v[i];
            if (p->len >= 0 && p->len < min_len && p->nsr < min_nsr) a.push_back(p);
        }
        introsort(a.size(), a.data(), VLess1());
        for (size_t i = 0; i < a.size(); ++i) v_transdel(*a[i], min_ovlp_);
        return (int)a.size();
    }
    void rm_edge(int min_ovlp_, double min_ratio, int min_len, int min_nsr)
    {
        std::vector<MagV *> a;
        for (size_t i = 0; i < v.size(); ++i) {
            MagV *p = &v[i];
            if (p->len < 0) continue;
            if ((p->nei[0].empty() || p->nei[1].empty()) && p->len < min_len && p->nsr < min_nsr) continue;
            a.push_back(p);
        }
        introsort(a.size(), a.data(), VLess1());
        for (i64 i = (i64)a.size() - 1; i >= 0; --i) {
            MagV *p = a[i];
            for (int j = 0; j < 2; ++j) {
                std::vector<MagEdge> &r = p->nei[j];
                int max_ovlp = min_ovlp_, max_k = -1;
                if (r.empty()) continue;
                for (size_t k = 0; k < r.size(); ++k)
                    if ((u64)(i64)max_ovlp < r[k].y) { max_ovlp = (int)r[k].y; max_k = (int)k; }
                if (max_k >= 0) {
                    u64 x = tid2idd(r[max_k].x);
                    const MagV *q = &v[x >> 1];
                    if (q->len >= 0 && (q->nei[0].empty() || q->nei[1].empty()) && q->len < min_len && q->nsr < min_nsr) max_ovlp = min_ovlp_;
                }
                for (size_t k = 0; k < r.size(); ++k) {
                    if (edge_is_del(r[k])) continue;
                    if (r[k].y < (u64)(i64)min_ovlp_ || (double)r[k].y / max_ovlp < min_ratio) {
                        eh_markdel(r[k].x, p->k[j]);
                        edge_mark_del(r[k]);
                    }
                }
            }
        }
    }

    // ------------------------------------------------------------------ bubbles (bubble.c:179-366)
    // f_ksw_align(..., xtra = 0) -> f_ksw_i16 (fermi-lite/f_ksw.c:247-358); only .score is consumed by the callers.
    // This fork's 16-bit kernel ADDS the gap penalties (`_mm_adds_epi16(h, gapoe)`, f_ksw.c:302-307,316-317) instead of
    // subtracting them and never clamps at zero, so the value it returns is not a Smith-Waterman score and depends on the
    // striped lane layout.  Identical unitigs need the identical number: the eight-lane vector code is replayed literally.
    struct V8 { int16_t v[8]; };
    static int16_t sat16(int x) { return (int16_t)(x > 32767 ? 32767 : x < -32768 ? -32768 : x); }
    static V8 v_adds(const V8 &a, const V8 &b) { V8 r; for (int i = 0; i < 8; ++i) r.v[i] = sat16((int)a.v[i] + b.v[i]); return r; }
    static V8 v_max(const V8 &a, const V8 &b) { V8 r; for (int i = 0; i < 8; ++i) r.v[i] = a.v[i] > b.v[i] ? a.v[i] : b.v[i]; return r; }
    static V8 v_shl(const V8 &a) { V8 r; r.v[0] = 0; for (int i = 1; i < 8; ++i) r.v[i] = a.v[i - 1]; return r; }   // _mm_slli_si128(x, 2)
    static bool v_any_gt(const V8 &a, const V8 &b) { for (int i = 0; i < 8; ++i) if (a.v[i] > b.v[i]) return true; return false; }
    static V8 v_set1(int x) { V8 r; for (int i = 0; i < 8; ++i) r.v[i] = (int16_t)x; return r; }
    static int sw_score(int qlen, const u8 *query, int tlen, const u8 *target, int gapo, int gape_)
    {
        const int slen = (qlen + 7) / 8;
        if (slen == 0) return 0;
        std::vector<V8> qp((size_t)slen * 4), H0(slen, v_set1(0)), H1(slen, v_set1(0)), E(slen, v_set1(0));
        for (int a = 0; a < 4; ++a)                                   // f_ksw_qinit, size 2 (f_ksw.c:105-113)
            for (int i = 0; i < slen; ++i)
                for (int k = i, lane = 0; lane < 8; k += slen, ++lane)
                    qp[(size_t)a * slen + i].v[lane] = (int16_t)(k >= qlen ? 0 : (query[k] == a ? 5 : -4));
        const V8 zero = v_set1(0), gapoe = v_set1(gapo + gape_), gape = v_set1(gape_);
        int gmax = 0;
        for (int i = 0; i < tlen; ++i) {
            V8 e, h, f = zero, max = zero;
            const V8 *S = &qp[(size_t)target[i] * slen];
            h = v_shl(H0[slen - 1]);
            for (int j = 0; j < slen; ++j) {
                h = v_adds(h, S[j]);
                e = E[j];
                h = v_max(h, e);
                h = v_max(h, f);
                max = v_max(max, h);
                H1[j] = h;
                h = v_adds(h, gapoe);
                e = v_adds(e, gape);
                e = v_max(e, h);
                E[j] = e;
                f = v_adds(f, gape);
                f = v_max(f, h);
                h = H0[j];
            }
            for (int k = 0; k < 16; ++k) {
                bool done = false;
                f = v_shl(f);
                for (int j = 0; j < slen; ++j) {
                    h = H1[j];
                    h = v_max(h, f);
                    H1[j] = h;
                    h = v_adds(h, gapoe);
                    f = v_adds(f, gape);
                    if (!v_any_gt(f, h)) { done = true; break; }
                }
                if (done) break;
            }
            int imax = max.v[0];
            for (int t = 1; t < 8; ++t) if (max.v[t] > imax) imax = max.v[t];
            if (imax > gmax) gmax = imax;
            H0.swap(H1);
        }
        return gmax;
    }

    int vh_pop_simple(u64 idd, float max_cov, float max_frac, int max_bdiff, int aggressive)
    {
        const double MAX_N_DIFF = 2.01, MAX_R_DIFF = 0.1, L_DIFF_COEF = 0.2;
        MagV *p = &v[idd >> 1], *q[2];
        int dir[2], l[2], ret = -1;
        std::string seq[2], cov[2];
        float n_diff, r_diff, avg[2] = {0, 0}, max_n_diff = aggressive ? MAX_N_DIFF * 2. : MAX_N_DIFF;
        if (p->len < 0 || p->nei[idd & 1].size() != 2) return ret;
        std::vector<MagEdge> &r = p->nei[idd & 1];
        for (int j = 0; j < 2; ++j) {
            if ((i64)r[j].x < 0) return ret;
            u64 x = tid2idd(r[j].x);
            dir[j] = (int)(x & 1);
            q[j] = &v[x >> 1];
            if (q[j]->nei[0].size() != 1 || q[j]->nei[1].size() != 1) return ret;
            l[j] = q[j]->len - (int)(q[j]->nei[0][0].y + q[j]->nei[1][0].y);
        }
        if (q[0]->nei[dir[0] ^ 1][0].x != q[1]->nei[dir[1] ^ 1][0].x) return ret;
        if (l[0] - l[1] > max_bdiff || l[1] - l[0] > max_bdiff) return 1;
        for (int j = 0; j < 2; ++j) {
            if (l[j] > 0) {
                seq[j].resize(l[j]); cov[j].resize(l[j]);
                for (int i = 0; i < l[j]; ++i) {
                    seq[j][i] = q[j]->seq[i + q[j]->nei[0][0].y];
                    cov[j][i] = q[j]->cov[i + q[j]->nei[0][0].y];
                }
                if (dir[j]) { revcomp6(seq[j]); reverse(cov[j]); }
                avg[j] = 0.;
                for (int i = 0; i < l[j]; ++i) { --seq[j][i]; avg[j] += cov[j][i] - 33; }
                avg[j] /= l[j];
            } else {
                int beg = (int)q[j]->nei[0][0].y, end = q[j]->len - (int)q[j]->nei[1][0].y;
                if (beg > end) std::swap(beg, end);
                if (beg < end) {
                    avg[j] = 0.;
                    for (int i = beg; i < end; ++i) avg[j] += q[j]->cov[i] - 33;
                    avg[j] /= end - beg;
                } else avg[j] = q[j]->cov[beg] - 33;
            }
        }
        ret = 1;
        if (l[0] > 0 && l[1] > 0) {
            int score = sw_score(l[0], (const u8 *)seq[0].data(), l[1], (const u8 *)seq[1].data(), 5, 2);
            n_diff = ((l[0] < l[1] ? l[0] : l[1]) * 5. - score) / (5. + 4.);
            r_diff = n_diff / ((l[0] + l[1]) / 2.);
        } else {
            n_diff = abs(l[0] - l[1]) * L_DIFF_COEF;
            r_diff = 1.;
        }
        if (n_diff < max_n_diff || r_diff < MAX_R_DIFF) {
            int j = avg[0] < avg[1] ? 0 : 1;
            if (aggressive || (avg[j] < max_cov && avg[j] / (avg[j ^ 1] + avg[j]) < max_frac)) {
                v_del(*q[j]);
                ret = 2;
            }
        }
        return ret;
    }
    void pop_simple(float max_cov, float max_frac, int min_merge_len, int max_bdiff, int aggressive)
    {
        for (size_t i = 0; i < v.size(); ++i) {
            vh_pop_simple((u64)i << 1 | 0, max_cov, max_frac, max_bdiff, aggressive);
            vh_pop_simple((u64)i << 1 | 1, max_cov, max_frac, max_bdiff, aggressive);
        }
        merge(0, min_merge_len);
    }

    void v_pop_open(MagV *p, int min_elen)
    {
        const double MAX_N_DIFF = 2.01, MAX_R_DIFF = 0.1;
        if (p->len < 0 || p->len >= min_elen) return;
        if (p->nei[0].size() + p->nei[1].size() != 1) return;
        int dir = !p->nei[0].empty() ? 0 : 1;
        std::vector<MagEdge> &s = p->nei[dir];
        for (size_t l = 0; l < s.size(); ++l) {
            if ((i64)s[l].x < 0) continue;
            u64 vv = tid2idd(s[l].x);
            MagV *q = &v[vv >> 1];
            if (q == p || q->nei[vv & 1].size() == 1) continue;
            int max_l = (p->len - (int)s[l].y) * 2;
            std::vector<u8> qseq, tseq((size_t)(max_l > 0 ? max_l : 0) + 1);
            if (dir == 0) { for (int j = (int)s[l].y; j < p->len; ++j) qseq.push_back((u8)(p->seq[j] - 1)); }
            else { for (int j = p->len - (int)s[l].y - 1; j >= 0; --j) qseq.push_back((u8)(4 - p->seq[j])); }
            int l_qry = (int)qseq.size();
            std::vector<MagEdge> &r = q->nei[vv & 1];
            size_t i;
            for (i = 0; i < r.size(); ++i) {
                if (r[i].x == p->k[dir] || (i64)r[i].x < 0) continue;
                u64 w = tid2idd(r[i].x);
                MagV *t = &v[w >> 1];
                int k = 0;
                if (w & 1) { for (int j = t->len - (int)r[i].y - 1; j >= 0 && k < max_l; --j) tseq[k++] = (u8)(4 - t->seq[j]); }
                else { for (int j = (int)r[i].y; j < t->len && k < max_l; ++j) tseq[k++] = (u8)(t->seq[j] - 1); }
                int score = sw_score(l_qry, qseq.data(), k, tseq.data(), 5, 2);
                if (score >= l_qry * 5 / 2) {
                    double n_diff = (l_qry * 5. - score) / (5. + 4.);
                    double r_diff = n_diff / l_qry;
                    if (n_diff < MAX_N_DIFF || r_diff < MAX_R_DIFF) break;
                }
            }
            if (i != r.size()) {
                edge_mark_del(s[l]);
                for (i = 0; i < r.size(); ++i) if (r[i].x == p->k[dir]) edge_mark_del(r[i]);
            }
        }
        size_t i;
        for (i = 0; i < s.size(); ++i) if (!edge_is_del(s[i])) break;
        if (i == s.size()) v_del(*p);
    }
    void pop_open(int min_elen)
    {
        for (size_t i = 0; i < v.size(); ++i) v_pop_open(&v[i], min_elen);
        merge(0, 0);
    }

    // ------------------------------------------------------------------ closed bubbles (bubble.c:22-176)
    struct TrInfo {                                  // trinfo_t / g_trinull
        u64 id; int cnt[2]; int n[2][2], d[2][2]; u64 v[2][2];
        explicit TrInfo(u64 id_) : id(id_)
        {
            cnt[0] = cnt[1] = 0;
            for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b) { n[a][b] = INT32_MIN; d[a][b] = INT32_MIN; v[a][b] = (u64)-1; }
        }
    };
    // mag_vh_simplify_bubble: keep the two best-supported paths of a bubble starting at end idd, delete the other vertices
    void vh_simplify_bubble(u64 idd, int max_vtx, int max_dist, std::vector<TrInfo> &pool, std::vector<u64> &stack,
                            std::vector<i32> &ptr, std::unordered_map<u64, int> &hh)
    {
        int n_pending = 0;
        MagV *p = &v[idd >> 1], *q;
        if (p->len < 0 || p->nei[idd & 1].size() < 2) return;
        stack.clear(); pool.clear(); hh.clear();
        auto tip_alloc = [&](u64 id) { pool.emplace_back(id); return (i32)pool.size() - 1; };
        ptr[idd >> 1] = tip_alloc(idd >> 1);
        pool[ptr[idd >> 1]].d[(idd & 1) ^ 1][0] = -p->len;
        pool[ptr[idd >> 1]].n[(idd & 1) ^ 1][0] = -p->nsr;
        stack.push_back(idd ^ 1);
        while (!stack.empty()) {
            u64 x, y;
            if (stack.size() == 1 && stack[0] != (idd ^ 1) && n_pending == 0) break;
            x = stack.back(); stack.pop_back();
            p = &v[x >> 1];
            std::vector<MagEdge> &r = p->nei[(x & 1) ^ 1];
            {
                const TrInfo &tp = pool[ptr[x >> 1]];
                if ((int)pool.size() > max_vtx || tp.d[x & 1][0] > max_dist || tp.d[x & 1][1] > max_dist || r.empty()) break;
            }
            for (size_t i = 0; i < r.size(); ++i) {
                int nsr, dist, which;
                if ((i64)r[i].x < 0) continue;
                y = tid2idd(r[i].x);
                if (y == (idd ^ 1)) { stack.clear(); break; }
                q = &v[y >> 1];
                if (ptr[y >> 1] < 0) {
                    ptr[y >> 1] = tip_alloc(y >> 1); ++n_pending;
                    v128_clean(q->nei[y & 1]);
                }
                const TrInfo tp = pool[ptr[x >> 1]];          // copy: tip_alloc may have moved the pool
                TrInfo &tq = pool[ptr[y >> 1]];
                nsr = tp.n[x & 1][0] + p->nsr; which = 0;
                dist = tp.d[x & 1][0] + p->len - (int)r[i].y;
                if (nsr > tq.n[y & 1][0]) {
                    tq.n[y & 1][1] = tq.n[y & 1][0]; tq.n[y & 1][0] = nsr;
                    tq.v[y & 1][1] = tq.v[y & 1][0]; tq.v[y & 1][0] = (x ^ 1) << 32 | (u64)i << 1 | (u64)which;
                    tq.d[y & 1][1] = tq.d[y & 1][0]; tq.d[y & 1][0] = dist;
                    nsr = tp.n[x & 1][1] + p->nsr; which = 1;
                    dist = tp.d[x & 1][1] + p->len - (int)r[i].y;
                }
                if (nsr > tq.n[y & 1][1]) { tq.n[y & 1][1] = nsr; tq.v[y & 1][1] = (x ^ 1) << 32 | (u64)i << 1 | (u64)which; tq.d[y & 1][1] = dist; }
                if (++tq.cnt[y & 1] == (int)q->nei[y & 1].size()) { stack.push_back(y); --n_pending; }
            }
        }
        u64 top = stack.empty() ? 0 : stack[0];
        if (n_pending == 0 && stack.size() == 1) {
            u64 x = stack[0];
            for (int w = 0; w < 2; ++w) {                 // backtrace along the best and the second best path
                u64 end = pool[ptr[x >> 1]].v[x & 1][w];
                while (end >> 32 != idd) {
                    hh.emplace(end >> 33, 1);
                    end = pool[ptr[end >> 33]].v[((end >> 32) ^ 1) & 1][end & 1];
                }
            }
        }
        for (size_t i = 0; i < pool.size(); ++i) ptr[pool[i].id] = -1;
        if (!hh.empty())
            for (size_t i = 1; i < pool.size(); ++i) {
                u64 id = pool[i].id;
                if (id != top >> 1 && hh.find(id) == hh.end()) v_del(v[id]);
            }
    }
    void simplify_bubble(int max_vtx, int max_dist)
    {
        std::vector<TrInfo> pool; std::vector<u64> stack; std::vector<i32> ptr(v.size(), -1); std::unordered_map<u64, int> hh;
        for (size_t i = 0; i < v.size(); ++i) {
            vh_simplify_bubble((u64)i << 1 | 0, max_vtx, max_dist, pool, stack, ptr, hh);
            vh_simplify_bubble((u64)i << 1 | 1, max_vtx, max_dist, pool, stack, ptr, hh);
        }
        merge(0, 0);
    }

    // ------------------------------------------------------------------ portal (mag.c:559-620, misc.c:130-137)
    void clean(const b200_magopt_t &o)
    {
        const int F_AGGRESSIVE = 0x20, F_POPOPEN = 0x40, F_NO_SIMPL = 0x80;
        if (min_ovlp < o.min_ovlp) min_ovlp = o.min_ovlp;
        for (int j = 2; j <= o.min_ensr; ++j) rm_vext(o.min_elen, j);
        merge(0, o.min_merge_len);
        rm_edge(min_ovlp, o.min_dratio1, o.min_elen, o.min_ensr);
        merge(1, o.min_merge_len);
        for (int j = 2; j <= o.min_ensr; ++j) rm_vext(o.min_elen, j);
        merge(0, o.min_merge_len);
        if ((o.flag & F_AGGRESSIVE) || (o.flag & F_POPOPEN)) pop_open(o.min_elen);
        if (!(o.flag & F_NO_SIMPL)) simplify_bubble(o.max_bvtx, o.max_bdist);
        pop_simple(o.max_bcov, o.max_bfrac, o.min_merge_len, o.max_bdiff, o.flag & F_AGGRESSIVE);
        rm_vint(o.min_elen, o.min_insr, min_ovlp);
        rm_edge(min_ovlp, o.min_dratio1, o.min_elen, o.min_ensr);
        merge(1, o.min_merge_len);
        rm_vext(o.min_elen, o.min_ensr);
        merge(0, o.min_merge_len);
        if ((o.flag & F_AGGRESSIVE) || (o.flag & F_POPOPEN)) pop_open(o.min_elen);
        rm_vext(o.min_elen, o.min_ensr);
        merge(0, o.min_merge_len);
    }
    void v_trim_open(MagV &p, int trim_len, int trim_depth)
    {
        int i, tl[2];
        if (!p.nei[0].empty() && !p.nei[1].empty()) return;
        if (p.nei[0].empty() && p.nei[1].empty() && p.len < trim_len * 3) { v_del(p); return; }
        for (int j = 0; j < 2; ++j) {
            std::vector<MagEdge> &r = p.nei[!j];
            int max_ovlp = 0;
            for (size_t k = 0; k < r.size(); ++k) max_ovlp = (u64)(i64)max_ovlp > r[k].y ? max_ovlp : (int)r[k].y;
            tl[j] = p.len - max_ovlp < trim_len ? p.len - max_ovlp : trim_len;
        }
        if (p.nei[0].empty()) {
            for (i = 0; i < tl[0] && p.cov[i] - 33 < trim_depth; ++i) {}
            tl[0] = i;
            p.len -= i;
            p.seq.erase(0, tl[0]); p.cov.erase(0, tl[0]);
        }
        if (p.nei[1].empty()) {
            for (i = p.len - 1; i >= p.len - tl[1] && p.cov[i] - 33 < trim_depth; --i) {}
            tl[1] = p.len - 1 - i;
            p.len -= tl[1];
            p.seq.resize(p.len); p.cov.resize(p.len);
        }
    }
    void trim_open(const b200_magopt_t &o)
    {
        if (o.trim_len == 0) return;
        for (size_t i = 0; i < v.size(); ++i) if (v[i].len >= 0) v_trim_open(v[i], o.trim_len, o.trim_depth);
    }
    // debugging aid for the parity tests: the first n passes of fml_mag_clean, in its order
    void fml_clean_steps(const b200_fml_opt_t &opt, int left)
    {
        b200_magopt_t o = opt.mag_opt;
        o.min_merge_len = opt.min_merge_len;
#define B200_STEP(x) do { if (left-- > 0) { x; } } while (0)
        B200_STEP(merge(1, opt.min_merge_len));
        for (int j = 2; j <= o.min_ensr; ++j) B200_STEP(rm_vext(o.min_elen, j));
        B200_STEP(merge(0, o.min_merge_len));
        B200_STEP(rm_edge(min_ovlp, o.min_dratio1, o.min_elen, o.min_ensr));
        B200_STEP(merge(1, o.min_merge_len));
        for (int j = 2; j <= o.min_ensr; ++j) B200_STEP(rm_vext(o.min_elen, j));
        B200_STEP(merge(0, o.min_merge_len));
        B200_STEP(pop_open(o.min_elen));
        B200_STEP(pop_simple(o.max_bcov, o.max_bfrac, o.min_merge_len, o.max_bdiff, 0));
        B200_STEP(rm_vint(o.min_elen, o.min_insr, min_ovlp));
        B200_STEP(rm_edge(min_ovlp, o.min_dratio1, o.min_elen, o.min_ensr));
        B200_STEP(merge(1, o.min_merge_len));
        B200_STEP(rm_vext(o.min_elen, o.min_ensr));
        B200_STEP(merge(0, o.min_merge_len));
        B200_STEP(pop_open(o.min_elen));
        B200_STEP(rm_vext(o.min_elen, o.min_ensr));
        B200_STEP(merge(0, o.min_merge_len));
#undef B200_STEP
    }

    // fml_mag_clean (misc.c:130-137)
    void fml_clean(const b200_fml_opt_t &opt)
    {
        b200_magopt_t o = opt.mag_opt;
        o.min_merge_len = opt.min_merge_len;
        merge(1, opt.min_merge_len);
        clean(o);
        trim_open(o);
    }

    // mag_v_write text of every live vertex (mag.c:151-176): the stage dump parity tests compare
    std::string text() const
    {
        std::string out;
        for (const MagV &p : v) {
            if (p.len <= 0) continue;
            out += '@'; out += std::to_string((long long)p.k[0]); out += ':'; out += std::to_string((long long)p.k[1]);
            out += '\t'; out += std::to_string(p.nsr);
            for (int j = 0; j < 2; ++j) {
                out += '\t';
                for (const MagEdge &e : p.nei[j]) {
                    if (edge_is_del(e)) continue;
                    out += std::to_string((long long)e.x); out += ','; out += std::to_string((int32_t)e.y); out += ';';
                }
                if (p.nei[j].empty()) out += '.';
            }
            out += '\n';
            for (int j = 0; j < p.len; ++j) out += "ACGT"[(int)p.seq[j] - 1];
            out += "\n+\n";
            out.append(p.cov.data(), p.len);
            out += '\n';
        }
        return out;
    }

    // fml_mag2utg (misc.c:153-213)
    std::vector<MagUtg> to_utg() const
    {
        std::unordered_map<u64, u64> hh;
        size_t j = 0;
        for (const MagV &p : v) {
            if (p.len < 0) continue;
            hh[p.k[0]] = (u64)j << 1 | 0;
            hh[p.k[1]] = (u64)j << 1 | 1;
            ++j;
        }
        std::vector<MagUtg> utg(j);
        j = 0;
        for (const MagV &p : v) {
            if (p.len < 0) continue;
            MagUtg &q = utg[j++];
            q.nsr = p.nsr;
            q.seq.resize(p.len); q.cov.assign(p.cov.data(), p.len);
            for (int a = 0; a < p.len; ++a) q.seq[a] = "$ACGTN"[(int)p.seq[a]];
            for (int from = 0; from < 2; ++from) {
                q.n_ovlp[from] = 0;
                for (const MagEdge &e : p.nei[from]) {
                    if (edge_is_del(e)) continue;
                    ++q.n_ovlp[from];
                    auto it = hh.find(e.x);
                    if (it == hh.end()) throw std::logic_error("mag: overlap with a missing unitig");
                    MagUtgOvlp o; o.id = (u32)(it->second >> 1); o.to = (int)(it->second & 1); o.len = (u32)e.y; o.from = from;
                    q.ovlp.push_back(o);
                }
            }
        }
        return utg;
    }
};

} // namespace b200
