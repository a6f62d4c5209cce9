// TEST HARNESS ONLY.  Builds the __host__ side of csrc/mathcore.cuh + csrc/select.cuh into a small shared library so
// the CPU test-suite (-m "not gpu") can check the very code the CUDA kernels execute against the oracle, before any
// GPU time is spent.  Nothing in the product loads this library.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>
#include "../../droplet_visual_odometry_b200/csrc/mathcore.cuh"
#include "../../droplet_visual_odometry_b200/csrc/select.cuh"

using namespace dvo;

struct HostItem { float r; int32_t idx; };

struct HostAcc {
    typedef HostItem Item;
    HostItem* d;
    Item get(int i) { return d[i]; }
    void set(int i, Item v) { d[i] = v; }
    static bool gt(Item a, Item b) { return a.r > b.r; }
    int scan_up(int first, Item pivot) { while (gt(d[first], pivot)) ++first; return first; }
    int scan_down(int last, Item pivot) { while (gt(pivot, d[last])) --last; return last; }
    int scan_up_ge(int first, int last, Item b) { while (first != last && d[first].r >= b.r) ++first; return first; }
    int scan_down_lt(int first, int last, Item b) { while (first != last && !(d[last].r >= b.r)) --last; return last; }
};

// Plain-loop accessor for the paired (data-parallel) formulation of select.cuh: builds the stopper lists explicitly.
struct HostPairAcc : HostAcc {
    template <class IsL, class IsR>
    int pair_swap(int first, int last, IsL isL, IsR isR, int* cut_hoare) {
        std::vector<int> L, R;
        for (int i = first; i < last; ++i) if (isL(d[i])) L.push_back(i);
        for (int i = last - 1; i >= first; --i) if (isR(d[i])) R.push_back(i);
        size_t K = 0;
        while (K < L.size() && K < R.size() && L[K] < R[K]) ++K;
        for (size_t k = 0; k < K; ++k) std::swap(d[L[k]], d[R[k]]);
        int lk = K < L.size() ? L[K] : 0x7fffffff, rk1 = K > 0 ? R[K - 1] : 0x7fffffff;
        *cut_hoare = std::min(lk, rk1);
        return (int)R.size();
    }
    void median_to_first(int result, int ia, int ib, int ic) { move_median_to_first(*this, result, ia, ib, ic); }
    int pair_swap_hoare(int first, int last, Item pivot) {
        int cut;
        pair_swap(first, last, [&](Item x) { return !gt(x, pivot); }, [&](Item x) { return !gt(pivot, x); }, &cut);
        return cut;
    }
    int pair_swap_ge(int first, int last, Item b) {
        int cut;
        int nr = pair_swap(first, last, [&](Item x) { return !(x.r >= b.r); }, [&](Item x) { return x.r >= b.r; }, &cut);
        return first + nr;
    }
    void sequential_tail(int first, int nth, int last, int depth) { nth_element_replay_depth(*this, first, nth, last, depth); }
};

extern "C" {

int hs_retain_best_paired(const float* resp, int count, int n_points, int seq_tail, int32_t* out_idx) {
    std::vector<HostItem> v(count);
    for (int i = 0; i < count; ++i) v[i] = HostItem{resp[i], i};
    HostPairAcc acc;
    acc.d = v.data();
    int k = retain_best_paired(acc, count, n_points, seq_tail);
    for (int i = 0; i < k; ++i) out_idx[i] = v[i].idx;
    return k;
}

int hs_retain_best(const float* resp, int count, int n_points, int32_t* out_idx) {
    std::vector<HostItem> v(count);
    for (int i = 0; i < count; ++i) v[i] = HostItem{resp[i], i};
    HostAcc acc{v.data()};
    int k = retain_best_replay(acc, count, n_points);
    for (int i = 0; i < k; ++i) out_idx[i] = v[i].idx;
    return k;
}

// heap_select path in isolation: mine vs libstdc++'s (internal) std::__heap_select + iter_swap as introselect does.
int hs_heap_select_check(const float* resp, int count, int nth) {
    std::vector<HostItem> mine(count), ref(count);
    for (int i = 0; i < count; ++i) mine[i] = ref[i] = HostItem{resp[i], i};
    HostAcc acc{mine.data()};
    heap_select(acc, 0, nth + 1, count);
    sel_swap(acc, 0, nth);
    auto comp = [](const HostItem& a, const HostItem& b) { return a.r > b.r; };
    std::__heap_select(ref.begin(), ref.begin() + nth + 1, ref.end(), __gnu_cxx::__ops::__iter_comp_iter(comp));
    std::iter_swap(ref.begin(), ref.begin() + nth);
    for (int i = 0; i < count; ++i)
        if (mine[i].idx != ref[i].idx) return i + 1;
    return 0;
}

void hs_fast_score_map(const uint8_t* img, int w, int h, int pitch, int t, uint8_t* out) {
    const int dx[16] = DVO_FAST_DX, dy[16] = DVO_FAST_DY;
    std::memset(out, 0, (size_t)w * h);
    for (int y = 3; y < h - 3; ++y)
        for (int x = 3; x < w - 3; ++x) {
            int p[16];
            for (int k = 0; k < 16; ++k) p[k] = img[(y + dy[k]) * pitch + x + dx[k]];
            out[y * w + x] = (uint8_t)fast_score16(img[y * pitch + x], p, t);
        }
}

// number of corners (fast_score16 > 0) the packed prefilter would reject: must be 0.  Also returns the pass count.
int hs_fast_prefilter_check(const uint8_t* img, int w, int h, int pitch, int t, int* passed) {
    const int dx[16] = DVO_FAST_DX, dy[16] = DVO_FAST_DY;
    int missed = 0, np_ = 0;
    for (int y = 3; y < h - 3; ++y)
        for (int x = 4; x + 4 < w - 3; x += 4) {
            auto quad = [&](int yy, int xx) {
                uint32_t r = 0;
                for (int k = 0; k < 4; ++k) r |= (uint32_t)img[yy * pitch + xx + k] << (8 * k);
                return r;
            };
            uint32_t pass = fast_prefilter_u8x4(quad(y, x), quad(y - 3, x), quad(y, x + 3), quad(y + 3, x), quad(y, x - 3), t);
            for (int k = 0; k < 4; ++k) {
                int p[16];
                for (int j = 0; j < 16; ++j) p[j] = img[(y + dy[j]) * pitch + x + k + dx[j]];
                bool corner = fast_score16(img[y * pitch + x + k], p, t) > 0;
                bool ps = (pass >> (8 * k + 7)) & 1;
                np_ += ps;
                if (corner && !ps) ++missed;
            }
        }
    *passed = np_;
    return missed;
}

// exhaustive check of the one-IMAD ring comparison used by the device build of fast_corner_polarity16: returns the number of
// (v, t, p) triples, all in 0..255, where a flag differs from the plain comparison (must be 0)
int hs_fast_ring_flags_check() {
    int bad = 0;
    for (int v = 0; v < 256; ++v)
        for (int t = 0; t < 256; ++t) {
            const int lo = v - t, hi = v + t;
            const uint32_t bias = fast_ring_flag_bias(lo, hi);
            for (int p = 0; p < 256; ++p) {
                const uint32_t f = fast_ring_flags((uint32_t)p, bias);
                if (((f >> 31) & 1u) != (uint32_t)(p > hi) || ((f >> 15) & 1u) != (uint32_t)(p < lo) || (f & ~0x80008000u)) ++bad;
            }
        }
    return bad;
}

// exhaustive check of the shift-and-AND 9-arc detector against a direct cyclic run-length count: number of 16-bit ring
// masks on which they disagree (must be 0)
int hs_ring_has9_check() {
    int bad = 0;
    for (uint32_t m = 0; m < 65536u; ++m) {
        int best = 0;
        for (int s = 0; s < 16; ++s) {
            int run = 0;
            while (run < 16 && ((m >> ((s + run) & 15)) & 1u)) ++run;
            if (run > best) best = run;
        }
        if (ring_has9(m) != (best >= 9)) ++bad;
    }
    return bad;
}

float hs_harris(int a, int b, int c) { return harris_from_sums(a, b, c); }
float hs_fast_atan2(float y, float x) { return fast_atan2_deg(y, x); }
int hs_five_point(const double* x1, const double* x2, double* models) { return five_point_solve(x1, x2, models); }
void hs_null_space(const double* Q, double* basis) { null_space_5x9(Q, basis); }
void hs_decompose(const double* E, double* R1, double* R2, double* t) { decompose_essential(E, R1, R2, t); }
int hs_cheirality(const double* R, const double* t, double x1, double y1, double x2, double y2, double dist) {
    return cheirality_ok(R, t, x1, y1, x2, y2, dist) ? 1 : 0;
}
int hs_cheirality_pair(const double* R, const double* t, double x1, double y1, double x2, double y2, double dist) {
    return cheirality_pair(R, t, x1, y1, x2, y2, dist);
}
void hs_triangulate(const double* R, const double* t, double x1, double y1, double x2, double y2, double* X) {
    triangulate_one(R, t, x1, y1, x2, y2, X);
}
float hs_sampson(const double* E, double x1, double y1, double x2, double y2) { return sampson_error_f32(E, x1, y1, x2, y2); }
int hs_sampson_inlier(const double* E, double x1, double y1, double x2, double y2, float t32) {
    double T = (double)t32;
    return sampson_inlier(E, x1, y1, x2, y2, t32, T * (1.0 - 1e-6), T * (1.0 + 1e-6)) ? 1 : 0;
}
int hs_update_iters(double p, double ep, int mp, int mx) { return ransac_update_num_iters(p, ep, mp, mx); }
void hs_rng_stream(uint32_t* out, int n) {
    uint64_t s = 0xFFFFFFFFFFFFFFFFull;
    for (int i = 0; i < n; ++i) out[i] = cvrng_next(s);
}
int hs_aberth_iters(const double* c, int deg) {
    cplx z[16];
    return aberth_roots(c, deg, z);
}
}
