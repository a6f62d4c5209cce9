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

extern "C" {

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

float hs_harris(int a, int b, int c) { return harris_from_sums(a, b, c); }
float hs_fast_atan2(float y, float x) { return fast_atan2_deg(y, x); }
int hs_five_point(const double* x1, const double* x2, double* models) { return five_point_solve(x1, x2, models); }
void hs_decompose(const double* E, double* R1, double* R2, double* t) { decompose_essential(E, R1, R2, t); }
int hs_cheirality(const double* R, const double* t, double x1, double y1, double x2, double y2, double dist) {
    return cheirality_ok(R, t, x1, y1, x2, y2, dist) ? 1 : 0;
}
void hs_triangulate(const double* R, const double* t, double x1, double y1, double x2, double y2, double* X) {
    triangulate_one(R, t, x1, y1, x2, y2, X);
}
float hs_sampson(const double* E, double x1, double y1, double x2, double y2) { return sampson_error_f32(E, x1, y1, x2, y2); }
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
