// Per-element arithmetic of the VO hot path, written once as __host__ __device__ so the exact code the kernels run
// can also be exercised on the CPU by tests/hostsim (test harness only -- the product never calls the host build).
//
// Each routine restates one step of the OpenCV algorithm the reference reaches through
// /root/reference/scripts/visual_odometry_v3.py:373 (ORB), :297-300 (findEssentialMat), :303-306 (recoverPose);
// the exact arithmetic contract is SURVEY.md Appendix A.  All float32 steps that must be contraction-exact use the
// explicit rn helpers below and this directory is compiled with -fmad=false.
#pragma once
#include <stdint.h>
#include <math.h>
#include <float.h>

#if defined(__CUDACC__)
#define DVO_HD __host__ __device__ __forceinline__
#define DVO_HDN __host__ __device__ inline
#else
#define DVO_HD inline
#define DVO_HDN inline
#endif

namespace dvo {

// ---- contraction-proof float helpers ---------------------------------------------------------------------------
DVO_HD float fmul(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fmul_rn(a, b);
#else
    volatile float r = a * b; return r;
#endif
}
DVO_HD float fadd(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fadd_rn(a, b);
#else
    volatile float r = a + b; return r;
#endif
}
DVO_HD float fsub(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fsub_rn(a, b);
#else
    volatile float r = a - b; return r;
#endif
}
DVO_HD float fdiv(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fdiv_rn(a, b);
#else
    volatile float r = a / b; return r;
#endif
}
DVO_HD float ffma(float a, float b, float c) {
#if defined(__CUDA_ARCH__)
    return __fmaf_rn(a, b, c);
#else
    return fmaf(a, b, c);
#endif
}
DVO_HD double dmul(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    volatile double r = a * b; return r;
#endif
}
DVO_HD double dadd(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    volatile double r = a + b; return r;
#endif
}

// ---- A.2 FAST-9/16 ---------------------------------------------------------------------------------------------
// circle offsets (dx, dy), k = 0..15
#define DVO_FAST_DX {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1}
#define DVO_FAST_DY {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3}

// true iff the 16-bit ring mask has >= 9 contiguous set bits (cyclic)
DVO_HD bool ring_has9(uint32_t m) {
    uint32_t x = m | (m << 16);
    x &= x >> 1;   // 2 contiguous
    x &= x >> 2;   // 4
    x &= x >> 4;   // 8
    x &= (m | (m << 16)) >> 8;  // 9
    return (x & 0xFFFFu) != 0;
}

DVO_HD int imin(int a, int b) { return a < b ? a : b; }
DVO_HD int imax(int a, int b) { return a > b ? a : b; }
// three-input min / max: one VIMNMX3 on sm_90+ (DPX), two compares elsewhere
DVO_HD int imin3(int a, int b, int c) {
#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 900
    return __vimin3_s32(a, b, c);
#else
    return imin(imin(a, b), c);
#endif
}
DVO_HD int imax3(int a, int b, int c) {
#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 900
    return __vimax3_s32(a, b, c);
#else
    return imax(imax(a, b), c);
#endif
}

// Both threshold comparisons of one ring pixel p (0..255) in one multiply-add: bit 31 <=> p > hi, bit 15 <=> p < lo, for
// lo = v - t, hi = v + t with v, t in 0..255 (checked exhaustively on the host: tests/hostsim hs_fast_ring_flags_check).
DVO_HD uint32_t fast_ring_flag_bias(int lo, int hi) { return ((uint32_t)(0x7FFF - hi) << 16) + (uint32_t)(0x8000 + lo - 1); }
DVO_HD uint32_t fast_ring_flags(uint32_t p, uint32_t bias) { return (p * 0xFFFFu + bias) & 0x80008000u; }

// Corner test of one pixel: true iff 9 contiguous ring pixels are all darker than v - t or all brighter than v + t.
// Returns the passing polarities: bit 0 = a 9-arc with every d > t, bit 1 = a 9-arc with every d < -t (0 = not a corner).
DVO_HD int fast_corner_polarity16(int v, const int* p, int t) {
    const int lo = v - t, hi = v + t;
#if defined(__CUDA_ARCH__)
    // Both comparisons of a ring pixel in one IMAD: W = p * 0xFFFF + C has p + 0x7FFF - hi in its high half (bit 31 <=> p > hi)
    // and 0x8000 + lo - 1 - p in its low half (bit 15 <=> p < lo); neither half leaves [0, 0xFFFF], so they do not interact.
    // acc = (acc >> 1) | (W & 0x80008000) collects the sixteen flag pairs: high half = "p > hi" ring mask, low half = "p < lo".
    const uint32_t C = fast_ring_flag_bias(lo, hi);
    uint32_t acc = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) acc = (acc >> 1) | fast_ring_flags((uint32_t)p[k], C);
    const uint32_t brighter = acc & 0xFFFFu, darker = acc >> 16;
#else
    // bit per ring pixel, shifted in at the bottom (ring order reversed -- contiguity is what matters):
    // sign(p - lo) <=> p < v - t <=> d > t ;  sign(hi - p) <=> p > v + t <=> d < -t
    uint32_t brighter = 0, darker = 0;
    for (int k = 0; k < 16; ++k) {
        brighter = (brighter << 1) | ((uint32_t)(p[k] - lo) >> 31);
        darker = (darker << 1) | ((uint32_t)(hi - p[k]) >> 31);
    }
#endif
    return (ring_has9(brighter) ? 1 : 0) | (ring_has9(darker) ? 2 : 0);
}
DVO_HD bool fast_is_corner16(int v, const int* p, int t) { return fast_corner_polarity16(v, p, t) != 0; }

// Score of a pixel already known to be a corner at threshold t: m-1 where m = max over the 16 arcs of
// max(min(d), min(-d)), d_k = v - p_k  (m > t  <=>  corner).  The arc maximum is evaluated pairwise (two 9-arcs share an
// 8-arc), starting the running bound at t.
// (A straightforward 16x9 min/max double loop was miscompiled by nvcc 12.9 for sm_100a -- returned max(d) -- so this
// formulation is deliberate; tests/test_gpu_* check it against the oracle on every level.)
// `pol` (fast_corner_polarity16) lets the loop of a polarity without a passing arc be skipped: it could not move its bound.
DVO_HD int fast_corner_score16(int v, const int* p, int t, int pol = 3) {
    int d[25];
#pragma unroll
    for (int k = 0; k < 16; ++k) d[k] = v - p[k];
#pragma unroll
    for (int k = 16; k < 25; ++k) d[k] = d[k - 16];
    int a0 = t;
    if (pol & 1)
#pragma unroll
    for (int k = 0; k < 16; k += 2) {
        int a = imin3(d[k + 1], d[k + 2], d[k + 3]);
        if (a <= a0) continue;
        a = imin3(a, imin3(d[k + 4], d[k + 5], d[k + 6]), imin(d[k + 7], d[k + 8]));
        a0 = imax3(a0, imin(a, d[k]), imin(a, d[k + 9]));
    }
    int b0 = -a0;
    if (pol & 2)
#pragma unroll
    for (int k = 0; k < 16; k += 2) {
        int b = imax3(d[k + 1], d[k + 2], d[k + 3]);
        if (b >= b0) continue;
        b = imax3(b, imax3(d[k + 4], d[k + 5], d[k + 6]), imax(d[k + 7], d[k + 8]));
        b0 = imin3(b0, imax(b, d[k]), imax(b, d[k + 9]));
    }
    return -b0 - 1;
}

// Corner score of one pixel given centre value v and the 16 ring values; 0 if not a corner at threshold t.
DVO_HD int fast_score16(int v, const int* p, int t) {
    const int pol = fast_corner_polarity16(v, p, t);
    return pol ? fast_corner_score16(v, p, t, pol) : 0;
}

// Packed prefilter on 4 horizontally adjacent pixels (one per byte): c = centres, n/e/s/w = the compass ring points
// (k = 0, 4, 8, 12) of each.  A 9-arc always holds two adjacent compass points, so a corner needs two adjacent compass
// points with |v - p| > t; the byte test used here, (|v-p| >> 1) >= (t >> 1), is implied by it (never rejects a corner).
// Returns 0x80 in every byte that may be a corner.
DVO_HD uint32_t absdiff_u8x4(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __vabsdiffu4(a, b);
#else
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i) {
        int x = (int)((a >> (8 * i)) & 0xFF) - (int)((b >> (8 * i)) & 0xFF);
        r |= (uint32_t)(x < 0 ? -x : x) << (8 * i);
    }
    return r;
#endif
}
DVO_HD uint32_t fast_prefilter_u8x4(uint32_t c, uint32_t n, uint32_t e, uint32_t s, uint32_t w, int t) {
    const uint32_t bias = 0x01010101u * (uint32_t)(128 - (t >> 1));
    uint32_t mn = (((absdiff_u8x4(c, n) >> 1) & 0x7f7f7f7fu) + bias);
    uint32_t me = (((absdiff_u8x4(c, e) >> 1) & 0x7f7f7f7fu) + bias);
    uint32_t ms = (((absdiff_u8x4(c, s) >> 1) & 0x7f7f7f7fu) + bias);
    uint32_t mw = (((absdiff_u8x4(c, w) >> 1) & 0x7f7f7f7fu) + bias);
    return (mn | ms) & (me | mw) & 0x80808080u;
}

// ---- A.3 Harris ------------------------------------------------------------------------------------------------
DVO_HD float harris_from_sums(int a, int b, int c) {
    const float scale = fdiv(1.0f, fmul(28.0f, 255.0f));
    const float s4 = fmul(fmul(fmul(scale, scale), scale), scale);
    float af = (float)a, bf = (float)b, cf = (float)c;
    float t1 = fmul(af, bf);
    float t2 = fmul(cf, cf);
    float sm = fadd(af, bf);
    float t3 = fmul(fmul(0.04f, sm), sm);
    return fmul(fsub(fsub(t1, t2), t3), s4);
}

// ---- A.5 cv::fastAtan2 -------------------------------------------------------------------------------------------
DVO_HD float fast_atan2_deg(float y, float x) {
    const float s = (float)(180.0 / 3.14159265358979323846);
    const float p1 = fmul(0.9997878412794807f, s);
    const float p3 = fmul(-0.3258083974640975f, s);
    const float p5 = fmul(0.1555786518463281f, s);
    const float p7 = fmul(-0.04432655554792128f, s);
    const float eps = (float)DBL_EPSILON;
    float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = fdiv(ay, fadd(ax, eps));
        c2 = fmul(c, c);
        a = fmul(fadd(fmul(fadd(fmul(fadd(fmul(p7, c2), p5), c2), p3), c2), p1), c);
    } else {
        c = fdiv(ax, fadd(ay, eps));
        c2 = fmul(c, c);
        a = fsub(90.0f, fmul(fadd(fmul(fadd(fmul(fadd(fmul(p7, c2), p5), c2), p3), c2), p1), c));
    }
    if (x < 0) a = fsub(180.0f, a);
    if (y < 0) a = fsub(360.0f, a);
    return a;
}

// ---- A.9 cv::RNG and the adaptive iteration rule -------------------------------------------------------------------
DVO_HD uint32_t cvrng_next(uint64_t& state) {
    state = (uint64_t)(uint32_t)state * 4164903690ull + (state >> 32);
    return (uint32_t)state;
}

DVO_HDN int ransac_update_num_iters(double p, double ep, int model_points, int max_iters) {
    p = p < 0.0 ? 0.0 : (p > 1.0 ? 1.0 : p);
    ep = ep < 0.0 ? 0.0 : (ep > 1.0 ? 1.0 : ep);
    double num = 1.0 - p;
    if (num < DBL_MIN) num = DBL_MIN;
    double denom = 1.0 - pow(1.0 - ep, (double)model_points);
    if (denom < DBL_MIN) return 0;
    num = log(num);
    denom = log(denom);
    if (denom >= 0 || -num >= max_iters * (-denom)) return max_iters;
    return (int)rint(num / denom);
}

// Sampson error exactly as cv2's EMEstimatorCallback::computeError: float64, no contraction, cast to float32.
DVO_HD float sampson_error_f32(const double* E, double x1, double y1, double x2, double y2) {
    double ex0 = dadd(dadd(dmul(E[0], x1), dmul(E[1], y1)), E[2]);
    double ex1 = dadd(dadd(dmul(E[3], x1), dmul(E[4], y1)), E[5]);
    double ex2 = dadd(dadd(dmul(E[6], x1), dmul(E[7], y1)), E[8]);
    double et0 = dadd(dadd(dmul(E[0], x2), dmul(E[3], y2)), E[6]);
    double et1 = dadd(dadd(dmul(E[1], x2), dmul(E[4], y2)), E[7]);
    double x2tEx1 = dadd(dadd(dmul(x2, ex0), dmul(y2, ex1)), ex2);
    double den = dadd(dadd(dadd(dmul(ex0, ex0), dmul(ex1, ex1)), dmul(et0, et0)), dmul(et1, et1));
    return (float)(dmul(x2tEx1, x2tEx1) / den);
}

// Same decision as `sampson_error_f32(...) <= t32`, with the division skipped when the ratio is not within 1e-6 of the
// threshold: tlo = (double)t32 * (1 - 1e-6), thi = (double)t32 * (1 + 1e-6).  float rounding is monotonic and t32 is a
// float, so num/den <= tlo implies (float)(num/den) <= t32, and num/den >= thi (> t32's upper rounding midpoint, relative
// 6e-8) implies (float)(num/den) > t32; anything in between, and every non-finite case, takes the exact path.
DVO_HD bool sampson_inlier(const double* E, double x1, double y1, double x2, double y2, float t32, double tlo, double thi) {
    double ex0 = dadd(dadd(dmul(E[0], x1), dmul(E[1], y1)), E[2]);
    double ex1 = dadd(dadd(dmul(E[3], x1), dmul(E[4], y1)), E[5]);
    double ex2 = dadd(dadd(dmul(E[6], x1), dmul(E[7], y1)), E[8]);
    double et0 = dadd(dadd(dmul(E[0], x2), dmul(E[3], y2)), E[6]);
    double et1 = dadd(dadd(dmul(E[1], x2), dmul(E[4], y2)), E[7]);
    double x2tEx1 = dadd(dadd(dmul(x2, ex0), dmul(y2, ex1)), ex2);
    double den = dadd(dadd(dadd(dmul(ex0, ex0), dmul(ex1, ex1)), dmul(et0, et0)), dmul(et1, et1));
    double num = dmul(x2tEx1, x2tEx1);
    if (den > 0.0 && den < 1e300 && num < 1e300) {
        if (num <= dmul(den, tlo)) return true;
        if (num >= dmul(den, thi)) return false;
    }
    return (float)(num / den) <= t32;
}

// cv.findEssentialMat's K-normalisation, bit for bit: OpenCV evaluates `(points.col(0) - cx) / fx` as the matrix expression
// alpha * p + beta with alpha = 1 / fx and beta = -cx * alpha, and its convertTo kernel (AVX2/FMA3 dispatch) fuses the
// multiply-add.  (p - cx) / fx differs from that in the last bit for most points, and a 1-ulp change of the normalised
// coordinates is enough to change the RANSAC winner on ~2 % of real frame pairs (ill-conditioned minimal samples).
// Verified against cv2 4.13.0: 120 of 120 five-point calls bit-identical with this form, 0 of 120 with any other.
DVO_HD double cv_normalize_coord(float p, double f, double c) {
    const double alpha = 1.0 / f;
    return fma((double)p, alpha, -(c * alpha));
}

// ---- small dense linear algebra ------------------------------------------------------------------------------------
// One-sided (Hestenes) Jacobi on the columns of an NxN matrix A (row-major, overwritten by U*Sigma); V accumulates the
// right singular vectors as columns.
template <int N>
DVO_HDN void jacobi_onesided(double* A, double* V) {
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) V[i * N + j] = (i == j) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 30; ++sweep) {
        bool changed = false;
        for (int p = 0; p < N - 1; ++p)
            for (int q = p + 1; q < N; ++q) {
                double alpha = 0, beta = 0, gamma = 0;
                for (int i = 0; i < N; ++i) {
                    alpha += A[i * N + p] * A[i * N + p];
                    beta += A[i * N + q] * A[i * N + q];
                    gamma += A[i * N + p] * A[i * N + q];
                }
                if (fabs(gamma) <= DBL_EPSILON * sqrt(alpha * beta) || gamma == 0.0) continue;
                changed = true;
                double zeta = (beta - alpha) / (2.0 * gamma);
                double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
                for (int i = 0; i < N; ++i) {
                    double ap = A[i * N + p], aq = A[i * N + q];
                    A[i * N + p] = c * ap - s * aq;
                    A[i * N + q] = s * ap + c * aq;
                    double vp = V[i * N + p], vq = V[i * N + q];
                    V[i * N + p] = c * vp - s * vq;
                    V[i * N + q] = s * vp + c * vq;
                }
            }
        if (!changed) break;
    }
}

// cv::SVD::compute(A, w, u, vt) for an NxN double matrix the way OpenCV runs it for small matrices (JacobiSVDImpl_ on the
// transposed copy; matrices with fewer than 25 rows never reach LAPACK): one-sided Jacobi on the columns of A with cv2's
// rotation formula, stop rule (|p| <= 10 eps sqrt(ab)) and sweep limit, then a selection sort by descending singular value.
// No sign normalisation happens anywhere, so the rows of Vt -- in particular the null vector in the last row -- carry the
// sign this exact sequence of rotations leaves them with; cv.triangulatePoints returns that row unscaled.
template <int N>
DVO_HDN void cv_jacobi_svd(const double* A /*NxN row-major*/, double* W, double* Vt /*NxN, rows = right singular vectors*/) {
    double At[N * N];
    for (int i = 0; i < N; ++i)
        for (int k = 0; k < N; ++k) At[i * N + k] = A[k * N + i];
    const double eps = DBL_EPSILON * 10;
    for (int i = 0; i < N; ++i) {
        double sd = 0;
        for (int k = 0; k < N; ++k) sd += At[i * N + k] * At[i * N + k];
        W[i] = sd;
        for (int k = 0; k < N; ++k) Vt[i * N + k] = (i == k) ? 1.0 : 0.0;
    }
    const int max_iter = N > 30 ? N : 30;
    for (int iter = 0; iter < max_iter; ++iter) {
        bool changed = false;
        for (int i = 0; i < N - 1; ++i)
            for (int j = i + 1; j < N; ++j) {
                double* Ai = At + i * N;
                double* Aj = At + j * N;
                double a = W[i], p = 0, b = W[j];
                for (int k = 0; k < N; ++k) p += Ai[k] * Aj[k];
                if (fabs(p) <= eps * sqrt(a * b)) continue;
                p *= 2;
                const double beta = a - b, gamma = hypot(p, beta);
                double c, s;
                if (beta < 0) {
                    const double delta = (gamma - beta) * 0.5;
                    s = sqrt(delta / gamma);
                    c = p / (gamma * s * 2);
                } else {
                    c = sqrt((gamma + beta) / (gamma * 2));
                    s = p / (gamma * c * 2);
                }
                a = b = 0;
                for (int k = 0; k < N; ++k) {
                    const double t0 = c * Ai[k] + s * Aj[k];
                    const double t1 = -s * Ai[k] + c * Aj[k];
                    Ai[k] = t0; Aj[k] = t1;
                    a += t0 * t0; b += t1 * t1;
                }
                W[i] = a; W[j] = b;
                changed = true;
                double* Vi = Vt + i * N;
                double* Vj = Vt + j * N;
                for (int k = 0; k < N; ++k) {
                    const double t0 = c * Vi[k] + s * Vj[k];
                    const double t1 = -s * Vi[k] + c * Vj[k];
                    Vi[k] = t0; Vj[k] = t1;
                }
            }
        if (!changed) break;
    }
    for (int i = 0; i < N; ++i) {
        double sd = 0;
        for (int k = 0; k < N; ++k) sd += At[i * N + k] * At[i * N + k];
        W[i] = sqrt(sd);
    }
    for (int i = 0; i < N - 1; ++i) {
        int j = i;
        for (int k = i + 1; k < N; ++k)
            if (W[j] < W[k]) j = k;
        if (i != j) {
            double tw = W[i]; W[i] = W[j]; W[j] = tw;
            for (int k = 0; k < N; ++k) { double tv = Vt[i * N + k]; Vt[i * N + k] = Vt[j * N + k]; Vt[j * N + k] = tv; }
        }
    }
}

// cv.triangulatePoints(P0, P1, x0, x1) for one correspondence: rows (x P[2] - P[0], y P[2] - P[1]) of both views, the last row
// of cv::SVD's Vt, unnormalised and with cv2's sign (visual_odometry_v3.py:265 measures distances on these raw vectors).
DVO_HDN void cv_triangulate_point(const double* P0 /*3x4*/, const double* P1, double x0, double y0, double x1, double y1, double* X) {
    double A[16], W[4], Vt[16];
    for (int k = 0; k < 4; ++k) {
        A[k] = x0 * P0[8 + k] - P0[k];
        A[4 + k] = y0 * P0[8 + k] - P0[4 + k];
        A[8 + k] = x1 * P1[8 + k] - P1[k];
        A[12 + k] = y1 * P1[8 + k] - P1[4 + k];
    }
    cv_jacobi_svd<4>(A, W, Vt);
    for (int k = 0; k < 4; ++k) X[k] = Vt[12 + k];
}

// cv2.decomposeEssentialMat: R1 = U W Vt, R2 = U Wt Vt, t = U[:,2] with det(U) = det(Vt) = +1.
DVO_HDN void decompose_essential(const double* E, double* R1, double* R2, double* t) {
    double A[9], V[9];
    for (int i = 0; i < 9; ++i) A[i] = E[i];
    jacobi_onesided<3>(A, V);
    // column norms = singular values; pick the two largest as (0,1)
    double n[3];
    for (int j = 0; j < 3; ++j) n[j] = sqrt(A[j] * A[j] + A[3 + j] * A[3 + j] + A[6 + j] * A[6 + j]);
    int o0 = 0, o1 = 1, o2 = 2;
    if (n[o0] < n[o1]) { int s = o0; o0 = o1; o1 = s; }
    if (n[o1] < n[o2]) { int s = o1; o1 = o2; o2 = s; }
    if (n[o0] < n[o1]) { int s = o0; o0 = o1; o1 = s; }
    double U[9], Vm[9];
    for (int i = 0; i < 3; ++i) {
        U[i * 3 + 0] = A[i * 3 + o0] / n[o0];
        U[i * 3 + 1] = A[i * 3 + o1] / n[o1];
        Vm[i * 3 + 0] = V[i * 3 + o0];
        Vm[i * 3 + 1] = V[i * 3 + o1];
    }
    // re-orthogonalise u1 against u0 (equal singular values make the pair well conditioned, this is cosmetic)
    // third columns by cross product => det = +1 for both
    U[2] = U[3] * U[7] - U[6] * U[4];
    U[5] = U[6] * U[1] - U[0] * U[7];
    U[8] = U[0] * U[4] - U[3] * U[1];
    Vm[2] = Vm[3] * Vm[7] - Vm[6] * Vm[4];
    Vm[5] = Vm[6] * Vm[1] - Vm[0] * Vm[7];
    Vm[8] = Vm[0] * Vm[4] - Vm[3] * Vm[1];
    // U W Vt with W = [[0,1,0],[-1,0,0],[0,0,1]]:  (U W) columns = (-u1, u0, u2);  U Wt columns = (u1, -u0, u2)
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double u0 = U[i * 3 + 0], u1 = U[i * 3 + 1], u2 = U[i * 3 + 2];
            double v0 = Vm[j * 3 + 0], v1 = Vm[j * 3 + 1], v2 = Vm[j * 3 + 2];
            R1[i * 3 + j] = -u1 * v0 + u0 * v1 + u2 * v2;
            R2[i * 3 + j] = u1 * v0 - u0 * v1 + u2 * v2;
        }
    t[0] = U[2]; t[1] = U[5]; t[2] = U[8];
}

// cv2.triangulatePoints for one correspondence, P0 = [I|0], P1 = [R|t]: right singular vector of the 4x4 DLT matrix
// belonging to the smallest singular value (sign arbitrary, the cheirality test is sign-invariant).
DVO_HDN void triangulate_one(const double* R, const double* t, double x1, double y1, double x2, double y2, double* X) {
    double A[16], V[16];
    A[0] = -1; A[1] = 0; A[2] = x1; A[3] = 0;
    A[4] = 0; A[5] = -1; A[6] = y1; A[7] = 0;
    for (int k = 0; k < 3; ++k) {
        A[8 + k] = x2 * R[6 + k] - R[k];
        A[12 + k] = y2 * R[6 + k] - R[3 + k];
    }
    A[11] = x2 * t[2] - t[0];
    A[15] = y2 * t[2] - t[1];
    jacobi_onesided<4>(A, V);
    int best = 0;
    double bn = 1e300;
    for (int j = 0; j < 4; ++j) {
        double nn = A[j] * A[j] + A[4 + j] * A[4 + j] + A[8 + j] * A[8 + j] + A[12 + j] * A[12 + j];
        if (nn < bn) { bn = nn; best = j; }
    }
    for (int i = 0; i < 4; ++i) X[i] = V[i * 4 + best];
}

// recoverPose's per-point cheirality test for one candidate (R, t); distance threshold as in cv2 (default 50).
DVO_HDN bool cheirality_ok(const double* R, const double* t, double x1, double y1, double x2, double y2, double dist) {
    double Q[4];
    triangulate_one(R, t, x1, y1, x2, y2, Q);
    bool m = (Q[2] * Q[3]) > 0;
    double qx = Q[0] / Q[3], qy = Q[1] / Q[3], qz = Q[2] / Q[3];
    m = m && (qz < dist);
    double z2 = R[6] * qx + R[7] * qy + R[8] * qz + t[2];   // row 2 of [R|t] * (qx,qy,qz,1)
    m = m && (z2 > 0) && (z2 < dist);
    return m;
}

// Both translation signs from ONE triangulation.  For P1' = [R | -t] the DLT matrix is A with its 4th column negated;
// one-sided Jacobi is sign-symmetric in a column (gamma, zeta, t and s flip sign together, every product is unchanged), so
// its null vector is +-(X0, X1, X2, -X3) and the tests become: q' = -q, z2' = -z2 (exact IEEE negations).
// Returns bit 0: (R, t) passes, bit 1: (R, -t) passes.
DVO_HDN int cheirality_pair(const double* R, const double* t, double x1, double y1, double x2, double y2, double dist) {
    double Q[4];
    triangulate_one(R, t, x1, y1, x2, y2, Q);
    const double s = Q[2] * Q[3];
    const double qx = Q[0] / Q[3], qy = Q[1] / Q[3], qz = Q[2] / Q[3];
    const double z2 = R[6] * qx + R[7] * qy + R[8] * qz + t[2];
    const bool mp = (s > 0) && (qz < dist) && (z2 > 0) && (z2 < dist);
    const bool mn = (-s > 0) && (-qz < dist) && (-z2 > 0) && (-z2 < dist);
    return (mp ? 1 : 0) | (mn ? 2 : 0);
}

// ---- Nister 5-point minimal solver -------------------------------------------------------------------------------
// cubic monomial order (first ten are eliminated): x3 y3 x2y xy2 x2z x2 y2z y2 xyz xy | xz2 xz x yz2 yz y z3 z2 z 1
// linear terms: 0:x 1:y 2:z 3:1 ; quadratic monomials: 0:x2 1:xy 2:xz 3:x 4:y2 5:yz 6:y 7:z2 8:z 9:1
DVO_HD int quad_index(int i, int j) {
    const int T[4][4] = {{0, 1, 2, 3}, {1, 4, 5, 6}, {2, 5, 7, 8}, {3, 6, 8, 9}};
    return T[i][j];
}
DVO_HD int cubic_index(int q, int l) {
    // rows: quadratic monomial, cols: linear term (x, y, z, 1)
    const int T[10][4] = {
        {0, 2, 4, 5},      // x2 * {x,y,z,1} = x3, x2y, x2z, x2
        {2, 3, 8, 9},      // xy          = x2y, xy2, xyz, xy
        {4, 8, 10, 11},    // xz          = x2z, xyz, xz2, xz
        {5, 9, 11, 12},    // x           = x2, xy, xz, x
        {3, 1, 6, 7},      // y2          = xy2, y3, y2z, y2
        {8, 6, 13, 14},    // yz          = xyz, y2z, yz2, yz
        {9, 7, 14, 15},    // y           = xy, y2, yz, y
        {10, 13, 16, 17},  // z2          = xz2, yz2, z3, z2
        {11, 14, 17, 18},  // z           = xz, yz, z2, z
        {12, 15, 18, 19},  // 1           = x, y, z, 1
    };
    return T[q][l];
}

DVO_HD void poly_mul11(const double* a, const double* b, double* out10, double sign) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) out10[quad_index(i, j)] += sign * a[i] * b[j];
}
DVO_HD void poly_mul21(const double* a10, const double* b, double* out20, double sign) {
#pragma unroll
    for (int i = 0; i < 10; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) out20[cubic_index(i, j)] += sign * a10[i] * b[j];
}

// Null-space basis of the 5x9 epipolar constraint matrix -- THE basis cv2's solver works in, which matters: the hidden-
// variable polynomial built from it is ill-conditioned for coplanar / clustered minimal samples, and how its roots move
// under rounding depends on the basis.  With any other orthonormal basis of the same null space (the Householder columns
// this code used before) 18 % of the minimal samples of a real sequence gave inlier counts different from cv2's; with
// cv2's basis 6 % do (the rest is rounding noise on samples where cv2 disagrees with itself under a 1-ulp change).
//
// cv2 runs cv::SVD::compute(Q, w, u, vt, FULL_UV) and takes rows 5..8 of vt (five-point.cpp).  Q has 5 rows, so those rows
// belong to no singular value: cv::SVD's Jacobi path fills them by Gram-Schmidt -- row i starts as the vector of +-1/9
// whose signs are bit 8 of successive draws of cv::RNG(0x12345678), is orthogonalised (twice) against every earlier row
// and normalised.  Mathematically: b_i = normalise(P s_i - sum_{j<i} (P s_i . b_j) b_j) with P the projector on the null
// space.  Here: Householder QR of Q^T gives H with H^T (null space) = span(e5..e8); the Gram-Schmidt runs on the four
// trailing coordinates of H^T s_i and the result is rotated back.
constexpr unsigned long long kCvSvdFillSigns = 0x74ec6fb84ull;   // bit (9 i + k): sign of component k of s_i is +
DVO_HD double cv_svd_fill_component(int i, int k) { return ((kCvSvdFillSigns >> (9 * i + k)) & 1ull) ? (1.0 / 9.0) : -(1.0 / 9.0); }

// basis[k*9 + c], k = 0..3.
DVO_HDN void null_space_5x9(const double* Q /*5x9*/, double* basis /*4x9*/) {
    double M[9 * 5];   // M = Q^T, column-major by constraint: M[r*5 + c]
    for (int r = 0; r < 9; ++r)
        for (int c = 0; c < 5; ++c) M[r * 5 + c] = Q[c * 9 + r];
    double vs[5][9];
    double betas[5];
    for (int k = 0; k < 5; ++k) {
        double norm = 0;
        for (int r = k; r < 9; ++r) norm += M[r * 5 + k] * M[r * 5 + k];
        norm = sqrt(norm);
        double alpha = M[k * 5 + k] > 0 ? -norm : norm;
        for (int r = 0; r < 9; ++r) vs[k][r] = 0;
        vs[k][k] = M[k * 5 + k] - alpha;
        for (int r = k + 1; r < 9; ++r) vs[k][r] = M[r * 5 + k];
        double vtv = 0;
        for (int r = k; r < 9; ++r) vtv += vs[k][r] * vs[k][r];
        betas[k] = vtv > 0 ? 2.0 / vtv : 0.0;
        for (int c = k; c < 5; ++c) {
            double dot = 0;
            for (int r = k; r < 9; ++r) dot += vs[k][r] * M[r * 5 + c];
            dot *= betas[k];
            for (int r = k; r < 9; ++r) M[r * 5 + c] -= dot * vs[k][r];
        }
    }
    // trailing coordinates of H^T s_i = H5 ... H1 s_i
    double u[4][4];
    for (int b = 0; b < 4; ++b) {
        double e[9];
        for (int r = 0; r < 9; ++r) e[r] = cv_svd_fill_component(b, r);
        for (int k = 0; k < 5; ++k) {
            double dot = 0;
            for (int r = k; r < 9; ++r) dot += vs[k][r] * e[r];
            dot *= betas[k];
            for (int r = k; r < 9; ++r) e[r] -= dot * vs[k][r];
        }
        for (int r = 0; r < 4; ++r) u[b][r] = e[5 + r];
    }
    // Gram-Schmidt (two passes, as cv::SVD does) and normalisation
    for (int b = 0; b < 4; ++b) {
        for (int pass = 0; pass < 2; ++pass)
            for (int j = 0; j < b; ++j) {
                double sd = 0;
                for (int r = 0; r < 4; ++r) sd += u[b][r] * u[j][r];
                for (int r = 0; r < 4; ++r) u[b][r] -= sd * u[j][r];
            }
        double nn = 0;
        for (int r = 0; r < 4; ++r) nn += u[b][r] * u[b][r];
        nn = nn > 0 ? 1.0 / sqrt(nn) : 0.0;
        for (int r = 0; r < 4; ++r) u[b][r] *= nn;
    }
    // back: b_i = H1 ... H5 [0; u_i]
    for (int b = 0; b < 4; ++b) {
        double e[9];
        for (int r = 0; r < 9; ++r) e[r] = r >= 5 ? u[b][r - 5] : 0.0;
        for (int k = 4; k >= 0; --k) {
            double dot = 0;
            for (int r = k; r < 9; ++r) dot += vs[k][r] * e[r];
            dot *= betas[k];
            for (int r = k; r < 9; ++r) e[r] -= dot * vs[k][r];
        }
        for (int r = 0; r < 9; ++r) basis[b * 9 + r] = e[r];
    }
}

struct cplx { double re, im; };
DVO_HD cplx cmul(cplx a, cplx b) { return cplx{a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
DVO_HD cplx cadd(cplx a, cplx b) { return cplx{a.re + b.re, a.im + b.im}; }
DVO_HD cplx csub(cplx a, cplx b) { return cplx{a.re - b.re, a.im - b.im}; }
DVO_HD cplx cinv(cplx a) { double d = a.re * a.re + a.im * a.im; return cplx{a.re / d, -a.im / d}; }

// All complex roots of c[0] z^deg + ... + c[deg] (c[0] != 0) by Aberth-Ehrlich iteration.  Returns iterations used.
DVO_HDN int aberth_roots(const double* c, int deg, cplx* z) {
    // Cauchy-style radius: 1 + max |c_k / c_0|
    double mx = 0;
    for (int k = 1; k <= deg; ++k) {
        double v = fabs(c[k] / c[0]);
        mx = v > mx ? v : mx;
    }
    // start on a circle of radius r = geometric mean magnitude of the roots, |c_deg/c_0|^(1/deg), kept in (1e-3, 1+mx)
    double r = pow(fabs(c[deg] / c[0]) + 1e-300, 1.0 / deg);
    if (!(r > 1e-3)) r = 1e-3;
    if (r > 1.0 + mx) r = 1.0 + mx;
    for (int k = 0; k < deg; ++k) {
        double ang = 2.0 * 3.14159265358979323846 * k / deg + 0.4;
        z[k] = cplx{r * cos(ang), r * sin(ang)};
    }
    int it = 0;
    for (; it < 40; ++it) {
        double worst = 0;
        for (int i = 0; i < deg; ++i) {
            cplx p{c[0], 0}, dp{0, 0};
            for (int k = 1; k <= deg; ++k) {
                dp = cadd(cmul(dp, z[i]), p);
                p = cadd(cmul(p, z[i]), cplx{c[k], 0});
            }
            double dpm = dp.re * dp.re + dp.im * dp.im;
            if (dpm == 0) { dp = cplx{1e-300, 0}; }
            cplx w = cmul(p, cinv(dp));
            cplx s{0, 0};
            for (int j = 0; j < deg; ++j)
                if (j != i) {
                    cplx d = csub(z[i], z[j]);
                    if (d.re == 0 && d.im == 0) d = cplx{1e-300, 0};
                    s = cadd(s, cinv(d));
                }
            cplx den = csub(cplx{1, 0}, cmul(w, s));
            if (den.re == 0 && den.im == 0) den = cplx{1e-300, 0};
            cplx step = cmul(w, cinv(den));
            z[i] = csub(z[i], step);
            double sm = fabs(step.re) + fabs(step.im);
            double zm = fabs(z[i].re) + fabs(z[i].im);
            double rel = sm / (zm > 1e-30 ? zm : 1e-30);
            worst = rel > worst ? rel : worst;
        }
        if (worst < 1e-14) { ++it; break; }
    }
    return it;
}

// x1, x2: 5 normalised correspondences (x,y interleaved, 10 doubles each).  Writes up to 10 models (9 doubles each,
// row-major, ||E||_F = 1, x2^T E x1 = 0) ordered by the hidden variable z ascending.  Returns the model count.
DVO_HDN int five_point_solve(const double* x1, const double* x2, double* models) {
    double Q[45];
    for (int i = 0; i < 5; ++i) {
        double a = x1[2 * i], b = x1[2 * i + 1], c = x2[2 * i], d = x2[2 * i + 1];
        double* q = Q + i * 9;
        q[0] = c * a; q[1] = c * b; q[2] = c; q[3] = d * a; q[4] = d * b; q[5] = d; q[6] = a; q[7] = b; q[8] = 1.0;
    }
    double EE[36];
    null_space_5x9(Q, EE);
    // entries as linear polynomials e[r*3+c][4] = (X, Y, Z, W)[rc]
    double e[9][4];
    for (int k = 0; k < 9; ++k)
        for (int b = 0; b < 4; ++b) e[k][b] = EE[b * 9 + k];
    double A[10][20];
    for (int i = 0; i < 10; ++i)
        for (int j = 0; j < 20; ++j) A[i][j] = 0;
    {
        double eet[3][3][10];
        for (int i = 0; i < 3; ++i)
            for (int j = i; j < 3; ++j) {
                for (int m = 0; m < 10; ++m) eet[i][j][m] = 0;
                for (int k = 0; k < 3; ++k) poly_mul11(e[i * 3 + k], e[j * 3 + k], eet[i][j], 1.0);
            }
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < i; ++j)
                for (int m = 0; m < 10; ++m) eet[i][j][m] = eet[j][i][m];
        for (int m = 0; m < 10; ++m) {
            double htr = 0.5 * (eet[0][0][m] + eet[1][1][m] + eet[2][2][m]);
            eet[0][0][m] -= htr; eet[1][1][m] -= htr; eet[2][2][m] -= htr;
        }
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j)
                for (int k = 0; k < 3; ++k) poly_mul21(eet[i][k], e[k * 3 + j], A[i * 3 + j], 1.0);
        // determinant
        double m2[10];
        const int cof[3][4] = {{4, 8, 5, 7}, {3, 8, 5, 6}, {3, 7, 4, 6}};   // e[a]*e[b] - e[c]*e[d]
        const double sg[3] = {1.0, -1.0, 1.0};
        for (int c0 = 0; c0 < 3; ++c0) {
            for (int m = 0; m < 10; ++m) m2[m] = 0;
            poly_mul11(e[cof[c0][0]], e[cof[c0][1]], m2, 1.0);
            poly_mul11(e[cof[c0][2]], e[cof[c0][3]], m2, -1.0);
            poly_mul21(m2, e[c0], A[9], sg[c0]);
        }
    }
    // Gauss-Jordan with partial pivoting on the left 10x10 block
    for (int col = 0; col < 10; ++col) {
        int piv = col;
        double pm = fabs(A[col][col]);
        for (int r = col + 1; r < 10; ++r)
            if (fabs(A[r][col]) > pm) { pm = fabs(A[r][col]); piv = r; }
        if (!(pm > 1e-300)) return 0;
        if (piv != col)
            for (int j = 0; j < 20; ++j) { double tmp = A[col][j]; A[col][j] = A[piv][j]; A[piv][j] = tmp; }
        double inv = 1.0 / A[col][col];
        for (int j = col; j < 20; ++j) A[col][j] *= inv;
        for (int r = 0; r < 10; ++r) {
            if (r == col) continue;
            double f = A[r][col];
            if (f == 0) continue;
            for (int j = col; j < 20; ++j) A[r][j] -= f * A[col][j];
        }
    }
    // B(z): rows from (x2z, x2), (y2z, y2), (xyz, xy):  <e> - z <f>; per row: px[4] (z^3..z^0), py[4], p1[5] (z^4..z^0)
    double B[3][13];
    for (int r = 0; r < 3; ++r) {
        const double* re = &A[4 + 2 * r][10];
        const double* rf = &A[5 + 2 * r][10];
        B[r][0] = -rf[0]; B[r][1] = re[0] - rf[1]; B[r][2] = re[1] - rf[2]; B[r][3] = re[2];
        B[r][4] = -rf[3]; B[r][5] = re[3] - rf[4]; B[r][6] = re[4] - rf[5]; B[r][7] = re[5];
        B[r][8] = -rf[6]; B[r][9] = re[6] - rf[7]; B[r][10] = re[7] - rf[8]; B[r][11] = re[8] - rf[9]; B[r][12] = re[9];
    }
    // det B(z), degree 10:  sum over the 6 permutations of (deg3 or deg3 or deg4 entries)
    double c[11];
    for (int k = 0; k < 11; ++k) c[k] = 0;
    {
        // helper lambdas are avoided for host/device portability: expand with small loops
        // term(a_row, a_off, a_len, b_row, b_off, b_len, c_row, c_off, c_len, sign)
        const int perm[6][3] = {{0, 1, 2}, {1, 2, 0}, {2, 0, 1}, {0, 2, 1}, {1, 0, 2}, {2, 1, 0}};  // column picked for rows 0,1,2
        const double psign[6] = {1, 1, 1, -1, -1, -1};
        const int off[3] = {0, 4, 8};
        const int len[3] = {4, 4, 5};
        for (int p = 0; p < 6; ++p) {
            int c0 = perm[p][0], c1 = perm[p][1], c2 = perm[p][2];
            double t2[9];
            for (int k = 0; k < 9; ++k) t2[k] = 0;
            for (int i = 0; i < len[c0]; ++i)
                for (int j = 0; j < len[c1]; ++j) t2[i + j] += B[0][off[c0] + i] * B[1][off[c1] + j];
            int l2 = len[c0] + len[c1] - 1;
            // total length l2 + len[c2] - 1 = 11 always (4+4+5-2)
            for (int i = 0; i < l2; ++i)
                for (int j = 0; j < len[c2]; ++j) c[i + j] += psign[p] * t2[i] * B[2][off[c2] + j];
        }
    }
    bool finite = true;
    for (int k = 0; k < 11; ++k) finite = finite && (fabs(c[k]) < 1e300);   // false for inf and NaN
    if (!finite || c[0] == 0.0) return 0;
    cplx roots[10];
    aberth_roots(c, 10, roots);
    double zr[10];
    int nr = 0;
    for (int i = 0; i < 10; ++i) {
        if (!(fabs(roots[i].im) <= 1e-10)) continue;
        double z = roots[i].re;
        for (int it = 0; it < 2; ++it) {   // Newton polish on the real polynomial
            double p = c[0], dp = 0;
            for (int k = 1; k <= 10; ++k) { dp = dp * z + p; p = p * z + c[k]; }
            if (dp != 0 && fabs(p / dp) < 1e300) z -= p / dp;
        }
        // insertion sort ascending
        int pos = nr;
        while (pos > 0 && zr[pos - 1] > z) { zr[pos] = zr[pos - 1]; --pos; }
        zr[pos] = z;
        ++nr;
    }
    int nm = 0;
    for (int i = 0; i < nr; ++i) {
        double z = zr[i], z2 = z * z, z3 = z2 * z, z4 = z3 * z;
        double bz[3][3];
        for (int r = 0; r < 3; ++r) {
            bz[r][0] = B[r][0] * z3 + B[r][1] * z2 + B[r][2] * z + B[r][3];
            bz[r][1] = B[r][4] * z3 + B[r][5] * z2 + B[r][6] * z + B[r][7];
            bz[r][2] = B[r][8] * z4 + B[r][9] * z3 + B[r][10] * z2 + B[r][11] * z + B[r][12];
        }
        // null vector of the (rank-2) 3x3: best-conditioned cross product of two rows
        double best[3] = {0, 0, 0}, bn = -1;
        for (int a = 0; a < 3; ++a) {
            int b = (a + 1) % 3;
            double v0 = bz[a][1] * bz[b][2] - bz[a][2] * bz[b][1];
            double v1 = bz[a][2] * bz[b][0] - bz[a][0] * bz[b][2];
            double v2 = bz[a][0] * bz[b][1] - bz[a][1] * bz[b][0];
            double nn = v0 * v0 + v1 * v1 + v2 * v2;
            if (nn > bn) { bn = nn; best[0] = v0; best[1] = v1; best[2] = v2; }
        }
        if (!(bn > 0)) continue;
        double inv = 1.0 / sqrt(bn);
        if (fabs(best[2] * inv) < 1e-10) continue;
        double x = best[0] / best[2], y = best[1] / best[2];
        double Ev[9], nrm = 0;
        for (int k = 0; k < 9; ++k) {
            Ev[k] = x * EE[k] + y * EE[9 + k] + z * EE[18 + k] + EE[27 + k];
            nrm += Ev[k] * Ev[k];
        }
        nrm = sqrt(nrm);
        if (!(nrm > 0) || !(nrm < 1e300)) continue;
        for (int k = 0; k < 9; ++k) models[nm * 9 + k] = Ev[k] / nrm;
        ++nm;
    }
    return nm;
}

}  // namespace dvo
