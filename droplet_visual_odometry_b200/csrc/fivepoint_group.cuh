// Nister 5-point minimal solver, one hypothesis per GROUP of 16 lanes (device only).
//
// Same mathematics as five_point_solve() in mathcore.cuh (which stays the scalar statement of the algorithm and is what
// tests/hostsim checks against cv2's own minimal solver), re-laid for the SIMT machine so that the critical path of one
// RANSAC hypothesis is ~1/20 of the scalar version's:
//   A  null space of the 5x9 constraint matrix: Householder QR with one matrix ROW per lane, dot products by shuffles
//   B  the ten cubic constraints: one constraint row (20 coefficients, registers) per lane
//   C  Gauss-Jordan with partial pivoting: rows stay in their lanes (implicit row exchange), pivot row broadcast by shuffles
//   D  hidden-variable determinant, degree-10 polynomial in z: every lane (registers)
//   E  all complex roots by Aberth-Ehrlich, one ROOT per lane (Jacobi-style simultaneous update)
//   F  real roots -> Newton polish -> back-substitution -> E, one root per lane; models ordered by z as in the scalar code
// Reference call this serves: cv.findEssentialMat (/root/reference/scripts/visual_odometry_v3.py:297-300); SURVEY A.9.
#pragma once
#include "mathcore.cuh"

namespace dvo {

constexpr int kGroupLanes = 16;

struct SolveScratch {        // per group, shared memory
    double EE[36];           // null-space basis, EE[b*9 + k]
    double eet[6][10];       // E E^T entries (00,01,02,11,12,22) as quadratic polynomials
    double red[10][10];      // right half of the reduced 10x20 system, indexed by pivot column
    double zr[10], zi[10];   // Aberth iterates, exchanged through shared memory
};

__device__ __forceinline__ double gshfl(unsigned gmask, double v, int src) { return __shfl_sync(gmask, v, src, kGroupLanes); }
__device__ __forceinline__ double gshfl_xor(unsigned gmask, double v, int m) { return __shfl_xor_sync(gmask, v, m, kGroupLanes); }
__device__ __forceinline__ double gsum(unsigned gmask, double v) {
#pragma unroll
    for (int m = 8; m > 0; m >>= 1) v += gshfl_xor(gmask, v, m);
    return v;
}

__device__ __forceinline__ cplx crcp(cplx a) {     // 1/a with one reciprocal
    const double inv = 1.0 / (a.re * a.re + a.im * a.im);
    return cplx{a.re * inv, -a.im * inv};
}

__device__ __forceinline__ int sym6(int i, int j) {   // index of eet[min][max]
    const int T[3][3] = {{0, 1, 2}, {1, 3, 4}, {2, 4, 5}};
    return T[i][j];
}

// x1, x2: the 5 normalised correspondences (every lane of the group passes the same values).  models: shared or global,
// room for 10 x 9 doubles.  Returns the model count (same value in every lane of the group).
__device__ int five_point_solve_group(const double* x1, const double* x2, SolveScratch& S, double* models, unsigned gmask) {
    const int gl = threadIdx.x & (kGroupLanes - 1);

    // ---- A: null space.  Lane r < 9 owns row r of M = Q^T (9 x 5): M[r][c] = Q[c][r].
    double Mr[5];
#pragma unroll
    for (int c = 0; c < 5; ++c) {
        const double a = x1[2 * c], b = x1[2 * c + 1], cc = x2[2 * c], d = x2[2 * c + 1];
        const double q[9] = {cc * a, cc * b, cc, d * a, d * b, d, a, b, 1.0};
        double v = 0.0;
#pragma unroll
        for (int r = 0; r < 9; ++r) v = (gl == r) ? q[r] : v;
        Mr[c] = v;
    }
    double vk[5], beta[5];   // this lane's component of Householder vector k
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        const double mine = (gl >= k && gl < 9) ? Mr[k] : 0.0;
        const double norm = sqrt(gsum(gmask, mine * mine));
        const double mkk = gshfl(gmask, Mr[k], k);
        const double alpha = mkk > 0 ? -norm : norm;
        vk[k] = (gl == k) ? (mkk - alpha) : mine;
        const double vtv = gsum(gmask, vk[k] * vk[k]);
        beta[k] = vtv > 0 ? 2.0 / vtv : 0.0;
#pragma unroll
        for (int c = 0; c < 5; ++c) {
            if (c < k) continue;
            const double dot = gsum(gmask, vk[k] * Mr[c]) * beta[k];
            Mr[c] -= dot * vk[k];
        }
    }
    // cv2's basis of that null space (see null_space_5x9 in mathcore.cuh): Gram-Schmidt of cv::SVD's fixed +-1/9 fill-in
    // vectors, done on the four trailing coordinates of H^T s_i (lanes 5..8), then rotated back with H.
    double u[4];
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        double e = gl < 9 ? cv_svd_fill_component(b, gl) : 0.0;
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            const double dot = gsum(gmask, vk[k] * e) * beta[k];
            e -= dot * vk[k];
        }
        u[b] = (gl >= 5 && gl < 9) ? e : 0.0;
    }
#pragma unroll
    for (int b = 0; b < 4; ++b) {
#pragma unroll
        for (int pass = 0; pass < 2; ++pass)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (j >= b) continue;
                const double sd = gsum(gmask, u[b] * u[j]);
                u[b] -= sd * u[j];
            }
        const double nn = gsum(gmask, u[b] * u[b]);
        u[b] *= nn > 0 ? 1.0 / sqrt(nn) : 0.0;
    }
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        double e = u[b];
#pragma unroll
        for (int k = 4; k >= 0; --k) {
            const double dot = gsum(gmask, vk[k] * e) * beta[k];
            e -= dot * vk[k];
        }
        if (gl < 9) S.EE[b * 9 + gl] = e;
    }
    __syncwarp(gmask);

    // ---- B: constraint rows.  e[k][b] = EE[b*9 + k].
    if (gl < 6) {
        const int I[6] = {0, 0, 0, 1, 1, 2}, J[6] = {0, 1, 2, 1, 2, 2};
        const int i = I[gl], j = J[gl];
        double acc[10];
#pragma unroll
        for (int m = 0; m < 10; ++m) acc[m] = 0.0;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            double a[4], b[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) { a[t] = S.EE[t * 9 + i * 3 + k]; b[t] = S.EE[t * 9 + j * 3 + k]; }
            poly_mul11(a, b, acc, 1.0);
        }
#pragma unroll
        for (int m = 0; m < 10; ++m) S.eet[gl][m] = acc[m];
    }
    __syncwarp(gmask);
    double row[20];
#pragma unroll
    for (int m = 0; m < 20; ++m) row[m] = 0.0;
    if (gl < 9) {
        const int i = gl / 3, j = gl - 3 * i;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            double p10[10], b[4];
            const int s = sym6(i, k);
#pragma unroll
            for (int m = 0; m < 10; ++m) {
                double v = S.eet[s][m];
                if (i == k) v -= 0.5 * (S.eet[0][m] + S.eet[3][m] + S.eet[5][m]);
                p10[m] = v;
            }
#pragma unroll
            for (int t = 0; t < 4; ++t) b[t] = S.EE[t * 9 + k * 3 + j];
            poly_mul21(p10, b, row, 1.0);
        }
    } else if (gl == 9) {
        const int cof[3][4] = {{4, 8, 5, 7}, {3, 8, 5, 6}, {3, 7, 4, 6}};   // e[a]*e[b] - e[c]*e[d]
        const double sg[3] = {1.0, -1.0, 1.0};
#pragma unroll
        for (int c0 = 0; c0 < 3; ++c0) {
            double m2[10], ea[4], eb[4], ec[4], ed[4], e0[4];
#pragma unroll
            for (int m = 0; m < 10; ++m) m2[m] = 0.0;
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                ea[t] = S.EE[t * 9 + cof[c0][0]]; eb[t] = S.EE[t * 9 + cof[c0][1]];
                ec[t] = S.EE[t * 9 + cof[c0][2]]; ed[t] = S.EE[t * 9 + cof[c0][3]];
                e0[t] = S.EE[t * 9 + c0];
            }
            poly_mul11(ea, eb, m2, 1.0);
            poly_mul11(ec, ed, m2, -1.0);
            poly_mul21(m2, e0, row, sg[c0]);
        }
    }

    // ---- C: Gauss-Jordan on the left 10x10 block, partial pivoting, rows stay in their lanes
    bool used = gl >= 10;
    int mycol = -1;
    bool singular = false;
#pragma unroll
    for (int col = 0; col < 10; ++col) {
        double best = used ? -1.0 : fabs(row[col]);
        int bl = gl;
#pragma unroll
        for (int m = 8; m > 0; m >>= 1) {
            const double ob = gshfl_xor(gmask, best, m);
            const int ol = __shfl_xor_sync(gmask, bl, m, kGroupLanes);
            if (ob > best || (ob == best && ol < bl)) { best = ob; bl = ol; }
        }
        if (!(best > 1e-300)) singular = true;
        if (gl == bl) {
            const double inv = 1.0 / row[col];
#pragma unroll
            for (int j = 0; j < 20; ++j)
                if (j >= col) row[j] *= inv;
            used = true;
            mycol = col;
        }
        const double f = row[col];
        const bool elim = (gl != bl) && gl < 10 && f != 0.0;
#pragma unroll
        for (int j = 0; j < 20; ++j) {
            if (j < col) continue;
            const double pj = gshfl(gmask, row[j], bl);
            if (elim) row[j] -= f * pj;
        }
    }
    if (singular) return 0;     // uniform in the group (every lane saw the same reductions)
    if (gl < 10) {
#pragma unroll
        for (int j = 0; j < 10; ++j) S.red[mycol][j] = row[10 + j];
    }
    __syncwarp(gmask);

    // ---- D: B(z) and det B(z) (degree 10), every lane
    double B[3][13];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const double* re = S.red[4 + 2 * r];
        const double* rf = S.red[5 + 2 * r];
        B[r][0] = -rf[0]; B[r][1] = re[0] - rf[1]; B[r][2] = re[1] - rf[2]; B[r][3] = re[2];
        B[r][4] = -rf[3]; B[r][5] = re[3] - rf[4]; B[r][6] = re[4] - rf[5]; B[r][7] = re[5];
        B[r][8] = -rf[6]; B[r][9] = re[6] - rf[7]; B[r][10] = re[7] - rf[8]; B[r][11] = re[8] - rf[9]; B[r][12] = re[9];
    }
    double c[11];
#pragma unroll
    for (int k = 0; k < 11; ++k) c[k] = 0.0;
    {
        const int perm[6][3] = {{0, 1, 2}, {1, 2, 0}, {2, 0, 1}, {0, 2, 1}, {1, 0, 2}, {2, 1, 0}};
        const double psign[6] = {1, 1, 1, -1, -1, -1};
        const int off[3] = {0, 4, 8};
        const int len[3] = {4, 4, 5};
#pragma unroll
        for (int p = 0; p < 6; ++p) {
            const int c0 = perm[p][0], c1 = perm[p][1], c2 = perm[p][2];
            double t2[9];
#pragma unroll
            for (int k = 0; k < 9; ++k) t2[k] = 0.0;
#pragma unroll
            for (int i = 0; i < 5; ++i)
#pragma unroll
                for (int j = 0; j < 5; ++j)
                    if (i < len[c0] && j < len[c1]) t2[i + j] += B[0][off[c0] + i] * B[1][off[c1] + j];
#pragma unroll
            for (int i = 0; i < 9; ++i)
#pragma unroll
                for (int j = 0; j < 5; ++j)
                    if (i < len[c0] + len[c1] - 1 && j < len[c2]) c[i + j] += psign[p] * t2[i] * B[2][off[c2] + j];
        }
    }
    bool finite = true;
#pragma unroll
    for (int k = 0; k < 11; ++k) finite = finite && (fabs(c[k]) < 1e300);
    if (!finite || c[0] == 0.0) return 0;     // uniform: every lane computed the same c[]

    // ---- E: Aberth-Ehrlich, root gl per lane (lanes >= 10 carry a dummy and are never read)
    cplx z;
    {
        double mx = 0;
#pragma unroll
        for (int k = 1; k <= 10; ++k) {
            const double v = fabs(c[k] / c[0]);
            mx = v > mx ? v : mx;
        }
        double r = pow(fabs(c[10] / c[0]) + 1e-300, 0.1);
        if (!(r > 1e-3)) r = 1e-3;
        if (r > 1.0 + mx) r = 1.0 + mx;
        const double ang = 2.0 * 3.14159265358979323846 * (gl < 10 ? gl : 0) / 10 + 0.4;
        z = cplx{r * cos(ang), r * sin(ang)};
    }
    // A root is frozen once its step is below 1e-13 relative or |p(z)| is within the rounding noise of Horner's rule
    // (|p| <= 64 eps * sum |c_k||z|^k); the loop ends when all ten are frozen.  F re-polishes the real ones.
    bool frozen = gl >= 10;
    for (int it = 0; it < 40; ++it) {
        if (gl < 10) { S.zr[gl] = z.re; S.zi[gl] = z.im; }
        __syncwarp(gmask);
        cplx p{c[0], 0}, dp{0, 0};
        const double az = sqrt(z.re * z.re + z.im * z.im);
        double pb = fabs(c[0]);
#pragma unroll
        for (int k = 1; k <= 10; ++k) {
            dp = cadd(cmul(dp, z), p);
            p = cadd(cmul(p, z), cplx{c[k], 0});
            pb = pb * az + fabs(c[k]);
        }
        if (fabs(p.re) + fabs(p.im) <= 64.0 * DBL_EPSILON * pb) frozen = true;
        if (dp.re * dp.re + dp.im * dp.im == 0) dp = cplx{1e-300, 0};
        const cplx w = cmul(p, crcp(dp));
        cplx s{0, 0};
#pragma unroll
        for (int j = 0; j < 10; ++j) {
            const cplx zj{S.zr[j], S.zi[j]};
            if (j != gl) {
                cplx d = csub(z, zj);
                if (d.re == 0 && d.im == 0) d = cplx{1e-300, 0};
                s = cadd(s, crcp(d));
            }
        }
        cplx den = csub(cplx{1, 0}, cmul(w, s));
        if (den.re == 0 && den.im == 0) den = cplx{1e-300, 0};
        const cplx step = cmul(w, crcp(den));
        if (!frozen) {
            z = csub(z, step);
            const double sm = fabs(step.re) + fabs(step.im);
            const double zm = fabs(z.re) + fabs(z.im);
            if (sm <= 1e-13 * (zm > 1e-30 ? zm : 1e-30)) frozen = true;     // false for NaN
        }
        const unsigned live = __ballot_sync(gmask, !frozen) & gmask;
        __syncwarp(gmask);                  // all reads of S.z* done before the next iteration overwrites them
        if (live == 0) break;               // uniform in the group
    }

    // ---- F: real roots -> models
    bool valid = gl < 10 && (fabs(z.im) <= 1e-10);
    double zr = z.re;
    double Ev[9];
    if (valid) {
#pragma unroll
        for (int it = 0; it < 2; ++it) {
            double p = c[0], dp = 0;
#pragma unroll
            for (int k = 1; k <= 10; ++k) { dp = dp * zr + p; p = p * zr + c[k]; }
            if (dp != 0 && fabs(p / dp) < 1e300) zr -= p / dp;
        }
        const double z2 = zr * zr, z3 = z2 * zr, z4 = z3 * zr;
        double bz[3][3];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            bz[r][0] = B[r][0] * z3 + B[r][1] * z2 + B[r][2] * zr + B[r][3];
            bz[r][1] = B[r][4] * z3 + B[r][5] * z2 + B[r][6] * zr + B[r][7];
            bz[r][2] = B[r][8] * z4 + B[r][9] * z3 + B[r][10] * z2 + B[r][11] * zr + B[r][12];
        }
        double best[3] = {0, 0, 0}, bn = -1;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const int b = (a + 1) % 3;
            const double v0 = bz[a][1] * bz[b][2] - bz[a][2] * bz[b][1];
            const double v1 = bz[a][2] * bz[b][0] - bz[a][0] * bz[b][2];
            const double v2 = bz[a][0] * bz[b][1] - bz[a][1] * bz[b][0];
            const double nn = v0 * v0 + v1 * v1 + v2 * v2;
            if (nn > bn) { bn = nn; best[0] = v0; best[1] = v1; best[2] = v2; }
        }
        valid = bn > 0;
        if (valid) {
            const double inv = 1.0 / sqrt(bn);
            valid = !(fabs(best[2] * inv) < 1e-10);
        }
        if (valid) {
            const double x = best[0] / best[2], y = best[1] / best[2];
            double nrm = 0;
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                Ev[k] = x * S.EE[k] + y * S.EE[9 + k] + zr * S.EE[18 + k] + S.EE[27 + k];
                nrm += Ev[k] * Ev[k];
            }
            nrm = sqrt(nrm);
            valid = (nrm > 0) && (nrm < 1e300);
            if (valid) {
#pragma unroll
                for (int k = 0; k < 9; ++k) Ev[k] /= nrm;
            }
        }
    }
    // The scalar code sorts ALL real roots by z (stable) and then drops the degenerate ones, so among the surviving
    // models the order is (z, root index) ascending.
    int rank = 0, count = 0;
#pragma unroll
    for (int j = 0; j < 10; ++j) {
        const double zj = gshfl(gmask, zr, j);
        const int vj = __shfl_sync(gmask, (int)valid, j, kGroupLanes);
        if (vj) {
            ++count;
            if (zj < zr || (zj == zr && j < gl)) ++rank;
        }
    }
    if (valid) {
#pragma unroll
        for (int k = 0; k < 9; ++k) models[rank * 9 + k] = Ev[k];
    }
    return count;
}

}  // namespace dvo
