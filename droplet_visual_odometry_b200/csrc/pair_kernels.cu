// Pair-level kernels on sm_100a: brute-force Hamming matching, match ordering, essential-matrix RANSAC, recoverPose.
// Replaces bf.match + sorted (/root/reference/scripts/visual_odometry_v3.py:219-221), cv.KeyPoint_convert (:355,:358),
// cv.findEssentialMat (:297-300) and cv.recoverPose (:303-306).  Contract: SURVEY.md Appendix A.8-A.10.
//
//   k_nn            A.8  tiled XOR+POPC nearest neighbour (both directions in one launch; 2-NN distance for the ratio test)
//   k_match_sort    A.8  cross-check / ratio+reverse check, bitonic sort on (distance, queryIdx), point gather, K-normalise
//   k_ransac        A.9/A.10  one CTA per pair: cv::RNG sample stream, 5-point solves (one per 16-lane group), Sampson scoring,
//                        cv2's strict-'>' update + adaptive stop replayed in order; final mask, SVD of E -> R1, R2, t
//   k_cheirality    A.10 per point x 4 candidates DLT triangulation + cheirality votes
//   k_pose_final    A.10 '>=' cascade, masks, dvo_pose record
#include "dvo_internal.cuh"
#include "mathcore.cuh"
#include "fivepoint_group.cuh"
#include <cooperative_groups.h>
#include <algorithm>
#include <cstdlib>

namespace dvo {

// Shared-memory counter update as one RED instruction: the callers have already reduced over the warp, so the
// warp-aggregation sequence the compiler wraps around atomicAdd() is overhead.
__device__ __forceinline__ void smem_red_add(int* p, int v) {
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(p)), "r"(v) : "memory");
}

void debug_sync(const char* name, cudaStream_t st);

// ransacState layout (ints)
enum { RS_MAXGOOD = 0, RS_NITERS = 1, RS_DONE = 2, RS_BESTITER = 3, RS_BESTMODEL = 4, RS_RNG_LO = 5, RS_RNG_HI = 6, RS_HASBEST = 7 };

// ================================================================================================ matching
// Every Hamming distance is computed once: the thread that owns query i keeps the row minimum (and runner-up) in
// registers, and the same distance feeds the column minimum through one redux.sync per (warp, train descriptor) on the
// packed key (distance << 16 | queryIdx) -- so ties go to the lowest index on both sides, as cv2's batchDistance does.
// Column keys are merged warp -> CTA (shared atomicMin) -> pair (global atomicMin); the launcher presets them to ~0.
// kSplit (cross-check mode, no runner-up needed): the train descriptors are dealt out in 128-column tiles to gridDim.y
// CTAs per query block, which multiplies the resident warps; row minima are then merged like the column minima, through
// atomicMin on (distance << 16 | trainIdx) in nnIdx[dir 0].
// 256-bit Hamming distance.  POPC runs on the XU pipe (16 lanes/clk/SM), which bounded this kernel (ncu: 92 % XU with 8
// POPCs per distance), so two carry-save adders on the ALU pipe replace two of them:
// popc(a)+popc(b)+popc(c) = popc(a^b^c) + 2 popc(maj(a,b,c)).  Measured on B200 (148 pairs x 2000^2 distances): 8 POPC
// 1.13 ms, 6 POPC (this) 0.90 ms, 5 POPC 0.90 ms, 4 POPC (full Harley-Seal tree) 1.10 ms -- issue-bound beyond this point.
__device__ __forceinline__ int hamming256(const uint32_t* q, const uint4& t0, const uint4& t1) {
    const uint32_t x0 = q[0] ^ t0.x, x1 = q[1] ^ t0.y, x2 = q[2] ^ t0.z, x3 = q[3] ^ t0.w;
    const uint32_t x4 = q[4] ^ t1.x, x5 = q[5] ^ t1.y, x6 = q[6] ^ t1.z, x7 = q[7] ^ t1.w;
    const uint32_t sa = x0 ^ x1 ^ x2, ca = (x0 & x1) | (x0 & x2) | (x1 & x2);
    const uint32_t sb = x3 ^ x4 ^ x5, cb = (x3 & x4) | (x3 & x5) | (x4 & x5);
    return (__popc(sa) + __popc(sb)) + 2 * (__popc(ca) + __popc(cb)) + (__popc(x6) + __popc(x7));
}

template <bool kSplit>
__global__ void __launch_bounds__(128) k_nn(OrbGeom og, OrbBuffers ob, PairGeom pg, PairBuffers pb, int slotA0, int pair0) {
    __shared__ __align__(16) uint32_t tile[128 * 8];
    __shared__ uint32_t colmin[128];
    const int pi = blockIdx.z;
    const int slotA = slotA0 + pi, slotB = slotA + 1;
    const int pair = pair0 + pi;
    const int nA = min(ob.featCount[slotA], pg.maxkp), nB = min(ob.featCount[slotB], pg.maxkp);
    if (blockIdx.x * 128 >= nA) return;
    const int i = blockIdx.x * 128 + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const bool valid = i < nA;
    const uint32_t* dA = reinterpret_cast<const uint32_t*>(ob.featDesc + (size_t)slotA * og.maxkp * 32);
    const uint32_t* dB = reinterpret_cast<const uint32_t*>(ob.featDesc + (size_t)slotB * og.maxkp * 32);
    uint32_t* colKey = reinterpret_cast<uint32_t*>(pb.nnIdx + ((size_t)pair * 2 + 1) * pg.maxkp);
    uint32_t q[8];
#pragma unroll
    for (int w = 0; w < 8; ++w) q[w] = valid ? dA[(size_t)i * 8 + w] : 0u;
    int best = 0x7fffffff, bestIdx = -1, second = 0x7fffffff;
    const uint32_t keyLow = valid ? (uint32_t)i : 0xFFFFFFFFu;     // rows past nA never win a column
    for (int j0 = blockIdx.y * 128; j0 < nB; j0 += 128 * gridDim.y) {
        __syncthreads();
        // stage 128 train descriptors: 256 uint4
        for (int v = threadIdx.x; v < 256; v += 128) {
            int j = j0 + (v >> 1);
            uint4 val = make_uint4(0, 0, 0, 0);
            if (j < nB) val = reinterpret_cast<const uint4*>(dB)[(size_t)j * 2 + (v & 1)];
            reinterpret_cast<uint4*>(tile)[v] = val;
        }
        colmin[threadIdx.x] = 0xFFFFFFFFu;
        __syncthreads();
        const int lim = min(128, nB - j0);
        for (int jg = 0; jg < lim; jg += 32) {
            uint32_t mycol = 0xFFFFFFFFu;    // lane l: this warp's minimum for column jg + l
#pragma unroll 4
            for (int jj = 0; jj < 32; ++jj) {
                const int j = jg + jj;       // columns >= lim are zero-filled tile rows: computed, never used
                const uint4 t0 = reinterpret_cast<const uint4*>(tile)[j * 2];
                const uint4 t1 = reinterpret_cast<const uint4*>(tile)[j * 2 + 1];
                const int d = hamming256(q, t0, t1);
                if (j < lim) {
                    if (d < best) { second = best; best = d; bestIdx = j0 + j; }
                    else if (d < second) second = d;
                }
                const uint32_t m = __reduce_min_sync(0xffffffffu, ((uint32_t)d << 16) | keyLow);
                if (lane == jj) mycol = m;
            }
            if (jg + lane < lim) atomicMin(&colmin[jg + lane], mycol);
        }
        __syncthreads();
        if (threadIdx.x < lim) atomicMin(&colKey[j0 + threadIdx.x], colmin[threadIdx.x]);
    }
    if (valid) {
        size_t o = (size_t)pair * 2 * pg.maxkp + i;
        if (kSplit) {
            if (bestIdx >= 0) atomicMin(reinterpret_cast<uint32_t*>(pb.nnIdx) + o, ((uint32_t)best << 16) | (uint32_t)bestIdx);
        } else {
            pb.nnIdx[o] = bestIdx;
            pb.nnDist[o] = best;
            pb.nn2Dist[(size_t)pair * pg.maxkp + i] = second;
        }
    }
}

__global__ void __launch_bounds__(1024) k_match_sort(OrbGeom og, OrbBuffers ob, PairGeom pg, PairBuffers pb, int slotA0, int pair0,
                                                     double fx, double fy, double cx, double cy) {
    extern __shared__ uint32_t keys[];
    __shared__ int s_count;
    const int pi = blockIdx.x;
    const int pair = pair0 + pi;
    const int slotA = slotA0 + pi, slotB = slotA + 1;
    const int nA = min(ob.featCount[slotA], pg.maxkp), nB = min(ob.featCount[slotB], pg.maxkp);
    const int tid = threadIdx.x;
    const int* fwd = pb.nnIdx + ((size_t)pair * 2 + 0) * pg.maxkp;
    const int* fwdD = pb.nnDist + ((size_t)pair * 2 + 0) * pg.maxkp;
    const uint32_t* bwdKey = reinterpret_cast<const uint32_t*>(pb.nnIdx + ((size_t)pair * 2 + 1) * pg.maxkp);   // d << 16 | queryIdx
    const int* d2 = pb.nn2Dist + (size_t)pair * pg.maxkp;
    if (tid == 0) s_count = 0;
    __syncthreads();
    int local = 0;
    for (int i = tid; i < pg.sortCap; i += 1024) {
        uint32_t key = 0xFFFFFFFFu;
        if (i < nA && nB > 0) {
            const bool packed = pg.matcher == DVO_MATCH_CROSSCHECK;     // k_nn<true> leaves (distance << 16 | trainIdx)
            const uint32_t fk = (uint32_t)fwd[i];
            const int j = packed ? (fk == 0xFFFFFFFFu ? -1 : (int)(fk & 0xFFFFu)) : fwd[i];
            const int dist = packed ? (int)(fk >> 16) : fwdD[i];
            bool ok = j >= 0 && (int)(bwdKey[j] & 0xFFFFu) == i;
            if (pg.matcher == DVO_MATCH_KNN_RATIO) {
                // knnMatch returns < 2 neighbours when the train set has one descriptor: the reference's unpacking needs two
                ok = ok && nB >= 2 && ((double)dist < (double)pg.ratio * (double)d2[i]);
            }
            if (ok) { key = ((uint32_t)dist << 16) | (uint32_t)i; ++local; }
        }
        keys[i] = key;
    }
    atomicAdd(&s_count, local);
    __syncthreads();
    // bitonic sort ascending
    for (int k = 2; k <= pg.sortCap; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < pg.sortCap; i += 1024) {
                int ixj = i ^ j;
                if (ixj > i) {
                    uint32_t a = keys[i], b = keys[ixj];
                    bool up = (i & k) == 0;
                    if ((a > b) == up) { keys[i] = b; keys[ixj] = a; }
                }
            }
            __syncthreads();
        }
    }
    const int M = s_count;
    if (tid == 0) pb.matchCount[pair] = M;
    const float* ptA = ob.featPt + (size_t)slotA * og.maxkp * 2;
    const float* ptB = ob.featPt + (size_t)slotB * og.maxkp * 2;
    const size_t o = (size_t)pair * pg.maxkp;
    for (int r = tid; r < M; r += 1024) {
        uint32_t key = keys[r];
        int i = key & 0xFFFF, d = key >> 16;
        int j = pg.matcher == DVO_MATCH_CROSSCHECK ? (int)((uint32_t)fwd[i] & 0xFFFFu) : fwd[i];
        pb.matches[(o + r) * 3 + 0] = i;
        pb.matches[(o + r) * 3 + 1] = j;
        pb.matches[(o + r) * 3 + 2] = d;
        float ax = ptA[i * 2], ay = ptA[i * 2 + 1], bx = ptB[j * 2], by = ptB[j * 2 + 1];
        pb.ptsPrev[(o + r) * 2] = ax; pb.ptsPrev[(o + r) * 2 + 1] = ay;
        pb.ptsCur[(o + r) * 2] = bx; pb.ptsCur[(o + r) * 2 + 1] = by;
        double* np_ = pb.normPts + (o + r) * 4;
        np_[0] = cv_normalize_coord(ax, fx, cx); np_[1] = cv_normalize_coord(ay, fy, cy);
        np_[2] = cv_normalize_coord(bx, fx, cx); np_[3] = cv_normalize_coord(by, fy, cy);
    }
    if (tid == 0) {
        int* rs = pb.ransacState + pair * 8;
        rs[RS_MAXGOOD] = 0;
        rs[RS_NITERS] = pg.maxIters;
        rs[RS_DONE] = (M <= 5) ? 1 : 0;      // < 5: no model; == 5: cv2 returns the stacked minimal solutions (not a 3x3 E)
        rs[RS_BESTITER] = -1;
        rs[RS_BESTMODEL] = -1;
        rs[RS_RNG_LO] = (int)0xFFFFFFFFu;
        rs[RS_RNG_HI] = (int)0xFFFFFFFFu;
        rs[RS_HASBEST] = 0;
    }
}

// ================================================================================================ RANSAC
// One CTA per frame pair runs cv2's whole findEssentialMat loop: per chunk of kRansacGroups iterations
//   thread 0: cv::RNG sample stream (5 distinct positions, single-index redraw)
//   16-lane groups: one 5-point solve each (fivepoint_group.cuh), models into shared memory
//   warps: Sampson error of every match against every model, warp-reduced inlier counts
//   thread 0: cv2's strict-'>' update + adaptive stop rule, in iteration order (speculated iterations past the stop are
//             discarded, so E, mask and iteration count are the ones the sequential loop produces)
// then the final mask and the SVD of E for recoverPose.  No launch is spent on pairs that have already stopped.
constexpr int kRansacGroups = 16;
constexpr int kWindowedMaxPairs = 8;      // at most this many pairs: windowed whole-GPU speculation instead of one CTA per pair
constexpr int kRansacThreads = kRansacGroups * kGroupLanes;

struct RansacShared {
    SolveScratch scratch[kRansacGroups];
    double models[kRansacGroups][kMaxModels][9];
    double bestE[9];
    unsigned long long rng;
    int samples[kRansacGroups][5];
    int modelCount[kRansacGroups];
    int modelGood[kRansacGroups][kMaxModels];
    int partial[kRansacGroups][kMaxModels];      // cluster mode: this CTA's share of the inlier counts
    int maxGood, niters, done, it0, bestIter, bestModel, hasBest, cnt;
};

// Final mask of the winning model and SVD(E) -> R1, R2, t (first half of recoverPose); whole CTA, any block size.
// bestE: 9 doubles in shared memory; s_cnt: a shared int.
__device__ void ransac_finish(const PairGeom& pg, const PairBuffers& pb, PoseScratch* ps, int pair, int M, const double* bestE,
                              int hasBest, float t32, int* s_cnt) {
    const int tid = threadIdx.x, lane = tid & 31;
    const double T = (double)t32, tlo = T * (1.0 - 1e-6), thi = T * (1.0 + 1e-6);
    PoseScratch& sc = ps[pair];
    uint8_t* mask = pb.ransacMask + (size_t)pair * pg.maxkp;
    const double4* np4 = reinterpret_cast<const double4*>(pb.normPts + (size_t)pair * pg.maxkp * 4);
    if (tid == 0) {
        *s_cnt = 0;
        for (int k = 0; k < 4; ++k) sc.good[k] = 0;
    }
    __syncthreads();
    if (!hasBest) {
        for (int i = tid; i < M; i += blockDim.x) mask[i] = 0;
        if (tid == 0) sc.nInl = 0;
        return;
    }
    double E[9];
#pragma unroll
    for (int j = 0; j < 9; ++j) E[j] = bestE[j];
    if (tid < 9) pb.bestE[pair * 9 + tid] = E[tid];
    int local = 0;
    for (int i = tid; i < M; i += blockDim.x) {
        const double4 p = np4[i];
        const int in = sampson_inlier(E, p.x, p.y, p.z, p.w, t32, tlo, thi) ? 1 : 0;
        mask[i] = (uint8_t)in;
        local += in;
    }
    local = __reduce_add_sync(0xffffffffu, local);
    if (lane == 0) atomicAdd(s_cnt, local);
    __syncthreads();
    if (tid == 0) {
        sc.nInl = *s_cnt;
        decompose_essential(E, sc.R1, sc.R2, sc.t);
    }
}

// kCluster: a thread-block cluster of up to 8 CTAs works on ONE pair (few pairs, many correspondences -- BASELINE configs[4]).
// Every CTA runs the identical, deterministic control flow (RNG, solves, replay), so the loop state stays in lock-step
// without communication; only the Sampson scoring is split -- each CTA scores a slice of the correspondences -- and the
// per-model counts are summed across the cluster through distributed shared memory between two cluster barriers.
template <bool kCluster>
__global__ void __launch_bounds__(kRansacThreads, 2) k_ransac(PairGeom pg, PairBuffers pb, PoseScratch* ps, int pair0, float t32) {
    __shared__ RansacShared sh;
    namespace cg = cooperative_groups;
    unsigned crank = 0, csize = 1;
    if (kCluster) {
        cg::cluster_group cluster = cg::this_cluster();
        crank = cluster.block_rank();
        csize = cluster.num_blocks();
    }
    const int pair = pair0 + (kCluster ? blockIdx.x / csize : blockIdx.x);
    int* rs = pb.ransacState + pair * 8;
    const int M = pb.matchCount[pair];
    const int tid = threadIdx.x, lane = tid & 31;
    const int grp = tid / kGroupLanes, gl = tid & (kGroupLanes - 1);
    const unsigned gmask = 0xFFFFu << (lane & 16);
    const double T = (double)t32, tlo = T * (1.0 - 1e-6), thi = T * (1.0 + 1e-6);
    if (tid == 0) {
        sh.maxGood = 0;
        sh.niters = pg.maxIters;
        sh.done = rs[RS_DONE];
        sh.it0 = 0;
        sh.bestIter = -1; sh.bestModel = -1; sh.hasBest = 0;
        sh.rng = ((unsigned long long)(uint32_t)rs[RS_RNG_HI] << 32) | (uint32_t)rs[RS_RNG_LO];
    }
    __syncthreads();
    const double* np_ = pb.normPts + (size_t)pair * pg.maxkp * 4;
    const double4* np4 = reinterpret_cast<const double4*>(np_);
    while (!sh.done) {
        const int it0 = sh.it0;
        const int nIt = min(kRansacGroups, min(sh.niters, pg.maxIters) - it0);
        __syncthreads();                     // everyone has read the loop state before thread 0 touches it again
        if (tid == 0) {
            uint64_t state = sh.rng;
            for (int it = 0; it < nIt; ++it) {
                int idx[5];
                int i = 0;
                while (i < 5) {
                    const int v = (int)(cvrng_next(state) % (uint32_t)M);
                    bool dup = false;
                    for (int j = 0; j < i; ++j) dup = dup || (idx[j] == v);
                    if (dup) continue;
                    idx[i++] = v;
                }
                for (int k = 0; k < 5; ++k) sh.samples[it][k] = idx[k];
            }
            sh.rng = state;
        }
        __syncthreads();
        if (grp < nIt) {
            double x1[10], x2[10];
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                const double4 p = np4[sh.samples[grp][k]];
                x1[2 * k] = p.x; x1[2 * k + 1] = p.y;
                x2[2 * k] = p.z; x2[2 * k + 1] = p.w;
            }
            const int n = five_point_solve_group(x1, x2, sh.scratch[grp], &sh.models[grp][0][0], gmask);
            if (gl == 0) sh.modelCount[grp] = n;
            if (gl < 5 && crank == 0) pb.samples[((size_t)pair * pg.maxIters + it0 + grp) * 5 + gl] = sh.samples[grp][gl];
        }
        __syncthreads();
        // Scoring: matches live in registers (two per thread per sweep), the models are broadcast from shared memory,
        // so each correspondence is read once per chunk and the loop is bound by the FP64 pipe, not by load latency.
        int (*acc)[kMaxModels] = kCluster ? sh.partial : sh.modelGood;
        for (int i = tid; i < kRansacGroups * kMaxModels; i += kRansacThreads) (&acc[0][0])[i] = 0;
        __syncthreads();
        const int per = (M + (int)csize - 1) / (int)csize;
        const int mBeg = min(M, (int)crank * per), mEnd = min(M, mBeg + per);
        for (int base = mBeg; base < mEnd; base += 2 * kRansacThreads) {
            const int i0 = base + tid, i1 = base + kRansacThreads + tid;
            const bool v0 = i0 < mEnd, v1 = i1 < mEnd;
            const double4 p0 = v0 ? np4[i0] : make_double4(0, 0, 0, 0);
            const double4 p1 = v1 ? np4[i1] : make_double4(0, 0, 0, 0);
            for (int h = 0; h < nIt; ++h) {
                const int nm = sh.modelCount[h];
                for (int k = 0; k < nm; ++k) {
                    double E[9];
#pragma unroll
                    for (int j = 0; j < 9; ++j) E[j] = sh.models[h][k][j];
                    int good = (v0 && sampson_inlier(E, p0.x, p0.y, p0.z, p0.w, t32, tlo, thi)) ? 1 : 0;
                    good += (v1 && sampson_inlier(E, p1.x, p1.y, p1.z, p1.w, t32, tlo, thi)) ? 1 : 0;
                    good = __reduce_add_sync(0xffffffffu, good);
                    if (lane == 0 && good) smem_red_add(&acc[h][k], good);
                }
            }
        }
        if (kCluster) {
            cg::cluster_group cluster = cg::this_cluster();
            cluster.sync();                               // every CTA's partial counts are complete
            for (int i = tid; i < kRansacGroups * kMaxModels; i += kRansacThreads) {
                int tot = 0;
                for (unsigned r = 0; r < csize; ++r) tot += cluster.map_shared_rank(&sh.partial[0][0], r)[i];
                (&sh.modelGood[0][0])[i] = tot;
            }
            cluster.sync();                               // nobody still reads a partial that the next chunk will reset
        } else {
            __syncthreads();
        }
        if (tid == 0) {
            int maxGood = sh.maxGood, niters = sh.niters;
            int it = it0;
            for (; it < it0 + nIt; ++it) {
                if (it >= niters) break;
                const int h = it - it0, nm = sh.modelCount[h];
                for (int k = 0; k < nm; ++k) {
                    const int good = sh.modelGood[h][k];
                    if (good > max(maxGood, 4)) {
                        maxGood = good;
                        for (int j = 0; j < 9; ++j) sh.bestE[j] = sh.models[h][k][j];
                        sh.bestIter = it; sh.bestModel = k; sh.hasBest = 1;
                        niters = ransac_update_num_iters(pg.prob, (double)(M - good) / M, 5, niters);
                    }
                }
            }
            sh.maxGood = maxGood;
            sh.niters = niters;
            sh.it0 = it;
            if (it >= niters || it >= pg.maxIters) sh.done = 1;
        }
        __syncthreads();
    }
    // ---- final state, mask of the winning model, SVD(E) -> R1, R2, t  (first half of recoverPose)
    if (kCluster && crank != 0) return;      // lock-step loop: the last cluster barrier is behind every CTA
    if (tid == 0) {
        rs[RS_MAXGOOD] = sh.maxGood; rs[RS_NITERS] = sh.niters; rs[RS_DONE] = 1;
        rs[RS_BESTITER] = sh.bestIter; rs[RS_BESTMODEL] = sh.bestModel; rs[RS_HASBEST] = sh.hasBest;
        rs[RS_RNG_LO] = (int)(uint32_t)sh.rng; rs[RS_RNG_HI] = (int)(uint32_t)(sh.rng >> 32);
    }
    ransac_finish(pg, pb, ps, pair, M, sh.bestE, sh.hasBest, t32, &sh.cnt);
}

// ================================================================================================ exhaustive RANSAC
// "All hypotheses scored" (dvo_config.ransac_exhaustive, BASELINE configs[4]): no adaptive stop, so nothing is sequential
// except the cv::RNG stream -- the whole GPU solves and scores:
//   k_ex_samples   one thread per pair walks the RNG (5 distinct positions per iteration, single-index redraw)
//   k_ex_solve     one 16-lane group per hypothesis (fivepoint_group.cuh), models to global memory
//   k_ex_score     CTA = 16 hypotheses x one slice of the correspondences: models in shared memory, matches in registers,
//                  warp-reduced counts, one atomicAdd per (CTA, model)
//   k_ex_pick      first model (iteration, then solver order) with the highest count > 4 -- what cv2's strict '>' update keeps
//                  when the iteration count never shrinks -- then mask + SVD(E)
// The cv::RNG stream does not depend on the data (cv2 reseeds it with -1 for every call), only the use made of it does
// (value % M, redraw on a duplicate).  So the 64-bit states are tabulated once per context on the host (rngStates), and a
// call only has to find where each iteration starts in that stream:
//   1. every thread simulates "one iteration starting at offset o" for its share of offsets -> draws consumed, len[o]
//   2. one thread follows o -> o + len[o] for maxIters steps (a shared-memory load per step)
//   3. every thread re-simulates its iterations from their start offsets and writes the 5 positions
// A serial walk of the generator itself took 3.4 ms for 4096 iterations; this takes ~0.1 ms.
__device__ __forceinline__ int ex_sample_from(const unsigned long long* states, int o, int nStates, uint32_t M, int* idx) {
    int i = 0, used = 0;
    while (i < 5 && o + used < nStates) {
        const int v = (int)((uint32_t)states[o + used] % M);
        ++used;
        bool dup = false;
        for (int j = 0; j < i; ++j) dup = dup || (idx[j] == v);
        if (!dup) idx[i++] = v;
    }
    return i == 5 ? used : 0;      // 0: ran off the table (cannot happen with the 16x margin unless M is tiny)
}

__global__ void __launch_bounds__(1024) k_ex_samples(PairGeom pg, PairBuffers pb, int pair0) {
    extern __shared__ unsigned char s_len[];           // [nStates] draws consumed by an iteration starting here
    __shared__ int s_ok;
    const int pair = pair0 + blockIdx.x;
    const int M = pb.matchCount[pair];
    if (M <= 5) return;
    const int nStates = pg.rngCount;
    int* starts = pb.exStart + (size_t)pair * pg.maxIters;
    int* out = pb.samples + (size_t)pair * pg.maxIters * 5;
    if (threadIdx.x == 0) s_ok = 1;
    for (int o = threadIdx.x; o < nStates; o += blockDim.x) {
        int idx[5];
        s_len[o] = (unsigned char)min(ex_sample_from(pb.rngStates, o, nStates, (uint32_t)M, idx), 255);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int o = 0;
        for (int it = 0; it < pg.maxIters; ++it) {
            starts[it] = o;
            const int l = o < nStates ? s_len[o] : 0;
            if (l == 0 || l == 255) { s_ok = 0; break; }
            o += l;
        }
    }
    __syncthreads();
    if (s_ok) {
        for (int it = threadIdx.x; it < pg.maxIters; it += blockDim.x) {
            int idx[5];
            ex_sample_from(pb.rngStates, starts[it], nStates, (uint32_t)M, idx);
            for (int k = 0; k < 5; ++k) out[it * 5 + k] = idx[k];
        }
    } else if (threadIdx.x == 0) {      // table exhausted (M barely above 5): plain serial walk of the generator
        uint64_t state = 0xFFFFFFFFFFFFFFFFull;
        for (int it = 0; it < pg.maxIters; ++it) {
            int idx[5];
            int i = 0;
            while (i < 5) {
                const int v = (int)(cvrng_next(state) % (uint32_t)M);
                bool dup = false;
                for (int j = 0; j < i; ++j) dup = dup || (idx[j] == v);
                if (dup) continue;
                idx[i++] = v;
            }
            for (int k = 0; k < 5; ++k) out[it * 5 + k] = idx[k];
        }
    }
}

// Window [it0, it0 + nIt) of iterations; a pair whose loop has already stopped (RS_DONE) is skipped, and so are iterations
// past the current adaptive bound (RS_NITERS) -- both only ever shrink.
__global__ void __launch_bounds__(kRansacThreads, 2) k_ex_solve(PairGeom pg, PairBuffers pb, int pair0, int it0, int nIt) {
    __shared__ SolveScratch scratch[kRansacGroups];
    const int pair = pair0 + blockIdx.y;
    const int* rs = pb.ransacState + pair * 8;
    if (rs[RS_DONE]) return;
    const int itEnd = min(min(it0 + nIt, pg.maxIters), rs[RS_NITERS]);
    const int M = pb.matchCount[pair];
    const int tid = threadIdx.x, lane = tid & 31, grp = tid / kGroupLanes, gl = tid & (kGroupLanes - 1);
    const unsigned gmask = 0xFFFFu << (lane & 16);
    const int itBase = it0 + blockIdx.x * kRansacGroups;
    const int it = itBase + grp;
    int* good = pb.exGood + ((size_t)pair * pg.maxIters + itBase) * kMaxModels;
    for (int i = tid; i < kRansacGroups * kMaxModels; i += kRansacThreads)
        if (itBase + i / kMaxModels < pg.maxIters) good[i] = 0;
    if (it >= itEnd) return;
    if (M <= 5) { if (gl == 0) pb.exCount[(size_t)pair * pg.maxIters + it] = 0; return; }
    const double4* np4 = reinterpret_cast<const double4*>(pb.normPts + (size_t)pair * pg.maxkp * 4);
    const int* smp = pb.samples + ((size_t)pair * pg.maxIters + it) * 5;
    double x1[10], x2[10];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        const double4 p = np4[smp[k]];
        x1[2 * k] = p.x; x1[2 * k + 1] = p.y;
        x2[2 * k] = p.z; x2[2 * k + 1] = p.w;
    }
    double* models = pb.exModels + ((size_t)pair * pg.maxIters + it) * kMaxModels * 9;
    const int n = five_point_solve_group(x1, x2, scratch[grp], models, gmask);
    if (gl == 0) pb.exCount[(size_t)pair * pg.maxIters + it] = n;
}

__global__ void __launch_bounds__(kRansacThreads) k_ex_score(PairGeom pg, PairBuffers pb, int pair0, int itw0, int nItw, float t32) {
    __shared__ double s_models[kRansacGroups][kMaxModels][9];
    __shared__ int s_count[kRansacGroups];
    __shared__ int s_good[kRansacGroups][kMaxModels];
    const int pair = pair0 + blockIdx.z;
    const int M = pb.matchCount[pair];
    const int* rs = pb.ransacState + pair * 8;
    if (M <= 5 || rs[RS_DONE]) return;
    const int tid = threadIdx.x, lane = tid & 31;
    const int it0 = itw0 + blockIdx.x * kRansacGroups;
    const int nIt = min(kRansacGroups, min(min(itw0 + nItw, pg.maxIters), rs[RS_NITERS]) - it0);
    if (nIt <= 0) return;
    const double T = (double)t32, tlo = T * (1.0 - 1e-6), thi = T * (1.0 + 1e-6);
    const double* gm = pb.exModels + ((size_t)pair * pg.maxIters + it0) * kMaxModels * 9;
    for (int i = tid; i < nIt * kMaxModels * 9; i += kRansacThreads) (&s_models[0][0][0])[i] = gm[i];
    for (int i = tid; i < kRansacGroups * kMaxModels; i += kRansacThreads) (&s_good[0][0])[i] = 0;
    if (tid < nIt) s_count[tid] = pb.exCount[(size_t)pair * pg.maxIters + it0 + tid];
    __syncthreads();
    const double4* np4 = reinterpret_cast<const double4*>(pb.normPts + (size_t)pair * pg.maxkp * 4);
    const int per = (M + gridDim.y - 1) / gridDim.y;
    const int mBeg = min(M, (int)blockIdx.y * per), mEnd = min(M, mBeg + per);
    for (int base = mBeg; base < mEnd; base += 2 * kRansacThreads) {
        const int i0 = base + tid, i1 = base + kRansacThreads + tid;
        const bool v0 = i0 < mEnd, v1 = i1 < mEnd;
        const double4 p0 = v0 ? np4[i0] : make_double4(0, 0, 0, 0);
        const double4 p1 = v1 ? np4[i1] : make_double4(0, 0, 0, 0);
        for (int h = 0; h < nIt; ++h) {
            const int nm = s_count[h];
            for (int k = 0; k < nm; ++k) {
                double E[9];
#pragma unroll
                for (int j = 0; j < 9; ++j) E[j] = s_models[h][k][j];
                int good = (v0 && sampson_inlier(E, p0.x, p0.y, p0.z, p0.w, t32, tlo, thi)) ? 1 : 0;
                good += (v1 && sampson_inlier(E, p1.x, p1.y, p1.z, p1.w, t32, tlo, thi)) ? 1 : 0;
                good = __reduce_add_sync(0xffffffffu, good);
                if (lane == 0 && good) smem_red_add(&s_good[h][k], good);
            }
        }
    }
    __syncthreads();
    int* gg = pb.exGood + ((size_t)pair * pg.maxIters + it0) * kMaxModels;
    for (int i = tid; i < nIt * kMaxModels; i += kRansacThreads) {
        const int v = (&s_good[0][0])[i];
        if (v) atomicAdd(gg + i, v);
    }
}

// cv2's update rule over one window of already solved and scored iterations, in order (one thread per pair: the rule is
// sequential, the window is short).  Sets RS_DONE when the adaptive bound has been reached.
__global__ void k_ex_replay(PairGeom pg, PairBuffers pb, int pair0, int nPairs, int it0, int nIt) {
    const int pi = blockIdx.x * blockDim.x + threadIdx.x;
    if (pi >= nPairs) return;
    const int pair = pair0 + pi;
    int* rs = pb.ransacState + pair * 8;
    if (rs[RS_DONE]) return;
    const int M = pb.matchCount[pair];
    int maxGood = rs[RS_MAXGOOD], niters = rs[RS_NITERS];
    const int itEnd = min(it0 + nIt, pg.maxIters);
    const int* cnt = pb.exCount + (size_t)pair * pg.maxIters;
    const int* good = pb.exGood + (size_t)pair * pg.maxIters * kMaxModels;
    int it = it0;
    for (; it < itEnd; ++it) {
        if (it >= niters) break;
        const int nm = cnt[it];
        for (int k = 0; k < nm; ++k) {
            const int g = good[it * kMaxModels + k];
            if (g > max(maxGood, 4)) {
                maxGood = g;
                const double* Eg = pb.exModels + ((size_t)pair * pg.maxIters * kMaxModels + (size_t)it * kMaxModels + k) * 9;
                for (int j = 0; j < 9; ++j) pb.bestE[pair * 9 + j] = Eg[j];
                rs[RS_BESTITER] = it;
                rs[RS_BESTMODEL] = k;
                rs[RS_HASBEST] = 1;
                niters = ransac_update_num_iters(pg.prob, (double)(M - g) / M, 5, niters);
            }
        }
    }
    rs[RS_MAXGOOD] = maxGood;
    rs[RS_NITERS] = niters;
    if (it >= niters || it >= pg.maxIters) rs[RS_DONE] = 1;
}

__global__ void __launch_bounds__(1024) k_ex_finish(PairGeom pg, PairBuffers pb, PoseScratch* ps, int pair0, float t32) {
    __shared__ double s_E[9];
    __shared__ int s_cnt;
    const int pair = pair0 + blockIdx.x;
    int* rs = pb.ransacState + pair * 8;
    const int hasBest = rs[RS_HASBEST];
    if (threadIdx.x < 9) s_E[threadIdx.x] = hasBest ? pb.bestE[pair * 9 + threadIdx.x] : 0.0;
    if (threadIdx.x == 0) rs[RS_DONE] = 1;
    __syncthreads();
    ransac_finish(pg, pb, ps, pair, pb.matchCount[pair], s_E, hasBest, t32, &s_cnt);
}

__global__ void __launch_bounds__(1024) k_ex_pick(PairGeom pg, PairBuffers pb, PoseScratch* ps, int pair0, float t32) {
    __shared__ unsigned long long s_best;
    __shared__ double s_E[9];
    __shared__ int s_cnt;
    const int pair = pair0 + blockIdx.x;
    int* rs = pb.ransacState + pair * 8;
    const int M = pb.matchCount[pair];
    const int tid = threadIdx.x;
    if (tid == 0) s_best = 0ull;
    __syncthreads();
    // key = count << 32 | ~flat index: the maximum is the highest count, and among equals the earliest (iteration, model)
    unsigned long long best = 0ull;
    if (M > 5) {
        const int* cnt = pb.exCount + (size_t)pair * pg.maxIters;
        const int* good = pb.exGood + (size_t)pair * pg.maxIters * kMaxModels;
        for (int f = tid; f < pg.maxIters * kMaxModels; f += blockDim.x) {
            const int it = f / kMaxModels, k = f - it * kMaxModels;
            if (k >= cnt[it]) continue;
            const int g = good[f];
            if (g > 4) best = max(best, ((unsigned long long)(uint32_t)g << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)f));
        }
    }
    for (int o = 16; o > 0; o >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, o));
    if ((tid & 31) == 0 && best) atomicMax(&s_best, best);
    __syncthreads();
    const unsigned long long b = s_best;
    const int hasBest = b != 0ull;
    const int f = hasBest ? (int)(0xFFFFFFFFu - (uint32_t)(b & 0xFFFFFFFFull)) : -1;
    if (hasBest && tid < 9) s_E[tid] = pb.exModels[((size_t)pair * pg.maxIters * kMaxModels + f) * 9 + tid];
    if (tid == 0) {
        rs[RS_MAXGOOD] = hasBest ? (int)(b >> 32) : 0;
        rs[RS_NITERS] = pg.maxIters;
        rs[RS_DONE] = 1;
        rs[RS_BESTITER] = hasBest ? f / kMaxModels : -1;
        rs[RS_BESTMODEL] = hasBest ? f % kMaxModels : -1;
        rs[RS_HASBEST] = hasBest;
    }
    __syncthreads();
    ransac_finish(pg, pb, ps, pair, M, s_E, hasBest, t32, &s_cnt);
}

// ================================================================================================ recoverPose
__global__ void __launch_bounds__(128) k_cheirality(PairGeom pg, PairBuffers pb, PoseScratch* ps, int pair0) {
    const int pi = blockIdx.y, pair = pair0 + pi;
    const int* rs = pb.ransacState + pair * 8;
    if (!rs[RS_HASBEST]) return;
    const int M = pb.matchCount[pair];
    if (blockIdx.x * 128 >= M) return;
    PoseScratch& sc = ps[pair];
    __shared__ double sR1[9], sR2[9], st[3];
    if (threadIdx.x < 9) { sR1[threadIdx.x] = sc.R1[threadIdx.x]; sR2[threadIdx.x] = sc.R2[threadIdx.x]; }
    if (threadIdx.x < 3) st[threadIdx.x] = sc.t[threadIdx.x];
    __syncthreads();
    const int i = blockIdx.x * 128 + threadIdx.x;
    int flags = 0;
    if (i < M) {
        const double* p = pb.normPts + ((size_t)pair * pg.maxkp + i) * 4;
        double x1 = p[0], y1 = p[1], x2 = p[2], y2 = p[3];
        const int a = cheirality_pair(sR1, st, x1, y1, x2, y2, pg.distThresh);     // candidates 0 (R1, t) and 2 (R1, -t)
        const int c = cheirality_pair(sR2, st, x1, y1, x2, y2, pg.distThresh);     // candidates 1 (R2, t) and 3 (R2, -t)
        flags = (a & 1) | ((c & 1) << 1) | ((a & 2) << 1) | ((c & 2) << 2);
        pb.poseMask[(size_t)pair * pg.maxkp + i] = (uint8_t)flags;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        int c = __reduce_add_sync(0xffffffffu, (flags >> k) & 1);
        if ((threadIdx.x & 31) == 0 && c) atomicAdd(&sc.good[k], c);
    }
}

__global__ void __launch_bounds__(256) k_pose_final(OrbGeom og, OrbBuffers ob, PairGeom pg, PairBuffers pb, PoseScratch* ps, int slotA0,
                                                    int pair0) {
    const int pi = blockIdx.x, pair = pair0 + pi;
    const int* rs = pb.ransacState + pair * 8;
    const int M = pb.matchCount[pair];
    const PoseScratch& sc = ps[pair];
    dvo_pose& out = pb.poses[pair];
    uint8_t* pm = pb.poseMask + (size_t)pair * pg.maxkp;
    int k = 0;
    if (rs[RS_HASBEST]) {
        const int g0 = sc.good[0], g1 = sc.good[1], g2 = sc.good[2], g3 = sc.good[3];
        if (g0 >= g1 && g0 >= g2 && g0 >= g3) k = 0;
        else if (g1 >= g0 && g1 >= g2 && g1 >= g3) k = 1;
        else if (g2 >= g0 && g2 >= g1 && g2 >= g3) k = 2;
        else k = 3;
        for (int i = threadIdx.x; i < M; i += 256) pm[i] = ((pm[i] >> k) & 1) ? 255 : 0;
    } else {
        for (int i = threadIdx.x; i < M; i += 256) pm[i] = 0;
    }
    if (threadIdx.x == 0) {
        out.n_matches = M;
        out.n_prev = slotA0 >= 0 ? ob.featCount[slotA0 + pi] : M;
        out.n_cur = slotA0 >= 0 ? ob.featCount[slotA0 + pi + 1] : M;
        out.ransac_iters = rs[RS_NITERS];
        out.best_iter = rs[RS_BESTITER];
        out.frame_flags = slotA0 >= 0 ? (ob.frameFlags[slotA0 + pi] | ob.frameFlags[slotA0 + pi + 1]) : 0;
        if (!rs[RS_HASBEST]) {
            out.status = (M < 5) ? DVO_PAIR_TOO_FEW_MATCHES : DVO_PAIR_NO_MODEL;
            for (int j = 0; j < 9; ++j) { out.R[j] = (j % 4 == 0) ? 1.0 : 0.0; out.E[j] = 0.0; }
            out.t[0] = out.t[1] = out.t[2] = 0.0;
            out.n_inliers = 0; out.n_good = 0; out.candidate = -1;
        } else {
            out.status = DVO_PAIR_OK;
            const double* R = (k & 1) ? sc.R2 : sc.R1;
            const double sgn = (k & 2) ? -1.0 : 1.0;
            for (int j = 0; j < 9; ++j) { out.R[j] = R[j]; out.E[j] = pb.bestE[pair * 9 + j]; }
            for (int j = 0; j < 3; ++j) out.t[j] = sgn * sc.t[j];
            out.n_inliers = sc.nInl;
            out.n_good = sc.good[k];
            out.candidate = k;
        }
    }
}

// ================================================================================================ launcher
static long long g_pair_launches = 0;
long long pair_launch_count() { return g_pair_launches; }

cudaError_t pair_kernels_init_exhaustive(int rngCount) {
    return cudaFuncSetAttribute(k_ex_samples, cudaFuncAttributeMaxDynamicSharedMemorySize, rngCount);
}

cudaError_t pair_kernels_init(int sortBytes) {
    // contexts too large for the in-smem sort (points-only use) never launch k_match_sort: dvo_pairs refuses them
    if (sortBytes > 48 * 1024 && sortBytes <= 200 * 1024)
        return cudaFuncSetAttribute(k_match_sort, cudaFuncAttributeMaxDynamicSharedMemorySize, sortBytes);
    return cudaSuccess;
}

// Points-only entry: normalise caller correspondences and reset the RANSAC state (no matching stage).
__global__ void __launch_bounds__(256) k_points_prep(PairGeom pg, PairBuffers pb, int pair, int n, double fx, double fy, double cx,
                                                     double cy) {
    const size_t o = (size_t)pair * pg.maxkp;
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
        float ax = pb.ptsPrev[(o + r) * 2], ay = pb.ptsPrev[(o + r) * 2 + 1];
        float bx = pb.ptsCur[(o + r) * 2], by = pb.ptsCur[(o + r) * 2 + 1];
        double* np_ = pb.normPts + (o + r) * 4;
        np_[0] = cv_normalize_coord(ax, fx, cx); np_[1] = cv_normalize_coord(ay, fy, cy);
        np_[2] = cv_normalize_coord(bx, fx, cx); np_[3] = cv_normalize_coord(by, fy, cy);
        pb.matches[(o + r) * 3 + 0] = r; pb.matches[(o + r) * 3 + 1] = r; pb.matches[(o + r) * 3 + 2] = 0;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        pb.matchCount[pair] = n;
        int* rs = pb.ransacState + pair * 8;
        rs[RS_MAXGOOD] = 0; rs[RS_NITERS] = pg.maxIters; rs[RS_DONE] = (n <= 5) ? 1 : 0;
        rs[RS_BESTITER] = -1; rs[RS_BESTMODEL] = -1;
        rs[RS_RNG_LO] = (int)0xFFFFFFFFu; rs[RS_RNG_HI] = (int)0xFFFFFFFFu; rs[RS_HASBEST] = 0;
    }
}

void launch_points_prep(const PairGeom& pg, const PairBuffers& pb, int pair, int n, const double* K, cudaStream_t st) {
    k_points_prep<<<64, 256, 0, st>>>(pg, pb, pair, n, K[0], K[4], K[2], K[5]);
    g_pair_launches += 1;
}

void launch_ransac_pose(const OrbGeom& og, const OrbBuffers& ob, const PairGeom& pg, const PairBuffers& pb, int slotA0, int pair0,
                        int nPairs, const double* K, cudaStream_t st) {
    if (nPairs <= 0) return;
    const double fx = K[0], fy = K[4];
    const double thr = pg.threshold / ((fx + fy) / 2.0);
    const float t32 = (float)(thr * thr);
    PoseScratch* ps = pb.poseScratch;
    if (pg.exhaustive) {
        ProfScope ps_(PF_RANSAC, st);
        const int chunks = (pg.maxIters + kRansacGroups - 1) / kRansacGroups;
        int slices = 1;      // enough CTAs to fill the machine a few times over, at least ~2048 correspondences per slice
        while (slices < 64 && (long long)chunks * nPairs * slices < (long long)pg.numSms * 8 && pg.maxkp / (slices * 2) >= 2048) slices *= 2;
        k_ex_samples<<<nPairs, 1024, pg.rngCount, st>>>(pg, pb, pair0);
        k_ex_solve<<<dim3(chunks, nPairs), kRansacThreads, 0, st>>>(pg, pb, pair0, 0, pg.maxIters);
        k_ex_score<<<dim3(chunks, slices, nPairs), kRansacThreads, 0, st>>>(pg, pb, pair0, 0, pg.maxIters, t32);
        k_ex_pick<<<nPairs, 1024, 0, st>>>(pg, pb, ps, pair0, t32);
        g_pair_launches += 4;
        debug_sync("k_ex_*", st);
    } else if (pb.exModels != nullptr && nPairs <= kWindowedMaxPairs && getenv("DVO_NO_WINDOWED") == nullptr) {
        // Few pairs (BASELINE configs[3], [4]): cv2's adaptive loop with speculation across the whole GPU -- windows of
        // iterations are solved and scored by the batched kernels above, then replayed in order; windows grow 128, 384,
        // 1536, ... so the typical pair (about a hundred iterations) stops after the first.
        ProfScope ps_(PF_RANSAC, st);
        k_ex_samples<<<nPairs, 1024, pg.rngCount, st>>>(pg, pb, pair0);
        ++g_pair_launches;
        for (int w0 = 0, wn = 128; w0 < pg.maxIters; w0 += wn, wn *= 3) {
            const int n = std::min(wn, pg.maxIters - w0);
            const int chunks = (n + kRansacGroups - 1) / kRansacGroups;
            int slices = 1;
            while (slices < 64 && (long long)chunks * nPairs * slices < (long long)pg.numSms * 4 && pg.maxkp / (slices * 2) >= 512) slices *= 2;
            k_ex_solve<<<dim3(chunks, nPairs), kRansacThreads, 0, st>>>(pg, pb, pair0, w0, n);
            k_ex_score<<<dim3(chunks, slices, nPairs), kRansacThreads, 0, st>>>(pg, pb, pair0, w0, n, t32);
            k_ex_replay<<<(nPairs + 31) / 32, 32, 0, st>>>(pg, pb, pair0, nPairs, w0, n);
            g_pair_launches += 3;
        }
        k_ex_finish<<<nPairs, 1024, 0, st>>>(pg, pb, ps, pair0, t32);
        ++g_pair_launches;
        debug_sync("k_ex_window", st);
    } else {
        ProfScope ps_(PF_RANSAC, st);
        // few pairs with many correspondences: a cluster of CTAs per pair shares the scoring (DSMEM count reduction)
        int csize = 1;
        if (pg.maxkp >= 4096 && getenv("DVO_NO_CLUSTER") == nullptr) {
            while (csize < 8 && nPairs * csize * 2 <= pg.numSms) csize *= 2;
        }
        if (csize > 1) {
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3(nPairs * csize);
            cfg.blockDim = dim3(kRansacThreads);
            cfg.dynamicSmemBytes = 0;
            cfg.stream = st;
            cudaLaunchAttribute attr{};
            attr.id = cudaLaunchAttributeClusterDimension;
            attr.val.clusterDim.x = csize; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
            cfg.attrs = &attr;
            cfg.numAttrs = 1;
            if (cudaLaunchKernelEx(&cfg, k_ransac<true>, pg, pb, ps, pair0, t32) != cudaSuccess) {
                (void)cudaGetLastError();      // the cluster launch was refused (resources): the single-CTA kernel does the same work
                k_ransac<false><<<nPairs, kRansacThreads, 0, st>>>(pg, pb, ps, pair0, t32);
            }
        } else {
            k_ransac<false><<<nPairs, kRansacThreads, 0, st>>>(pg, pb, ps, pair0, t32);
        }
    }
    debug_sync("k_ransac", st);
    { ProfScope ps_(PF_CHEIRALITY, st); k_cheirality<<<dim3((pg.maxkp + 127) / 128, nPairs), 128, 0, st>>>(pg, pb, ps, pair0); }
    debug_sync("k_cheirality", st);
    { ProfScope ps_(PF_POSE_FINAL, st); k_pose_final<<<nPairs, 256, 0, st>>>(og, ob, pg, pb, ps, slotA0, pair0); }
    g_pair_launches += 3;
    debug_sync("k_pose_final", st);
}

void launch_match(const OrbGeom& og, const OrbBuffers& ob, const PairGeom& pg, const PairBuffers& pb, int slotA0, int pair0,
                  int nPairs, const double* K, cudaStream_t st) {
    if (nPairs <= 0) return;
    const double fx = K[0], fy = K[4], cx = K[2], cy = K[5];
    {
        ProfScope ps_(PF_NN, st);
        const int qBlocks = (pg.maxkp + 127) / 128;
        if (pg.nnTensor) {
            launch_nn_tensor(og, ob, pg, pb, slotA0, pair0, nPairs, pg.numSms, st);      // writes every row key itself
        } else if (cudaMemsetAsync(pb.nnIdx + (size_t)pair0 * 2 * pg.maxkp, 0xFF, sizeof(int) * 2 * (size_t)nPairs * pg.maxkp, st),
                   pg.matcher == DVO_MATCH_CROSSCHECK) {
            const int nSplit = qBlocks >= 4 ? 4 : 1;
            k_nn<true><<<dim3(qBlocks, nSplit, nPairs), 128, 0, st>>>(og, ob, pg, pb, slotA0, pair0);
        } else {
            k_nn<false><<<dim3(qBlocks, 1, nPairs), 128, 0, st>>>(og, ob, pg, pb, slotA0, pair0);
        }
    }
    { ProfScope ps_(PF_SORT, st); k_match_sort<<<nPairs, 1024, pg.sortCap * sizeof(uint32_t), st>>>(og, ob, pg, pb, slotA0, pair0, fx, fy, cx, cy); }
    g_pair_launches += pg.nnTensor ? 3 : 2;      // (k_expand_desc + k_nn_tensor | k_nn) + k_match_sort
    debug_sync("k_nn+sort", st);
}

void launch_pairs(const OrbGeom& og, const OrbBuffers& ob, const PairGeom& pg, const PairBuffers& pb, int slotA0, int pair0,
                  int nPairs, const double* K, cudaStream_t st) {
    if (nPairs <= 0) return;
    launch_match(og, ob, pg, pb, slotA0, pair0, nPairs, K, st);
    launch_ransac_pose(og, ob, pg, pb, slotA0, pair0, nPairs, K, st);
}

}  // namespace dvo
