// ORB detect + describe on sm_100a: the cv2 stages behind feature_detector.detectAndCompute
// (/root/reference/scripts/visual_odometry_v3.py:373).  Bit-exact contract: SURVEY.md Appendix A.1-A.7.
//
// Kernels (one launch each per batch of frame slots; blockIdx.y / .z carries the slot):
//   k_pyr_down      A.1  INTER_LINEAR_EXACT chained down-scale, one launch per level
//   k_fast_nms      A.2  FAST-9/16 score + strict 3x3 NMS, tile staged by TMA (cp.async.bulk.tensor) into smem
//   k_compact       A.2  raster-order compaction of NMS survivors (warp ballot-free prefix by shuffles)
//   k_select        A.3/A.4  retainBest(2N) on FAST score, Harris, retainBest(N), order-exact (select.cuh)
//   k_angle_pack    A.5  intensity-centroid angle + per-frame feature packing
//   k_blur          A.6  7x7 sigma-2 separable Gaussian, float32 FMA order of cv2's filter engine
//   k_brief         A.7  steered BRIEF-256
#include "dvo_internal.cuh"
#include "mathcore.cuh"
#include "select.cuh"
#include <cstdio>
#include <cstdlib>

namespace dvo {

// Shared-memory atomics as single instructions: the callers already aggregate per warp (or hit distinct addresses), so
// the warp-aggregation sequence the compiler wraps around atomicAdd() is pure overhead here.
__device__ __forceinline__ int smem_atom_add(int* p, int v) {
    int old;
    asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"((uint32_t)__cvta_generic_to_shared(p)), "r"(v) : "memory");
    return old;
}
__device__ __forceinline__ void smem_red_add(int* p, int v) {
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(p)), "r"(v) : "memory");
}


static long long g_launches = 0;

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ int find_level_by_tile(const OrbGeom& g, int tile) {
    int L = 0;
#pragma unroll
    for (int i = 1; i < kMaxLevels; ++i)
        if (i < g.nlevels && tile >= g.lv[i].tileBase) L = i;
    return L;
}

// =========================================================================================== A.1 pyramid
// CTA = 128 output columns x (8 * rows) output rows.  The source window of the tile (<= 171 x 78 px at the 1.2 ratio) is first
// copied into shared memory with coalesced 128-bit loads; the four taps of every output pixel are then shared-memory byte
// loads with 32-bit addressing (tap +1 and the next row are immediate offsets), which costs a third of the instructions of
// gathering bytes from global memory with 64-bit addresses.  Each thread owns 4 adjacent output columns (x coefficients in
// registers) and walks `rows` output rows (8 on the large levels, 2 on the small ones so that the grid fills the machine).
// Clamping at the right / bottom edge is implicit: the coefficient tables give weight 0 to the tap past the last pixel.
// kUseTma: the window is one cp.async.bulk.tensor.3d box (176 x 82, or 176 x 24 on the two-row launches) of the source level's
// tensor map, landed through an mbarrier; out-of-range rows / columns arrive as zeros and only ever meet a zero weight.
template <bool kUseTma>
__global__ void __launch_bounds__(256) k_pyr_down(OrbGeom g, OrbBuffers b, const __grid_constant__ CUtensorMap srcMap, int L, int slot0,
                                                  int rows) {
    __shared__ __align__(128) uint8_t tile[kPyrSrcH * kPyrSrcW];
    __shared__ __align__(8) unsigned long long bar;
    const LevelGeom& d = g.lv[L];
    const LevelGeom& s = g.lv[L - 1];
    const int slot = slot0 + blockIdx.z;
    const uint32_t* tx = b.resizeTab + b.resizeTabOff[L][0];
    const uint32_t* ty = b.resizeTab + b.resizeTabOff[L][1];
    const uint8_t* src = b.pyr + (size_t)slot * g.slotStride + s.off;
    uint8_t* dst = b.pyr + (size_t)slot * g.slotStride + d.off;
    const int x0 = blockIdx.x * 128, y0 = blockIdx.y * (8 * rows);
    const int yLast = min(y0 + 8 * rows - 1, d.h - 1);
    const int sx0 = (int)(__ldg(tx + x0) >> 16);
    const int sy0 = (int)(__ldg(ty + y0) >> 16), sy1 = (int)(__ldg(ty + yLast) >> 16) + 1;
    const int ax0 = sx0 & ~15;
    const int tid = threadIdx.y * 32 + threadIdx.x;
    if (kUseTma) {
        const uint32_t barAddr = smem_u32(&bar);
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(barAddr), "r"(1));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(barAddr),
                         "r"(kPyrSrcW * (rows >= 8 ? kPyrSrcH : kPyrSrcHSmall))
                         : "memory");
            asm volatile(
                "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
                    smem_u32(tile)),
                "l"(reinterpret_cast<uint64_t>(&srcMap)), "r"(barAddr), "r"(ax0), "r"(sy0), "r"(slot)
                : "memory");
        }
        __syncthreads();                       // the barrier is initialised before anyone polls it
        uint32_t done = 0;
        for (uint32_t spin = 0; !done; ++spin) {      // bounded: a lost copy traps instead of hanging the GPU
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(done)
                : "r"(barAddr), "r"(0)
                : "memory");
            if (spin > (1u << 24)) __trap();
        }
    } else {
        // always the full 176-byte rows (compile-time divisor below); what lies beyond sx1 is never used, and reading it is
        // safe: rows are pitch-padded and the buffers end with slack
        constexpr int nVec = kPyrSrcW / 16;
        const int nRow = min(sy1 - sy0 + 1, kPyrSrcH);
        for (int i = tid; i < nRow * nVec; i += 256) {
            const int r = i / nVec, v = i - r * nVec;
            // row s.h (one past the image) is allocated padding; it only ever meets a zero weight
            const uint4 q = __ldg(reinterpret_cast<const uint4*>(src + (size_t)min(sy0 + r, s.h) * s.pitch + ax0) + v);
            *reinterpret_cast<uint4*>(tile + r * kPyrSrcW + v * 16) = q;
        }
    }
    __syncthreads();
    const int x4 = x0 + threadIdx.x * 4;
    if (x4 >= d.w) return;
    // The 4 columns' taps (ox, ox+1) lie within 8 bytes of ox[0] (ratio 1.2: ox[3] - ox[0] <= 4; checked on the host), so a
    // row needs three aligned 32-bit words, two funnel shifts to bring byte ox[0] to position 0, and per column one PRMT
    // (pick the two taps) + one DP2A (two 16-bit weights x two bytes): 3 shared loads per source row instead of 8.
    uint32_t wgt[4], sel[4];
    int ox0 = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t ex = __ldg(tx + min(x4 + k, d.w - 1));
        const int o = (int)(ex >> 16) - ax0;
        if (k == 0) ox0 = o;
        const uint32_t c1 = ex & 0xFFFF, rel = (uint32_t)(o - ox0);
        wgt[k] = (256u - c1) | (c1 << 16);
        sel[k] = rel | ((rel + 1u) << 4);
    }
    const int wofs = ox0 & ~3;
    const uint32_t sh = (uint32_t)(ox0 & 3) * 8u;
    for (int r = 0; r < rows; ++r) {
        const int y = y0 + threadIdx.y + 8 * r;
        if (y >= d.h) break;
        const uint32_t ey = __ldg(ty + y);
        const uint32_t cy1 = ey & 0xFFFF, cy0 = 256u - cy1;
        const uint32_t* p0 = reinterpret_cast<const uint32_t*>(tile + ((int)(ey >> 16) - sy0) * kPyrSrcW + wofs);
        const uint32_t* p1 = p0 + kPyrSrcW / 4;
        const uint32_t a0 = p0[0], a1 = p0[1], a2 = p0[2], b0 = p1[0], b1 = p1[1], b2 = p1[2];
        const uint32_t alo = __funnelshift_r(a0, a1, sh), ahi = __funnelshift_r(a1, a2, sh);
        const uint32_t blo = __funnelshift_r(b0, b1, sh), bhi = __funnelshift_r(b1, b2, sh);
        uint32_t out = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t h0 = __dp2a_lo(wgt[k], __byte_perm(alo, ahi, sel[k]), 0u);
            const uint32_t h1 = __dp2a_lo(wgt[k], __byte_perm(blo, bhi, sel[k]), 0u);
            const uint32_t v = (h0 * cy0 + h1 * cy1 + (1u << 15)) >> 16;
            out |= v << (8 * k);
        }
        *reinterpret_cast<uint32_t*>(dst + (size_t)y * d.pitch + x4) = out;
    }
}

// =========================================================================================== A.2 FAST + NMS

template <bool kUseTma>
__global__ void __launch_bounds__(256, 7) k_fast_nms(OrbGeom g, OrbBuffers b, const __grid_constant__ TensorMaps tm, int slot0) {
    __shared__ __align__(128) uint8_t raw[kFastBoxH * kFastBoxW];
    __shared__ __align__(16) uint8_t sc[(kTileH + 2) * 136];
    __shared__ __align__(16) uint16_t list1[(kTileH + 2) * (kTileW + 2)];
    constexpr int kList2Cap = (kTileH + 2) * (kTileW + 2);
    __shared__ uint16_t list2[kList2Cap];
    __shared__ int s_n1, s_n2;
    __shared__ int s_rowcnt[2 * kTileH];
    __shared__ __align__(8) unsigned long long bar;

    const uint32_t ti = __ldg(b.tileInfo + blockIdx.x);
    const int L = ti & 15;
    const LevelGeom lv = g.lv[L];
    const int x0 = (int)((ti >> 4) & 0xFFF) * kTileW, y0 = (int)(ti >> 16) * kTileH;
    const int slot = slot0 + blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31;

    if (kUseTma) {
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (tid == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)),
                         "r"(kFastBoxH * kFastBoxW)
                         : "memory");
            asm volatile(
                "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
                    smem_u32(raw)),
                "l"(reinterpret_cast<uint64_t>(&tm.pyr[L])), "r"(smem_u32(&bar)), "r"(x0 - kFastHaloL), "r"(y0 - 4), "r"(slot)
                : "memory");
        }
        for (int i = tid; i < (kTileH + 2) * 136 / 16; i += 256) reinterpret_cast<uint4*>(sc)[i] = make_uint4(0, 0, 0, 0);
        if (tid == 0) { s_n1 = 0; s_n2 = 0; }
        // all threads wait for the bytes to land (phase 0)
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "WAIT_%=:\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
            "@p bra DONE_%=;\n"
            "bra WAIT_%=;\n"
            "DONE_%=:\n"
            "}\n" ::"r"(smem_u32(&bar)),
            "r"(0)
            : "memory");
    } else {
        const uint8_t* img = b.pyr + (size_t)slot * g.slotStride + lv.off;
        for (int i = tid; i < kFastBoxH * (kFastBoxW / 4); i += 256) {
            int ry = i / (kFastBoxW / 4), rw = i % (kFastBoxW / 4);
            int gy = y0 - 4 + ry, gx = x0 - kFastHaloL + rw * 4;
            uint32_t v = 0;
            if (gy >= 0 && gy < lv.h && gx >= 0 && gx < lv.pitch) {
                v = *reinterpret_cast<const uint32_t*>(img + (size_t)gy * lv.pitch + gx);
                // zero bytes at x >= w like TMA's out-of-bounds fill
                if (gx + 3 >= lv.w) {
                    uint32_t keep = 0;
                    for (int k = 0; k < 4; ++k)
                        if (gx + k < lv.w) keep |= 0xFFu << (8 * k);
                    v &= keep;
                }
            }
            *reinterpret_cast<uint32_t*>(raw + ry * kFastBoxW + rw * 4) = v;
        }
        for (int i = tid; i < (kTileH + 2) * 136 / 16; i += 256) reinterpret_cast<uint4*>(sc)[i] = make_uint4(0, 0, 0, 0);
        if (tid == 0) { s_n1 = 0; s_n2 = 0; }
    }
    __syncthreads();

    // ---- phase 1: packed prefilter, 4 adjacent positions per thread (quads aligned to 4 raw columns).
    // Ring positions are 130 x 34 (tile + 1): raw rows 3..36, raw cols 15..144; a position is coded by its raw offset.
    // Thread layout: 34 quad columns x 7 rows per sweep (238 of 256 threads), 5 sweeps cover the 34 ring rows.
    const int th = g.fastThreshold;
    constexpr int kRowWords = kFastBoxW / 4;
    constexpr int kQuadsX = 34, kQuadsY = kTileH + 2, kRowsPerSweep = 7;
    {
        const int qx = tid % kQuadsX, qrow = tid / kQuadsX;          // qrow == 7 for the 18 spare threads
        // bytes of this thread's quads that are ring columns (raw 15..144) inside the level's FAST domain (3 <= x < w - 3):
        // cut the leading / trailing bytes outside [cLo, cHi]
        const int rc0 = 12 + 4 * qx;
        const int cLo = max(15, 3 - (x0 - kFastHaloL)), cHi = min(144, lv.w - 4 - (x0 - kFastHaloL));
        const int nLead = min(max(cLo - rc0, 0), 4), nTrail = min(max(rc0 + 3 - cHi, 0), 4);
        uint32_t colmask = 0x80808080u;
        colmask &= nLead >= 4 ? 0u : (0xFFFFFFFFu << (8 * nLead));
        colmask &= nTrail >= 4 ? 0u : (0xFFFFFFFFu >> (8 * nTrail));
        if (qrow >= kRowsPerSweep) colmask = 0;
        constexpr int kSweeps = (kQuadsY + kRowsPerSweep - 1) / kRowsPerSweep;
        uint32_t passw[kSweeps];
        int cnt = 0;
#pragma unroll
        for (int sw = 0; sw < kSweeps; ++sw) {
            const int qy = sw * kRowsPerSweep + qrow;
            const int sy = y0 - 1 + qy;
            uint32_t pass = 0;
            if (colmask != 0 && qy < kQuadsY && sy >= 3 && sy < lv.h - 3) {
                const uint32_t* rp = reinterpret_cast<const uint32_t*>(raw + (qy + 3) * kFastBoxW + 12 + 4 * qx);
                const uint32_t c = rp[0], lw = rp[-1], rw = rp[1];
                const uint32_t up = rp[-3 * kRowWords], dn = rp[3 * kRowWords];
                const uint32_t e = __byte_perm(c, rw, 0x6543), w = __byte_perm(lw, c, 0x4321);
                pass = fast_prefilter_u8x4(c, up, e, dn, w, th) & colmask;
            }
            passw[sw] = pass;
            cnt += __popc(pass);
        }
        // warp-wide exclusive scan of the per-thread survivor counts; each warp reserves its share of list1 with one
        // shared atomic (the list order is irrelevant: every later phase writes through maps), then every thread appends
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int n = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += n;
        }
        int base = 0;
        if (lane == 31 && incl > 0) base = smem_atom_add(&s_n1, incl);
        base = __shfl_sync(0xffffffffu, base, 31);
        // four predicated appends per sweep through a 32-bit shared address (no divergent walk over the set bits, and no
        // re-derivation of the list base under every predicate)
        uint32_t la = (uint32_t)__cvta_generic_to_shared(list1) + 2u * (uint32_t)(base + incl - cnt);
#pragma unroll
        for (int sw = 0; sw < kSweeps; ++sw) {
            const uint32_t m = passw[sw];
            const int code0 = (sw * kRowsPerSweep + qrow + 3) * kFastBoxW + 12 + 4 * qx;
#pragma unroll
            for (int q = 0; q < 4; ++q)
                asm volatile(
                    "{\n\t.reg .pred p;\n\t.reg .b32 t;\n\t"
                    "and.b32 t, %2, %3;\n\t"
                    "setp.ne.u32 p, t, 0;\n\t"
                    "@p st.shared.u16 [%0], %1;\n\t"
                    "@p add.u32 %0, %0, 2;\n\t}"
                    : "+r"(la)
                    : "h"((uint16_t)(code0 + q)), "r"(m), "r"(0x80u << (8 * q))
                    : "memory");
        }
    }
    __syncthreads();

    // ---- phase 2: exact 16-point corner test on the survivors
    const int dxs[16] = DVO_FAST_DX, dys[16] = DVO_FAST_DY;
    const int n1 = s_n1;
    for (int i0 = 0; i0 < n1; i0 += 256) {
        const int i = i0 + tid;
        int pol = 0;
        int code = 0;
        if (i < n1) {
            code = list1[i];
            const uint8_t* c = raw + code;
            int pr[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) pr[k] = c[dys[k] * kFastBoxW + dxs[k]];
            pol = fast_corner_polarity16(*c, pr, th);
        }
        const bool corner = pol != 0;
        const unsigned m = __ballot_sync(0xffffffffu, corner);
        int base = 0;
        if (lane == 0 && m) base = smem_atom_add(&s_n2, __popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (corner) list2[base + __popc(m & ((1u << lane) - 1))] = (uint16_t)(code | (pol << 13));   // code < 6400 < 2^13
    }
    __syncthreads();

    // ---- phase 3: scores of the corners into the ring score tile; list1 is dead now and becomes the survivor staging area
    uint32_t* stage = reinterpret_cast<uint32_t*>(list1);                    // [kTileListCap] packed survivors, any order
    uint16_t* stageIdx = reinterpret_cast<uint16_t*>(stage + kTileListCap);  // [kTileListCap] arrival index within the row
    static_assert(sizeof(list1) >= kTileListCap * 6, "survivor staging does not fit the dead prefilter list");
    if (tid < kTileH) s_rowcnt[tid] = 0;
    if (tid == 0) s_n1 = 0;                                                  // reused as the survivor counter
    const int n2 = s_n2;
    for (int i = tid; i < n2; i += 256) {
        const int e = list2[i];
        const int code = e & 0x1FFF, pol = e >> 13;
        const uint8_t* c = raw + code;
        int pr[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) pr[k] = c[dys[k] * kFastBoxW + dxs[k]];
        const int ry = code / kFastBoxW, rx = code - ry * kFastBoxW;
        sc[(ry - 3) * 136 + (rx - 15)] = (uint8_t)fast_corner_score16(*c, pr, th, pol);
    }
    __syncthreads();

    // ---- phase 4: strict 3x3 NMS, driven by the corner list (non-corners score 0), 31-px border cull.  A survivor takes the
    // next place of its row (shared atomic) and is parked in the staging list.
    const int border = 31;
    for (int i = tid; i < n2; i += 256) {
        const int code = list2[i] & 0x1FFF;
        const int ry = code / kFastBoxW, rx = code - ry * kFastBoxW;
        const int cy = ry - 3, cx = rx - 15;                   // ring coordinates: tile pixel (cx-1, cy-1)
        const int x = x0 + cx - 1, y = y0 + cy - 1;
        // inside the tile (ring positions only lend their scores) and inside the 31-px border: four unsigned range checks
        const bool in = ((unsigned)(cx - 1) < (unsigned)kTileW) & ((unsigned)(cy - 1) < (unsigned)kTileH) &
                        ((unsigned)(x - border) < (unsigned)max(lv.w - 2 * border, 0)) &
                        ((unsigned)(y - border) < (unsigned)max(lv.h - 2 * border, 0));
        if (!in) continue;
        const uint8_t* q = sc + cy * 136 + cx;
        const int sv = *q;
        const int nmax = imax3(imax3(q[-136 - 1], q[-136], q[-136 + 1]), imax3(q[-1], q[1], q[136 - 1]), imax(q[136], q[136 + 1]));
        if (sv > nmax) {
            const int k = smem_atom_add(&s_rowcnt[cy - 1], 1);
            const int p = smem_atom_add(&s_n1, 1);
            stage[p] = ((uint32_t)sv << 24) | ((uint32_t)y << 12) | (uint32_t)x;
            stageIdx[p] = (uint16_t)k;
        }
    }
    __syncthreads();

    // ---- phase 5: the tile's survivors -> this tile's list in HBM, grouped by row (k_gather orders the few entries of a
    // (tile, row) run by x).  Nothing else is written: the full-resolution score map this kernel used to store -- and k_compact
    // to re-read -- was 2/3 of its DRAM traffic.
    {
        const size_t tileIdx = (size_t)slot * g.tilesPerFrame + blockIdx.x;
        int* s_rowOff = s_rowcnt + kTileH;       // 32 exclusive offsets
        if (tid < 32) {
            const int c = s_rowcnt[tid];
            int incl = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int n = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += n;
            }
            s_rowOff[tid] = incl - c;
            b.tileCnt[tileIdx * kTileH + tid] = (uint16_t)c;
            if (tid == 31) b.tileTot[tileIdx] = incl;
        }
        __syncthreads();
        const int n3 = s_n1;
        uint32_t* list = b.tileList + tileIdx * kTileListCap;
        for (int i = tid; i < n3; i += 256) {
            const uint32_t e = stage[i];
            list[s_rowOff[((e >> 12) & 0xFFFu) - y0] + stageIdx[i]] = e;
        }
    }
}

// =========================================================================================== raster-order gather
// The per-tile survivor lists of one 32-row band of a level -> the level's raster-order candidate list (cv2's FAST output
// order, which retainBest's order-exact replay depends on).  Only survivors move: ~8 bytes each.
__global__ void __launch_bounds__(256) k_gather(OrbGeom g, OrbBuffers b, int slot0) {
    __shared__ int s_red[8];
    __shared__ int s_rowOff[33];
    __shared__ uint16_t s_cnt[32][33];      // [tile in band][row]; a level is at most 4096 / 128 = 32 tiles wide
    __shared__ uint16_t s_src[32][33];      // offset of (tile, row) inside the tile's list
    __shared__ uint16_t s_dst[32][33];      // offset of (tile, row) inside the row's run of the level list
    const int rb = blockIdx.x;
    int L = 0;
#pragma unroll
    for (int i = 1; i < kMaxLevels; ++i)
        if (i < g.nlevels && rb >= g.lv[i].rbBase) L = i;
    const LevelGeom lv = g.lv[L];
    const int slot = slot0 + blockIdx.y;
    const int band = rb - lv.rbBase, tilesX = lv.tilesX;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const size_t tile0 = (size_t)slot * g.tilesPerFrame + lv.tileBase;
    // base = survivors of every band above this one
    int acc = 0;
    for (int i = tid; i < band * tilesX; i += 256) acc += b.tileTot[tile0 + i];
    acc = __reduce_add_sync(0xffffffffu, acc);
    if (lane == 0) s_red[warp] = acc;
    const uint16_t* cnt = b.tileCnt + (tile0 + (size_t)band * tilesX) * kTileH;
    for (int i = tid; i < tilesX * kTileH; i += 256) s_cnt[i >> 5][i & 31] = cnt[i];
    __syncthreads();
    if (warp == 0) {          // lane = row: run of the row in the level list
        int c = 0;
        for (int t = 0; t < tilesX; ++t) { s_dst[t][lane] = (uint16_t)c; c += s_cnt[t][lane]; }
        int base = 0;
        for (int i = 0; i < 8; ++i) base += s_red[i];
        int incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int n = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += n;
        }
        s_rowOff[lane] = base + incl - c;
        if (lane == 31) s_rowOff[32] = base + incl;
    } else if (warp == 1 && lane < tilesX) {      // lane = tile: start of each row inside the tile's list
        int c = 0;
        for (int r = 0; r < kTileH; ++r) { s_src[lane][r] = (uint16_t)c; c += s_cnt[lane][r]; }
    }
    __syncthreads();
    if (band == lv.tilesY - 1 && tid == 0) b.candCount[slot * kMaxLevels + L] = s_rowOff[32];
    uint32_t* cand = b.cand + (size_t)slot * g.candPerSlot + lv.candBase;
    const uint32_t* lists = b.tileList + (tile0 + (size_t)band * tilesX) * kTileListCap;
    for (int i = tid; i < kTileH * tilesX; i += 256) {      // row-major over (row, tile): neighbouring threads write neighbouring runs
        const int r = i / tilesX, t = i - r * tilesX;
        const int n = s_cnt[t][r];
        if (n == 0) continue;
        const uint32_t* src = lists + (size_t)t * kTileListCap + s_src[t][r];
        const int dst = s_rowOff[r] + s_dst[t][r];
        if (n == 1) {
            if (dst < lv.candCap) cand[dst] = src[0];
            continue;
        }
        if (n <= 4) {        // the common case from registers: four loads, six comparisons
            const uint32_t e0 = src[0], e1 = src[1], e2 = n > 2 ? src[2] : 0xFFFFFFFFu, e3 = n > 3 ? src[3] : 0xFFFFFFFFu;
            const uint32_t a0 = e0 & 0xFFFu, a1 = e1 & 0xFFFu, a2 = n > 2 ? (e2 & 0xFFFu) : 0x1000u, a3 = n > 3 ? (e3 & 0xFFFu) : 0x1001u;
            const int l01 = a0 < a1, l02 = a0 < a2, l03 = a0 < a3, l12 = a1 < a2, l13 = a1 < a3, l23 = a2 < a3;
            const int r0 = 3 - l01 - l02 - l03, r1 = l01 + 2 - l12 - l13, r2 = l02 + l12 + 1 - l23, r3 = l03 + l13 + l23;
            if (dst + n <= lv.candCap) {
                cand[dst + r0] = e0;
                cand[dst + r1] = e1;
                if (n > 2) cand[dst + r2] = e2;
                if (n > 3) cand[dst + r3] = e3;
            }
            continue;
        }
        for (int k = 0; k < n; ++k) {        // a run arrives in atomic order: place each entry by its rank in x (runs are a few entries)
            const uint32_t e = src[k];
            int rank = 0;
            for (int q = 0; q < n; ++q) rank += (src[q] & 0xFFFu) < (e & 0xFFFu) ? 1 : 0;
            if (dst + rank < lv.candCap) cand[dst + rank] = e;
        }
    }
}

// =========================================================================================== selection
// Accessor used by select.cuh on the device: the whole warp executes the replay with identical scalar state; element
// writes are done by lane 0, the two Hoare scan loops look at 32 elements per step.
template <class T, class Key>
struct WarpAcc {
    typedef T Item;
    T* d;
    int n;   // array length: scans never read outside [0, n)
    __device__ __forceinline__ T get(int i) const { return d[i]; }
    __device__ __forceinline__ void set(int i, T v) {
        if (lane_id() == 0) d[i] = v;
        __syncwarp();
    }
    __device__ __forceinline__ static bool gt(T a, T b) { return Key::key(a) > Key::key(b); }
    __device__ __forceinline__ int scan_up(int first, T pivot) const {
        while (true) {
            int i = first + lane_id();
            bool stop = (i >= n) || !gt(d[i], pivot);
            unsigned m = __ballot_sync(0xffffffffu, stop);
            if (m) return first + __ffs(m) - 1;
            first += 32;
        }
    }
    __device__ __forceinline__ int scan_down(int last, T pivot) const {
        while (true) {
            int i = last - lane_id();
            bool stop = (i < 0) || !gt(pivot, d[i]);
            unsigned m = __ballot_sync(0xffffffffu, stop);
            if (m) return last - (__ffs(m) - 1);
            last -= 32;
        }
    }
    __device__ __forceinline__ int scan_up_ge(int first, int last, T bnd) const {
        while (true) {
            int i = first + lane_id();
            bool stop = (i >= last) || !(Key::key(d[i]) >= Key::key(bnd));
            unsigned m = __ballot_sync(0xffffffffu, stop);
            if (m) return min(first + __ffs(m) - 1, last);
            first += 32;
        }
    }
    __device__ __forceinline__ int scan_down_lt(int first, int last, T bnd) const {
        while (true) {
            int i = last - lane_id();
            bool stop = (i <= first) || (Key::key(d[i]) >= Key::key(bnd));
            unsigned m = __ballot_sync(0xffffffffu, stop);
            if (m) return max(last - (__ffs(m) - 1), first);
            last -= 32;
        }
    }
};

struct KeyScore {   // packed candidate: FAST score in the top byte
    __device__ __forceinline__ static int key(uint32_t v) { return (int)(v >> 24); }
};
struct KeyHarris {  // harris float bits in the high word
    __device__ __forceinline__ static float key(unsigned long long v) { return __uint_as_float((uint32_t)(v >> 32)); }
};

// 7x7 block of Sobel-3 products for kN keypoints at once: lanes 0..48 (two rounds) each take one block pixel of every
// keypoint, and the 18 * kN byte loads are all issued before the first use (one memory latency for the group -- the
// one-keypoint-at-a-time form left each warp waiting on L2/DRAM ~27 times per level).
template <int kN>
__device__ __forceinline__ void harris_sums_warp(const uint8_t* img, int pitch, const int (&xs)[kN], const int (&ys)[kN], int lane,
                                                 int (&a_out)[kN], int (&b_out)[kN], int (&c_out)[kN]) {
    int px[kN][2][8];
#pragma unroll
    for (int q = 0; q < kN; ++q)
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int i = lane + 32 * r;
            const bool on = i < 49;
            const int dy = i / 7 - 3, dx = i % 7 - 3;
            const uint8_t* p = img + (size_t)(ys[q] + dy) * pitch + (xs[q] + dx);
            px[q][r][0] = on ? p[-pitch - 1] : 0; px[q][r][1] = on ? p[-pitch] : 0; px[q][r][2] = on ? p[-pitch + 1] : 0;
            px[q][r][3] = on ? p[-1] : 0; px[q][r][4] = on ? p[1] : 0;
            px[q][r][5] = on ? p[pitch - 1] : 0; px[q][r][6] = on ? p[pitch] : 0; px[q][r][7] = on ? p[pitch + 1] : 0;
        }
#pragma unroll
    for (int q = 0; q < kN; ++q) {
        int a = 0, bb = 0, c = 0;
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int* v = px[q][r];
            const int ix = (v[4] - v[3]) * 2 + (v[2] - v[0]) + (v[7] - v[5]);
            const int iy = (v[6] - v[1]) * 2 + (v[5] - v[0]) + (v[7] - v[2]);
            a += ix * ix;
            bb += iy * iy;
            c += ix * iy;
        }
        a_out[q] = __reduce_add_sync(0xffffffffu, a);
        b_out[q] = __reduce_add_sync(0xffffffffu, bb);
        c_out[q] = __reduce_add_sync(0xffffffffu, c);
    }
}

// Single-thread accessor (median-of-3 step).
template <class T, class Key>
struct OneAcc {
    typedef T Item;
    T* d;
    __device__ __forceinline__ T get(int i) const { return d[i]; }
    __device__ __forceinline__ void set(int i, T v) { d[i] = v; }
    __device__ __forceinline__ static bool gt(T a, T b) { return Key::key(a) > Key::key(b); }
};

struct SelShared { int warpL[32], warpR[32]; int K; };

// Whole-CTA accessor for the paired (data-parallel) form in select.cuh: each partition pass = two ballot sweeps
// (stopper counts, then ranks + stopper lists + K) and K independent swaps.  Every thread runs the same scalar control
// flow; element writes outside the swap phase are done by one thread / one warp and fenced with a CTA barrier.
template <class T, class Key>
struct BlockAcc {
    typedef T Item;
    T* d;
    int n;
    uint32_t* listL;     // scratch (global): left / right stopper positions by rank, >= n/2 + 2 entries each
    uint32_t* listR;
    SelShared* sh;
    __device__ __forceinline__ T get(int i) const { return d[i]; }
    __device__ __forceinline__ static bool gt(T a, T b) { return Key::key(a) > Key::key(b); }

    __device__ void median_to_first(int result, int ia, int ib, int ic) {
        if (threadIdx.x == 0) {
            OneAcc<T, Key> one{d};
            move_median_to_first(one, result, ia, ib, ic);
        }
        __syncthreads();
    }
    __device__ void sequential_tail(int first, int nth, int last, int depth) {
        if (threadIdx.x < 32) {
            WarpAcc<T, Key> w{d, n};
            nth_element_replay_depth(w, first, nth, last, depth);
        }
        __syncthreads();
    }
    // kMode 0: Hoare pass against `ref` as pivot, returns the cut.  kMode 1: std::partition(x >= ref), returns the
    // partition point.
    template <int kMode>
    __device__ int pair_swap(int first, int last, T ref) {
        const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nW = blockDim.x >> 5;
        const int m = last - first;
        const int seg = (((m + nW - 1) / nW) + 31) & ~31;
        const int s0 = min(first + warp * seg, last), s1 = min(s0 + seg, last);
        const auto kref = Key::key(ref);
        int cL = 0, cR = 0;
        for (int p = s0; p < s1; p += 32) {
            const int i = p + lane;
            bool l = false, r = false;
            if (i < s1) {
                const auto kx = Key::key(d[i]);
                if (kMode == 0) { l = !(kx > kref); r = !(kref > kx); }
                else { r = kx >= kref; l = !r; }
            }
            cL += __popc(__ballot_sync(0xffffffffu, l));
            cR += __popc(__ballot_sync(0xffffffffu, r));
        }
        if (lane == 0) { sh->warpL[warp] = cL; sh->warpR[warp] = cR; }
        if (tid == 0) sh->K = 0;
        __syncthreads();
        const int vL = lane < nW ? sh->warpL[lane] : 0, vR = lane < nW ? sh->warpR[lane] : 0;
        int iL = vL, iR = vR;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int a = __shfl_up_sync(0xffffffffu, iL, o), c = __shfl_up_sync(0xffffffffu, iR, o);
            if (lane >= o) { iL += a; iR += c; }
        }
        const int totalL = __shfl_sync(0xffffffffu, iL, 31), totalR = __shfl_sync(0xffffffffu, iR, 31);
        int runL = __shfl_sync(0xffffffffu, iL - vL, warp), runR = __shfl_sync(0xffffffffu, iR - vR, warp);
        const int cap = m / 2 + 1;
        const unsigned lt = (1u << lane) - 1u;
        int kc = 0;
        for (int p = s0; p < s1; p += 32) {
            const int i = p + lane;
            bool l = false, r = false;
            if (i < s1) {
                const auto kx = Key::key(d[i]);
                if (kMode == 0) { l = !(kx > kref); r = !(kref > kx); }
                else { r = kx >= kref; l = !r; }
            }
            const unsigned ml = __ballot_sync(0xffffffffu, l), mr = __ballot_sync(0xffffffffu, r);
            const int rl = runL + __popc(ml & lt), rrl = runR + __popc(mr & lt);
            if (l) {
                // pair rl exists iff more than rl right stoppers lie to the right of i
                if (totalR - (rrl + (r ? 1 : 0)) >= rl + 1) ++kc;
                if (rl <= cap) listL[rl] = (uint32_t)i;
            }
            if (r) {
                const int rr = totalR - 1 - rrl;
                if (rr <= cap) listR[rr] = (uint32_t)i;
            }
            runL += __popc(ml);
            runR += __popc(mr);
        }
        kc = __reduce_add_sync(0xffffffffu, kc);
        if (lane == 0 && kc) atomicAdd(&sh->K, kc);
        __syncthreads();
        const int K = sh->K;
        for (int k = tid; k < K; k += blockDim.x) {
            const int i = (int)listL[k], j = (int)listR[k];
            const T x = d[i], y = d[j];
            d[i] = y;
            d[j] = x;
        }
        int res;
        if (kMode == 0) {
            const int lk = K < totalL ? (int)listL[K] : 0x7fffffff;
            const int rk1 = K > 0 ? (int)listR[K - 1] : 0x7fffffff;
            res = min(lk, rk1);
        } else {
            res = first + totalR;
        }
        __syncthreads();
        return res;
    }
    __device__ int pair_swap_hoare(int first, int last, T pivot) { return pair_swap<0>(first, last, pivot); }
    __device__ int pair_swap_ge(int first, int last, T boundary) { return pair_swap<1>(first, last, boundary); }
};

// Launch shapes.  Few frames in flight (two-frame calls, the high-density configuration): big CTAs -- 1024 threads and the whole
// candidate list of a level in 160 KB of shared memory (512 threads / 64 KB for the small levels) -- finish a level fastest.
// Batches: SMALL CTAs -- 128 threads and a 32 KB window, longer lists run from the L2-resident global working copy.  A big CTA
// needs a whole SM's registers, starves behind the small CTAs of FAST / blur and blocks them while it runs; small CTAs co-reside
// with everything, which is what lets the image half of the next batch run under this kernel (dvo_api.cu,
// sequence_step_pipelined): 35.8 k -> 36.9 k pairs/s, and 148 x 8 of them also finish sooner than 444 big ones (0.52 -> 0.31 ms).
constexpr int kSelectThreads = 1024;          // launch bound (64 registers)
constexpr int kSelectBatchThreads = 128;
constexpr int kSelectBatchSmemBytes = 32 * 1024;
constexpr int kSelectBatchMinSlots = 24;
constexpr int kSelectSeqTail = 32;    // ranges this short are finished by one warp running the scalar replay

__global__ void __launch_bounds__(kSelectThreads) k_select(OrbGeom g, OrbBuffers b, int slot0, int smemBytes, int level0) {
    extern __shared__ __align__(16) unsigned char s_dyn[];
    __shared__ SelShared sh;
    const int L = level0 + blockIdx.y;        // the largest level of the launch (the long CTAs) is dispatched first
    const int slot = slot0 + blockIdx.x;
    const LevelGeom lv = g.lv[L];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t* cand = b.cand + (size_t)slot * g.candPerSlot + lv.candBase;
    unsigned long long* pairs = b.pairs + (size_t)slot * g.candPerSlot + lv.candBase;
    uint32_t* listL = b.selList + ((size_t)slot * g.candPerSlot + lv.candBase) * 2;
    uint32_t* listR = listL + lv.candCap;
    const int nRaw = b.candCount[slot * kMaxLevels + L];
    const int n = min(nRaw, lv.candCap);

    // ---- pass 1: retainBest(2 * quota) on the FAST score, order-exact
    const bool inSmem1 = (size_t)n * sizeof(uint32_t) <= (size_t)smemBytes;
    uint32_t* work1 = inSmem1 ? reinterpret_cast<uint32_t*>(s_dyn) : b.selWork + (size_t)slot * g.candPerSlot + lv.candBase;
    for (int i = tid; i < n; i += blockDim.x) work1[i] = cand[i];
    __syncthreads();
    int n1;
    {
        BlockAcc<uint32_t, KeyScore> acc{work1, n, listL, listR, &sh};
        n1 = retain_best_paired(acc, n, 2 * lv.quota, kSelectSeqTail);
    }

    // ---- Harris response of the survivors (un-blurred level)
    const uint8_t* img = b.pyr + (size_t)slot * g.slotStride + lv.off;
    {
        constexpr int kH = 2;                      // keypoints per warp per round (3 spills at 64 registers)
        const int nW = blockDim.x >> 5;
        for (int i0 = warp; i0 < n1; i0 += nW * kH) {
            uint32_t cc[kH];
            int xs[kH], ys[kH], sa[kH], sb[kH], sc_[kH];
#pragma unroll
            for (int q = 0; q < kH; ++q) {
                const int i = min(i0 + q * nW, n1 - 1);      // past the end: recompute the last one, result unused
                cc[q] = work1[i];
                xs[q] = cc[q] & 0xFFF; ys[q] = (cc[q] >> 12) & 0xFFF;
            }
            harris_sums_warp<kH>(img, lv.pitch, xs, ys, lane, sa, sb, sc_);
            if (lane < kH) {
                int a = sa[0], bq = sb[0], cq = sc_[0];
                uint32_t c = cc[0];
#pragma unroll
                for (int q = 1; q < kH; ++q)
                    if (lane == q) { a = sa[q]; bq = sb[q]; cq = sc_[q]; c = cc[q]; }
                const int i = i0 + lane * nW;
                if (i < n1) {
                    const float r = harris_from_sums(a, bq, cq);
                    pairs[i] = ((unsigned long long)__float_as_uint(r) << 32) | (unsigned long long)(c & 0xFFFFFFu);
                }
            }
        }
    }
    __syncthreads();   // pairs[] written with plain stores by this block, read back below after the barrier

    // ---- pass 2: retainBest(quota) on Harris, order-exact
    const bool inSmem2 = (size_t)n1 * sizeof(unsigned long long) <= (size_t)smemBytes;
    unsigned long long* work2 = inSmem2 ? reinterpret_cast<unsigned long long*>(s_dyn) : pairs;
    if (inSmem2) {
        for (int i = tid; i < n1; i += blockDim.x) work2[i] = pairs[i];
    }
    __syncthreads();
    int n2;
    {
        BlockAcc<unsigned long long, KeyHarris> acc{work2, n1, listL, listR, &sh};
        n2 = retain_best_paired(acc, n1, lv.quota, kSelectSeqTail);
    }
    // cv2 keeps EVERY tie at the Harris boundary (retainBest); room for quota + kFinSlack of them exists per level.  More
    // than that (periodic or saturated images) cannot be represented: the frame is flagged and every reader of its
    // features fails loudly (dvo_get_frame_flags, dvo_pose.frame_flags) instead of returning a truncated keypoint set.
    int flags = nRaw > lv.candCap ? DVO_FRAME_CANDIDATES_TRUNCATED : 0;
    if (n2 > lv.finCap) { n2 = lv.finCap; flags |= DVO_FRAME_TIES_TRUNCATED; }
    uint32_t* finXY = b.finXY + (size_t)slot * g.finPerSlot + lv.finBase;
    float* finResp = b.finResp + (size_t)slot * g.finPerSlot + lv.finBase;
    for (int i = tid; i < n2; i += blockDim.x) {
        unsigned long long v = work2[i];
        finXY[i] = (uint32_t)(v & 0xFFFFFFu);
        finResp[i] = __uint_as_float((uint32_t)(v >> 32));
    }
    if (tid == 0) {
        b.finCount[slot * kMaxLevels + L] = n2;
        int* dbg = b.selDbg + (slot * kMaxLevels + L) * 4;
        dbg[0] = n; dbg[1] = n1; dbg[2] = n2; dbg[3] = flags;
    }
}

// =========================================================================================== A.5 angle + pack
__constant__ int c_umax[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};

__global__ void __launch_bounds__(256) k_angle_pack(OrbGeom g, OrbBuffers b, int slot0) {
    const int slot = slot0 + blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gi = blockIdx.x * 8 + warp;
    // level of keypoint gi: lane i holds level i's survivor count, a three-step scan gives the running totals and one
    // ballot counts the levels that end at or before gi (was an 8-iteration scalar loop in every lane: 20 % of the kernel)
    const int* finCount = b.finCount + slot * kMaxLevels;
    int incl = lane < g.nlevels ? finCount[lane] : 0;
#pragma unroll
    for (int o = 1; o < kMaxLevels; o <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += n;
    }
    const int total = __shfl_sync(0xffffffffu, incl, g.nlevels - 1);
    const int L = __popc(__ballot_sync(0xffffffffu, lane < g.nlevels && gi >= incl));
    const int before = __shfl_sync(0xffffffffu, incl, max(L - 1, 0));
    const int idx = gi - (L > 0 ? before : 0);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        b.featCount[slot] = min(total, g.maxkp);
        int f = total > g.maxkp ? DVO_FRAME_KEYPOINTS_TRUNCATED : 0;
        for (int l = 0; l < g.nlevels; ++l) f |= b.selDbg[(slot * kMaxLevels + l) * 4 + 3];
        b.frameFlags[slot] = f;
    }
    if (gi >= total || gi >= g.maxkp) return;
    const LevelGeom lv = g.lv[L];
    const uint32_t xy = b.finXY[(size_t)slot * g.finPerSlot + lv.finBase + idx];
    const float resp = b.finResp[(size_t)slot * g.finPerSlot + lv.finBase + idx];
    const int x = xy & 0xFFF, y = (xy >> 12) & 0xFFF;
    const uint8_t* img = b.pyr + (size_t)slot * g.slotStride + lv.off;
    // lanes 0..30 own column u = lane - 15
    const int u = lane - 15;
    int m10 = 0, m01 = 0;
    if (lane < 31) {
        // all 31 loads of the column are issued back to back (the whole 31x31 window is inside the level: keypoints keep a
        // 31-px border), rows outside the disc are masked afterwards -- one memory latency per keypoint instead of fifteen
        const int au = abs(u);
        const uint8_t* c = img + (size_t)y * lv.pitch + x + u;
        int vp[16], vm[16];
        vp[0] = c[0];
#pragma unroll
        for (int v = 1; v <= 15; ++v) { vp[v] = c[v * lv.pitch]; vm[v] = c[-v * lv.pitch]; }
        m10 = u * vp[0];
#pragma unroll
        for (int v = 1; v <= 15; ++v) {
            const int in = au <= c_umax[v] ? 1 : 0;
            m10 += in * u * (vp[v] + vm[v]);
            m01 += in * v * (vp[v] - vm[v]);
        }
    }
    m10 = __reduce_add_sync(0xffffffffu, m10);
    m01 = __reduce_add_sync(0xffffffffu, m01);
    if (lane == 0) {
        float ang = fast_atan2_deg((float)m01, (float)m10);
        size_t o = (size_t)slot * g.maxkp + gi;
        b.featPt[o * 2 + 0] = fmul((float)x, lv.scale);
        b.featPt[o * 2 + 1] = fmul((float)y, lv.scale);
        b.featResp[o] = resp;
        b.featAngle[o] = ang;
        b.featOctave[o] = L;
        b.featXY[o] = xy;
    }
}

// =========================================================================================== A.6 blur
// Row pass straight from global memory (aligned 32-bit words, four outputs per thread), float rows in shared memory,
// column pass with 128-bit shared loads and one 32-bit store per four pixels.
__device__ __forceinline__ int reflect101(int p, int n) {
    if (p < 0) p = -p;
    if (p >= n) p = 2 * n - 2 - p;
    return max(0, min(p, n - 1));
}

__global__ void __launch_bounds__(256) k_blur(OrbGeom g, OrbBuffers b, int slot0) {
    __shared__ __align__(16) float hrow[(kTileH + 6) * kTileW];
    const float k0 = __uint_as_float(0x3d8fafb1u), k1 = __uint_as_float(0x3e06387eu), k2 = __uint_as_float(0x3e434a39u),
                k3 = __uint_as_float(0x3e5d4ae0u);
    const uint32_t ti = __ldg(b.tileInfo + blockIdx.x);
    const int L = ti & 15;
    const LevelGeom lv = g.lv[L];
    const int x0 = (int)((ti >> 4) & 0xFFF) * kTileW, y0 = (int)(ti >> 16) * kTileH;
    const int slot = slot0 + blockIdx.y;
    const int tid = threadIdx.x;
    const uint8_t* img = b.pyr + (size_t)slot * g.slotStride + lv.off;
    uint8_t* out = b.blur + (size_t)slot * g.slotStride + lv.off;
    // row pass: s = k0*p[-3]; s = fma(k_i, p_i, s), i = 1..6  (BORDER_REFLECT_101 on both axes).
    // Interior octets (four aligned words cover the 14 taps of 8 outputs) go first with uniform warps; the few quads
    // that touch the left / right image border (at most six per row) are done afterwards by a handful of threads.
    auto row_quad = [&](const float* p, int ry, int xq) {
        float4 o;
        float* ov = reinterpret_cast<float*>(&o);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float s = fmul(k0, p[k]);
            s = ffma(k1, p[k + 1], s);
            s = ffma(k2, p[k + 2], s);
            s = ffma(k3, p[k + 3], s);
            s = ffma(k2, p[k + 4], s);
            s = ffma(k1, p[k + 5], s);
            s = ffma(k0, p[k + 6], s);
            ov[k] = s;
        }
        *reinterpret_cast<float4*>(hrow + ry * kTileW + xq) = o;
    };
    auto interior = [&](int gxo) { return gxo >= 4 && gxo + 11 < lv.w; };      // gxo: first column of an aligned octet
    for (int i = tid; i < (kTileH + 6) * (kTileW / 8); i += 256) {
        const int ry = i / (kTileW / 8), xo = (i - ry * (kTileW / 8)) * 8;
        const int gx = x0 + xo;
        if (!interior(gx)) continue;      // border octet (or outside the level): second loop / never read
        const int gy = reflect101(y0 - 3 + ry, lv.h);
        const uint32_t* wp = reinterpret_cast<const uint32_t*>(img + (size_t)gy * lv.pitch + gx);
        const uint32_t w0 = __ldg(wp - 1), w1 = __ldg(wp), w2 = __ldg(wp + 1), w3 = __ldg(wp + 2);
        float p[14];
        p[0] = (float)((w0 >> 8) & 0xFF); p[1] = (float)((w0 >> 16) & 0xFF); p[2] = (float)(w0 >> 24);
        p[3] = (float)(w1 & 0xFF); p[4] = (float)((w1 >> 8) & 0xFF); p[5] = (float)((w1 >> 16) & 0xFF); p[6] = (float)(w1 >> 24);
        p[7] = (float)(w2 & 0xFF); p[8] = (float)((w2 >> 8) & 0xFF); p[9] = (float)((w2 >> 16) & 0xFF); p[10] = (float)(w2 >> 24);
        p[11] = (float)(w3 & 0xFF); p[12] = (float)((w3 >> 8) & 0xFF); p[13] = (float)((w3 >> 16) & 0xFF);
        row_quad(p, ry, xo);
        row_quad(p + 4, ry, xo + 4);
    }
    for (int i = tid; i < (kTileH + 6) * 6; i += 256) {
        const int ry = i / 6, c = i - ry * 6;
        const int lasto = (lv.w - 1) & ~7;
        const int go = (c >> 1) == 0 ? 0 : ((c >> 1) == 1 ? lasto : lasto - 8);      // candidate border octets
        const int gx = go + 4 * (c & 1);
        if (go < 0 || ((c >> 1) > 0 && go == 0) || gx < x0 || gx >= x0 + kTileW || gx >= lv.w) continue;
        if (interior(go)) continue;                     // an interior octet after all
        const int gy = reflect101(y0 - 3 + ry, lv.h);
        const uint8_t* rowp = img + (size_t)gy * lv.pitch;
        float p[10];
#pragma unroll
        for (int k = 0; k < 10; ++k) p[k] = (float)rowp[reflect101(gx - 3 + k, lv.w)];
        row_quad(p, ry, gx - x0);
    }
    __syncthreads();
    // column pass: a warp owns four output rows of the tile, a lane four adjacent columns; the ten row-pass rows those
    // outputs need are read once into registers (10 LDS.128 for 16 pixels instead of 7 per 4)
    static_assert(kTileW == 128 && kTileH == 32, "column pass layout: 32 lanes x 4 columns, 8 warps x 4 rows");
    {
        const int xq = (tid & 31) * 4, r0 = (tid >> 5) * 4;
        const int x = x0 + xq;
        if (x < lv.w && y0 + r0 < lv.h) {
            float4 w[10];
#pragma unroll
            for (int j = 0; j < 10; ++j) w[j] = *reinterpret_cast<const float4*>(hrow + (r0 + j) * kTileW + xq);
            uint8_t* op = out + (size_t)(y0 + r0) * lv.pitch + x;
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) {
                if (y0 + r0 + rr >= lv.h) break;
                const float4 c = w[rr + 3], u1 = w[rr + 2], d1 = w[rr + 4], u2 = w[rr + 1], d2 = w[rr + 5], u3 = w[rr], d3 = w[rr + 6];
                uint32_t word = 0;
#define DVO_BLUR_COL(F, SH)                                        \
                {                                                  \
                    float s = fmul(k3, c.F);                       \
                    s = ffma(k2, fadd(d1.F, u1.F), s);             \
                    s = ffma(k1, fadd(d2.F, u2.F), s);             \
                    s = ffma(k0, fadd(d3.F, u3.F), s);             \
                    /* 0 <= s <= 255 * (sum of weights)^2 < 255.5: no saturation needed */ \
                    word |= (uint32_t)__float2int_rn(s) << SH;     \
                }
                DVO_BLUR_COL(x, 0) DVO_BLUR_COL(y, 8) DVO_BLUR_COL(z, 16) DVO_BLUR_COL(w, 24)
#undef DVO_BLUR_COL
                *reinterpret_cast<uint32_t*>(op + (size_t)rr * lv.pitch) = word;
            }
        }
    }
}

// =========================================================================================== A.7 rBRIEF
static const signed char h_brief_pattern[256][4] = {
#include "brief_pattern.inc"
};
// Transposed for the warp: entry [j][lane] = pair (lane * 8 + j) packed x0 | y0 << 8 | x1 << 16 | y1 << 24, so the 32 lanes of
// a warp read 128 contiguous bytes per pair index instead of 32 sectors.  Filled by orb_kernels_init().
__device__ uint32_t d_brief_pattern_t[8 * 32];

// cos / sin of every keypoint's angle, one THREAD per keypoint: cv2 steers the pattern with float32 cos/sin of the float32
// angle; computing them in float64 and rounding reproduces those values, and doing it here rather than per warp in
// k_brief keeps the FP64 pipe out of that kernel (32x fewer FP64 instructions).
__global__ void __launch_bounds__(256) k_trig(OrbGeom g, OrbBuffers b, int slot0) {
    const int slot = slot0 + blockIdx.y;
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= b.featCount[slot]) return;
    const size_t o = (size_t)slot * g.maxkp + i;
    const float ang = fmul(b.featAngle[o], (float)(3.14159265358979323846 / 180.0));
    double sd, cd;
    sincos((double)ang, &sd, &cd);
    b.featCS[o * 2] = (float)cd;
    b.featCS[o * 2 + 1] = (float)sd;
}

// Warp per keypoint.  The rotated pattern stays within 18 px of the centre (|(x,y)| <= sqrt(13^2+13^2)), so the warp first
// copies the 37-row x 64-byte window of the smoothed level into shared memory with 128-bit loads (4 per row, 8 rows per
// warp instruction: 5 instructions) and then gathers its 512 samples from there, instead of 512 scattered byte loads.
constexpr int kBriefR = 18, kBriefRows = 2 * kBriefR + 1, kBriefWords = 16;
__global__ void __launch_bounds__(256) k_brief(OrbGeom g, OrbBuffers b, int slot0) {
    __shared__ __align__(16) uint32_t s_patch[8][kBriefRows * kBriefWords];
    const int slot = slot0 + blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gi = blockIdx.x * 8 + warp;
    const int count = b.featCount[slot];
    if (gi >= count) return;
    const size_t o = (size_t)slot * g.maxkp + gi;
    const int L = b.featOctave[o];
    const LevelGeom lv = g.lv[L];
    // cv2: centre = (cvRound(pt.x * (1/scale)), cvRound(pt.y * (1/scale))) on the blurred level
    const float px = b.featPt[o * 2], py = b.featPt[o * 2 + 1];
    const int cx = __float2int_rn(fmul(px, lv.invScale)), cy = __float2int_rn(fmul(py, lv.invScale));
    const float ca = b.featCS[o * 2], sa = b.featCS[o * 2 + 1];
    const uint8_t* img = b.blur + (size_t)slot * g.slotStride + lv.off;
    // window: rows cy-18..cy+18, 64 bytes from xa = (cx-18) & ~15 (16-byte aligned; xa + 63 >= cx + 30)
    const int xa = (cx - kBriefR) & ~15;
    uint32_t* patch = s_patch[warp];
    const bool inside = xa >= 0 && xa + 4 * kBriefWords <= lv.pitch && cy - kBriefR >= 0 && cy + kBriefR < lv.h;
    if (inside) {
        const int sub = lane >> 2, qi = lane & 3;          // 8 rows x 4 uint4 per instruction
        for (int r = 0; r < kBriefRows; r += 8) {
            const int rr = r + sub;
            if (rr < kBriefRows)
                reinterpret_cast<uint4*>(patch + rr * kBriefWords)[qi] =
                    __ldg(reinterpret_cast<const uint4*>(img + (size_t)(cy - kBriefR + rr) * lv.pitch + xa) + qi);
        }
    }
    __syncwarp();
    const uint8_t* pb8 = reinterpret_cast<const uint8_t*>(patch) + kBriefR * (4 * kBriefWords) + (cx - xa);
    const uint8_t* center = img + (size_t)cy * lv.pitch + cx;
    unsigned byte = 0;
    // two separate loops (the test is warp-uniform): with both sources in one loop body the compiler computed the global
    // fallback addresses for every sample and predicated the loads -- a sixth of the kernel's instructions
    auto rotate = [&](uint32_t pq, int& ix0, int& iy0, int& ix1, int& iy1) {
        const float x0f = (float)(signed char)(pq & 0xFF), y0f = (float)(signed char)((pq >> 8) & 0xFF);
        const float x1f = (float)(signed char)((pq >> 16) & 0xFF), y1f = (float)(signed char)(pq >> 24);
        ix0 = __float2int_rn(fsub(fmul(x0f, ca), fmul(y0f, sa)));
        iy0 = __float2int_rn(fadd(fmul(x0f, sa), fmul(y0f, ca)));
        ix1 = __float2int_rn(fsub(fmul(x1f, ca), fmul(y1f, sa)));
        iy1 = __float2int_rn(fadd(fmul(x1f, sa), fmul(y1f, ca)));
    };
    if (inside) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            int ix0, iy0, ix1, iy1;
            rotate(__ldg(&d_brief_pattern_t[j * 32 + lane]), ix0, iy0, ix1, iy1);
            const int t0 = pb8[iy0 * (4 * kBriefWords) + ix0], t1 = pb8[iy1 * (4 * kBriefWords) + ix1];
            byte |= (unsigned)(t0 < t1) << j;
        }
    } else {                                   // never for cv2-selected keypoints (31-px border); keeps odd inputs safe
        for (int j = 0; j < 8; ++j) {
            int ix0, iy0, ix1, iy1;
            rotate(__ldg(&d_brief_pattern_t[j * 32 + lane]), ix0, iy0, ix1, iy1);
            const int t0 = center[iy0 * lv.pitch + ix0], t1 = center[iy1 * lv.pitch + ix1];
            byte |= (unsigned)(t0 < t1) << j;
        }
    }
    b.featDesc[o * 32 + lane] = (uint8_t)byte;
}

// =========================================================================================== launcher
long long orb_launch_count() { return g_launches; }

struct ProfRec { int id; cudaEvent_t a, b; };
static bool g_profOn = false;
// DVO_TIMELINE=1: record the per-kernel events WITHOUT serialising the stages; prof_collect prints start/end of every record
static bool timeline_on() { static int v = -1; if (v < 0) v = getenv("DVO_TIMELINE") ? 1 : 0; return v == 1; }
static std::vector<ProfRec> g_profRecs;
static std::vector<cudaEvent_t> g_profPool;
static cudaEvent_t prof_event() {
    if (!g_profPool.empty()) { cudaEvent_t e = g_profPool.back(); g_profPool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
}
void prof_begin(int id, cudaStream_t st) {
    if (!g_profOn && !timeline_on()) return;
    ProfRec r{id, prof_event(), prof_event()};
    cudaEventRecord(r.a, st);
    g_profRecs.push_back(r);
}
void prof_end(int id, cudaStream_t st) {
    if (!g_profOn && !timeline_on()) return;
    for (size_t i = g_profRecs.size(); i-- > 0;)
        if (g_profRecs[i].id == id) { cudaEventRecord(g_profRecs[i].b, st); return; }
}
void prof_enable(bool on) { g_profOn = on; }
bool prof_enabled() { return g_profOn; }
// total milliseconds and launch-group counts per id since the last collect (synchronises the device)
void prof_collect(double* ms, int* count, int n) {
    cudaDeviceSynchronize();
    for (int i = 0; i < n; ++i) { ms[i] = 0; count[i] = 0; }
    if (timeline_on() && !g_profRecs.empty()) {
        const size_t first = g_profRecs.size() > 60 ? g_profRecs.size() - 60 : 0;
        for (size_t i = first; i < g_profRecs.size(); ++i) {
            float t0 = 0, t1 = 0;
            cudaEventElapsedTime(&t0, g_profRecs[first].a, g_profRecs[i].a);
            cudaEventElapsedTime(&t1, g_profRecs[first].a, g_profRecs[i].b);
            fprintf(stderr, "[timeline] id %2d  reached %8.3f  done %8.3f ms\n", g_profRecs[i].id, t0, t1);
        }
    }
    for (auto& r : g_profRecs) {
        float t = 0;
        if (cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess && r.id < n) { ms[r.id] += t; count[r.id] += 1; }
        g_profPool.push_back(r.a); g_profPool.push_back(r.b);
    }
    g_profRecs.clear();
}

// DVO_DEBUG_SYNC=1: synchronise and report after every kernel (debug only)
bool debug_sync_enabled() {
    static int v = -1;
    if (v < 0) v = getenv("DVO_DEBUG_SYNC") ? 1 : 0;
    return v == 1;
}
void debug_sync(const char* name, cudaStream_t st) {
    if (!debug_sync_enabled()) return;
    cudaError_t e = cudaStreamSynchronize(st);
    if (e == cudaSuccess) e = cudaGetLastError();
    fprintf(stderr, "[dvo] %-16s %s\n", name, e == cudaSuccess ? "ok" : cudaGetErrorString(e));
}

// Frames (device memory, arbitrary pitch) -> level 0 of the slots, one launch for the whole batch.  A CTA moves
// kLoadRows rows (one-row CTAs were launch-rate bound: 0.37 ms for 148 frames, 1 TB/s).
constexpr int kLoadRows = 32;
__global__ void __launch_bounds__(256) k_load_frames(OrbGeom g, OrbBuffers b, const uint8_t* __restrict__ src, size_t pitch,
                                                     size_t frameStride, int slot0, int vec16) {
    const LevelGeom& l0 = g.lv[0];
    const int y0 = blockIdx.x * kLoadRows, f = blockIdx.y;
    const int rows = min(kLoadRows, l0.h - y0);
    const uint8_t* s = src + (size_t)f * frameStride + (size_t)y0 * pitch;
    uint8_t* d = b.pyr + (size_t)(slot0 + f) * g.slotStride + l0.off + (size_t)y0 * l0.pitch;
    const int cpr = (l0.w + 15) >> 4;            // 16-byte chunks per row
    const int total = rows * cpr;
#pragma unroll 4
    for (int i = threadIdx.x; i < total; i += 256) {
        const int r = i / cpr, x = (i - r * cpr) * 16;
        const uint8_t* sp = s + (size_t)r * pitch + x;
        uint8_t* dp = d + (size_t)r * l0.pitch + x;
        if (vec16 && x + 16 <= l0.w) {
            *reinterpret_cast<uint4*>(dp) = __ldg(reinterpret_cast<const uint4*>(sp));
        } else {
            for (int k = 0; k < 16 && x + k < l0.w; ++k) dp[k] = sp[k];
        }
    }
}

void launch_load_frames(const OrbGeom& g, const OrbBuffers& b, const uint8_t* d_src, int n, size_t pitch, size_t frameStride,
                        int slot0, cudaStream_t st) {
    if (n <= 0) return;
    const int vec16 = ((reinterpret_cast<uintptr_t>(d_src) | pitch | frameStride) & 15) == 0 ? 1 : 0;
    dim3 grid((g.lv[0].h + kLoadRows - 1) / kLoadRows, n);
    ProfScope ps_(PF_INGEST, st);
    k_load_frames<<<grid, 256, 0, st>>>(g, b, d_src, pitch, frameStride, slot0, vec16);
    ++g_launches;
}

// =========================================================================================== image ingest
// cv.undistort == remap(src, initUndistortRectifyMap(K, dist, I, newK, size, CV_16SC2), INTER_LINEAR, BORDER_CONSTANT 0):
// the map (float64, rounded to 1/32 px) is built once per calibration; per frame one gather kernel does grey conversion
// (3735 B + 19235 G + 9798 R, >> 15) of the four taps and the int16-weight bilinear blend, straight into pyramid level 0.
// Restated and checked against cv2 in oracle/ingest_np.py.
__global__ void __launch_bounds__(256) k_build_undistort_map(OrbGeom g, IngestBuffers ib, IngestParams p) {
    const int w = g.lv[0].w, h = g.lv[0].h;
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (j >= w || i >= h) return;
    const double di = (double)i, dj = (double)j;
    const double X = dadd(dadd(dmul(di, p.ir[1]), p.ir[2]), dmul(dj, p.ir[0]));
    const double Y = dadd(dadd(dmul(di, p.ir[4]), p.ir[5]), dmul(dj, p.ir[3]));
    const double W = dadd(dadd(dmul(di, p.ir[7]), p.ir[8]), dmul(dj, p.ir[6]));
    const double wv = 1.0 / W;
    const double x = dmul(X, wv), y = dmul(Y, wv);
    const double x2 = dmul(x, x), y2 = dmul(y, y), r2 = dadd(x2, y2), xy2 = dmul(dmul(2.0, x), y);
    const double num = dadd(1.0, dmul(dadd(dmul(dadd(dmul(p.k[4], r2), p.k[1]), r2), p.k[0]), r2));
    const double den = dadd(1.0, dmul(dadd(dmul(dadd(dmul(p.k[7], r2), p.k[6]), r2), p.k[5]), r2));
    const double kr = num / den;
    const double xd = dadd(dadd(dmul(x, kr), dmul(p.k[2], xy2)), dmul(p.k[3], dadd(r2, dmul(2.0, x2))));
    const double yd = dadd(dadd(dmul(y, kr), dmul(p.k[2], dadd(r2, dmul(2.0, y2)))), dmul(p.k[3], xy2));
    const double u = dadd(dmul(p.fx, xd), p.cx), v = dadd(dmul(p.fy, yd), p.cy);
    const int iu = __double2int_rn(dmul(u, 32.0)), iv = __double2int_rn(dmul(v, 32.0));
    const int sx = max(-32768, min(32767, iu >> 5)), sy = max(-32768, min(32767, iv >> 5));
    const uint2 wq = __ldg(ib.wtab + (((iv & 31) << 5) | (iu & 31)));      // the four tap weights ride along in the map entry
    ib.map[(size_t)i * w + j] = make_uint4((uint32_t)(uint16_t)(short)sx | ((uint32_t)(uint16_t)(short)sy << 16), wq.x, wq.y, 0u);
}

template <int kChannels>
__device__ __forceinline__ uint32_t ingest_tap(const uint8_t* frame, size_t pitch, int w, int h, int y, int x) {
    if ((unsigned)x >= (unsigned)w || (unsigned)y >= (unsigned)h) return 0;      // BORDER_CONSTANT, value 0
    const uint8_t* q = frame + (size_t)y * pitch + (size_t)x * kChannels;
    if (kChannels == 1) return q[0];
    return (uint32_t)((3735 * (int)q[0] + 19235 * (int)q[1] + 9798 * (int)q[2] + (1 << 14)) >> 15);
}

template <int kChannels>
__global__ void __launch_bounds__(256) k_ingest(OrbGeom g, OrbBuffers b, IngestBuffers ib, const uint8_t* __restrict__ src, size_t pitch,
                                                size_t frameStride, int slot0) {
    const LevelGeom& l0 = g.lv[0];
    const int y = blockIdx.y, f = blockIdx.z;
    const int x4 = (blockIdx.x * 256 + threadIdx.x) * 4;
    if (x4 >= l0.w) return;
    const uint8_t* frame = src + (size_t)f * frameStride;
    uint32_t out = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int x = x4 + k;
        if (x >= l0.w) break;
        const uint4 m = __ldg(ib.map + (size_t)y * l0.w + x);
        const int sx = (int)(short)(m.x & 0xFFFFu), sy = (int)(short)(m.x >> 16);
        // two taps per DP2A: (w00, w01) x (p00, p01) and (w10, w11) x (p10, p11); weights are uint16 (0 .. 32768)
        const uint32_t t0 = ingest_tap<kChannels>(frame, pitch, l0.w, l0.h, sy, sx) |
                            (ingest_tap<kChannels>(frame, pitch, l0.w, l0.h, sy, sx + 1) << 8);
        const uint32_t t1 = ingest_tap<kChannels>(frame, pitch, l0.w, l0.h, sy + 1, sx) |
                            (ingest_tap<kChannels>(frame, pitch, l0.w, l0.h, sy + 1, sx + 1) << 8);
        const int acc = (int)__dp2a_lo(m.y, t0, __dp2a_lo(m.z, t1, 0u));
        const int v = max(0, min(255, (acc + (1 << 14)) >> 15));
        out |= (uint32_t)v << (8 * k);
    }
    *reinterpret_cast<uint32_t*>(b.pyr + (size_t)(slot0 + f) * g.slotStride + l0.off + (size_t)y * l0.pitch + x4) = out;
}

void launch_build_undistort_map(const OrbGeom& g, const IngestBuffers& ib, const IngestParams& prm, cudaStream_t st) {
    dim3 grid((g.lv[0].w + 255) / 256, g.lv[0].h);
    k_build_undistort_map<<<grid, 256, 0, st>>>(g, ib, prm);
    ++g_launches;
}

void launch_ingest(const OrbGeom& g, const OrbBuffers& b, const IngestBuffers& ib, const uint8_t* d_src, int n, size_t pitch,
                   size_t frameStride, int slot0, cudaStream_t st) {
    if (n <= 0) return;
    dim3 grid((g.lv[0].w + 1023) / 1024, g.lv[0].h, n);
    ProfScope ps_(PF_INGEST, st);
    if (ib.channels == 3) k_ingest<3><<<grid, 256, 0, st>>>(g, b, ib, d_src, pitch, frameStride, slot0);
    else k_ingest<1><<<grid, 256, 0, st>>>(g, b, ib, d_src, pitch, frameStride, slot0);
    ++g_launches;
}

// The ORB stage in two halves so that a sequence runner can put them on different streams: the IMAGE half (pyramid + FAST: dense,
// issue-bound, small CTAs) reads only the frame and writes the pyramid and the tile lists; the KEYPOINT half (gather, select,
// angle, blur, brief) is mostly latency-bound.
void launch_orb(const OrbGeom& g, const OrbBuffers& b, const TensorMaps* tmaps, bool useTma, int slot0, int nSlots,
                cudaStream_t st, const SideStreams* ss) {
    launch_orb_image(g, b, tmaps, useTma, slot0, nSlots, st);
    launch_orb_keypoints(g, b, slot0, nSlots, st, ss);
}

void launch_orb_image(const OrbGeom& g, const OrbBuffers& b, const TensorMaps* tmaps, bool useTma, int slot0, int nSlots, cudaStream_t st) {
    if (nSlots <= 0) return;
    for (int L = 1; L < g.nlevels; ++L) {
        const long long ctas8 = (long long)((g.lv[L].w + 127) / 128) * ((g.lv[L].h + 63) / 64) * nSlots;
        const int rows = ctas8 >= 148 * 16 ? 8 : 2;
        dim3 grid((g.lv[L].w + 127) / 128, (g.lv[L].h + 8 * rows - 1) / (8 * rows), nSlots);
        ProfScope ps_(PF_PYR, st);
        if (useTma) k_pyr_down<true><<<grid, dim3(32, 8), 0, st>>>(g, b, tmaps->pyrSrc[rows >= 8 ? 0 : 1][L - 1], L, slot0, rows);
        else k_pyr_down<false><<<grid, dim3(32, 8), 0, st>>>(g, b, tmaps->pyr[0], L, slot0, rows);
        ++g_launches;
        debug_sync("k_pyr_down", st);
    }
    {
        dim3 grid(g.tilesPerFrame, nSlots);
        ProfScope ps_(PF_FAST, st);
        if (useTma) k_fast_nms<true><<<grid, 256, 0, st>>>(g, b, *tmaps, slot0);
        else k_fast_nms<false><<<grid, 256, 0, st>>>(g, b, *tmaps, slot0);
        ++g_launches;
        debug_sync("k_fast_nms", st);
    }
}

void launch_orb_keypoints(const OrbGeom& g, const OrbBuffers& b, int slot0, int nSlots, cudaStream_t st, const SideStreams* ss) {
    if (nSlots <= 0) return;
    const bool fork = ss != nullptr && ss->side != nullptr;
    if (fork) {      // k_blur beside compact -> select -> angle
        cudaEventRecord(ss->evFork, st);
        cudaStreamWaitEvent(ss->side, ss->evFork, 0);
        { ProfScope ps_(PF_BLUR, ss->side); k_blur<<<dim3(g.tilesPerFrame, nSlots), 256, 0, ss->side>>>(g, b, slot0); }
        ++g_launches;
        cudaEventRecord(ss->evJoin, ss->side);
    }
    { ProfScope ps_(PF_COMPACT, st); k_gather<<<dim3(g.rowBlocksPerFrame, nSlots), 256, 0, st>>>(g, b, slot0); }
    ++g_launches;
    debug_sync("k_gather", st);
    {
        // Large and small levels as two launches on two streams, so that the short small-level CTAs do not queue behind the long ones.
        // Launch shape by the number of frames in flight (see kSelectBatch*).
        int nBig = 0;
        while (nBig < g.nlevels && (long long)g.lv[nBig].w * g.lv[nBig].h > kSelectBigLevelPixels) ++nBig;
        const bool split = fork && ss->side2 != nullptr && nBig > 0 && nBig < g.nlevels;
        ProfScope ps_(PF_SELECT, st);
        const bool batch = nSlots >= kSelectBatchMinSlots;
        const int bigThreads = batch ? kSelectBatchThreads : kSelectThreads, bigSmem = batch ? kSelectBatchSmemBytes : kSelectSmemBytes;
        const int smallThreads = batch ? kSelectBatchThreads : 512, smallSmem = batch ? kSelectBatchSmemBytes : kSelectSmallSmemBytes;
        if (split) {
            cudaEventRecord(ss->evFork2, st);
            cudaStreamWaitEvent(ss->side2, ss->evFork2, 0);
            k_select<<<dim3(nSlots, g.nlevels - nBig), smallThreads, smallSmem, ss->side2>>>(g, b, slot0, smallSmem, nBig);
            cudaEventRecord(ss->evJoin2, ss->side2);
            k_select<<<dim3(nSlots, nBig), bigThreads, bigSmem, st>>>(g, b, slot0, bigSmem, 0);
            cudaStreamWaitEvent(st, ss->evJoin2, 0);
            ++g_launches;
        } else {
            k_select<<<dim3(nSlots, g.nlevels), bigThreads, bigSmem, st>>>(g, b, slot0, bigSmem, 0);
        }
    }
    ++g_launches;
    debug_sync("k_select", st);
    {
        ProfScope ps_(PF_ANGLE, st);
        k_angle_pack<<<dim3((g.maxkp + 7) / 8, nSlots), 256, 0, st>>>(g, b, slot0);
        k_trig<<<dim3((g.maxkp + 255) / 256, nSlots), 256, 0, st>>>(g, b, slot0);
    }
    g_launches += 2;
    debug_sync("k_angle_pack", st);
    if (fork) {
        cudaStreamWaitEvent(st, ss->evJoin, 0);
    } else {
        { ProfScope ps_(PF_BLUR, st); k_blur<<<dim3(g.tilesPerFrame, nSlots), 256, 0, st>>>(g, b, slot0); }
        ++g_launches;
    }
    debug_sync("k_blur", st);
    { ProfScope ps_(PF_BRIEF, st); k_brief<<<dim3((g.maxkp + 7) / 8, nSlots), 256, 0, st>>>(g, b, slot0); }
    ++g_launches;
    debug_sync("k_brief", st);
}

cudaError_t orb_kernels_init() {
    cudaError_t e = cudaFuncSetAttribute(k_select, cudaFuncAttributeMaxDynamicSharedMemorySize, kSelectSmemBytes);
    if (e != cudaSuccess) return e;
    uint32_t t[8 * 32];
    for (int lane = 0; lane < 32; ++lane)
        for (int j = 0; j < 8; ++j) {
            const signed char* p = h_brief_pattern[lane * 8 + j];
            t[j * 32 + lane] = (uint32_t)(uint8_t)p[0] | ((uint32_t)(uint8_t)p[1] << 8) | ((uint32_t)(uint8_t)p[2] << 16) |
                               ((uint32_t)(uint8_t)p[3] << 24);
        }
    return cudaMemcpyToSymbol(d_brief_pattern_t, t, sizeof t);
}

}  // namespace dvo
