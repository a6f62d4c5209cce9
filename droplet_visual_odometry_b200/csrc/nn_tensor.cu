// Brute-force Hamming nearest neighbour on the 5th-generation tensor cores (cross-check matcher, SURVEY A.8;
// replaces cv.BFMatcher(NORM_HAMMING, crossCheck=True).match -- visual_odometry_v3.py:219).
//
// A 256-bit ORB descriptor becomes 256 int8 values, +1 for a clear bit and -1 for a set bit.  For two descriptors
//   dot(a, b) = (#equal bits) - (#different bits) = 256 - 2 * hamming(a, b),
// so the whole 2000 x 2000 distance matrix of a frame pair is one int8 GEMM with K = 256 and exact int32 accumulators,
// and the XOR+POPC kernel's bound (POPC runs on the 16-lane XU pipe: 0.79 ms per 148 pairs at best) disappears.
//
//   k_expand_desc   bits -> +-1 bytes, 256 B per descriptor, written in the order the MMA reads its operands
//   k_nn_tensor     persistent, one CTA per SM, warp-specialised (320 threads):
//       warps 0-7   epilogue.  Warp w reads TMEM lanes 32 (w % 4) .. +31 (a warp can only reach its own lane quarter), i.e.
//                   thread t of the quarter owns accumulator row t, and columns [128 (w / 4), +128) of every tile:
//                   tcgen05.ld 32 columns at a time (the next load in flight while one chunk is reduced), one IMAD per
//                   element builds the packed key (distance << 16 | column) and VIMNMX3 keeps the row minimum; ties go to
//                   the lowest column, as cv2's batchDistance does.  The two partial minima of a row meet in shared
//                   memory once per work item.  No cross-thread reduction inside a tile and no atomics: the reverse
//                   direction (train -> query) is simply the transposed product, scheduled as its own work items.
//                   kSecond (ratio matcher): forward rows also keep the runner-up key (knnMatch k = 2).
//       warp 8      one lane issues tcgen05.mma.kind::i8 (M 128 x N 256 x K 32, eight per tile) into a double-buffered
//                   TMEM accumulator (2 x 256 columns) and commits to the mbarriers that free the operand stages
//       warp 9      producer: one lane streams the operand tiles with cp.async.bulk (mbarrier complete_tx).  k_expand_desc
//                   already stores the rows in the canonical no-swizzle K-major core-matrix order (8 rows x 16 B
//                   contiguous; K-adjacent core matrices 128 B apart, 8-row groups 2 KB apart): a tile is one block
//   A work item is (pair, direction, 128-row block); it streams every 256-column tile of the other frame past its rows.
//   Measured (148 pairs of 2000 x 2000): 0.21 ms + 0.03 ms expansion = 2.55 int8 POP/s; bound by L2 -> SM operand traffic
//   (2.6 GB per batch).  Every mbarrier wait is a bounded spin that traps, so a protocol error fails the launch loudly.
#include "dvo_internal.cuh"

namespace dvo {

namespace {

constexpr int kRowBytes = 256;                 // one expanded descriptor
constexpr int kTileM = 128, kTileN = 256;
constexpr int kABytes = kTileM * kRowBytes;    // 32 KB
constexpr int kBBytes = kTileN * kRowBytes;    // 64 KB
constexpr int kAStages = 2, kBStages = 2;
constexpr int kEpiWarps = 8;                   // two warps per TMEM lane quarter, each takes half of a tile's columns
constexpr int kEpiCols = kTileN / (kEpiWarps / 4);
constexpr int kThreads = (kEpiWarps + 2) * 32;                  // epilogue, MMA warp, producer warp
constexpr int kBulkBytes = 16384;              // one cp.async.bulk
constexpr uint32_t kLbo = 128, kSbo = 2048;    // bytes: K-adjacent core matrices / 8-row groups
constexpr size_t kSmemBytes = (size_t)kAStages * kABytes + (size_t)kBStages * kBBytes + 1024 /*alignment slack*/ + 256;

// instruction descriptor, kind::i8: D s32 (bits 4-5 = 2), A and B signed 8-bit (bits 7-9, 10-12 = 1), both K-major,
// N >> 3 at bits 17-22, M >> 4 at bits 24-28
constexpr uint32_t kIdesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kTileN >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}
// Bounded spin: a protocol error traps (the launch fails loudly) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spin = 0; !done; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (spin > (1u << 24)) __trap();
    }
}

__device__ __forceinline__ uint64_t smem_desc(uint32_t addr) {
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(kLbo >> 4) << 16) | ((uint64_t)(kSbo >> 4) << 32) | (1ull << 46);
}

__device__ __forceinline__ void mma_i8(uint32_t tmemD, uint64_t descA, uint64_t descB, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmemD), "l"(descA), "l"(descB), "r"(kIdesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, int (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
// The registers are in/out operands so that nothing reading them is scheduled above the wait.
__device__ __forceinline__ void tmem_ld_wait(int (&v)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]),
                   "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]), "+r"(v[16]),
                   "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]), "+r"(v[24]),
                   "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
                 :
                 : "memory");
}

// Packed key of the chunk's best column: (256 - dot) << 15 == distance << 16, plus the column number within the chunk.
template <bool kMasked>
__device__ __forceinline__ int chunk_min(const int (&v)[32], int negScale, int lim) {
    int m = 0x7fffffff;
#pragma unroll
    for (int c = 0; c < 32; c += 2) {
        int k0 = v[c] * negScale + ((256 << 15) + c);
        int k1 = v[c + 1] * negScale + ((256 << 15) + c + 1);
        if (kMasked) {
            if (c >= lim) k0 = 0x7fffffff;
            if (c + 1 >= lim) k1 = 0x7fffffff;
        }
        m = __vimin3_s32(m, k0, k1);
    }
    return m;
}

// Two smallest keys of the chunk (ratio matcher: knnMatch(k=2)); different columns have different keys, so the pair is
// the two nearest neighbours in cv2's tie order.
template <bool kMasked>
__device__ __forceinline__ void chunk_min2(const int (&v)[32], int negScale, int lim, int& m1, int& m2) {
    m1 = 0x7fffffff; m2 = 0x7fffffff;
#pragma unroll
    for (int c = 0; c < 32; ++c) {
        int k = v[c] * negScale + ((256 << 15) + c);
        if (kMasked && c >= lim) k = 0x7fffffff;
        const int t = max(m1, k);
        m1 = min(m1, k);
        m2 = min(m2, t);
    }
}

// One 32-column chunk into the running row minimum (and runner-up).  lim = valid columns left at the chunk's start.
template <bool kTwo>
__device__ __forceinline__ void epi_chunk(const int (&v)[32], int lim, int colBase, int negScale, int& best, int& second) {
    if (!kTwo) {
        if (lim >= 32) best = min(best, chunk_min<false>(v, negScale, 32) + colBase);
        else if (lim > 0) best = min(best, chunk_min<true>(v, negScale, lim) + colBase);
    } else if (lim > 0) {
        int a1, a2;
        if (lim >= 32) chunk_min2<false>(v, negScale, 32, a1, a2); else chunk_min2<true>(v, negScale, lim, a1, a2);
        a1 += colBase;                                   // a chunk always has its first column: a1 is a real key
        if (a2 != 0x7fffffff) a2 += colBase;
        second = min(max(best, a1), min(second, a2));
        best = min(best, a1);
    }
}

// This thread's row of one accumulator stage, kEpiCols columns from taddr: tcgen05.ld of chunk k+1 is in flight while
// chunk k is reduced.
template <bool kTwo>
__device__ __forceinline__ void epi_tile(uint32_t taddr, int valid, int col0, int negScale, int& best, int& second) {
    int va[32], vb[32];
    tmem_ld32(taddr, va);
#pragma unroll
    for (int k = 0; k < kEpiCols / 32; k += 2) {
        tmem_ld_wait(va);
        tmem_ld32(taddr + (k + 1) * 32, vb);
        epi_chunk<kTwo>(va, valid - k * 32, col0 + k * 32, negScale, best, second);
        tmem_ld_wait(vb);
        if (k + 2 < kEpiCols / 32) tmem_ld32(taddr + (k + 2) * 32, va);
        epi_chunk<kTwo>(vb, valid - (k + 1) * 32, col0 + (k + 1) * 32, negScale, best, second);
    }
}

struct Item { int pair, dir, mt, nX, nY, slotX, slotY; bool rows, work; };

__device__ __forceinline__ Item decode_item(int item, int mTiles, const int* featCount, int slotA0, int maxkp) {
    Item it;
    const int perPair = 2 * mTiles;
    it.pair = item / perPair;
    const int rem = item - it.pair * perPair;
    it.dir = rem / mTiles;
    it.mt = rem - it.dir * mTiles;
    const int nA = min(featCount[slotA0 + it.pair], maxkp), nB = min(featCount[slotA0 + it.pair + 1], maxkp);
    it.nX = it.dir ? nB : nA;
    it.nY = it.dir ? nA : nB;
    it.slotX = it.pair + it.dir;          // local slot in the expanded buffer
    it.slotY = it.pair + 1 - it.dir;
    it.rows = it.mt * kTileM < it.nX;
    it.work = it.rows && it.nY > 0;
    return it;
}

}  // namespace

// bits -> bytes: bit clear -> 0x01 (+1), bit set -> 0xFF (-1).  One thread per 16 output bytes, written in the order the
// MMA's no-swizzle K-major descriptor reads them: [8-row group][16-byte K chunk][row in group][16 B], so that a tile of
// 128 or 256 rows is one contiguous block for cp.async.bulk.
__global__ void __launch_bounds__(256) k_expand_desc(OrbGeom og, OrbBuffers ob, PairGeom pg, PairBuffers pb, int slotA0, int nSlots) {
    const size_t id = (size_t)blockIdx.x * 256 + threadIdx.x;
    const int r8 = (int)(id & 7), chunk = (int)((id >> 3) & 15);
    const size_t grpId = id >> 7;
    const int groups = pb.descXRows >> 3;
    const int ls = (int)(grpId / groups), g = (int)(grpId - (size_t)ls * groups);
    if (ls >= nSlots) return;
    const int slot = slotA0 + ls, r = g * 8 + r8;
    if (r >= min(ob.featCount[slot], pg.maxkp)) return;
    const uint32_t bits = reinterpret_cast<const uint16_t*>(ob.featDesc + ((size_t)slot * og.maxkp + r) * 32)[chunk];
    uint32_t w[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint32_t nib = (bits >> (4 * q)) & 15u;
        const uint32_t ones = (nib * 0x00204081u) & 0x01010101u;     // bit k of the nibble -> byte k
        w[q] = (ones * 0xFFu) | 0x01010101u;
    }
    reinterpret_cast<uint4*>(pb.descX + (size_t)ls * pb.descXRows * kRowBytes)[(size_t)g * 128 + chunk * 8 + r8] =
        make_uint4(w[0], w[1], w[2], w[3]);
}

template <bool kSecond>      // kSecond: forward rows also keep the runner-up distance (ratio matcher)
__global__ void __launch_bounds__(kThreads, 1)
k_nn_tensor(OrbBuffers ob, PairGeom pg, PairBuffers pb, int slotA0, int pair0, int nPairs, int mTiles, int negScale) {
    extern __shared__ uint8_t smemRaw[];
    const uint32_t base = (smem_u32(smemRaw) + 1023u) & ~1023u;
    const uint32_t sA = base, sB = base + kAStages * kABytes;
    const uint32_t sBar = sB + kBStages * kBBytes;
    // barriers: fullA[2] emptyA[2] fullB[2] emptyB[2] accFull[2] accEmpty[2]
    const uint32_t fullA = sBar, emptyA = sBar + 16, fullB = sBar + 32, emptyB = sBar + 48, accFull = sBar + 64, accEmpty = sBar + 80;
    const uint32_t sTmem = sBar + 96;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __shared__ int sPart[2 * kTileM];            // partial (best, runner-up) of the upper column half

    if (threadIdx.x == 0) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(fullA + 8 * s, 1);
            mbar_init(emptyA + 8 * s, 1);
            mbar_init(fullB + 8 * s, 1);
            mbar_init(emptyB + 8 * s, 1);
            mbar_init(accFull + 8 * s, 1);
            mbar_init(accEmpty + 8 * s, kEpiWarps * 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kEpiWarps) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(sTmem) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t tmemBase;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmemBase) : "r"(sTmem));

    const int total = nPairs * 2 * mTiles;
    const size_t slotStride = (size_t)pb.descXRows * kRowBytes;

    if (warp < kEpiWarps) {
        // ---------------------------------------------------------------- epilogue
        // warp w reads TMEM lanes 32 (w % 4) .. +31 (the hardware's lane quarter of a warp) and columns
        // [kEpiCols (w / 4), +kEpiCols) of every tile; the two partial minima of a row meet in shared memory per item
        uint32_t accUse = 0;
        const int quarter = warp & 3, half = warp >> 2;
        const int rowInTile = quarter * 32 + lane;
        const uint32_t laneBase = ((uint32_t)(quarter * 32)) << 16;
        for (int item = blockIdx.x; item < total; item += gridDim.x) {
            const Item it = decode_item(item, mTiles, ob.featCount, slotA0, pg.maxkp);
            if (!it.rows) continue;
            const int row = it.mt * kTileM + rowInTile;
            int best = 0x7fffffff, second = 0x7fffffff;
            const bool two = kSecond && it.dir == 0;
            if (it.work) {
                const int nTiles = (it.nY + kTileN - 1) / kTileN;
                for (int nt = 0; nt < nTiles; ++nt, ++accUse) {
                    const uint32_t cs = accUse & 1;
                    mbar_wait(accFull + 8 * cs, (accUse >> 1) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t taddr = tmemBase + laneBase + cs * kTileN + half * kEpiCols;
                    const int col0 = nt * kTileN + half * kEpiCols;
                    const int valid = it.nY - col0;          // may be <= 0 for the upper half of the last tile
                    if (kSecond && two) epi_tile<true>(taddr, valid, col0, negScale, best, second);
                    else epi_tile<false>(taddr, valid, col0, negScale, best, second);
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    mbar_arrive(accEmpty + 8 * cs);
                }
            }
            if (kEpiWarps > 4) {
                if (half == 1) { sPart[rowInTile] = best; sPart[kTileM + rowInTile] = second; }
                asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");
                if (half == 0) {
                    const int b1 = sPart[rowInTile], b2 = sPart[kTileM + rowInTile];
                    second = min(max(best, b1), min(second, b2));
                    best = min(best, b1);
                }
                asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");      // sPart is free for the next item
            }
            if (half == 0 && row < it.nX) {
                const size_t o = ((size_t)(pair0 + it.pair) * 2 + it.dir) * pg.maxkp + row;
                if (two) {            // k_match_sort's ratio path: plain index, distance, runner-up distance
                    pb.nnIdx[o] = best == 0x7fffffff ? -1 : (best & 0xFFFF);
                    pb.nnDist[o] = best == 0x7fffffff ? 0x7fffffff : (best >> 16);
                    pb.nn2Dist[(size_t)(pair0 + it.pair) * pg.maxkp + row] = second == 0x7fffffff ? 0x7fffffff : (second >> 16);
                } else {
                    reinterpret_cast<uint32_t*>(pb.nnIdx)[o] = best == 0x7fffffff ? 0xFFFFFFFFu : (uint32_t)best;
                }
            }
        }
    } else if (warp == kEpiWarps) {
        // ---------------------------------------------------------------- MMA issue (one lane)
        if (lane == 0) {
            uint32_t aUse = 0, bUse = 0, accUse = 0;
            for (int item = blockIdx.x; item < total; item += gridDim.x) {
                const Item it = decode_item(item, mTiles, ob.featCount, slotA0, pg.maxkp);
                if (!it.work) continue;
                const uint32_t as = aUse & 1;
                mbar_wait(fullA + 8 * as, (aUse >> 1) & 1);
                const int nTiles = (it.nY + kTileN - 1) / kTileN;
                for (int nt = 0; nt < nTiles; ++nt, ++bUse, ++accUse) {
                    const uint32_t bs = bUse & 1, cs = accUse & 1;
                    mbar_wait(fullB + 8 * bs, (bUse >> 1) & 1);
                    mbar_wait(accEmpty + 8 * cs, ((accUse >> 1) & 1) ^ 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t d = tmemBase + cs * kTileN;
#pragma unroll
                    for (int k = 0; k < kRowBytes / 32; ++k)
                        mma_i8(d, smem_desc(sA + as * kABytes + k * 2 * kLbo), smem_desc(sB + bs * kBBytes + k * 2 * kLbo), k > 0);
                    mma_commit(emptyB + 8 * bs);
                    mma_commit(accFull + 8 * cs);
                }
                mma_commit(emptyA + 8 * as);
                ++aUse;
            }
        }
        __syncwarp();
    } else {
        // ---------------------------------------------------------------- producer (one lane): bulk copies
        // k_expand_desc stores the descriptors in the core-matrix order the MMA reads, so a tile is one contiguous block
        if (lane == 0) {
            uint32_t aUse = 0, bUse = 0;
            for (int item = blockIdx.x; item < total; item += gridDim.x) {
                const Item it = decode_item(item, mTiles, ob.featCount, slotA0, pg.maxkp);
                if (!it.work) continue;
                {
                    const uint32_t as = aUse & 1;
                    mbar_wait(emptyA + 8 * as, ((aUse >> 1) & 1) ^ 1);
                    mbar_expect_tx(fullA + 8 * as, kABytes);
                    const int8_t* src = pb.descX + (size_t)it.slotX * slotStride + (size_t)it.mt * kABytes;
                    for (int q = 0; q < kABytes; q += kBulkBytes) bulk_load(sA + as * kABytes + q, src + q, kBulkBytes, fullA + 8 * as);
                    ++aUse;
                }
                const int nTiles = (it.nY + kTileN - 1) / kTileN;
                for (int nt = 0; nt < nTiles; ++nt, ++bUse) {
                    const uint32_t bs = bUse & 1;
                    mbar_wait(emptyB + 8 * bs, ((bUse >> 1) & 1) ^ 1);
                    mbar_expect_tx(fullB + 8 * bs, kBBytes);
                    const int8_t* src = pb.descX + (size_t)it.slotY * slotStride + (size_t)nt * kBBytes;
                    for (int q = 0; q < kBBytes; q += kBulkBytes) bulk_load(sB + bs * kBBytes + q, src + q, kBulkBytes, fullB + 8 * bs);
                }
            }
        }
        __syncwarp();
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (warp == kEpiWarps) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmemBase) : "memory");
}

// ---- int8 tensor-pipe rate (bench.py roofline denominator): every SM issues `iters` M128 x N256 x K32 MMAs back to back from
// one resident pair of operand slabs into two alternating TMEM accumulators; no loads, no epilogue.
__global__ void __launch_bounds__(128, 1) k_int8_mma_rate(int iters) {
    extern __shared__ uint8_t smemRaw[];
    const uint32_t base = (smem_u32(smemRaw) + 1023u) & ~1023u;
    const uint32_t sA = base, sB = base + kABytes, sBar = sB + kBBytes, sTmem = sBar + 16;
    for (int i = threadIdx.x; i < (kABytes + kBBytes) / 4; i += blockDim.x)
        reinterpret_cast<uint32_t*>(smemRaw + (base - smem_u32(smemRaw)))[i] = 0x01FF01FFu;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(sBar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(sTmem) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t tmemBase;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmemBase) : "r"(sTmem));
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < iters; ++i)
            mma_i8(tmemBase + (i & 1) * kTileN, smem_desc(sA + (i & 7) * 2 * kLbo), smem_desc(sB + (i & 7) * 2 * kLbo), i > 1);
        mma_commit(sBar);
        mbar_wait(sBar, 0);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmemBase) : "memory");
}

double nn_tensor_peak_tops(int numSms, int iters) {
    const int smem = kABytes + kBBytes + 1024 + 64;
    if (cudaFuncSetAttribute(k_int8_mma_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return 0.0;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 0;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        k_int8_mma_rate<<<numSms, 128, smem>>>(iters);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { best = 0; break; }
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double ops = 2.0 * kTileM * kTileN * 32 * (double)iters * numSms;
        if (rep > 0 && ms > 0) best = best > ops / (ms * 1e-3) ? best : ops / (ms * 1e-3);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return best;
}

int nn_tensor_rows(int maxkp) { return ((maxkp + kTileN - 1) / kTileN) * kTileN; }

cudaError_t nn_tensor_init() {
    cudaError_t e = cudaFuncSetAttribute(k_nn_tensor<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_nn_tensor<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
}

void launch_nn_tensor(const OrbGeom& og, const OrbBuffers& ob, const PairGeom& pg, const PairBuffers& pb, int slotA0, int pair0,
                      int nPairs, int numSms, cudaStream_t st) {
    const int nSlots = nPairs + 1;
    const size_t pieces = (size_t)nSlots * pb.descXRows * 16;
    k_expand_desc<<<(unsigned)((pieces + 255) / 256), 256, 0, st>>>(og, ob, pg, pb, slotA0, nSlots);
    const int mTiles = (pg.maxkp + kTileM - 1) / kTileM;
    const int total = nPairs * 2 * mTiles;
    if (pg.matcher == DVO_MATCH_KNN_RATIO)
        k_nn_tensor<true><<<std::min(total, numSms), kThreads, kSmemBytes, st>>>(ob, pg, pb, slotA0, pair0, nPairs, mTiles, -32768);
    else
        k_nn_tensor<false><<<std::min(total, numSms), kThreads, kSmemBytes, st>>>(ob, pg, pb, slotA0, pair0, nPairs, mTiles, -32768);
}

}  // namespace dvo
