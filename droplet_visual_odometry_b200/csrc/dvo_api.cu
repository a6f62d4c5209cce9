// libdvo C ABI (include/dvo.h): context, HBM buffers, TMA tensor maps, stage taps, sequence runner.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include "dvo_internal.cuh"
#include "mathcore.cuh"

namespace dvo {
long long orb_launch_count();
long long pair_launch_count();
cudaError_t orb_kernels_init();
cudaError_t pair_kernels_init(int sortBytes);
cudaError_t pair_kernels_init_exhaustive(int rngCount);
void prof_enable(bool on);
bool prof_enabled();
void prof_collect(double* ms, int* count, int n);
}  // namespace dvo

using namespace dvo;

struct dvo_ctx {
    dvo_config cfg;
    int device = 0;
    std::string err;
    OrbGeom og{};
    OrbBuffers ob{};
    PairGeom pg{};
    PairBuffers pb{};
    TensorMaps tmaps{};
    bool useTma = false;
    int nSlots = 0, nPairs = 0;
    std::vector<void*> allocs;
    uint32_t* d_resizeTab = nullptr;
    double* d_K = nullptr;
    dvo_pose* h_poseStage = nullptr;   // pinned staging for the host sequence runner
    size_t poseStageCap = 0;
    uint8_t* h_frameStage = nullptr;
    long long launchBase = 0;
    int carrySlot = -1;                // slot holding the last frame of the previous dvo_sequence_step
    SideStreams ss;
    size_t stageBytes[2] = {0, 0};
    IngestBuffers ingest;              // undistort map + weights; channels == 0: frames arrive grey and undistorted
    // Two-lane sequence pipeline (cfg.pipeline): lane 1 is a second set of ORB buffers so that the ORB stage of batch
    // s+1 (stream sOrb) runs while the pair stage of batch s (stream sPair) still reads batch s's features.
    OrbBuffers ob1{};
    TensorMaps tmaps1{};
    bool lane1Ready = false;
    std::vector<void*> lane1Allocs;
    cudaStream_t sOrb = nullptr, sPair = nullptr, sImg = nullptr;
    cudaEvent_t evCall = nullptr, evOrbDone[2] = {nullptr, nullptr}, evPairsDone[2] = {nullptr, nullptr},
                evCarryCopied[2] = {nullptr, nullptr}, evImgDone[2] = {nullptr, nullptr};
    bool pairsPending[2] = {false, false}, carryPending[2] = {false, false};
    int carryLane = 0, lastLane = -1;
    bool pipeOutstanding = false;      // sPair holds work the caller's stream has not been joined with yet
};

#define CK(call)                                                                                     \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess) {                                                                     \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_);                           \
            return DVO_E_CUDA;                                                                       \
        }                                                                                            \
    } while (0)

static size_t round_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

template <class T>
static int dalloc(dvo_ctx* ctx, T** p, size_t count) {
    void* q = nullptr;
    size_t bytes = std::max<size_t>(count * sizeof(T), 256);
    cudaError_t e = cudaMalloc(&q, bytes);
    if (e != cudaSuccess) {
        ctx->err = std::string("cudaMalloc: ") + cudaGetErrorString(e);
        return DVO_E_CUDA;
    }
    cudaMemset(q, 0, bytes);
    ctx->allocs.push_back(q);
    *p = reinterpret_cast<T*>(q);
    return 0;
}

static int alloc_orb_buffers(dvo_ctx* ctx, OrbBuffers& b) {
    const OrbGeom& g = ctx->og;
    const size_t S = ctx->nSlots;
    int rc;
#define DA(p, n) if ((rc = dalloc(ctx, &(p), (n))) != 0) return rc
    DA(b.pyr, S * g.slotStride + 4096);
    DA(b.blur, S * g.slotStride + 4096);
    DA(b.tileCnt, S * g.tilesPerFrame * kTileH);
    DA(b.tileTot, S * g.tilesPerFrame);
    DA(b.tileList, S * (size_t)g.tilesPerFrame * kTileListCap);
    DA(b.cand, S * g.candPerSlot);
    DA(b.candCount, S * kMaxLevels);
    DA(b.pairs, S * g.candPerSlot);
    DA(b.selWork, S * g.candPerSlot);
    DA(b.selList, S * g.candPerSlot * 2);
    DA(b.finXY, S * g.finPerSlot);
    DA(b.finResp, S * g.finPerSlot);
    DA(b.finCount, S * kMaxLevels);
    DA(b.selDbg, S * kMaxLevels * 4);
    DA(b.featPt, S * g.maxkp * 2);
    DA(b.featResp, S * g.maxkp);
    DA(b.featAngle, S * g.maxkp);
    DA(b.featOctave, S * g.maxkp);
    DA(b.featXY, S * g.maxkp);
    DA(b.featCS, S * g.maxkp * 2);
    DA(b.featDesc, S * g.maxkp * 32);
    DA(b.featCount, S);
    DA(b.frameFlags, S);
#undef DA
    return 0;
}

// A.0 / A.4 geometry, in the float32 arithmetic cv2 uses
static void build_geometry(dvo_ctx* ctx) {
    const dvo_config& c = ctx->cfg;
    OrbGeom& g = ctx->og;
    g.nlevels = c.nlevels;
    g.fastThreshold = c.fast_threshold;
    const double sf = (double)1.2f;
    // nfeaturesPerLevel
    float factor = (float)(1.0 / sf);
    float nd = (float)c.nfeatures * (1.0f - factor) / (1.0f - (float)std::pow((double)factor, (double)c.nlevels));
    int quota[kMaxLevels], sum = 0;
    for (int L = 0; L < c.nlevels - 1; ++L) {
        quota[L] = (int)std::nearbyint(nd);
        sum += quota[L];
        nd = nd * factor;
    }
    quota[c.nlevels - 1] = std::max(c.nfeatures - sum, 0);
    size_t off = 0;
    int rowBase = 0, candBase = 0, finBase = 0, tileBase = 0, rbBase = 0;
    for (int L = 0; L < c.nlevels; ++L) {
        LevelGeom& lv = g.lv[L];
        float scale = (float)std::pow(sf, (double)L);
        lv.scale = scale;
        lv.invScale = 1.0f / scale;
        lv.w = (int)std::nearbyint(c.width * lv.invScale);
        lv.h = (int)std::nearbyint(c.height * lv.invScale);
        if (L == 0) { lv.w = c.width; lv.h = c.height; }
        lv.pitch = (int)round_up(lv.w, 128);
        lv.off = off;
        off += round_up((size_t)lv.pitch * (lv.h + 1), 256);
        lv.rowBase = rowBase;
        rowBase += lv.h;
        lv.quota = quota[L];
        lv.candCap = ((lv.w + 1) / 2) * ((lv.h + 1) / 2);
        lv.candBase = candBase;
        candBase += (int)round_up(lv.candCap, 4);
        lv.finCap = lv.quota + kFinSlack;
        lv.finBase = finBase;
        finBase += lv.finCap;
        lv.tilesX = (lv.w + kTileW - 1) / kTileW;
        lv.tilesY = (lv.h + kTileH - 1) / kTileH;
        lv.tileBase = tileBase;
        tileBase += lv.tilesX * lv.tilesY;
        lv.rbBase = rbBase;
        rbBase += (lv.h + 31) / 32;
    }
    g.slotStride = round_up(off, 1024);
    g.rowsPerSlot = rowBase;
    g.candPerSlot = candBase;
    g.finPerSlot = finBase;
    g.maxkp = finBase;
    g.tilesPerFrame = tileBase;
    g.rowBlocksPerFrame = rbBase;
}

// A.1 coefficient tables: per destination index, source offset and 8-bit weight of the right/lower neighbour
static void resize_axis_table(int dst, int src, std::vector<uint32_t>& out) {
    volatile double inv_scale = (double)dst / (double)src;
    volatile double scale = 1.0 / inv_scale;
    for (int d = 0; d < dst; ++d) {
        volatile double a = scale * (d + 0.5);
        volatile double f = a - 0.5;
        int i = (int)std::floor(f);
        uint32_t ofs, c1;
        if (i < 0) { ofs = 0; c1 = 0; }
        else if (i >= src - 1) { ofs = (uint32_t)(src - 1); c1 = 0; }
        else {
            volatile double fr = (f - i) * 256.0;
            ofs = (uint32_t)i;
            c1 = (uint32_t)std::nearbyint(fr);
        }
        out.push_back((ofs << 16) | c1);
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int build_tensor_maps(dvo_ctx* ctx, const OrbBuffers& ob, TensorMaps& tmaps) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || fn == nullptr || qres != cudaDriverEntryPointSuccess) {
        ctx->err = "cuTensorMapEncodeTiled not available from the driver";
        return DVO_E_CUDA;
    }
    EncodeTiledFn encode = reinterpret_cast<EncodeTiledFn>(fn);
    for (int L = 0; L < ctx->og.nlevels; ++L) {
        const LevelGeom& lv = ctx->og.lv[L];
        cuuint64_t dims[3] = {(cuuint64_t)lv.w, (cuuint64_t)lv.h, (cuuint64_t)ctx->nSlots};
        cuuint64_t strides[2] = {(cuuint64_t)lv.pitch, (cuuint64_t)ctx->og.slotStride};
        cuuint32_t box[3] = {(cuuint32_t)kFastBoxW, (cuuint32_t)kFastBoxH, 1};
        cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = encode(&tmaps.pyr[L], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, ob.pyr + lv.off, dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            char buf[128];
            snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled failed for level %d: CUresult %d", L, (int)r);
            ctx->err = buf;
            return DVO_E_CUDA;
        }
        for (int v = 0; v < 2 && L + 1 < ctx->og.nlevels; ++v) {      // level L as the source of k_pyr_down(L + 1)
            cuuint32_t sbox[3] = {(cuuint32_t)kPyrSrcW, (cuuint32_t)(v == 0 ? kPyrSrcH : kPyrSrcHSmall), 1};
            r = encode(&tmaps.pyrSrc[v][L], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, ob.pyr + lv.off, dims, strides, sbox, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) {
                char buf[128];
                snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled failed for pyramid source level %d: CUresult %d", L, (int)r);
                ctx->err = buf;
                return DVO_E_CUDA;
            }
        }
    }
    return 0;
}

extern "C" {

const char* dvo_version(void) { return "libdvo 0.1 (sm_100a)"; }

int dvo_sizeof(int which) {
    switch (which) {
        case 0: return (int)sizeof(dvo_config);
        case 1: return (int)sizeof(dvo_pose);
        case 2: return (int)sizeof(dvo_features);
        case 3: return (int)sizeof(dvo_pair_arrays);
        default: return DVO_E_INVALID;
    }
}

void dvo_default_config(dvo_config* c) {
    memset(c, 0, sizeof *c);
    c->width = 1280; c->height = 1024;
    c->nfeatures = 500; c->nlevels = 8; c->fast_threshold = 20;
    c->max_frames = 2;
    c->matcher = DVO_MATCH_CROSSCHECK;
    c->ransac_max_iters = 1000; c->ransac_prob = 0.999; c->ransac_threshold = 1.0;
    c->distance_thresh = 50.0; c->ratio = 0.75f; c->use_tma = 1; c->pipeline = 1; c->ransac_exhaustive = 0; c->nn_engine = 0;
}

int dvo_create(const dvo_config* cfg, int device, dvo_ctx** out) {
    if (!cfg || !out) return DVO_E_INVALID;
    *out = nullptr;
    if (cfg->width < 64 || cfg->height < 64 || cfg->width > kMaxImageDim || cfg->height > kMaxImageDim ||
        cfg->nlevels < 1 || cfg->nlevels > kMaxLevels || cfg->nfeatures < 1 || cfg->nfeatures > 65000 || cfg->max_frames < 2 ||
        cfg->ransac_max_iters < 1)
        return DVO_E_INVALID;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return DVO_E_NODEVICE;
    dvo_ctx* ctx = new dvo_ctx();
    ctx->cfg = *cfg;
    ctx->device = device;
    *out = ctx;   // returned even on failure so the caller can read dvo_last_error, then dvo_destroy
    CK(cudaSetDevice(device));
    ctx->nSlots = cfg->max_frames;
    ctx->nPairs = cfg->max_frames - 1;
    build_geometry(ctx);
    OrbGeom& g = ctx->og;
    OrbBuffers& b = ctx->ob;
    int rc;
#define DA(p, n) if ((rc = dalloc(ctx, &(p), (n))) != 0) return rc
    if ((rc = alloc_orb_buffers(ctx, b)) != 0) return rc;
    // resize tables
    {
        std::vector<uint32_t> tab;
        for (int L = 0; L < g.nlevels; ++L) {
            if (L == 0) { b.resizeTabOff[0][0] = b.resizeTabOff[0][1] = 0; continue; }
            b.resizeTabOff[L][0] = (int)tab.size();
            resize_axis_table(g.lv[L].w, g.lv[L - 1].w, tab);
            b.resizeTabOff[L][1] = (int)tab.size();
            resize_axis_table(g.lv[L].h, g.lv[L - 1].h, tab);
        }
        // k_pyr_down assumes the taps of 4 adjacent output columns span at most 8 source bytes
        for (int L = 1; L < g.nlevels; ++L)
            for (int x = 0; x + 3 < g.lv[L].w; x += 4) {
                const uint32_t* t = tab.data() + b.resizeTabOff[L][0];
                if ((int)(t[x + 3] >> 16) - (int)(t[x] >> 16) > 6) { ctx->err = "pyramid ratio too large for k_pyr_down"; return DVO_E_INVALID; }
            }
        tab.push_back(0);
        DA(ctx->d_resizeTab, tab.size());
        CK(cudaMemcpy(ctx->d_resizeTab, tab.data(), tab.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
        b.resizeTab = ctx->d_resizeTab;
    }
    {   // tile -> (level, tileX, tileY) table for the tiled image kernels
        std::vector<uint32_t> ti(g.tilesPerFrame);
        for (int L = 0; L < g.nlevels; ++L)
            for (int t = 0; t < g.lv[L].tilesX * g.lv[L].tilesY; ++t)
                ti[g.lv[L].tileBase + t] = (uint32_t)L | ((uint32_t)(t % g.lv[L].tilesX) << 4) | ((uint32_t)(t / g.lv[L].tilesX) << 16);
        uint32_t* d_ti = nullptr;
        DA(d_ti, ti.size());
        CK(cudaMemcpy(d_ti, ti.data(), ti.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
        b.tileInfo = d_ti;
    }
    // pair buffers
    PairGeom& pg = ctx->pg;
    PairBuffers& pb = ctx->pb;
    pg.maxkp = g.maxkp;
    pg.maxMatches = g.maxkp;
    pg.sortCap = 1;
    while (pg.sortCap < g.maxkp) pg.sortCap <<= 1;
    pg.matcher = cfg->matcher;
    pg.maxIters = cfg->ransac_max_iters;
    pg.exhaustive = cfg->ransac_exhaustive != 0;
    pg.nnTensor = cfg->nn_engine == 0;
    pg.numSms = 148;
    cudaDeviceGetAttribute(&pg.numSms, cudaDevAttrMultiProcessorCount, ctx->device);
    pg.prob = cfg->ransac_prob;
    pg.threshold = cfg->ransac_threshold;
    pg.distThresh = cfg->distance_thresh;
    pg.ratio = cfg->ratio;
    const size_t P = ctx->nPairs, M = g.maxkp;
    DA(pb.nnIdx, P * 2 * M);
    DA(pb.nnDist, P * 2 * M);
    DA(pb.nn2Dist, P * M);
    pb.descX = nullptr;
    pb.descXRows = nn_tensor_rows(g.maxkp);
    if (pg.nnTensor) {
        DA(pb.descX, (P + 1) * (size_t)pb.descXRows * 256);
        CK(nn_tensor_init());
    }
    DA(pb.matches, P * M * 3);
    DA(pb.matchCount, P);
    DA(pb.ptsPrev, P * M * 2);
    DA(pb.ptsCur, P * M * 2);
    DA(pb.normPts, P * M * 4);
    DA(pb.samples, P * (size_t)pg.maxIters * 5);
    if (pg.exhaustive || (g.maxkp >= 1500 && P <= 8)) {
        // cv::RNG(-1) state table: 16 draws of margin per iteration (5 are needed, more only after duplicate redraws)
        pg.rngCount = std::min(16 * pg.maxIters, 200 * 1024);
        std::vector<unsigned long long> st(pg.rngCount);
        unsigned long long state = 0xFFFFFFFFFFFFFFFFull;
        for (int i = 0; i < pg.rngCount; ++i) {
            state = (unsigned long long)(uint32_t)state * 4164903690ull + (state >> 32);
            st[i] = state;
        }
        unsigned long long* d_st = nullptr;
        DA(d_st, st.size());
        CK(cudaMemcpy(d_st, st.data(), st.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice));
        pb.rngStates = d_st;
        DA(pb.exStart, P * (size_t)pg.maxIters);
        CK(pair_kernels_init_exhaustive(pg.rngCount));
        DA(pb.exModels, P * (size_t)pg.maxIters * kMaxModels * 9);
        DA(pb.exCount, P * (size_t)pg.maxIters);
        DA(pb.exGood, P * (size_t)pg.maxIters * kMaxModels);
    }
    DA(pb.ransacState, P * 8);
    DA(pb.bestE, P * 9);
    DA(pb.ransacMask, P * M);
    DA(pb.poseMask, P * M);
    DA(pb.poses, P);
    DA(pb.poseScratch, P);
    DA(ctx->d_K, 16);
#undef DA
    ctx->useTma = cfg->use_tma != 0 && getenv("DVO_NO_TMA") == nullptr;
    if (ctx->useTma) {
        rc = build_tensor_maps(ctx, ctx->ob, ctx->tmaps);
        if (rc != 0) return rc;
    }
    if (getenv("DVO_NO_SIDE_STREAMS") == nullptr) {
        CK(cudaStreamCreateWithFlags(&ctx->ss.side, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&ctx->ss.side2, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&ctx->ss.copy, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&ctx->ss.evFork2, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&ctx->ss.evJoin2, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&ctx->ss.evFork, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&ctx->ss.evJoin, cudaEventDisableTiming));
        for (int i = 0; i < 2; ++i) {
            CK(cudaEventCreateWithFlags(&ctx->ss.evCopied[i], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&ctx->ss.evStageFree[i], cudaEventDisableTiming));
        }
    }
    if (cfg->pipeline != 0 && getenv("DVO_NO_PIPELINE") == nullptr) {
        CK(cudaStreamCreateWithFlags(&ctx->sOrb, cudaStreamNonBlocking));
        {   // the pair stage is short and latency-bound: give its CTAs first pick of freed SM resources
            int lo = 0, hi = 0;
            CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
            CK(cudaStreamCreateWithPriority(&ctx->sPair, cudaStreamNonBlocking, hi));
            if (getenv("DVO_NO_IMAGE_STREAM") == nullptr) CK(cudaStreamCreateWithPriority(&ctx->sImg, cudaStreamNonBlocking, lo));
        }
        CK(cudaEventCreateWithFlags(&ctx->evCall, cudaEventDisableTiming));
        for (int i = 0; i < 2; ++i) {
            CK(cudaEventCreateWithFlags(&ctx->evOrbDone[i], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&ctx->evPairsDone[i], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&ctx->evCarryCopied[i], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&ctx->evImgDone[i], cudaEventDisableTiming));
        }
    }
    CK(orb_kernels_init());
    CK(pair_kernels_init((int)(pg.sortCap * sizeof(uint32_t))));
    CK(cudaDeviceSynchronize());
    ctx->launchBase = orb_launch_count() + pair_launch_count();
    return DVO_OK;
}

void dvo_destroy(dvo_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    for (void* p : ctx->allocs) cudaFree(p);
    if (ctx->evCall) cudaEventDestroy(ctx->evCall);
    for (int i = 0; i < 2; ++i) {
        if (ctx->evOrbDone[i]) cudaEventDestroy(ctx->evOrbDone[i]);
        if (ctx->evPairsDone[i]) cudaEventDestroy(ctx->evPairsDone[i]);
        if (ctx->evCarryCopied[i]) cudaEventDestroy(ctx->evCarryCopied[i]);
        if (ctx->evImgDone[i]) cudaEventDestroy(ctx->evImgDone[i]);
    }
    if (ctx->sImg) cudaStreamDestroy(ctx->sImg);
    if (ctx->sOrb) cudaStreamDestroy(ctx->sOrb);
    if (ctx->sPair) cudaStreamDestroy(ctx->sPair);
    for (int i = 0; i < 2; ++i) {
        if (ctx->ss.stage[i]) cudaFree(ctx->ss.stage[i]);
        if (ctx->ss.evCopied[i]) cudaEventDestroy(ctx->ss.evCopied[i]);
        if (ctx->ss.evStageFree[i]) cudaEventDestroy(ctx->ss.evStageFree[i]);
    }
    if (ctx->ss.evFork) cudaEventDestroy(ctx->ss.evFork);
    if (ctx->ss.evJoin) cudaEventDestroy(ctx->ss.evJoin);
    if (ctx->ss.evFork2) cudaEventDestroy(ctx->ss.evFork2);
    if (ctx->ss.evJoin2) cudaEventDestroy(ctx->ss.evJoin2);
    if (ctx->ss.side2) cudaStreamDestroy(ctx->ss.side2);
    if (ctx->ss.side) cudaStreamDestroy(ctx->ss.side);
    if (ctx->ss.copy) cudaStreamDestroy(ctx->ss.copy);
    if (ctx->h_poseStage) cudaFreeHost(ctx->h_poseStage);
    if (ctx->h_frameStage) cudaFreeHost(ctx->h_frameStage);
    delete ctx;
}

const char* dvo_last_error(const dvo_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
int dvo_max_keypoints(const dvo_ctx* ctx) { return ctx ? ctx->og.maxkp : DVO_E_INVALID; }
int dvo_max_frames(const dvo_ctx* ctx) { return ctx ? ctx->nSlots : DVO_E_INVALID; }
long long dvo_kernel_launches(const dvo_ctx* ctx) {
    return ctx ? orb_launch_count() + pair_launch_count() - ctx->launchBase : 0;
}

int dvo_level_size(const dvo_ctx* ctx, int level, int* w, int* h, int* quota) {
    if (!ctx || level < 0 || level >= ctx->og.nlevels) return DVO_E_INVALID;
    if (w) *w = ctx->og.lv[level].w;
    if (h) *h = ctx->og.lv[level].h;
    if (quota) *quota = ctx->og.lv[level].quota;
    return DVO_OK;
}

static int load_frames_into(dvo_ctx* ctx, const OrbBuffers& ob, const uint8_t* frames, int n, size_t pitch, size_t frame_stride,
                            int slot0, int kind, cudaStream_t st);

int dvo_load_frames(dvo_ctx* ctx, const uint8_t* frames, int n, size_t pitch, size_t frame_stride, int slot0, int kind,
                    void* stream) {
    if (!ctx || !frames || n < 0 || slot0 < 0 || slot0 + n > ctx->nSlots) {
        if (ctx) ctx->err = "dvo_load_frames: bad arguments";
        return DVO_E_INVALID;
    }
    return load_frames_into(ctx, ctx->ob, frames, n, pitch, frame_stride, slot0, kind, (cudaStream_t)stream);
}

static int load_frames_into(dvo_ctx* ctx, const OrbBuffers& ob, const uint8_t* frames, int n, size_t pitch, size_t frame_stride,
                            int slot0, int kind, cudaStream_t st) {
    const LevelGeom& l0 = ctx->og.lv[0];
    if (n == 0) return DVO_OK;
    const int ch = ctx->ingest.channels > 0 ? ctx->ingest.channels : 1;
    const size_t rowBytes = (size_t)l0.w * ch, frameBytes = rowBytes * l0.h;
    if (pitch < rowBytes) { ctx->err = "dvo_load_frames: pitch smaller than width * channels"; return DVO_E_INVALID; }
    auto to_slots = [&](const uint8_t* d_src, size_t p, size_t fs) {
        if (ctx->ingest.channels > 0) launch_ingest(ctx->og, ob, ctx->ingest, d_src, n, p, fs, slot0, st);
        else launch_load_frames(ctx->og, ob, d_src, n, p, fs, slot0, st);
    };
    if (kind == 0) {
        to_slots(frames, pitch, frame_stride);
        CK(cudaGetLastError());
        return DVO_OK;
    }
    SideStreams& ss = ctx->ss;
    // Host frames: H2D (on the copy stream when there is one, so the upload of the next batch overlaps the kernels of this
    // one) into one of two staging buffers, then one kernel on the compute stream moves / ingests them into the slots.
    const int bsel = ss.stageIdx;
    ss.stageIdx ^= 1;
    const size_t need = frameBytes * ctx->nSlots;
    if (ctx->stageBytes[bsel] < need) {
        if (ss.stage[bsel]) { CK(cudaDeviceSynchronize()); CK(cudaFree(ss.stage[bsel])); ss.stage[bsel] = nullptr; }
        CK(cudaMalloc(&ss.stage[bsel], need));
        ctx->stageBytes[bsel] = need;
    }
    cudaStream_t cs = ss.copy != nullptr ? ss.copy : st;
    if (ss.copy != nullptr && ss.stageUsed[bsel]) CK(cudaStreamWaitEvent(ss.copy, ss.evStageFree[bsel], 0));
    if (pitch == rowBytes && frame_stride == frameBytes) {
        CK(cudaMemcpyAsync(ss.stage[bsel], frames, frameBytes * n, cudaMemcpyHostToDevice, cs));
    } else {
        for (int i = 0; i < n; ++i)
            CK(cudaMemcpy2DAsync(ss.stage[bsel] + (size_t)i * frameBytes, rowBytes, frames + (size_t)i * frame_stride, pitch, rowBytes,
                                 l0.h, cudaMemcpyHostToDevice, cs));
    }
    if (ss.copy != nullptr) {
        CK(cudaEventRecord(ss.evCopied[bsel], ss.copy));
        CK(cudaStreamWaitEvent(st, ss.evCopied[bsel], 0));
    }
    to_slots(ss.stage[bsel], rowBytes, frameBytes);
    if (ss.copy != nullptr) {
        CK(cudaEventRecord(ss.evStageFree[bsel], st));
        ss.stageUsed[bsel] = true;
    }
    CK(cudaGetLastError());
    return DVO_OK;
}

// Grey conversion + undistortion while loading (replaces cv.cvtColor + cv.undistort of visual_odometry_v3.py:110-135).
int dvo_set_undistort(dvo_ctx* ctx, const double* K, const double* dist, int n_dist, const double* newK, int channels, void* stream) {
    if (!ctx) return DVO_E_INVALID;
    if (channels == 0) { ctx->ingest.channels = 0; return DVO_OK; }
    if (!K || !newK || (n_dist > 0 && !dist) || n_dist < 0 || (channels != 1 && channels != 3)) {
        ctx->err = "dvo_set_undistort: bad arguments (channels must be 0, 1 or 3)";
        return DVO_E_INVALID;
    }
    IngestParams prm{};
    {   // inv(newK) by cofactors, float64
        const double* m = newK;
        const double c00 = m[4] * m[8] - m[5] * m[7], c01 = m[5] * m[6] - m[3] * m[8], c02 = m[3] * m[7] - m[4] * m[6];
        const double det = m[0] * c00 + m[1] * c01 + m[2] * c02;
        if (!(std::fabs(det) > 0)) { ctx->err = "dvo_set_undistort: newCameraMatrix is singular"; return DVO_E_INVALID; }
        const double id = 1.0 / det;
        prm.ir[0] = c00 * id; prm.ir[1] = (m[2] * m[7] - m[1] * m[8]) * id; prm.ir[2] = (m[1] * m[5] - m[2] * m[4]) * id;
        prm.ir[3] = c01 * id; prm.ir[4] = (m[0] * m[8] - m[2] * m[6]) * id; prm.ir[5] = (m[2] * m[3] - m[0] * m[5]) * id;
        prm.ir[6] = c02 * id; prm.ir[7] = (m[1] * m[6] - m[0] * m[7]) * id; prm.ir[8] = (m[0] * m[4] - m[1] * m[3]) * id;
    }
    prm.fx = K[0]; prm.fy = K[4]; prm.cx = K[2]; prm.cy = K[5];
    for (int i = 0; i < 8; ++i) prm.k[i] = i < n_dist ? dist[i] : 0.0;
    int rc;
    if (ctx->ingest.map == nullptr) {
        uint2* wt = nullptr;
        if ((rc = dalloc(ctx, &ctx->ingest.map, (size_t)ctx->og.lv[0].w * ctx->og.lv[0].h)) != 0) return rc;
        if ((rc = dalloc(ctx, &wt, 1024)) != 0) return rc;
        // cv2's BilinearTab_i: float32 products of (1 - a | a), rounded to 1/32768, largest/smallest patched to sum 32768
        std::vector<uint2> tab(1024);
        const float s = 1.0f / 32;
        for (int fy = 0; fy < 32; ++fy)
            for (int fx = 0; fx < 32; ++fx) {
                volatile float ay = fy * s, ax = fx * s;
                volatile float cy[2] = {1.0f - ay, ay}, cx[2] = {1.0f - ax, ax};
                int it[4], sum = 0;
                for (int a = 0; a < 2; ++a)
                    for (int b = 0; b < 2; ++b) {
                        volatile float v = cy[a] * cx[b];
                        volatile float sc = v * 32768.0f;
                        it[a * 2 + b] = (int)std::lrintf(sc);
                        sum += it[a * 2 + b];
                    }
                const int diff = sum - 32768;
                if (diff != 0) {
                    int mk = 0, Mk = 0;
                    for (int k = 0; k < 4; ++k) {
                        if (it[k] < it[mk]) mk = k;
                        else if (it[k] > it[Mk]) Mk = k;
                    }
                    if (diff < 0) it[Mk] -= diff; else it[mk] -= diff;
                }
                tab[fy * 32 + fx] = make_uint2((uint32_t)(uint16_t)it[0] | ((uint32_t)(uint16_t)it[1] << 16),
                                               (uint32_t)(uint16_t)it[2] | ((uint32_t)(uint16_t)it[3] << 16));   // 0 .. 32768
            }
        CK(cudaMemcpy(wt, tab.data(), sizeof(uint2) * 1024, cudaMemcpyHostToDevice));
        ctx->ingest.wtab = wt;
    }
    launch_build_undistort_map(ctx->og, ctx->ingest, prm, (cudaStream_t)stream);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize((cudaStream_t)stream));    // the map may be used from internal streams next
    ctx->ingest.channels = channels;
    return DVO_OK;
}

int dvo_orb(dvo_ctx* ctx, int slot0, int n, void* stream) {
    if (!ctx || n < 0 || slot0 < 0 || slot0 + n > ctx->nSlots) {
        if (ctx) ctx->err = "dvo_orb: slot range out of bounds";
        return DVO_E_INVALID;
    }
    launch_orb(ctx->og, ctx->ob, &ctx->tmaps, ctx->useTma, slot0, n, (cudaStream_t)stream, &ctx->ss);
    CK(cudaGetLastError());
    return DVO_OK;
}

__global__ void k_fill_size(const int* octave, const int* count, float* size, int cap, float s0, float s1, float s2, float s3,
                            float s4, float s5, float s6, float s7) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cap || i >= *count) return;
    const float sc[8] = {s0, s1, s2, s3, s4, s5, s6, s7};
    size[i] = __fmul_rn(31.0f, sc[octave[i] & 7]);
}

int dvo_get_features(dvo_ctx* ctx, int slot, const dvo_features* out, void* stream) {
    if (!ctx || !out || slot < 0 || slot >= ctx->nSlots) return DVO_E_INVALID;
    if (out->capacity < ctx->og.maxkp) { ctx->err = "dvo_get_features: capacity < dvo_max_keypoints"; return DVO_E_CAPACITY; }
    cudaStream_t st = (cudaStream_t)stream;
    const OrbGeom& g = ctx->og;
    const OrbBuffers& b = ctx->ob;
    const size_t M = g.maxkp, o = (size_t)slot * M;
    if (out->d_pt) CK(cudaMemcpyAsync(out->d_pt, b.featPt + o * 2, M * 2 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (out->d_angle) CK(cudaMemcpyAsync(out->d_angle, b.featAngle + o, M * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (out->d_response) CK(cudaMemcpyAsync(out->d_response, b.featResp + o, M * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (out->d_octave) CK(cudaMemcpyAsync(out->d_octave, b.featOctave + o, M * sizeof(int), cudaMemcpyDeviceToDevice, st));
    if (out->d_desc) CK(cudaMemcpyAsync(out->d_desc, b.featDesc + o * 32, M * 32, cudaMemcpyDeviceToDevice, st));
    if (out->d_count) CK(cudaMemcpyAsync(out->d_count, b.featCount + slot, sizeof(int), cudaMemcpyDeviceToDevice, st));
    if (out->d_size) {
        const LevelGeom* lv = g.lv;
        k_fill_size<<<(int)((M + 255) / 256), 256, 0, st>>>(b.featOctave + o, b.featCount + slot, out->d_size, (int)M, lv[0].scale,
                                                            lv[1].scale, lv[2].scale, lv[3].scale, lv[4].scale, lv[5].scale,
                                                            lv[6].scale, lv[7].scale);
        CK(cudaGetLastError());
    }
    return DVO_OK;
}

int dvo_get_frame_flags(dvo_ctx* ctx, int slot0, int n, int32_t* h_flags, void* stream) {
    if (!ctx || !h_flags || n < 0 || slot0 < 0 || slot0 + n > ctx->nSlots) return DVO_E_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    if (n > 0) CK(cudaMemcpyAsync(h_flags, ctx->ob.frameFlags + slot0, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return DVO_OK;
}

int dvo_tap_image(dvo_ctx* ctx, int slot, int level, int which, uint8_t* d_dst, void* stream) {
    if (!ctx || !d_dst || slot < 0 || slot >= ctx->nSlots || level < 0 || level >= ctx->og.nlevels || which < 0 || which > 1)
        return DVO_E_INVALID;
    const LevelGeom& lv = ctx->og.lv[level];
    const uint8_t* base = which == 0 ? ctx->ob.pyr : ctx->ob.blur;
    CK(cudaMemcpy2DAsync(d_dst, lv.w, base + (size_t)slot * ctx->og.slotStride + lv.off, lv.pitch, lv.w, lv.h,
                         cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return DVO_OK;
}

int dvo_tap_candidates(dvo_ctx* ctx, int slot, int level, uint32_t* d_dst, int capacity, int* h_count, void* stream) {
    if (!ctx || slot < 0 || slot >= ctx->nSlots || level < 0 || level >= ctx->og.nlevels || !h_count) return DVO_E_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaMemcpyAsync(h_count, ctx->ob.candCount + slot * kMaxLevels + level, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (*h_count > capacity) { ctx->err = "dvo_tap_candidates: capacity too small"; return DVO_E_CAPACITY; }
    const LevelGeom& lv = ctx->og.lv[level];
    if (d_dst && *h_count > 0)
        CK(cudaMemcpyAsync(d_dst, ctx->ob.cand + (size_t)slot * ctx->og.candPerSlot + lv.candBase, sizeof(uint32_t) * (*h_count),
                           cudaMemcpyDeviceToDevice, st));
    return DVO_OK;
}

int dvo_pairs(dvo_ctx* ctx, int slot0, int pair0, int n, const double* K, void* stream) {
    if (!ctx || !K || n < 0 || slot0 < 0 || slot0 + n + (n > 0 ? 1 : 0) > ctx->nSlots || pair0 < 0 || pair0 + n > ctx->nPairs) {
        if (ctx) ctx->err = "dvo_pairs: slot/pair range out of bounds";
        return DVO_E_INVALID;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (ctx->pg.sortCap * sizeof(uint32_t) > 200 * 1024) {
        ctx->err = "dvo_pairs: nfeatures too large for the in-shared-memory match sort (max ~49000 keypoints)";
        return DVO_E_CAPACITY;
    }
    launch_pairs(ctx->og, ctx->ob, ctx->pg, ctx->pb, slot0, pair0, n, K, st);
    CK(cudaGetLastError());
    return DVO_OK;
}

int dvo_match(dvo_ctx* ctx, int slot0, int pair0, int n, const double* K, void* stream) {
    if (!ctx || !K || n < 0 || slot0 < 0 || slot0 + n + (n > 0 ? 1 : 0) > ctx->nSlots || pair0 < 0 || pair0 + n > ctx->nPairs) {
        if (ctx) ctx->err = "dvo_match: slot/pair range out of bounds";
        return DVO_E_INVALID;
    }
    if (ctx->pg.sortCap * sizeof(uint32_t) > 200 * 1024) {
        ctx->err = "dvo_match: nfeatures too large for the in-shared-memory match sort (max ~49000 keypoints)";
        return DVO_E_CAPACITY;
    }
    launch_match(ctx->og, ctx->ob, ctx->pg, ctx->pb, slot0, pair0, n, K, (cudaStream_t)stream);
    CK(cudaGetLastError());
    return DVO_OK;
}

int dvo_pose_pairs(dvo_ctx* ctx, int slot0, int pair0, int n, const double* K, void* stream) {
    if (!ctx || !K || n < 0 || pair0 < 0 || pair0 + n > ctx->nPairs || (slot0 >= 0 && slot0 + n + (n > 0 ? 1 : 0) > ctx->nSlots)) {
        if (ctx) ctx->err = "dvo_pose_pairs: slot/pair range out of bounds";
        return DVO_E_INVALID;
    }
    launch_ransac_pose(ctx->og, ctx->ob, ctx->pg, ctx->pb, slot0 < 0 ? -1 : slot0, pair0, n, K, (cudaStream_t)stream);
    CK(cudaGetLastError());
    return DVO_OK;
}

int dvo_get_match_count(dvo_ctx* ctx, int pair, int* h_count, void* stream) {
    if (!ctx || !h_count || pair < 0 || pair >= ctx->nPairs) return DVO_E_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaMemcpyAsync(h_count, ctx->pb.matchCount + pair, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return DVO_OK;
}

int dvo_set_features(dvo_ctx* ctx, int slot, const float* pt, const uint8_t* desc, int n, int kind, void* stream) {
    if (!ctx || slot < 0 || slot >= ctx->nSlots || n < 0 || (n > 0 && (!pt || !desc))) return DVO_E_INVALID;
    if (n > ctx->og.maxkp) { ctx->err = "dvo_set_features: n exceeds dvo_max_keypoints"; return DVO_E_CAPACITY; }
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemcpyKind k = kind == 0 ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    const size_t o = (size_t)slot * ctx->og.maxkp;
    if (n > 0) {
        CK(cudaMemcpyAsync(ctx->ob.featPt + o * 2, pt, sizeof(float) * 2 * n, k, st));
        CK(cudaMemcpyAsync(ctx->ob.featDesc + o * 32, desc, 32 * (size_t)n, k, st));
    }
    CK(cudaMemcpyAsync(ctx->ob.featCount + slot, &n, sizeof(int), cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(ctx->ob.frameFlags + slot, 0, sizeof(int), st));
    CK(cudaStreamSynchronize(st));   // &n is a stack variable
    return DVO_OK;
}

int dvo_pose_points(dvo_ctx* ctx, int pair, const float* pts_prev, const float* pts_cur, int n, const double* K, int kind,
                    void* stream) {
    if (!ctx || !K || pair < 0 || pair >= ctx->nPairs || n < 0 || (n > 0 && (!pts_prev || !pts_cur))) return DVO_E_INVALID;
    if (n > ctx->og.maxkp) { ctx->err = "dvo_pose_points: n exceeds dvo_max_keypoints"; return DVO_E_CAPACITY; }
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemcpyKind k = kind == 0 ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    const size_t o = (size_t)pair * ctx->og.maxkp;
    if (n > 0) {
        CK(cudaMemcpyAsync(ctx->pb.ptsPrev + o * 2, pts_prev, sizeof(float) * 2 * n, k, st));
        CK(cudaMemcpyAsync(ctx->pb.ptsCur + o * 2, pts_cur, sizeof(float) * 2 * n, k, st));
    }
    launch_points_prep(ctx->pg, ctx->pb, pair, n, K, st);
    launch_ransac_pose(ctx->og, ctx->ob, ctx->pg, ctx->pb, -1, pair, 1, K, st);
    CK(cudaGetLastError());
    return DVO_OK;
}

int dvo_get_poses(dvo_ctx* ctx, int pair0, int n, dvo_pose* dst, int kind, void* stream) {
    if (!ctx || !dst || n < 0 || pair0 < 0 || pair0 + n > ctx->nPairs) return DVO_E_INVALID;
    CK(cudaMemcpyAsync(dst, ctx->pb.poses + pair0, sizeof(dvo_pose) * n, kind == 0 ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost,
                       (cudaStream_t)stream));
    return DVO_OK;
}

int dvo_get_pair_arrays(dvo_ctx* ctx, int pair, const dvo_pair_arrays* out, void* stream) {
    if (!ctx || !out || pair < 0 || pair >= ctx->nPairs) return DVO_E_INVALID;
    if (out->capacity < ctx->og.maxkp) { ctx->err = "dvo_get_pair_arrays: capacity < dvo_max_keypoints"; return DVO_E_CAPACITY; }
    cudaStream_t st = (cudaStream_t)stream;
    const size_t M = ctx->og.maxkp, o = (size_t)pair * M;
    const PairBuffers& pb = ctx->pb;
    if (out->d_matches) CK(cudaMemcpyAsync(out->d_matches, pb.matches + o * 3, M * 3 * sizeof(int), cudaMemcpyDeviceToDevice, st));
    if (out->d_pts_prev) CK(cudaMemcpyAsync(out->d_pts_prev, pb.ptsPrev + o * 2, M * 2 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (out->d_pts_cur) CK(cudaMemcpyAsync(out->d_pts_cur, pb.ptsCur + o * 2, M * 2 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (out->d_ransac_mask) CK(cudaMemcpyAsync(out->d_ransac_mask, pb.ransacMask + o, M, cudaMemcpyDeviceToDevice, st));
    if (out->d_pose_mask) CK(cudaMemcpyAsync(out->d_pose_mask, pb.poseMask + o, M, cudaMemcpyDeviceToDevice, st));
    return DVO_OK;
}

int dvo_tap_ransac(dvo_ctx* ctx, int pair, int32_t* h_state8, void* stream) {
    if (!ctx || !h_state8 || pair < 0 || pair >= ctx->nPairs) return DVO_E_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaMemcpyAsync(h_state8, ctx->pb.ransacState + pair * 8, 8 * sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return DVO_OK;
}

__global__ void k_copy_features(OrbGeom g, OrbBuffers a, OrbBuffers b, int src, int dst) {
    // carry the last frame of a batch (buffers a, slot src) into slot dst of buffers b for the next batch
    const int M = g.maxkp;
    const size_t so = (size_t)src * M, d0 = (size_t)dst * M;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < M * 8; i += gridDim.x * blockDim.x) {
        reinterpret_cast<uint32_t*>(b.featDesc + d0 * 32)[i] = reinterpret_cast<const uint32_t*>(a.featDesc + so * 32)[i];
        if (i < M * 2) b.featPt[d0 * 2 + i] = a.featPt[so * 2 + i];
        if (i < M) {
            b.featResp[d0 + i] = a.featResp[so + i];
            b.featAngle[d0 + i] = a.featAngle[so + i];
            b.featOctave[d0 + i] = a.featOctave[so + i];
            b.featXY[d0 + i] = a.featXY[so + i];
        }
        if (i == 0) { b.featCount[dst] = a.featCount[src]; b.frameFlags[dst] = a.frameFlags[src]; }
    }
}

// Two-lane, three-stage pipeline.  Per call: image half of this batch on sImg and keypoint half on sOrb into lane L = (previous
// lane ^ 1); pair stage on sPair after it.
// Hazards and the events that order them:
//   ORB(L) overwrites lane L's features       -> waits evPairsDone[L] (pairs of the batch before last) and
//                                                evCarryCopied[L] (the carry copy out of lane L issued with the last batch)
//   carry copy  last slot of lane L^1 -> slot 0 of lane L runs on sPair after the previous pair stage (stream order)
//   pairs(L) needs ORB(L)                      -> sPair waits evOrbDone[L]
// The caller's stream is joined with the PREVIOUS batch's pair stage only (so consecutive calls overlap); the current
// batch's pose records are complete once the next call, or dvo_sequence_flush, has been enqueued on that stream.
static int sequence_step_pipelined(dvo_ctx* ctx, const uint8_t* frames, int n_new, size_t pitch, size_t frame_stride, const double* K,
                                   dvo_pose* poses, int kind, int first, cudaStream_t st) {
    const bool fresh = first || ctx->carrySlot < 0 || ctx->lastLane < 0;
    if (fresh ? n_new > ctx->nSlots : n_new > ctx->nSlots - 1) {
        ctx->err = "dvo_sequence_step: more frames than slots";
        return DVO_E_CAPACITY;
    }
    const int L = fresh ? 0 : (ctx->lastLane ^ 1);
    int rc;
    if (L == 1 && !ctx->lane1Ready) {
        size_t mark = ctx->allocs.size();
        if ((rc = alloc_orb_buffers(ctx, ctx->ob1)) != 0) return rc;
        (void)mark;
        ctx->ob1.resizeTab = ctx->ob.resizeTab;
        ctx->ob1.tileInfo = ctx->ob.tileInfo;
        memcpy(ctx->ob1.resizeTabOff, ctx->ob.resizeTabOff, sizeof(ctx->ob.resizeTabOff));
        if (ctx->useTma && (rc = build_tensor_maps(ctx, ctx->ob1, ctx->tmaps1)) != 0) return rc;
        CK(cudaDeviceSynchronize());
        ctx->lane1Ready = true;
    }
    const OrbBuffers& ob = L == 0 ? ctx->ob : ctx->ob1;
    const TensorMaps& tm = L == 0 ? ctx->tmaps : ctx->tmaps1;
    const int slot0 = fresh ? 0 : 1, nPairs = fresh ? n_new - 1 : n_new;
    // ---- ORB stage.  Its image half (upload, pyramid, FAST) touches only lane L's pyramid and tile lists, which nothing has read
    // since lane L's previous ORB stage ended: it goes on the low-priority stream sImg and may run under the keypoint half of
    // the batch before (lane L ^ 1, stream sOrb).  The keypoint half overwrites lane L's features and therefore also waits for
    // the pair stage and the carry copy that read them.
    const bool serial = prof_enabled();     // per-kernel timing pass: no overlap between stages, no side streams
    cudaStream_t sImg = (serial || ctx->sImg == nullptr) ? ctx->sOrb : ctx->sImg;
    CK(cudaEventRecord(ctx->evCall, st));
    CK(cudaStreamWaitEvent(ctx->sOrb, ctx->evCall, 0));
    if (sImg != ctx->sOrb) {
        CK(cudaStreamWaitEvent(sImg, ctx->evCall, 0));
        CK(cudaStreamWaitEvent(sImg, ctx->evOrbDone[L], 0));      // lane L's previous ORB stage (no-op before the first one)
    }
    if (ctx->pairsPending[L]) CK(cudaStreamWaitEvent(ctx->sOrb, ctx->evPairsDone[L], 0));
    if (serial && !fresh && ctx->pairsPending[L ^ 1]) CK(cudaStreamWaitEvent(ctx->sOrb, ctx->evPairsDone[L ^ 1], 0));
    if (ctx->carryPending[L]) CK(cudaStreamWaitEvent(ctx->sOrb, ctx->evCarryCopied[L], 0));
    if ((rc = load_frames_into(ctx, ob, frames, n_new, pitch, frame_stride, slot0, kind, sImg)) != 0) return rc;
    launch_orb_image(ctx->og, ob, &tm, ctx->useTma, slot0, n_new, sImg);
    if (sImg != ctx->sOrb) {
        CK(cudaEventRecord(ctx->evImgDone[L], sImg));
        CK(cudaStreamWaitEvent(ctx->sOrb, ctx->evImgDone[L], 0));
    }
    launch_orb_keypoints(ctx->og, ob, slot0, n_new, ctx->sOrb, serial ? nullptr : &ctx->ss);
    CK(cudaGetLastError());
    CK(cudaEventRecord(ctx->evOrbDone[L], ctx->sOrb));
    // ---- pair stage
    CK(cudaStreamWaitEvent(ctx->sPair, ctx->evOrbDone[L], 0));
    if (!fresh) {
        const OrbBuffers& prev = ctx->carryLane == 0 ? ctx->ob : ctx->ob1;
        k_copy_features<<<32, 256, 0, ctx->sPair>>>(ctx->og, prev, ob, ctx->carrySlot, 0);
        CK(cudaGetLastError());
        CK(cudaEventRecord(ctx->evCarryCopied[ctx->carryLane], ctx->sPair));
        ctx->carryPending[ctx->carryLane] = true;
    }
    if (nPairs > 0) {
        if (ctx->pg.sortCap * sizeof(uint32_t) > 200 * 1024) {
            ctx->err = "dvo_sequence_step: nfeatures too large for the in-shared-memory match sort";
            return DVO_E_CAPACITY;
        }
        launch_pairs(ctx->og, ob, ctx->pg, ctx->pb, 0, 0, nPairs, K, ctx->sPair);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(poses, ctx->pb.poses, sizeof(dvo_pose) * nPairs, kind == 0 ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost,
                           ctx->sPair));
    }
    CK(cudaEventRecord(ctx->evPairsDone[L], ctx->sPair));
    ctx->pairsPending[L] = true;
    // ---- deferred join: the caller's stream waits for the batch before this one
    if (!fresh && ctx->pairsPending[L ^ 1]) CK(cudaStreamWaitEvent(st, ctx->evPairsDone[L ^ 1], 0));
    ctx->pipeOutstanding = true;
    ctx->lastLane = L;
    ctx->carryLane = L;
    ctx->carrySlot = fresh ? n_new - 1 : n_new;
    return nPairs;
}

int dvo_sequence_flush(dvo_ctx* ctx, void* stream) {
    if (!ctx) return DVO_E_INVALID;
    if (ctx->sOrb != nullptr && ctx->pipeOutstanding && ctx->lastLane >= 0) {
        CK(cudaStreamWaitEvent((cudaStream_t)stream, ctx->evPairsDone[ctx->lastLane], 0));
        ctx->pipeOutstanding = false;
    }
    return DVO_OK;
}

int dvo_sequence_step(dvo_ctx* ctx, const uint8_t* frames, int n_new, size_t pitch, size_t frame_stride, const double* K,
                      dvo_pose* poses, int kind, int first, void* stream) {
    if (!ctx || !frames || !K || !poses || n_new < 1) {
        if (ctx) ctx->err = "dvo_sequence_step: bad arguments";
        return DVO_E_INVALID;
    }
    cudaStream_t st = (cudaStream_t)stream;
    int rc;
    if (ctx->sOrb != nullptr) return sequence_step_pipelined(ctx, frames, n_new, pitch, frame_stride, K, poses, kind, first, st);
    if (first || ctx->carrySlot < 0) {
        if (n_new > ctx->nSlots) { ctx->err = "dvo_sequence_step: more frames than slots"; return DVO_E_CAPACITY; }
        if ((rc = dvo_load_frames(ctx, frames, n_new, pitch, frame_stride, 0, kind, st)) != 0) return rc;
        if ((rc = dvo_orb(ctx, 0, n_new, st)) != 0) return rc;
        if (n_new > 1) {
            if ((rc = dvo_pairs(ctx, 0, 0, n_new - 1, K, st)) != 0) return rc;
            if ((rc = dvo_get_poses(ctx, 0, n_new - 1, poses, kind, st)) != 0) return rc;
        }
        ctx->carrySlot = n_new - 1;
        return n_new - 1;
    }
    if (n_new > ctx->nSlots - 1) { ctx->err = "dvo_sequence_step: more new frames than slots - 1"; return DVO_E_CAPACITY; }
    if (ctx->carrySlot != 0) {
        k_copy_features<<<32, 256, 0, st>>>(ctx->og, ctx->ob, ctx->ob, ctx->carrySlot, 0);
        CK(cudaGetLastError());
    }
    if ((rc = dvo_load_frames(ctx, frames, n_new, pitch, frame_stride, 1, kind, st)) != 0) return rc;
    if ((rc = dvo_orb(ctx, 1, n_new, st)) != 0) return rc;
    if ((rc = dvo_pairs(ctx, 0, 0, n_new, K, st)) != 0) return rc;
    if ((rc = dvo_get_poses(ctx, 0, n_new, poses, kind, st)) != 0) return rc;
    ctx->carrySlot = n_new;
    return n_new;
}

int dvo_sequence(dvo_ctx* ctx, const uint8_t* frames, int n_frames, size_t pitch, size_t frame_stride, const double* K,
                 dvo_pose* poses, int kind, void* stream) {
    if (!ctx || !frames || !K || !poses || n_frames < 2) {
        if (ctx) ctx->err = "dvo_sequence: bad arguments";
        return DVO_E_INVALID;
    }
    const int B = ctx->nSlots - 1;   // new frames per batch after the first
    // Host destination: the records go through a pinned staging array, so the per-batch D2H copies stay asynchronous (a
    // copy into pageable memory would block the host and with it the overlap between batches).
    dvo_pose* dst = poses;
    if (kind == 1) {
        const size_t need = (size_t)(n_frames - 1);
        if (ctx->poseStageCap < need) {
            if (ctx->h_poseStage) { CK(cudaStreamSynchronize((cudaStream_t)stream)); CK(cudaFreeHost(ctx->h_poseStage)); ctx->h_poseStage = nullptr; }
            CK(cudaMallocHost(&ctx->h_poseStage, need * sizeof(dvo_pose)));
            ctx->poseStageCap = need;
        }
        dst = ctx->h_poseStage;
    }
    int done = std::min(n_frames, ctx->nSlots);
    int rc = dvo_sequence_step(ctx, frames, done, pitch, frame_stride, K, dst, kind, 1, stream);
    if (rc < 0) return rc;
    while (done < n_frames) {
        int nb = std::min(B, n_frames - done);
        rc = dvo_sequence_step(ctx, frames + (size_t)done * frame_stride, nb, pitch, frame_stride, K, dst + (done - 1), kind, 0, stream);
        if (rc < 0) return rc;
        done += nb;
    }
    if ((rc = dvo_sequence_flush(ctx, stream)) != 0) return rc;
    if (kind == 1) {
        CK(cudaStreamSynchronize((cudaStream_t)stream));
        memcpy(poses, dst, (size_t)(n_frames - 1) * sizeof(dvo_pose));
    }
    return DVO_OK;
}

// Host-side: cv.triangulatePoints of get_scaling_factor_from_triangulation (visual_odometry_v3.py:265) for a handful of
// fiducial corners.  Same arithmetic as OpenCV (4x4 DLT, cv::SVD's Jacobi path), so the unnormalised vectors -- sign
// included -- are the ones the reference measures its marker distance on.
int dvo_triangulate_points_host(const double* P0, const double* P1, const double* pts0, const double* pts1, int n, double* X) {
    if (!P0 || !P1 || n < 0 || (n > 0 && (!pts0 || !pts1 || !X))) return DVO_E_INVALID;
    for (int i = 0; i < n; ++i) {
        double q[4];
        dvo::cv_triangulate_point(P0, P1, pts0[2 * i], pts0[2 * i + 1], pts1[2 * i], pts1[2 * i + 1], q);
        for (int k = 0; k < 4; ++k) X[(size_t)k * n + i] = q[k];
    }
    return DVO_OK;
}

void dvo_profile_enable(int on) { prof_enable(on != 0); }
int dvo_profile_collect(double* ms, int* count, int n) {
    if (!ms || !count || n < 1) return DVO_E_INVALID;
    prof_collect(ms, count, n);
    return PF_COUNT;
}
const char* dvo_profile_name(int id) {
    static const char* names[PF_COUNT] = {"k_pyr_down", "k_fast_nms", "k_gather", "k_select", "k_angle_pack", "k_blur", "k_brief", "k_nn",
                                          "k_match_sort", "k_ransac", "k_cheirality", "k_pose_final", "k_load_or_ingest"};
    return (id >= 0 && id < PF_COUNT) ? names[id] : "";
}

}  // extern "C"
