// Internal layout of a dvo context: HBM buffers, per-level geometry, kernel launch entry points.
//
// HBM layout (DESIGN.md "Data layout"): everything is per frame SLOT (one slot per frame in flight).
//   pyr / blur       : u8 images, 8 levels each, row pitch = roundup(w, 128), levels 256-B aligned, slot stride fixed
//   tileList/Cnt/Tot : NMS survivors of every 128x32 FAST tile, in (row, x) order, with per-(tile,row) and per-tile counts
//   cand             : u32 per candidate, packed score<<24 | y<<12 | x, raster order per level
//   pairs            : u64 per first-cut survivor, harris f32 bits <<32 | packed xy, cv2 order
//   fin*             : final keypoints per level in cv2 order
//   feat*            : per-frame compact feature arrays (what cv2 returns from detectAndCompute)
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>
#include <stdint.h>
#include <string>
#include <vector>
#include "../../include/dvo.h"

namespace dvo {

constexpr int kMaxLevels = 8;
constexpr int kTileW = 128;          // image-kernel tile (outputs)
constexpr int kTileH = 32;
constexpr int kFastHaloL = 16;       // TMA needs the inner start coordinate 16-byte aligned: left halo is 16 px (4 used)
constexpr int kFastBoxW = 160;       // TMA box: 16 + tile + 16 (multiple of 16 bytes)
constexpr int kFastBoxH = 40;
constexpr int kTileListCap = (kTileW / 2) * (kTileH / 2);   // strict 3x3 NMS: no two survivors touch
constexpr int kFinSlack = 64;        // extra per-level room for ties at the Harris boundary
constexpr int kMaxImageDim = 4096;   // 12-bit packed coordinates
constexpr int kSelectSmallSmemBytes = 64 * 1024;     // few-frame k_select launch for the small pyramid levels (3 CTAs per SM)
constexpr long long kSelectBigLevelPixels = 600000;  // levels above this go to the big-level launch (~0.1 B of list per pixel)
constexpr int kSelectSmemBytes = 160 * 1024;   // few-frame k_select working array: 40960 candidates per level stay on chip
constexpr int kMaxModels = 10;

struct LevelGeom {
    int w, h, pitch;
    int rowBase;         // first row in the per-slot row arrays
    int quota;           // nfeaturesPerLevel
    int candCap, candBase;
    int finCap, finBase;
    int tilesX, tilesY, tileBase;
    int rbBase;          // first 32-row block in the per-slot row-block index space
    float scale, invScale;
    unsigned long long off;   // byte offset inside a slot's image buffer
};

struct OrbGeom {
    int nlevels;
    int rowsPerSlot, candPerSlot, finPerSlot, maxkp;
    int tilesPerFrame, rowBlocksPerFrame;
    int fastThreshold;
    unsigned long long slotStride;   // bytes between slots in pyr/blur/map
    LevelGeom lv[kMaxLevels];
};

struct OrbBuffers {
    uint8_t* pyr;
    uint8_t* blur;
    uint16_t* tileCnt;       // [slots][tilesPerFrame][kTileH]  NMS survivors per (tile, row)
    int* tileTot;            // [slots][tilesPerFrame]          ... per tile
    uint32_t* tileList;      // [slots][tilesPerFrame][kTileListCap] survivors of a tile in (row, x) order, packed like cand
    uint32_t* cand;          // [slots][candPerSlot]
    int* candCount;          // [slots][8]
    unsigned long long* pairs;   // [slots][candPerSlot]
    uint32_t* selWork;       // [slots][candPerSlot]     k_select pass-1 working copy when a level does not fit in smem
    uint32_t* selList;       // [slots][2*candPerSlot]   k_select stopper lists (left | right) per level
    uint32_t* finXY;         // [slots][finPerSlot]
    float* finResp;          // [slots][finPerSlot]
    int* finCount;           // [slots][8]
    int* selDbg;             // [slots][8][4]: n candidates, n after fast cut, n final, flags
    // compact per-frame features
    float* featPt;           // [slots][maxkp][2]
    float* featResp;         // [slots][maxkp]
    float* featAngle;        // [slots][maxkp]
    int* featOctave;         // [slots][maxkp]
    uint32_t* featXY;        // [slots][maxkp] packed level coords
    float* featCS;           // [slots][maxkp][2] cos, sin of the keypoint angle (float32 roundings of the float64 values)
    uint8_t* featDesc;       // [slots][maxkp][32]
    int* featCount;          // [slots]
    int* frameFlags;         // [slots] DVO_FRAME_* bits of the last dvo_orb on the slot (0 = the feature set is complete)
    const uint32_t* resizeTab;   // per level: x table then y table, packed ofs<<16 | c1
    const uint32_t* tileInfo;    // [tilesPerFrame] level | tileX << 4 | tileY << 16 for the 128x32 image-kernel tiles
    int resizeTabOff[kMaxLevels][2];
};

struct TensorMaps {
    CUtensorMap pyr[kMaxLevels];          // FAST tile boxes (kFastBoxW x kFastBoxH), one map per level
    CUtensorMap pyrSrc[2][kMaxLevels];    // k_pyr_down source windows of level L: [0] kPyrSrcW x kPyrSrcH, [1] kPyrSrcW x kPyrSrcHSmall
};
constexpr int kPyrSrcW = 176, kPyrSrcH = 82, kPyrSrcHSmall = 24;   // source window of a 128 x 64 / 128 x 16 output tile at ratio 1.2

struct PairGeom {
    int maxkp;            // descriptor capacity per slot
    int maxMatches;       // == maxkp
    int sortCap;          // power of two >= maxkp
    int matcher;          // 0 crosscheck, 1 knn ratio + reverse check
    int maxIters;
    int exhaustive;       // score all maxIters hypotheses (no adaptive stop)
    int nnTensor;         // cross-check matcher on the tensor cores (nn_tensor.cu) instead of XOR+POPC
    int numSms;
    int rngCount;         // entries of PairBuffers::rngStates
    double prob, threshold, distThresh;
    float ratio;
};

// per-pair scratch carried between the three pose kernels
struct PoseScratch { double R1[9], R2[9], t[3]; int good[4]; int nInl; int pad[3]; };

struct PairBuffers {
    // per pair slot
    int* nnIdx;           // [pairs][2 dirs][maxkp]      best index
    int* nnDist;          // [pairs][2][maxkp]
    int* nn2Dist;         // [pairs][maxkp]              second-best distance (forward only, knn mode)
    int8_t* descX;        // [pairs + 1][descXRows][256] descriptors of one dvo_pairs call as +-1 bytes (tensor-core matcher)
    int descXRows;        // maxkp rounded up to the matcher's column tile
    int* matches;         // [pairs][maxkp][3]           (queryIdx, trainIdx, distance) sorted by (distance, queryIdx)
    int* matchCount;      // [pairs]
    float* ptsPrev;       // [pairs][maxkp][2]
    float* ptsCur;        // [pairs][maxkp][2]
    double* normPts;      // [pairs][maxkp][4]           x1 y1 x2 y2 normalised
    int* samples;         // [pairs][maxIters][5]
    int* ransacState;     // [pairs][8]: maxGood, niters, done, bestIter, bestModel, rng lo, rng hi, hasBest
    double* bestE;        // [pairs][9]
    uint8_t* ransacMask;  // [pairs][maxkp]
    uint8_t* poseMask;    // [pairs][maxkp]
    const unsigned long long* rngStates;   // [rngCount] cv::RNG(-1) states after 1, 2, ... steps (exhaustive mode only)
    int* exStart;         // [pairs][maxIters] offset of each iteration's first draw in rngStates
    double* exModels;     // [pairs][maxIters][10][9]  exhaustive mode only
    int* exCount;         // [pairs][maxIters]
    int* exGood;          // [pairs][maxIters][10]
    dvo_pose* poses;      // [pairs]
    PoseScratch* poseScratch;   // [pairs]
};

// ---- image ingest (cv.cvtColor BGR2GRAY + cv.undistort; /root/reference/scripts/visual_odometry_v3.py:110-135)
struct IngestParams {
    double ir[9];            // inv(newCameraMatrix)
    double fx, fy, cx, cy;   // cameraMatrix
    double k[8];             // k1 k2 p1 p2 k3 k4 k5 k6
};
struct IngestBuffers {
    uint4* map = nullptr;        // [h][w]: x = (u16)sx | (u16)sy << 16 (integer source position, int16), y = w00 | w01 << 16, z = w10 | w11 << 16
    const uint2* wtab = nullptr; // [32*32]: four uint16 bilinear weights (taps 00 01 10 11), sum 32768
    int channels = 0;            // 0 = ingest disabled (frames are already grey + undistorted); 1 grey, 3 BGR
};

// ---- per-kernel CUDA-event profiler (bench.py's roofline pass; off by default, zero cost when off)
enum ProfId { PF_PYR = 0, PF_FAST, PF_COMPACT, PF_SELECT, PF_ANGLE, PF_BLUR, PF_BRIEF, PF_NN, PF_SORT, PF_RANSAC,
              PF_CHEIRALITY, PF_POSE_FINAL, PF_INGEST, PF_COUNT };
void prof_begin(int id, cudaStream_t st);
void prof_end(int id, cudaStream_t st);
struct ProfScope {
    int id; cudaStream_t st;
    ProfScope(int i, cudaStream_t s) : id(i), st(s) { prof_begin(id, st); }
    ~ProfScope() { prof_end(id, st); }
};

// Side stream + events of a context: k_blur depends only on the pyramid, so it runs beside the latency-bound
// compact -> select -> angle chain; host frames are staged through a double buffer on a copy stream.
struct SideStreams {
    cudaStream_t side = nullptr, side2 = nullptr, copy = nullptr;
    cudaEvent_t evFork = nullptr, evJoin = nullptr, evFork2 = nullptr, evJoin2 = nullptr;
    cudaEvent_t evCopied[2] = {nullptr, nullptr}, evStageFree[2] = {nullptr, nullptr};
    bool stageUsed[2] = {false, false};
    uint8_t* stage[2] = {nullptr, nullptr};
    int stageIdx = 0;
};

// ---- launchers (orb_kernels.cu / pair_kernels.cu) ---------------------------------------------------------------
void launch_orb(const OrbGeom& g, const OrbBuffers& b, const TensorMaps* tmaps, bool useTma, int slot0, int nSlots,
                cudaStream_t st, const SideStreams* ss);
void launch_orb_image(const OrbGeom& g, const OrbBuffers& b, const TensorMaps* tmaps, bool useTma, int slot0, int nSlots, cudaStream_t st);
void launch_orb_keypoints(const OrbGeom& g, const OrbBuffers& b, int slot0, int nSlots, cudaStream_t st, const SideStreams* ss);
void launch_build_undistort_map(const OrbGeom& g, const IngestBuffers& ib, const IngestParams& prm, cudaStream_t st);
void launch_ingest(const OrbGeom& g, const OrbBuffers& b, const IngestBuffers& ib, const uint8_t* d_src, int n, size_t pitch,
                   size_t frameStride, int slot0, cudaStream_t st);
void launch_load_frames(const OrbGeom& g, const OrbBuffers& b, const uint8_t* d_src, int n, size_t pitch, size_t frameStride,
                        int slot0, cudaStream_t st);
int nn_tensor_rows(int maxkp);
cudaError_t nn_tensor_init();
void launch_nn_tensor(const OrbGeom& og, const OrbBuffers& ob, const PairGeom& pg, const PairBuffers& pb, int slotA0, int pair0,
                      int nPairs, int numSms, cudaStream_t st);
void launch_match(const OrbGeom& og, const OrbBuffers& ob, const PairGeom& pg, const PairBuffers& pb, int slotA0,
                  int pair0, int nPairs, const double* K, cudaStream_t st);
void launch_pairs(const OrbGeom& og, const OrbBuffers& ob, const PairGeom& pg, const PairBuffers& pb, int slotA0,
                  int pair0, int nPairs, const double* K, cudaStream_t st);

void launch_ransac_pose(const OrbGeom& og, const OrbBuffers& ob, const PairGeom& pg, const PairBuffers& pb, int slotA0,
                        int pair0, int nPairs, const double* K, cudaStream_t st);
void launch_points_prep(const PairGeom& pg, const PairBuffers& pb, int pair, int n, const double* K, cudaStream_t st);

}  // namespace dvo
