/* libdvo -- C ABI of the B200-native visual-odometry hot path.
 *
 * Drop-in boundary for the per-frame-pair chain of theivyzhang/droplet_visual_odometry.  The reference has no FFI: its
 * hot path is five cv2 calls made from Python (scripts/visual_odometry_v3.py).  Each entry point below names the
 * reference call(s) it replaces; the Python host side (droplet_visual_odometry_b200/visual_odometry_v3.py) binds
 * them with ctypes and keeps the reference's class and method names.
 *
 * Conventions
 *   - plain C types only; `stream` is a cudaStream_t passed as void* (NULL = default stream).
 *   - return 0 on success, a negative DVO_E_* code on error; dvo_last_error(ctx) gives the message.  No exceptions
 *     cross this boundary.  Per-pair soft failures (too few matches, no model) are reported in dvo_pose.status.
 *   - "d_" pointers are device memory owned by the caller (torch tensors on the Python side); "h_" are host pointers.
 *     The library never frees caller memory.  Variable-length outputs are written up to the given capacity and the
 *     true count returned; a count above capacity fails with DVO_E_CAPACITY.
 *   - a context is bound to one device, is not thread-safe, and enqueues all work on the given stream; calls are
 *     asynchronous unless their name ends in _host or says they synchronise.
 *   - there is no CPU fallback anywhere behind this interface.
 */
#ifndef DVO_H_
#define DVO_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DVO_OK 0
#define DVO_E_INVALID (-1)     /* bad argument */
#define DVO_E_CUDA (-2)        /* CUDA runtime / driver error */
#define DVO_E_CAPACITY (-3)    /* caller buffer or context capacity too small */
#define DVO_E_NODEVICE (-4)    /* no usable CUDA device */

#define DVO_MATCH_CROSSCHECK 0 /* cv.BFMatcher(NORM_HAMMING, crossCheck=True).match   (visual_odometry_v3.py:75,219) */
#define DVO_MATCH_KNN_RATIO 1  /* knnMatch(k=2) + 0.75 ratio (:203,:227) + reverse 1-NN check (BASELINE config 4)   */

/* per-pair status */
#define DVO_PAIR_OK 0
#define DVO_PAIR_TOO_FEW_MATCHES 1 /* < 5 correspondences: cv.findEssentialMat returns None */
#define DVO_PAIR_NO_MODEL 2        /* RANSAC found no model with > 4 inliers                */

/* per-frame flags: cv2 keeps every tie at a retainBest boundary, the context has room for quota + 64 per level */
#define DVO_FRAME_TIES_TRUNCATED 1        /* more Harris-boundary ties than a level can hold: keypoint set != cv2's */
#define DVO_FRAME_CANDIDATES_TRUNCATED 2  /* FAST survivor list of a level overflowed (cannot happen with 3x3 NMS)  */
#define DVO_FRAME_KEYPOINTS_TRUNCATED 4   /* more keypoints than dvo_max_keypoints                                  */

typedef struct dvo_ctx dvo_ctx;

typedef struct dvo_config {
    int width, height;       /* frame size in pixels, each <= 4096                                         */
    int nfeatures;           /* cv.ORB_create(nfeatures=...)   reference literal: 500 (visual_odometry_v3.py:96) */
    int nlevels;             /* ORB pyramid levels, 1..8 (cv2 default 8)                                    */
    int fast_threshold;      /* cv2 default 20                                                             */
    int max_frames;          /* frame slots resident at once (batch size + 1 for the sequence runner)      */
    int matcher;             /* DVO_MATCH_*                                                                */
    int ransac_max_iters;    /* cv.findEssentialMat maxIters, cv2 default 1000                             */
    double ransac_prob;      /* 0.999  (visual_odometry_v3.py:300)                                         */
    double ransac_threshold; /* 1.0 px (visual_odometry_v3.py:300)                                         */
    double distance_thresh;  /* cv.recoverPose distanceThresh, cv2 default 50                              */
    float ratio;             /* 0.75   (visual_odometry_v3.py:227)                                         */
    int use_tma;             /* 1: stage FAST tiles with TMA (default); 0: plain loads (debug)             */
    int pipeline;            /* 1 (default): dvo_sequence / dvo_sequence_step run three batches at once on internal streams --
                                image half (upload, pyramid, FAST) of batch s+1, keypoint half of batch s, pair stage of
                                batch s-1 (second buffer set, allocated on first use); 0: every stage in order on the
                                caller's stream                                                                       */
    int ransac_exhaustive;   /* 0 (default): cv.findEssentialMat's adaptive stop.  1: every one of ransac_max_iters hypotheses
                                is solved and scored (whole-GPU batched solve + Sampson sweep; BASELINE configs[4] "all
                                hypotheses scored"); the first model with the highest inlier count wins                */
    int nn_engine;           /* cross-check matcher: 0 (default) int8 tensor-core GEMM (tcgen05, 256 - 2*hamming = dot of +-1 bytes);
                                1: XOR + POPC kernel (k_nn).  Both give cv2's matches bit for bit, and both serve either matcher: the ratio
                                matcher (DVO_MATCH_KNN_RATIO) keeps the runner-up per row on whichever engine is selected */
} dvo_config;

/* Result of one frame pair: what cv.findEssentialMat + cv.recoverPose return (visual_odometry_v3.py:297-306). */
typedef struct dvo_pose {
    double R[9];             /* recoverPose rotation, row-major             */
    double t[3];             /* recoverPose translation (unit norm)         */
    double E[9];             /* findEssentialMat result, row-major          */
    int32_t status;          /* DVO_PAIR_*                                  */
    int32_t n_matches;       /* correspondences fed to findEssentialMat     */
    int32_t n_inliers;       /* RANSAC mask population                      */
    int32_t n_good;          /* recoverPose return value (cheirality count) */
    int32_t ransac_iters;    /* iterations cv2's loop would have executed   */
    int32_t best_iter;       /* iteration that produced E                   */
    int32_t candidate;       /* 0..3: which (R1|R2, +-t) recoverPose chose   */
    int32_t n_prev, n_cur;   /* keypoints in the two frames                 */
    int32_t frame_flags;     /* DVO_FRAME_* bits of the pair's two frames, OR-ed: non-zero = a feature set was truncated */
} dvo_pose;

/* Device-side views of one frame's features: cv2 detectAndCompute output (visual_odometry_v3.py:373). */
typedef struct dvo_features {
    float* d_pt;        /* [cap][2] KeyPoint.pt                 */
    float* d_size;      /* [cap]    KeyPoint.size               */
    float* d_angle;     /* [cap]    KeyPoint.angle (degrees)    */
    float* d_response;  /* [cap]    KeyPoint.response (Harris)  */
    int32_t* d_octave;  /* [cap]    KeyPoint.octave             */
    uint8_t* d_desc;    /* [cap][32] descriptors                */
    int32_t* d_count;   /* [1]      number of keypoints         */
    int32_t capacity;
} dvo_features;

/* Device-side views of one pair's correspondences and masks. */
typedef struct dvo_pair_arrays {
    int32_t* d_matches;  /* [cap][3] (queryIdx, trainIdx, distance), order = sorted(bf.match(...), key=distance) */
    float* d_pts_prev;   /* [cap][2] cv.KeyPoint_convert of the previous-frame keypoints (visual_odometry_v3.py:355) */
    float* d_pts_cur;    /* [cap][2]                                                         (:358) */
    uint8_t* d_ransac_mask; /* [cap] findEssentialMat mask (0/1)   */
    uint8_t* d_pose_mask;   /* [cap] recoverPose mask (0/255)       */
    int32_t capacity;
} dvo_pair_arrays;

void dvo_default_config(dvo_config* cfg);
int dvo_create(const dvo_config* cfg, int device, dvo_ctx** out);
void dvo_destroy(dvo_ctx* ctx);
const char* dvo_last_error(const dvo_ctx* ctx);
const char* dvo_version(void);
int dvo_sizeof(int which);                    /* 0 dvo_config, 1 dvo_pose, 2 dvo_features, 3 dvo_pair_arrays: binding self-check */
int dvo_max_keypoints(const dvo_ctx* ctx);   /* capacity needed for dvo_features / dvo_pair_arrays */
int dvo_max_frames(const dvo_ctx* ctx);
long long dvo_kernel_launches(const dvo_ctx* ctx);   /* kernels launched by this context so far */

/* Upload n frames (device or host memory, row pitch `pitch`, `frame_stride` bytes apart) into slots
 * [slot0, slot0+n).  kind: 0 = source is device memory, 1 = source is (preferably pinned) host memory. */
int dvo_load_frames(dvo_ctx* ctx, const uint8_t* frames, int n, size_t pitch, size_t frame_stride, int slot0, int kind,
                    void* stream);

/* Image ingest while loading: replaces cv.cvtColor(BGR2GRAY) + cv.undistort(grey, cameraMatrix, distCoeffs,
 * newCameraMatrix) of ros_img_msg_to_opencv_image / undistort_image (visual_odometry_v3.py:110-135), bit-exact with cv2.
 * K, newK: 3x3 row-major (host); dist: k1 k2 p1 p2 [k3 [k4 k5 k6]] (host, n_dist <= 8).  channels = 1: frames passed to
 * dvo_load_frames / dvo_sequence* are distorted grey images; 3: distorted BGR (pitch counts bytes, >= 3 * width);
 * 0: switch ingest off again (frames are grey and already undistorted -- the default).  Builds the undistortion map
 * for the context's frame size; synchronises `stream`. */
int dvo_set_undistort(dvo_ctx* ctx, const double* K, const double* dist, int n_dist, const double* newK, int channels, void* stream);

/* ORB detect+describe on slots [slot0, slot0+n): replaces feature_detector.detectAndCompute
 * (visual_odometry_v3.py:373, called from compute_current_image_elements :370-379). */
int dvo_orb(dvo_ctx* ctx, int slot0, int n, void* stream);

/* DVO_FRAME_* flags of slots [slot0, slot0+n) after dvo_orb (host destination, synchronises).  A non-zero flag means the
 * slot's keypoint set is NOT what cv2 returns (truncated); callers must treat it as DVO_E_CAPACITY.  The same bits ride in
 * dvo_pose.frame_flags for the sequence runner. */
int dvo_get_frame_flags(dvo_ctx* ctx, int slot0, int n, int32_t* h_flags, void* stream);

/* Copy one slot's features into caller device buffers (async on stream). */
int dvo_get_features(dvo_ctx* ctx, int slot, const dvo_features* out, void* stream);

/* Match + pose for pairs (slot0+i, slot0+i+1), i < n, results in pair slots [pair0, pair0+n): replaces bf.match +
 * sorted (:219-221), cv.KeyPoint_convert (:355,:358), cv.findEssentialMat (:297-300), cv.recoverPose (:303-306).
 * K is the 3x3 row-major camera matrix (host pointer, read before the call returns). */
int dvo_pairs(dvo_ctx* ctx, int slot0, int pair0, int n, const double* K, void* stream);

/* The two halves of dvo_pairs as separate calls, for callers that use the reference's method-level surface.
 * dvo_match: bf.match + sorted (:219-221) + cv.KeyPoint_convert (:355,:358) only -- what get_matches_between_two_frames
 * (:191-239) does -- leaving the sorted (queryIdx, trainIdx, distance) list and the matched coordinates in pair slots
 * [pair0, pair0+n) (read them with dvo_get_match_count + dvo_get_pair_arrays); no RANSAC runs.
 * dvo_pose_pairs: cv.findEssentialMat + cv.recoverPose (:297-306) on the correspondences already in those pair slots (from
 * dvo_match); slot0 names the frame slots the pairs came from (keypoint counts of the record) or -1. */
int dvo_match(dvo_ctx* ctx, int slot0, int pair0, int n, const double* K, void* stream);
int dvo_pose_pairs(dvo_ctx* ctx, int slot0, int pair0, int n, const double* K, void* stream);
int dvo_get_match_count(dvo_ctx* ctx, int pair, int* h_count /*host, synchronises*/, void* stream);

/* Overwrite one slot's keypoint coordinates and descriptors with caller data (kind 0 device, 1 host source;
 * synchronises): lets get_matches_between_two_frames (:191-239) run on caller-supplied descriptors. */
int dvo_set_features(dvo_ctx* ctx, int slot, const float* pt, const uint8_t* desc, int n, int kind, void* stream);

/* findEssentialMat + recoverPose on caller correspondences (n x 2 float32 each; kind 0 device, 1 host source), result
 * in pair slot `pair`: replaces get_transformation_between_two_frames' two cv2 calls (:297-306) and serves the
 * RANSAC-heavy configuration (synthetic correspondences, no images). */
int dvo_pose_points(dvo_ctx* ctx, int pair, const float* pts_prev, const float* pts_cur, int n, const double* K, int kind,
                    void* stream);

/* Copy n pair results (device->device or device->host, async on stream). kind: 0 device dst, 1 host dst. */
int dvo_get_poses(dvo_ctx* ctx, int pair0, int n, dvo_pose* dst, int kind, void* stream);
int dvo_get_pair_arrays(dvo_ctx* ctx, int pair, const dvo_pair_arrays* out, void* stream);

/* cv.triangulatePoints(projMatr1, projMatr2, projPoints1, projPoints2) as get_scaling_factor_from_triangulation calls it
 * (visual_odometry_v3.py:263-291) on the fiducial corners of two frames.  Host arithmetic (four points per pair; no
 * context, no device work): P0, P1 are 3x4 row-major projection matrices, pts0/pts1 are n x 2, X receives the 4 x n
 * homogeneous points row-major, UNNORMALISED and with cv2's sign (the last row of cv::SVD's Vt): the reference measures
 * the marker edge on these raw vectors (:272-279), so the sign convention is part of the contract. */
int dvo_triangulate_points_host(const double* P0, const double* P1, const double* pts0, const double* pts1, int n, double* X);

/* Whole-sequence runner: consecutive-pair VO over n_frames frames; writes n_frames-1 dvo_pose records.
 * Frames are processed in batches of max_frames-1 with a one-frame carry, each frame's ORB computed once.
 * kind: 0 = frames and poses in device memory (async), 1 = host memory (H2D/D2H inside; synchronises). */
int dvo_sequence(dvo_ctx* ctx, const uint8_t* frames, int n_frames, size_t pitch, size_t frame_stride, const double* K,
                 dvo_pose* poses, int kind, void* stream);

/* One batch of a longer sequence: `first` != 0 starts a sequence (n_new frames -> n_new-1 pairs); otherwise the last
 * frame of the previous call is carried (its features are kept, not recomputed) and n_new frames give n_new pairs.
 * Returns the number of pose records written (>= 0) or a negative error.  Asynchronous; see dvo_sequence_flush. */
int dvo_sequence_step(dvo_ctx* ctx, const uint8_t* frames, int n_new, size_t pitch, size_t frame_stride, const double* K,
                      dvo_pose* poses, int kind, int first, void* stream);

/* With cfg.pipeline = 1 the pose records written by a dvo_sequence_step call are complete once the NEXT
 * dvo_sequence_step call, or this flush, has been enqueued and `stream` has reached that point.  dvo_sequence flushes by
 * itself.  A no-op without the pipeline.  Also call it before using the per-stage entry points (dvo_orb, dvo_pairs, taps)
 * after a pipelined sequence. */
int dvo_sequence_flush(dvo_ctx* ctx, void* stream);

/* Per-kernel CUDA-event timing for the roofline report (process-wide switch; collect synchronises the device and
 * returns the number of kernel ids; ms/count are totals since the previous collect). */
void dvo_profile_enable(int on);
int dvo_profile_collect(double* ms, int* count, int n);
const char* dvo_profile_name(int id);

/* Pipe-rate microbenchmarks for the roofline denominators that MEASURED_PEAKS.json does not carry: register-only
 * loops on `device`, best of a few launches, CUDA events.  out[DVO_PEAK_*] in flop/s (a multiply-add counts 2, fused or not),
 * POPC32/s, int8 op/s (tcgen05.mma.kind::i8, M128 N256 K32 issued back to back from resident shared-memory tiles).
 * Synchronises the device; takes ~50 ms. */
#define DVO_PEAK_FP32_FMA 0
#define DVO_PEAK_FP32_MUL_ADD 1   /* FMUL + FADD, no contraction: what the bit-exact float32 stages execute   */
#define DVO_PEAK_FP64_FMA 2
#define DVO_PEAK_FP64_MUL_ADD 3   /* DMUL + DADD, no contraction: what k_ransac / k_cheirality execute        */
#define DVO_PEAK_POPC 4
#define DVO_PEAK_INT8_TENSOR 5
#define DVO_PEAK_COUNT 6
int dvo_measure_peaks(int device, double* out, int n);

/* Stage taps for the parity tests (device destination, tightly packed rows of `w` bytes unless noted). */
int dvo_level_size(const dvo_ctx* ctx, int level, int* w, int* h, int* quota);
int dvo_tap_image(dvo_ctx* ctx, int slot, int level, int which /*0 pyramid, 1 blurred*/,
                  uint8_t* d_dst, void* stream);
int dvo_tap_candidates(dvo_ctx* ctx, int slot, int level, uint32_t* d_dst /*packed score<<24|y<<12|x*/, int capacity,
                       int* h_count /*host, synchronises*/, void* stream);
int dvo_tap_ransac(dvo_ctx* ctx, int pair, int32_t* h_state8 /*host, synchronises*/, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DVO_H_ */
