/* stereo_b200.h — C ABI of the B200-native stereo-matching hot path.
 *
 * This is the drop-in boundary for the path BASELINE.json names: the reference's
 * CStereoMatching (NCC matching, constraint filters, rematch, median, refinement) and the
 * per-point triangulation DisparityToCloud.  The reference has no FFI; its boundary is the C++
 * class surface.  Every entry point below cites the reference code it replaces,
 * all paths relative to /root/reference/reconstruction/.
 *
 * Conventions: plain pointers and sizes, no C++/torch types; every call returns an sb200 status
 * (0 = OK); one context per GPU and per host thread; no global state.  All host buffers are
 * caller-owned, C-contiguous, reference layout: images are interleaved BGR u8 rows of 3*W bytes
 * (cv::Mat CV_8UC3), masks u8 rows of W bytes, disparity maps row-major s16 or f64 (cv::Mat
 * CV_16S / CV_64F, the two types MatchOneLayer alternates between).
 *
 * There is NO CPU fallback behind this ABI: a machine without a CUDA device gets
 * SB200_ERR_NO_DEVICE from sb200_ctx_create and nothing else works.
 */
#ifndef STEREO_B200_H
#define STEREO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define SB200_API __attribute__((visibility("default")))
#else
#define SB200_API
#endif

#define SB200_NOMATCH (-10000) /* CStereoMatching.h:9 */

enum sb200_status {
  SB200_OK = 0,
  SB200_ERR_NO_DEVICE = 1,        /* no CUDA device / wrong architecture */
  SB200_ERR_BAD_ARG = 2,
  SB200_ERR_CUDA = 3,             /* a CUDA call failed; sb200_last_error() has the text */
  SB200_ERR_STATE = 4,            /* call order violated (e.g. match before upload) */
  SB200_ERR_DEGENERATE_MARGIN = 5 /* the reference exit(0)s here: CStereoMatching.cpp:827-830,954-958 */
};

/* Stage ids of MatchOneLayer, CStereoMatching.cpp:51-109 (same numbering as SURVEY.md §3.3). */
enum sb200_stage {
  SB200_STAGE_FIND_MARGIN = 1,   /* :51-52   FindMargin x2                          */
  SB200_STAGE_INITIAL_MATCH = 2, /* :53-62   Lowest/HighLevelInitialMatch x2        */
  SB200_STAGE_SMOOTH = 3,        /* :66-67   SmoothConstraint x2                    */
  SB200_STAGE_ORDER = 4,         /* :71-72   OrderConstraint x2                     */
  SB200_STAGE_UNIQUE_1 = 5,      /* :75      UniquenessContraint<short>             */
  SB200_STAGE_REMATCH = 6,       /* :80-81   Rematch x2 (SetBoundary_smooth + NCC)  */
  SB200_STAGE_UNIQUE_2 = 7,      /* :86      UniquenessContraint<short>             */
  SB200_STAGE_MEDIAN = 8,        /* :89-90   MedianFilter x2                        */
  SB200_STAGE_REFINE = 9,        /* :95-98   DisparityRefine x2                     */
  SB200_STAGE_UNIQUE_3 = 10      /* :109     UniquenessContraint<double>            */
};

/* Boundary, CManageData.h:10-14 (same field order as the reference struct). */
typedef struct sb200_boundary {
  int32_t YL, YR, XL, XR, width, height;
} sb200_boundary;

typedef struct sb200_ctx sb200_ctx;

/* ---- context ---------------------------------------------------------------------------------
 * Replaces CStereoMatching::Init (CStereoMatching.cpp:5-13: radii, ws, disparity_offset) plus the
 * sizes CManageData::Init reads from config.yml (CManageData.cpp:26-42: PyrmNum, LowestLevelWidth/
 * Height) and m_OriginSize (CManageData.cpp:68-69), which sets the triangulation scale
 * (CStereoMatching.cpp:692).  Allocates every device buffer for one camera pair once. */
SB200_API int sb200_ctx_create(sb200_ctx** out, int device, int pyrm_num, int lowest_w, int lowest_h, int origin_w,
                     int origin_h, int radius, double ws, int offset);
SB200_API void sb200_ctx_destroy(sb200_ctx* ctx);
SB200_API const char* sb200_last_error(const sb200_ctx* ctx);
SB200_API const char* sb200_status_string(int status);

/* ---- staging a pair ----------------------------------------------------------------------------
 * What CStereoMatching::Rectify leaves in cam[pair][k].image / .mask (CStereoMatching.cpp:154-158)
 * for the two views, at the top pyramid level (lowest << (pyrm_num-1)).  The call copies them to
 * HBM and runs ConstructPyrm (:1040-1053, pyrDown of image and mask per level) and FindMargin for
 * every level on the device.  Host buffers may be pageable or pinned. */
SB200_API int sb200_pair_upload(sb200_ctx* ctx, const uint8_t* bgr0, const uint8_t* bgr1, const uint8_t* mask0,
                      const uint8_t* mask1);
/* Same, from buffers that already live in this GPU's memory (device pointers, same layout): the
 * "images staged into HBM once" case of the north star; used for the device-resident bench figure. */
SB200_API int sb200_pair_stage_device(sb200_ctx* ctx, const void* bgr0_dev, const void* bgr1_dev, const void* mask0_dev,
                            const void* mask1_dev);
/* Q (4x4, AFTER the sign flip at :138), R_final (3x3, :132), T_final (3, :133), row-major f64. */
SB200_API int sb200_pair_set_calib(sb200_ctx* ctx, const double* Q, const double* R_final, const double* T_final);

/* ---- Rectify (the step before the hot path, CStereoMatching.cpp:117-168) -------------------------------
 * Calibration half (:121-145), host, double precision, no context needed: from the two cameras' intrinsics K (3x3) and
 * extrinsics [R|t] (3x4) as CManageData::Init reads them (CManageData.cpp:59-60) to what the rest of the path uses.
 * R_new[2][9] = R_new[0..1] of cv::stereoRectify(K0, 0, K1, 0, OriginSize, R, T, ..., flags 0, alpha -1, OriginSize);
 * P_scaled[2][12] = P with rows 0-1 times scale (:143, the newCameraMatrix of initUndistortRectifyMap);
 * P_final[2][12] = cam[pair][k].P after :145; Q after the sign flip (:138); R_final (:132); T_final (:133). */
SB200_API int sb200_rectify_calib(const double* K0, const double* Rt0, const double* K1, const double* Rt1, int origin_w, int origin_h,
                                  int lowest_w, int pyrm_num, double* R_new, double* P_scaled, double* P_final, double* Q,
                                  double* R_final, double* T_final);
/* cv::stereoRectify changed between the OpenCV the reference links (2.4.5: the SMALLER focal length of the two cameras, image
 * corners taken at (nx, ny)) and current OpenCV (4.13: the MEAN focal length, corners at (nx-1, ny-1)).  opencv_compat = 245 or
 * 413 selects the behaviour; sb200_rectify_calib uses 245 - the reference's own dependency - unless the environment says
 * SB200_OPENCV_COMPAT=413.  Only the 413 behaviour can be pinned by vectors here (tests/golden/rectify_cv2.npz). */
SB200_API int sb200_rectify_calib_compat(const double* K0, const double* Rt0, const double* K1, const double* Rt1, int origin_w, int origin_h,
                                         int lowest_w, int pyrm_num, int opencv_compat, double* R_new, double* P_scaled, double* P_final,
                                         double* Q, double* R_final, double* T_final);
/* cv::stereoRectify alone, for pinning against OpenCV vectors. */
SB200_API int sb200_stereo_rectify_host(const double* K1, const double* K2, int nx, int ny, const double* R, const double* T,
                                        int opencv_compat, double* R1, double* R2, double* P1, double* P2, double* Q);
/* Image half for one view (:144-158) on the device: initUndistortRectifyMap(K, 0, R_new, P_scaled, top size, CV_16SC2), remap
 * (INTER_LINEAR) of the colour image and of the mask, erode of the mask with the 3*2^(L-1) ellipse.  src_* are the ORIGINAL
 * frames (src_w x src_h, interleaved BGR / grey) in host memory; the results land in the context's top-level buffers
 * (sb200_get_level reads them back).  use_given_maps != 0 skips the map computation and uses sb200_set_rectify_maps. */
SB200_API int sb200_rectify_view(sb200_ctx* ctx, int view, const uint8_t* src_bgr, const uint8_t* src_mask, int src_w, int src_h,
                                 const double* K, const double* R_new, const double* P_scaled, int use_given_maps);
/* ConstructPyrm + FindMargin after both views were rectified (what sb200_pair_upload does after its copy). */
SB200_API int sb200_pair_build(sb200_ctx* ctx);
/* Parity hooks: the fixed-point maps of the view rectified last (map1: H*W*2 s16 = (x, y), map2: H*W u16), the remapped mask
 * before erosion. */
SB200_API int sb200_get_rectify_maps(sb200_ctx* ctx, int16_t* map1, uint16_t* map2);
SB200_API int sb200_set_rectify_maps(sb200_ctx* ctx, const int16_t* map1, const uint16_t* map2);
SB200_API int sb200_get_remapped_mask(sb200_ctx* ctx, uint8_t* out);

/* ---- the hot path ------------------------------------------------------------------------------
 * MatchAllLayer's per-pair body, CStereoMatching.cpp:21-29: MatchOneLayer for every level
 * (stages 1-10) then DisparityToCloud.  Asynchronous work is finished when the call returns.
 * n_points receives the number of emitted 3-D points. */
SB200_API int sb200_match_pair(sb200_ctx* ctx, int64_t* n_points);
/* MatchOneLayer(disparity, level), CStereoMatching.cpp:36-113. */
SB200_API int sb200_match_one_layer(sb200_ctx* ctx, int level);
/* One stage of MatchOneLayer (the reference's SaveMat dump points, :63-111) — parity hook. */
SB200_API int sb200_run_stage(sb200_ctx* ctx, int level, int stage);
/* DisparityRefine normally runs 30+30*level sweeps (:95); n >= 0 overrides, n < 0 restores. */
SB200_API int sb200_set_refine_iters(sb200_ctx* ctx, int n);

/* ---- disparity maps (cv::Mat disparity[2] of MatchAllLayer, :22) --------------------------------
 * elem_size is 2 (s16, stages 2-8) or 8 (f64, stages 9-10), 0 before the first match. */
SB200_API int sb200_disparity_info(const sb200_ctx* ctx, int* width, int* height, int* elem_size);
SB200_API int sb200_get_disparity(sb200_ctx* ctx, int dir, void* host_out);
/* Teacher forcing for stage-by-stage parity: replace disparity[dir] (both dirs must share a type). */
SB200_API int sb200_set_disparity(sb200_ctx* ctx, int dir, const void* host_in, int width, int height, int elem_size);
/* BL / BR of SetBoundary_smooth (CStereoMatching.cpp:817-942), the commented bl.dat dump (:515). */
SB200_API int sb200_get_rematch_bounds(sb200_ctx* ctx, int dir, int16_t* bl_out, int16_t* br_out);

/* ---- pyramid and margins ------------------------------------------------------------------------
 * imagePyrm[level][view] / maskPyrm[level][view] (CManageData.h) and `margin[view]` at `level`
 * (FindMargin, CStereoMatching.cpp:1011-1038).  Either output pointer may be NULL. */
SB200_API int sb200_get_level(sb200_ctx* ctx, int level, int view, uint8_t* bgr_out, uint8_t* mask_out);
SB200_API int sb200_get_margin(const sb200_ctx* ctx, int level, int view, sb200_boundary* out);

/* ---- triangulation ------------------------------------------------------------------------------
 * DisparityToCloud<double>(disparity[0], maskPyrm[L-1][0], Q, L-1, true, pair), :682-761: one point
 * per masked, matched pixel of view 0 in row-major order — the sequence of InsertPoint calls (:751).
 * xyz is what InsertPoint receives (f64), bgr the three source bytes the PLY branch writes
 * (:754-756), pix the flat index y*W+x of the source pixel.  Any output pointer may be NULL. */
SB200_API int sb200_triangulate(sb200_ctx* ctx, int64_t* n_points);
SB200_API int sb200_get_points(sb200_ctx* ctx, double* xyz_out, uint8_t* bgr_out, int32_t* pix_out);
/* Device-resident point buffers of the last triangulation, for the per-pair all-gather over NCCL
 * (SURVEY.md §8e); valid until the next sb200_triangulate / sb200_match_pair on this context. */
SB200_API int sb200_points_device(sb200_ctx* ctx, void** xyz_dev, void** bgr_dev, void** pix_dev, int64_t* n_points);

/* ---- one-call host-buffer entry (what the C++ CStereoMatching::MatchAllLayer mirror calls) ------
 * upload + set_calib + match_pair + download of the points, H2D/D2H included.  xyz_out/bgr_out/
 * pix_out must hold capacity points; returns SB200_ERR_BAD_ARG if more are produced. */
SB200_API int sb200_match_pair_host(sb200_ctx* ctx, const uint8_t* bgr0, const uint8_t* bgr1, const uint8_t* mask0,
                          const uint8_t* mask1, const double* Q, const double* R_final, const double* T_final,
                          double* xyz_out, uint8_t* bgr_out, int32_t* pix_out, int64_t capacity,
                          int64_t* n_points);

/* ---- instrumentation ----------------------------------------------------------------------------
 * The stream all kernels of this context are launched on (a cudaStream_t), so callers can time
 * with CUDA events on the launching stream. */
SB200_API void* sb200_stream(sb200_ctx* ctx);
/* Number of kernels enqueued by this context since creation (bench.py's gpu_launches).  sb200_match_pair replays the fixed
 * stage order of a pair (MatchOneLayer x PyrmNum + DisparityToCloud, CStereoMatching.cpp:21-29) as ONE CUDA graph launch whose
 * nodes are these kernels; sb200_graph_info counts the graph launches and how often the executable graph had to be rebuilt
 * rather than updated in place. */
SB200_API int64_t sb200_launch_count(const sb200_ctx* ctx);
SB200_API int sb200_graph_info(const sb200_ctx* ctx, int64_t* graph_launches, int64_t* graph_instantiations);
/* Accumulated device time (ms, CUDA events) per stage id 0..15 since the last reset; index 0 is the
 * pyramid build, 1-10 the MatchOneLayer stages, 11 the triangulation.  Enabled by
 * sb200_set_profiling(ctx, 1); adds an event pair around each stage. */
SB200_API int sb200_set_profiling(sb200_ctx* ctx, int enable);
SB200_API int sb200_get_stage_ms(sb200_ctx* ctx, double* ms16, int reset);
/* The same for one stage at one pyramid level (e.g. stage 2 at the top level = the HighLevelInitialMatch NCC search,
 * CStereoMatching.cpp:231-308, both directions). */
SB200_API int sb200_get_stage_level_ms(sb200_ctx* ctx, int stage, int level, double* ms, int reset);
/* The dominant kernel (the DisparityRefine sweep, CStereoMatching.cpp:590-674) for the roofline line of
 * bench.py: device time of all sweeps since the last reset (ms, CUDA events on the context stream,
 * profiling must be enabled), the number of sweeps, and the algorithmic pixel-iterations they covered
 * (sweeps x margin.width x margin.height; 22 algorithmic bytes each), for one pyramid
 * level or, with level < 0, summed over all levels. */
SB200_API int sb200_get_refine_profile(sb200_ctx* ctx, int level, double* sweep_ms, int64_t* sweep_launches, int64_t* px_iters, int reset);
/* Counters: [0] = pixels the integer screening pass of the NCC searches left to the exact FP64 search (near ties,
 * flat windows, wide ranges), [1] = out-of-table pixel evaluations of the refinement kernel. */
SB200_API int sb200_get_refine_counters(sb200_ctx* ctx, int64_t* out2, int reset);
/* NCC searches (CStereoMatching.cpp:202-218, :268-300, :535-562): [0] = pixels the tile kernel of HighLevelInitialMatch handed to the
 * list kernels (holes with a carried range, near ties, target strips outside the staged box), [1] = pixels that reached the exact
 * FP64 pass. */
SB200_API int sb200_get_search_counters(sb200_ctx* ctx, int64_t* out2, int reset);

/* ---- the sink's per-pair filter, on the device (next row f-3) ---------------------------------------
 * What CCloudOptimization::filter(idx) does to the points InsertPoint collected for one camera pair before it meshes them
 * (CloudOptimization/CCloudOptimization.cpp:64-121): pcl::StatisticalOutlierRemoval (setMeanK / setStddevMulThresh, :79-83),
 * pcl::NormalEstimationOMP with a radius search (:99-106) and the flip of every normal towards CamCenter[idx] (:109-116).
 * Points are narrowed to float32 as InsertPoint does (:59-62).  PCL is not under /root/reference: parity is against
 * oracle/sink_oracle.py.  xyz: n x 3 f64 in InsertPoint order (host).  Output: the kept points in their original
 * order, 7 floats each (x y z normal_x normal_y normal_z curvature = the PointNormal fields savePLYFileBinary writes, :119),
 * kept_index (optional): their indices in the input.  stats5 (optional): mean, stddev, threshold of the SOR distances, device
 * milliseconds, number of queries whose search ring had to grow.  No CPU path: SB200_ERR_NO_DEVICE without a GPU. */
SB200_API int sb200_sink_filter(int device, const double* xyz, int64_t n, int sor_meank, double sor_std_mul, double normal_radius,
                                const double* cam_center, float* out_xyz_normal_curv, int32_t* kept_index, int64_t capacity,
                                int64_t* n_kept, double* stats5);
SB200_API const char* sb200_sink_last_error(void);

/* ---- the exchange step: per-pair point buffers gathered on every rank ------------------------
 * The reference processes the camera pairs serially in one process (CStereoMatching.cpp:17-33) and appends every pair's cloud
 * to the sink in pair order (CloudOptimization/CCloudOptimization.cpp:123).  Here the pairs shard one per GPU; after
 * DisparityToCloud every rank contributes its pair's points and receives everybody's, rank-major (= pair order), not padded:
 * one NCCL all-gather of the counts, then one grouped collective in which each rank broadcasts its own xyz / bgr / pix block.
 * One communicator per GPU: one process per GPU (the unique id travels over the launcher's own bootstrap) or one host thread
 * per GPU inside a process.  NCCL is bound at run time (libnccl.so.2); without it these calls return SB200_ERR_NO_DEVICE and
 * sb200_comm_last_error(NULL) says why.  `producers` = contexts per rank that submit pairs (1 for the synchronous call),
 * `slots` = snapshots a producer may have waiting for the exchange (2). */
#define SB200_UNIQUE_ID_BYTES 128
typedef struct sb200_comm sb200_comm;
SB200_API int sb200_comm_unique_id(void* id_out /* SB200_UNIQUE_ID_BYTES, call on one rank and distribute */);
SB200_API int sb200_comm_init(sb200_comm** out, int device, int rank, int nranks, const void* unique_id, int producers, int slots);
SB200_API void sb200_comm_destroy(sb200_comm* comm);
SB200_API const char* sb200_comm_last_error(const sb200_comm* comm);
/* Synchronous form: gathers the points of `ctx`'s last triangulation from every rank.  counts_out[nranks]; the host buffers
 * (any may be NULL) receive the rank-major concatenation and must hold `capacity` points; total_out = sum of the counts. */
SB200_API int sb200_allgather_points(sb200_comm* comm, sb200_ctx* ctx, int64_t* counts_out, double* xyz_host, uint8_t* bgr_host,
                                     int32_t* pix_host, int64_t capacity, int64_t* total_out);
/* Overlapped form.  submit: snapshot ctx's points on ctx's own stream and return at once; an exchange thread issues the
 * collectives in ticket order (ticket = seq * producers + producer, the same sequence on every rank) on its own low-priority
 * stream, so a context never waits for another rank and runs up to `slots` pairs ahead.  wait: block until `ticket` is gathered
 * (optionally copy it to host buffers); the two most recent results stay on the device (sb200_exchange_device).  drain: wait
 * until tickets [0, n_tickets) are gathered and the exchange stream is idle. */
SB200_API int sb200_exchange_submit(sb200_comm* comm, sb200_ctx* ctx, int producer, int64_t seq);
SB200_API int sb200_exchange_wait(sb200_comm* comm, int64_t ticket, int64_t* counts_out, double* xyz_host, uint8_t* bgr_host,
                                  int32_t* pix_host, int64_t capacity, int64_t* total_out);
SB200_API int sb200_exchange_device(sb200_comm* comm, int64_t ticket, void** xyz_dev, void** bgr_dev, void** pix_dev, int64_t* total);
SB200_API int sb200_exchange_drain(sb200_comm* comm, int64_t n_tickets);
/* Flow control for a consumer that reads every result (sb200_exchange_wait for every ticket, in order): when enabled, the exchange
 * thread overwrites the result of ticket t-2 only after sb200_exchange_wait(t-2) has returned.  Off by default (a producer-only
 * user such as bench.py never waits for single tickets). */
SB200_API int sb200_comm_set_consumer(sb200_comm* comm, int enable);
/* Device time spent inside the collectives (CUDA events on the exchange stream), bytes received, exchanges done. */
SB200_API int sb200_comm_stats(sb200_comm* comm, double* collective_ms, int64_t* bytes_received, int64_t* n_exchanges, int reset);

/* glibc-compatible exp() used by the refinement weights (see DESIGN.md "exp"); host twin of the
 * device function, exported so the CPU tests can pin it against the C library. */
SB200_API double sb200_exp_host(double x);

#ifdef __cplusplus
}
#endif
#endif /* STEREO_B200_H */
