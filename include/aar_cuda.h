/*
 * aar_cuda.h — C ABI of the B200 (sm_100a) joint-optimisation path of automatic-ar.
 *
 * The reference (HSarham/automatic-ar) has no FFI layer: the seam of its hot path is the C++
 * call `MultiCamMapper::solve()` / `track()` (libs/multicam_mapper.cpp:419-443) which binds
 * `error_function` / `jacobian_function` into `ucoslam::SparseLevMarq<double>::solve`
 * (libs/sparselevmarq.h:80, 118).  This header is what a binding of that seam looks like:
 * plain pointers and sizes, opaque handle, int status codes, no exceptions, no C++ / torch types.
 * Every entry point cites the reference interface it replaces.  INTEGRATION.md shows the
 * reference-side stub.
 *
 * Conventions
 *   - all matrices are row-major 4x4 doubles (cv::Mat CV_64F as the reference holds them);
 *   - ids are the reference's ids (camera folder number, ArUco marker id, frame ordinal);
 *     indices are ranks of ids in ascending order (MatArray::operator=, multicam_mapper.h:95-105);
 *   - host buffers are caller-owned; one handle = one CUDA device + one stream; not re-entrant;
 *   - there is NO CPU fallback: every function that computes returns AAR_ERR_CUDA when the
 *     device path is unavailable.
 */
#ifndef AAR_CUDA_H
#define AAR_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct aar_problem aar_problem;

enum {
    AAR_OK = 0,
    AAR_ERR_INVALID = 1,      /* bad argument / inconsistent description */
    AAR_ERR_CUDA = 2,         /* CUDA runtime error (no device, launch failure, out of memory) */
    AAR_ERR_UNSUPPORTED = 3,  /* e.g. a non-pinhole camera matrix, more than 4095 cameras */
    AAR_ERR_COMM = 4,         /* NCCL error, or a peer rank that stopped answering the peer-memory reduction */
    AAR_ERR_NUMERIC = 5       /* non-finite cost or non-positive Cholesky pivot */
};

/* What MultiCamMapper::init (8 arguments, multicam_mapper.cpp:281-335) receives from the Initializer,
 * flattened.  Detections are in aruco.detections file order: frame, camera, detection order. */
typedef struct {
    int32_t num_cams, num_markers, num_frames;
    const int32_t *cam_ids, *marker_ids, *frame_ids;  /* strictly ascending */
    int32_t root_cam, root_marker;                    /* ids (Initializer::get_root_cam/_marker) */
    float marker_size;                                /* narrowed to float like the reference ctor (multicam_mapper.h:17) */
    const double *cam_T;                              /* [num_cams][16]    camera -> root camera     */
    const double *marker_T;                           /* [num_markers][16] marker -> root marker     */
    const double *frame_T;                            /* [num_frames][16]  root marker -> root camera */
    const double *cam_K;                              /* [num_cams][9] camera matrices (CamConfig::getCamMat) */
    const double *cam_dist;                           /* [num_cams][5] k1 k2 p1 p2 k3 (CamConfig::getDistCoeffs) */
    int64_t num_detections;
    const int32_t *det_frame, *det_cam, *det_marker;  /* ids */
    const float *det_xy;                              /* [num_detections][8] x0 y0 .. x3 y3, raw pixels */
    /* MultiCamMapper::Config (multicam_mapper.h:75-81) + set_with_huber (:40) */
    uint8_t optimize_cam_poses, optimize_marker_poses, optimize_object_poses, optimize_cam_intrinsics;
    uint8_t with_huber;
    uint8_t corners_undistorted;                      /* detections come from a .solution file (already undistorted, multicam_mapper.cpp:1085-1088):
                                                       * no remove_distortions pass, and the Jacobian's "raw" corners are these corners too */
    uint8_t analytic_jacobian;                        /* 0: the reference's central differences on float32-rounded projections (parity path).
                                                       * 1: analytic Jacobian, residuals kept in double (include/aar_analytic.h) — NOT the reference's
                                                       *    arithmetic; pose blocks only (with optimize_cam_intrinsics: AAR_ERR_UNSUPPORTED); track ignores it */
    uint8_t reserved[1];
    double J_delta;                                   /* multicam_mapper.h:189, 0 -> 1e-3 */
    /* placement */
    int32_t device;                                   /* CUDA device ordinal */
    void *stream;                                     /* cudaStream_t to launch on; NULL -> the handle creates its own */
    int32_t rank, world_size;                         /* frame shard of this handle (world_size <= 1: everything) */
} aar_problem_desc;

/* ucoslam::SparseLevMarq<double>::Params (sparselevmarq.h:30-50) as MultiCamMapper sets them
 * (multicam_mapper.cpp:326-330) */
typedef struct {
    int32_t max_iters;                 /* 10000 */
    double min_error;                  /* 1e-5  */
    double min_step_error_diff;        /* 0     */
    double min_average_step_error_diff;/* 1e-4  */
    double tau;                        /* 1     */
    double der_epsilon;                /* 1e-3 (track: FD step of calcDerivates) */
    int32_t ignore_stop_rules;         /* benchmark aid: run exactly max_iters iterations */
    int32_t verbose;                   /* print the reference's per-iteration line (sparselevmarq.h:421) */
} aar_lm_params;

typedef struct { double cost, mu, gain; int32_t tries, accepted; float huber_delta; int32_t pad; } aar_lm_trace;

typedef struct {
    double initial_cost, final_cost;
    int32_t iterations;        /* LM iterations executed */
    int32_t exit_code;         /* mustExit of sparselevmarq.h:453-465: 0 max iters, 1 minError, 2 small step / rejected, 3 error grew */
    int64_t total_tries;
    aar_lm_trace *trace;       /* optional caller buffer */
    int32_t trace_capacity, trace_len;
} aar_lm_report;

void aar_lm_default_params(aar_lm_params *p);

/* MultiCamMapper ctor / init (multicam_mapper.cpp:252-259, 281-335): builds the row map
 * (fill_iteration_arrays :345-377), uploads the observations packed SoA, undistorts them on the
 * device (remove_distortions :554-578). */
int aar_problem_create(const aar_problem_desc *desc, aar_problem **out);
void aar_problem_destroy(aar_problem *p);
const char *aar_last_error(void);

/* sizes: get_num_vars (multicam_mapper.cpp:239-250), num_point_xys/8, structural nnz of J */
int64_t aar_num_vars(const aar_problem *p);
int64_t aar_num_observations(const aar_problem *p);       /* global, after the erasures of :353-367 */
int64_t aar_num_local_observations(const aar_problem *p); /* this rank's shard */
int64_t aar_jacobian_nnz(const aar_problem *p);

/* Bit-exact index contract.  Any pointer may be NULL.
 * obs_*: per global observation, in row order (row0 = 8*ordinal).  col_*: first column of each
 * block in io_vec, -1 for the root / a non-optimised group. */
int aar_index_maps(const aar_problem *p, int32_t *obs_frame_idx, int32_t *obs_cam_idx, int32_t *obs_marker_idx,
                   int32_t *obs_has_jacobian, int64_t *col_cam, int64_t *col_marker, int64_t *col_frame,
                   int32_t *shard_frame_begin, int32_t *shard_frame_end);

/* undistorted (device pass) and raw corners of this rank's observations, [n_local][8] floats */
int aar_get_observations(aar_problem *p, float *und_xy, float *raw_xy);

/* mats2eVec (multicam_mapper.cpp:445-461): host side, R -> r like cv::Rodrigues(3x3 -> 3x1) */
int aar_mats2evec(const aar_problem *p, double *z);
/* eVec2Mats (multicam_mapper.cpp:595-606): device expansion read back; 4x4 per camera/marker/frame
 * (frames: this rank's shard only when sharded; others are left untouched) */
int aar_evec2mats(aar_problem *p, const double *z, double *cam_T, double *marker_T, double *frame_T);

/* MultiCamMapper::error_function (multicam_mapper.cpp:731-737): r has 8*n_local entries */
int aar_eval_residual(aar_problem *p, const double *z, float huber_delta, double *r, double *sum_sq);
/* MultiCamMapper::jacobian_function (multicam_mapper.cpp:739-801) in the compressed-column form
 * Eigen's setFromTriplets produces: colptr[num_vars+1], rowidx[nnz], vals[nnz].  Single-rank handles only. */
int aar_eval_jacobian(aar_problem *p, const double *z, int64_t *colptr, int32_t *rowidx, double *vals);
/* frame-eliminated normal equations of this rank's shard at z for damping mu (parity hook of the
 * Schur stage): S [n_r*n_r] row-major, UPPER block triangle valid, WITHOUT mu on its diagonal;
 * b [n_r]; cost = sum of squared residuals of the shard.  n_r = 6 (cameras - 1) + 6 (markers - 1) of the optimised groups; with
 * optimize_cam_intrinsics 12 more per camera, in the internal order [fx cx fy cy k1 k2 | p1 p2 k3 . . .] (csrc/aar_intrinsics.cuh). */
int aar_reduced_system(aar_problem *p, const double *z, double mu, double *S, double *b, double *cost);

/* MultiCamMapper::solve() = SparseLevMarq::solve(z, f, J) (multicam_mapper.cpp:419-428,
 * sparselevmarq.h:439-472): z is in/out like io_vec; the loop is device resident. */
int aar_lm_solve(aar_problem *p, double *z_inout, const aar_lm_params *params, aar_lm_report *report);
/* the same in three steps, z staying on the device in between (bench / step-by-step mode,
 * SparseLevMarq::init + step, sparselevmarq.h:88-96).  aar_lm_begin(p, NULL, params) restarts from the z0 of
 * the previous aar_lm_begin, which is still resident on the device (no host->device copy). */
int aar_lm_begin(aar_problem *p, const double *z0, const aar_lm_params *params);
int aar_lm_iterate(aar_problem *p, int32_t max_iters, aar_lm_report *report);
int aar_lm_end(aar_problem *p, double *z_out);

/* MultiCamMapper::track() (multicam_mapper.cpp:430-443) for every frame of the handle at once:
 * per-frame 6-dof LM against the fixed rig, Jacobian by central differences on z
 * (calcDerivates, sparselevmarq.h:164-220).  z6 [num_frames_local][6] in/out. */
int aar_track_batch(aar_problem *p, double *z6_inout, const aar_lm_params *params, double *final_cost, int32_t *iterations);
/* the same in three steps — start poses to the device / all per-frame solves from the uploaded poses (repeatable) / results back —
 * so that a caller can keep the poses resident; aar_track_ms: device time of the solves since aar_set_profiling(p, 1) */
int aar_track_upload(aar_problem *p, const double *z6);
int aar_track_run(aar_problem *p, const aar_lm_params *params);
int aar_track_download(aar_problem *p, double *z6, double *final_cost, int32_t *iterations);
int aar_track_ms(const aar_problem *p, double *ms, int64_t *runs);

/* The frame shard a handle created from `desc` (rank, world_size) would own, computed on the host without touching a
 * device: contiguous frame-index range [frame_begin, frame_end) balanced by observation count, and its observation
 * (row / 8) range.  Any output pointer may be NULL. */
int aar_shard_plan(const aar_problem_desc *desc, int32_t *frame_begin, int32_t *frame_end, int64_t *obs_begin, int64_t *obs_end, int64_t *num_observations);

/* The observation -> residual-row map of fill_iteration_arrays (multicam_mapper.cpp:345-377) computed on the host
 * without touching a device: per kept observation, in row order (row0 = 8 * ordinal), the frame / camera / marker
 * INDEX and whether it owns Jacobian rows.  Arrays may be NULL (count only); capacity = their length. */
int aar_row_map(const aar_problem_desc *desc, int64_t capacity, int32_t *obs_frame_idx, int32_t *obs_cam_idx, int32_t *obs_marker_idx,
                int32_t *obs_has_jacobian, int64_t *num_observations);

/* multi-GPU (one process per GPU of one box): one handle per rank; id is the 128-byte ncclUniqueId created on rank 0.  Per LM try the
 * reduced system and three scalars are all-reduced with NCCL.  With AAR_PEER=1 in the environment aar_comm_init also maps every rank's
 * reduced system through cudaIpc and the all-reduce is fused into its consumers instead (the reduced Cholesky sums all ranks' pieces over
 * NVLink peer memory while loading them; no collective-library call inside the try, graph-resident loop on sharded handles). */
int aar_comm_unique_id(void *id128);
int aar_comm_init(aar_problem *p, const void *id128);

/* instrumentation for bench.py: kernels launched by this handle so far, and device time (ms, CUDA events on
 * the handle's stream) accumulated per phase since aar_set_profiling(p, 1):
 *   [0] expansion + Jacobian/normal-equation assembly  [1] Schur + all-reduce + reduced solve  [2] back-substitution
 *   [3] trial residual  [4] cost all-reduce + decision  [5] k_jac_project alone  [6] number of launches in [5]
 *   [7] the assembly kernels (k_asm_pairs + k_asm_mruns) alone  [8] k_schur_syrk alone  [9] number of launches in [8]
 *   [10] k_asm_pairs alone  [11] k_asm_mruns alone */
#define AAR_NUM_PHASES 12
int64_t aar_kernel_launches(const aar_problem *p);
/* sizes of this rank's shard: [0] W slots  [1] (frame, camera) pairs  [2] (frame, marker) runs  [3] fused multiply-adds of the upper
 * triangle of the Schur update S -= E E^T per try  [4] bit 0: peer-memory reduction active (sharded handles), bit 1: graph-resident loop built, bits 8..: LM iterations run inside the graph  [5] most blocks seen by one frame  [6] [7] frame index range */
int aar_problem_stats(const aar_problem *p, int64_t *out /* [8] */);
int aar_set_profiling(aar_problem *p, int32_t on);
int aar_get_phase_ms(const aar_problem *p, double *ms /* [AAR_NUM_PHASES] */);

#ifdef __cplusplus
}
#endif
#endif /* AAR_CUDA_H */
