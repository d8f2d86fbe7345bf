/*
 * aar_init.h — C ABI of the B200 (sm_100a) initialisation path of automatic-ar: what the
 * reference runs between `aruco.detections` + `calib.yml` and the MultiCamMapper constructor.
 *
 * The reference's seam is the C++ class `Initializer` (/root/reference/libs/initializer.h:9-73) as used by
 * apps/find_solution.cpp:122 (`Initializer initializer(detections, marker_size, cam_configs, excluded_cams)` followed by
 * `MultiCamMapper mcm(initializer)`) and apps/track.cpp:128-131 (`set_detections`, `obtain_pose_estimations`,
 * `init_object_transforms` per frame).  Each entry point cites the method it replaces; INTEGRATION.md shows the
 * reference-side stub.  Same conventions as aar_cuda.h: plain pointers and sizes, opaque handle, int status
 * (AAR_OK ... of aar_cuda.h), row-major 4x4 doubles, caller-owned host buffers, NO CPU fallback.
 *
 *   device:  one IPPE solve per detection (aruco::solvePnP_, 3rdparty/aruco/aruco/ippe.cpp:118-219), the candidate
 *            transformation triples (initializer.cpp:74-146) and the O(n^2) consensus of find_best_transformation
 *            (initializer.cpp:156-205) for every camera pair, marker pair and frame;
 *   host:    the integer bookkeeping (which detections pair up, in the reference's list order), the spanning tree over
 *            <= cameras / markers nodes and the chaining of the chosen transforms (initializer.cpp:237-314).
 */
#ifndef AAR_INIT_H
#define AAR_INIT_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct aar_init aar_init;

/* Initializer(detections, marker_size, cam_configs, excluded_cams) (initializer.cpp:64-72), detections flattened in
 * aruco.detections file order: frame ordinal, camera index, detection order. */
typedef struct {
    int32_t num_cams;                 /* cameras of the detections file; camera id = index into cam_K / cam_dist (initializer.cpp:401) */
    const double *cam_K;              /* [num_cams][9] */
    const double *cam_dist;           /* [num_cams][5] k1 k2 p1 p2 k3 */
    double marker_size;               /* metres; narrowed to float for the IPPE solve like the reference's call (ippe.cpp:118) */
    int32_t num_frames;               /* frame groups in the file */
    int64_t num_detections;
    const int32_t *det_frame, *det_cam, *det_marker;
    const float *det_xy;              /* [num_detections][8] raw pixels */
    const uint8_t *excluded_cams;     /* optional [num_cams], non-zero = excluded (find_solution -exclude-cams) */
    double threshold;                 /* second IPPE solution kept when err1 / err0 < threshold; 0 -> 2.0 (initializer.h:55) */
    int32_t min_detections;           /* frames with fewer detections are skipped; 0 -> 2 (initializer.h:53) */
    int32_t consensus_max;            /* 0 = the reference's exhaustive consensus; k > 0 = lists longer than k are reduced to the k
                                       * candidates at positions floor(i * n / k) (the O(n^2) consensus does not scale) */
    int32_t device;
    void *stream;                     /* cudaStream_t or NULL */
} aar_init_desc;

/* Initializer ctor + obtain_pose_estimations (initializer.cpp:364-419): uploads the detections and runs the IPPE kernel */
int aar_init_create(const aar_init_desc *desc, aar_init **out);
void aar_init_destroy(aar_init *h);

/* parity hook: per flat detection both IPPE poses [n][2][16] (float32-valued doubles, ippe.cpp:124 + initializer.cpp:403),
 * their reprojection errors [n][2] and how many the Initializer keeps (0 = frame skipped / camera excluded, 1, 2) */
int aar_init_get_estimations(aar_init *h, double *T, double *err, uint8_t *ncand);

/* init_transforms_cam + init_transforms_marker (initializer.cpp:422-449) */
int aar_init_transforms(aar_init *h);
/* set_transforms_to_root_cam / _marker (initializer.cpp:14-20): the track app's flow (rig from a solved MultiCamMapper) */
int aar_init_set_rig(aar_init *h, int32_t num_cams, const int32_t *cam_ids, const double *cam_T, int32_t num_markers, const int32_t *marker_ids, const double *marker_T);
/* init_object_transforms (initializer.cpp:451-463): one consensus per kept frame */
int aar_init_object_transforms(aar_init *h);

/* get_cam_ids / get_marker_ids / get_root_cam / get_root_marker / get_transforms_to_root_cam / _marker / get_object_transforms:
 * counts[0..6] = #cam ids, #marker ids, #cameras with a transform, #markers with a transform, #frames with an object transform,
 * root camera id, root marker id (-1 before aar_init_transforms / aar_init_set_rig) */
int aar_init_counts(const aar_init *h, int32_t *counts /* [7] */);
int aar_init_get_ids(const aar_init *h, int32_t *cam_ids, int32_t *marker_ids);
int aar_init_get_rig(const aar_init *h, int32_t *cam_ids, double *cam_T, int32_t *marker_ids, double *marker_T);
int aar_init_get_object_transforms(const aar_init *h, int32_t *frame_ids, double *T);
/* the consensus graph of find_best_transformations (initializer.cpp:207-235): per edge id1 < id2, the length of its candidate
 * list and the consensus error of the chosen transformation; returns the number of edges through *num_edges */
int aar_init_edges(const aar_init *h, int32_t cams, int32_t capacity, int32_t *id1, int32_t *id2, int64_t *list_len, double *weight, int32_t *num_edges);

/* find_best_transformation (initializer.cpp:156-205) on caller-provided candidates, [n][16] each: index of the winner and its error */
int aar_init_consensus(int32_t device, double marker_size, int64_t n, const double *T, const double *T1inv, const double *T2inv, int32_t *best, double *weight);

/* device time (ms) of the last aar_init_create (IPPE), aar_init_transforms and aar_init_object_transforms calls, kernels launched */
int aar_init_timings(const aar_init *h, double *ms /* [3] */, int64_t *launches);

#ifdef __cplusplus
}
#endif
#endif /* AAR_INIT_H */
