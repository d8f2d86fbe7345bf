// initializer.h — host-side mirror of the reference's Initializer (/root/reference/libs/initializer.h:9-73): same method names
// and argument meaning, cv::Mat replaced by aar::Mat44, the arithmetic forwarded to the CUDA path through the C ABI of
// include/aar_init.h (IPPE per detection, the consensus of find_best_transformation; spanning tree on the host inside the library).
#pragma once
#include <map>
#include <set>
#include <string>
#include <vector>

#include "../../include/aar_init.h"
#include "aar_host_types.h"

namespace aar {

class Initializer {
public:
    typedef std::vector<std::vector<std::vector<Marker>>> Detections;      // [frame][camera][detection]
    // initializer.cpp:58-72
    Initializer(double marker_s, const std::vector<CamConfig> &cam_c, const std::set<int> &excluded_cs = std::set<int>());
    Initializer(const Detections &dts, double marker_s, const std::vector<CamConfig> &cam_c, const std::set<int> &excluded_cs = std::set<int>());
    ~Initializer();
    Initializer(const Initializer &) = delete;
    Initializer &operator=(const Initializer &) = delete;

    static Detections read_detections_file(const std::string &path, const std::vector<int> &subseqs = std::vector<int>());   // initializer.cpp:316-362

    std::set<int> get_marker_ids() const { return marker_ids; }
    std::set<int> get_cam_ids() const { return cam_ids; }
    int get_root_cam() const { return root_cam; }
    int get_root_marker() const { return root_marker; }
    std::map<int, Mat44> get_transforms_to_root_cam() const { return transforms_to_root_cam; }
    std::map<int, Mat44> get_transforms_to_root_marker() const { return transforms_to_root_marker; }
    std::map<int, Mat44> get_object_transforms() const { return object_transforms; }
    FrameCamMarkers get_frame_cam_markers() const { return frame_cam_markers; }
    std::vector<CamConfig> get_cam_configs() const { return cam_configs; }
    double get_marker_size() const { return marker_size; }
    void set_transforms_to_root_cam(const std::map<int, Mat44> &t) { transforms_to_root_cam = t; rig_dirty = true; }
    void set_transforms_to_root_marker(const std::map<int, Mat44> &t) { transforms_to_root_marker = t; rig_dirty = true; }
    void set_detections(const Detections &dts) { detections = dts; drop_handle(); }
    void obtain_pose_estimations();      // initializer.cpp:364-419 (device: one IPPE solve per detection)
    void init_object_transforms();       // initializer.cpp:451-463 (device: one consensus per frame)
    void init_transforms();              // initializer.cpp:465-469: cameras, markers, then objects

    double threshold = 2.0;              // initializer.h:55 (find_solution -thresh)
    int consensus_max = 0;               // 0 = the reference's exhaustive consensus (include/aar_init.h)
    int device = 0;

private:
    void drop_handle();
    void check(int rc, const char *what) const;
    void pull_rig();
    Detections detections;
    std::vector<CamConfig> cam_configs;
    std::set<int> excluded_cams, marker_ids, cam_ids;
    std::map<int, Mat44> transforms_to_root_cam, transforms_to_root_marker, object_transforms;
    FrameCamMarkers frame_cam_markers;
    int root_cam = -1, root_marker = -1;
    double marker_size = 0;
    bool rig_dirty = false;
    aar_init *handle = nullptr;
};

} // namespace aar
