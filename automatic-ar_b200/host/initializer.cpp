// initializer.cpp — see initializer.h.  Reference line numbers are into /root/reference/libs/initializer.cpp.
#include "initializer.h"

#include <cstring>
#include <stdexcept>

#include "../../include/aar_cuda.h"
#include "multicam_mapper.h"

namespace aar {

Initializer::Initializer(double marker_s, const std::vector<CamConfig> &cam_c, const std::set<int> &excluded_cs)
    : cam_configs(cam_c), excluded_cams(excluded_cs), marker_size(marker_s) {}

Initializer::Initializer(const Detections &dts, double marker_s, const std::vector<CamConfig> &cam_c, const std::set<int> &excluded_cs)
    : detections(dts), cam_configs(cam_c), excluded_cams(excluded_cs), marker_size(marker_s) {
    obtain_pose_estimations();
    init_transforms();
}

Initializer::~Initializer() { drop_handle(); }
void Initializer::drop_handle() { if (handle) { aar_init_destroy(handle); handle = nullptr; } }
void Initializer::check(int rc, const char *what) const { if (rc != AAR_OK) throw std::runtime_error(std::string(what) + ": " + aar_last_error()); }

Initializer::Detections Initializer::read_detections_file(const std::string &path, const std::vector<int> &subseqs) {
    return MultiCamMapper::read_detections_file(path, subseqs);
}

void Initializer::obtain_pose_estimations() {
    drop_handle();
    frame_cam_markers.clear(); cam_ids.clear(); marker_ids.clear();
    const int num_cams = (int)cam_configs.size();
    std::vector<double> K((size_t)num_cams * 9), dist((size_t)num_cams * 5);
    for (int c = 0; c < num_cams; c++) { std::memcpy(&K[9 * c], cam_configs[c].getCamMat(), 72); std::memcpy(&dist[5 * c], cam_configs[c].getDistCoeffs(), 40); }
    std::vector<int32_t> df, dc, dm; std::vector<float> xy;
    for (size_t f = 0; f < detections.size(); f++) {
        if ((int)detections[f].size() > num_cams) throw std::runtime_error("Initializer: more cameras in the detections than camera configurations");
        for (size_t c = 0; c < detections[f].size(); c++)
            for (const Marker &m : detections[f][c]) { df.push_back((int)f); dc.push_back((int)c); dm.push_back(m.id); xy.insert(xy.end(), m.xy, m.xy + 8); }
    }
    std::vector<uint8_t> ex(num_cams, 0);
    for (int c : excluded_cams) if (c >= 0 && c < num_cams) ex[c] = 1;
    aar_init_desc d; std::memset(&d, 0, sizeof d);
    d.num_cams = num_cams; d.cam_K = K.data(); d.cam_dist = dist.data(); d.marker_size = marker_size; d.num_frames = (int)detections.size();
    d.num_detections = (int64_t)df.size(); d.det_frame = df.data(); d.det_cam = dc.data(); d.det_marker = dm.data(); d.det_xy = xy.data();
    d.excluded_cams = ex.data(); d.threshold = threshold; d.min_detections = 2; d.consensus_max = consensus_max; d.device = device;
    check(aar_init_create(&d, &handle), "aar_init_create");
    std::vector<uint8_t> ncand(df.size());
    check(aar_init_get_estimations(handle, nullptr, nullptr, ncand.data()), "aar_init_get_estimations");
    // frame_cam_markers: the detections of the frames that were kept, excluded cameras dropped (:385-396)
    size_t i = 0;
    for (size_t f = 0; f < detections.size(); f++)
        for (size_t c = 0; c < detections[f].size(); c++)
            for (const Marker &m : detections[f][c]) {
                if (ncand[i++]) { frame_cam_markers[(int)f][(int)c].push_back(m); cam_ids.insert((int)c); marker_ids.insert(m.id); }
            }
    rig_dirty = true;
}

void Initializer::pull_rig() {
    int32_t c[7];
    check(aar_init_counts(handle, c), "aar_init_counts");
    std::vector<int32_t> ci(c[2]), mi(c[3]); std::vector<double> cT((size_t)c[2] * 16), mT((size_t)c[3] * 16);
    check(aar_init_get_rig(handle, ci.data(), cT.data(), mi.data(), mT.data()), "aar_init_get_rig");
    transforms_to_root_cam.clear(); transforms_to_root_marker.clear();
    for (int i = 0; i < c[2]; i++) std::memcpy(transforms_to_root_cam[ci[i]].m, &cT[16 * (size_t)i], 128);
    for (int i = 0; i < c[3]; i++) std::memcpy(transforms_to_root_marker[mi[i]].m, &mT[16 * (size_t)i], 128);
    root_cam = c[5]; root_marker = c[6];
    rig_dirty = false;
}

void Initializer::init_object_transforms() {
    if (!handle) throw std::runtime_error("Initializer::init_object_transforms before obtain_pose_estimations");
    if (rig_dirty) {
        std::vector<int32_t> ci, mi; std::vector<double> cT, mT;
        for (auto &kv : transforms_to_root_cam) { ci.push_back(kv.first); cT.insert(cT.end(), kv.second.m, kv.second.m + 16); }
        for (auto &kv : transforms_to_root_marker) { mi.push_back(kv.first); mT.insert(mT.end(), kv.second.m, kv.second.m + 16); }
        check(aar_init_set_rig(handle, (int)ci.size(), ci.data(), cT.data(), (int)mi.size(), mi.data(), mT.data()), "aar_init_set_rig");
        rig_dirty = false;
    }
    check(aar_init_object_transforms(handle), "aar_init_object_transforms");
    int32_t c[7];
    check(aar_init_counts(handle, c), "aar_init_counts");
    std::vector<int32_t> fi(c[4]); std::vector<double> fT((size_t)c[4] * 16);
    check(aar_init_get_object_transforms(handle, fi.data(), fT.data()), "aar_init_get_object_transforms");
    object_transforms.clear();
    for (int i = 0; i < c[4]; i++) std::memcpy(object_transforms[fi[i]].m, &fT[16 * (size_t)i], 128);
}

void Initializer::init_transforms() {
    if (!handle) throw std::runtime_error("Initializer::init_transforms before obtain_pose_estimations");
    check(aar_init_transforms(handle), "aar_init_transforms");
    pull_rig();
    init_object_transforms();
}

} // namespace aar
