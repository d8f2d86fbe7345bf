// aar_host_types.h — OpenCV-free stand-ins for the types the reference's public interface uses
// (cv::Mat 4x4 CV_64F, aruco::Marker, CamConfig of /root/reference/libs/cam_config.h).
#pragma once
#include <array>
#include <cstring>
#include <map>
#include <string>
#include <vector>

namespace aar {

// 4x4 row-major double matrix (what the reference keeps as cv::Mat(4,4,CV_64FC1))
struct Mat44 {
    double m[16];
    Mat44() { std::memset(m, 0, sizeof m); m[0] = m[5] = m[10] = m[15] = 1.0; }
    static Mat44 eye() { return Mat44(); }
    double &at(int r, int c) { return m[r * 4 + c]; }
    double at(int r, int c) const { return m[r * 4 + c]; }
};

// aruco::Marker as the path uses it: id + 4 corners (3rdparty/aruco/aruco/marker.h:47-59); 36 bytes on disk
struct Marker {
    int id = -1;
    float xy[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // x0 y0 x1 y1 x2 y2 x3 y3
};

// frame id -> camera id -> detections in detection order (the reference's frame_cam_markers)
typedef std::map<int, std::map<int, std::vector<Marker>>> FrameCamMarkers;

// libs/cam_config.h: camera matrix, distortion coefficients (k1 k2 p1 p2 k3), image size
class CamConfig {
public:
    CamConfig() { std::memset(K, 0, sizeof K); std::memset(dist, 0, sizeof dist); K[0] = K[4] = K[8] = 1; }
    bool read_from_file(const std::string &path);                       // OpenCV FileStorage YAML / XML is not parsed: YAML only
    static std::vector<CamConfig> read_cam_configs(const std::string &folder_path);   // <folder>/<cam index>/calib.{yml,yaml}, numeric order
    const double *getCamMat() const { return K; }
    const double *getDistCoeffs() const { return dist; }
    void setCamMat(const double *k9) { std::memcpy(K, k9, sizeof K); }
    void setDistCoeffs(const double *d5) { std::memcpy(dist, d5, sizeof dist); }
    int width = 0, height = 0;
    double K[9], dist[5];
};

} // namespace aar
