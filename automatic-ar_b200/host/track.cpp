// track — the caller of MultiCamMapper::track() (/root/reference/apps/track.cpp:60-156) without the image pipeline:
//   track <rig.solution> <aruco.detections> <out.solution>     detections of new frames against a solved rig: per frame
//                                                              Initializer::obtain_pose_estimations + init_object_transforms
//                                                              (track.cpp:128-131) and MultiCamMapper::track(), all frames batched
//   track <frames.solution> <out.solution>                     starting poses and detections taken from a .solution file
// The reference detects the markers in the images of every frame (aruco, out of scope) and refines one frame at a time.
#include <iostream>

#include "initializer.h"
#include "multicam_mapper.h"

int main(int argc, char **argv) {
    if (argc < 3) { std::cout << "Usage: track <rig.solution> <aruco.detections> <out.solution> | track <frames.solution> <out.solution>" << std::endl; return -1; }
    try {
        aar::MultiCamMapper mcm;
        if (!mcm.read_solution_file(argv[1])) return 1;
        if (argc >= 4) {
            std::vector<aar::CamConfig> cam_configs;                   // camera id = index (cam_configs[cam_id], initializer.cpp:401)
            for (auto &kv : mcm.cam_configs) { if ((size_t)kv.first >= cam_configs.size()) cam_configs.resize((size_t)kv.first + 1); cam_configs[(size_t)kv.first] = kv.second; }
            aar::Initializer initializer(mcm.get_marker_size(), cam_configs);
            initializer.set_transforms_to_root_cam(mcm.transforms_to_root_cam);          // track.cpp:85-86
            initializer.set_transforms_to_root_marker(mcm.transforms_to_root_marker);
            initializer.set_detections(aar::Initializer::read_detections_file(argv[2]));
            initializer.obtain_pose_estimations();
            initializer.init_object_transforms();
            mcm.init(initializer.get_object_transforms(), initializer.get_frame_cam_markers());   // track.cpp:132
        }
        mcm.track();
        std::cout << "tracked " << mcm.object_to_global.size() << " frames, sum of final errors " << mcm.final_error << ", max iterations " << mcm.iterations << std::endl;
        mcm.write_solution_file(argv[argc >= 4 ? 3 : 2]);
    } catch (const std::exception &e) { std::cerr << "track: " << e.what() << std::endl; return 2; }
    return 0;
}
