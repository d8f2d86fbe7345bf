// track — the caller of MultiCamMapper::track() (/root/reference/apps/track.cpp:102-156) without the image pipeline:
//   track <solution_file> <aruco.detections> <initial_poses.solution> <out.solution>
// The reference detects markers per image and initialises the object pose with the Initializer; here the detections come
// from an aruco.detections file and the per-frame starting poses from a .solution file (same rig, frames to track),
// and all frames are refined in one batched call.
#include <iostream>

#include "multicam_mapper.h"

int main(int argc, char **argv) {
    if (argc < 3) { std::cout << "Usage: track <frames.solution> <out.solution>" << std::endl; return -1; }
    try {
        aar::MultiCamMapper mcm;
        if (!mcm.read_solution_file(argv[1])) return 1;         // rig + frames (initial object poses + undistorted detections)
        mcm.track();
        std::cout << "tracked " << mcm.object_to_global.size() << " frames, sum of final errors " << mcm.final_error << ", max iterations " << mcm.iterations << std::endl;
        mcm.write_solution_file(argv[2]);
    } catch (const std::exception &e) { std::cerr << "track: " << e.what() << std::endl; return 2; }
    return 0;
}
