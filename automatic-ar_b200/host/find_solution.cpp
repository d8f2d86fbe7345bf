// find_solution — the reference's app (/root/reference/apps/find_solution.cpp:28-181) on the CUDA path, OpenCV-free:
//   find_solution <path_to_data_folder> <marker_size> [-subseqs] [-exclude-cams <cam_id> ...] [-with-huber] [-thresh <t>]
//                 [-consensus-max <k>] [-init <initial.solution>] [-analytic]
// Reads <folder>/<cam>/calib.yml and <folder>/aruco.detections, builds the starting point with the Initializer (device: IPPE per
// detection + consensus; include/aar_init.h), writes initial<suffix>.solution(.yaml), solves on the GPU (include/aar_cuda.h) and
// writes final<suffix>.solution(.yaml) — the file names of find_solution.cpp:76-100.  Two additions: -consensus-max (SURVEY 8(f) row 3,
// the reference's exhaustive consensus is O(n^2) in the number of co-observations) and -init (start from an existing .solution
// instead of the detections) and -analytic (analytic Jacobian, residuals in double: include/aar_analytic.h; not the reference's arithmetic).  Unlike the reference (whose option loop starts at argv[4]) options are read from argv[3] on.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <iostream>
#include <set>

#include "initializer.h"
#include "multicam_mapper.h"

int main(int argc, char **argv) {
    if (argc < 3) {
        std::cout << "Usage: find_solution <path_to_data_folder> <marker_size> [-subseqs] [-exclude-cams <cam_id> ...] [-with-huber] [-thresh <t>] [-consensus-max <k>] [-init <file>] [-analytic]" << std::endl;
        return -1;
    }
    const std::string folder = argv[1];
    const double marker_size = std::stod(argv[2]);
    bool use_subseqs = false, with_huber = false, set_threshold = false, analytic = false; double threshold = 2.0; int consensus_max = 0;
    std::set<int> excluded_cams; std::string init_path;
    enum { NONE, EXCLUDE, THRESH, CMAX, INIT } flag = NONE;
    for (int i = 3; i < argc; i++) {
        const std::string a = argv[i];
        if (a == "-subseqs") { use_subseqs = true; flag = NONE; }
        else if (a == "-exclude-cams") flag = EXCLUDE;
        else if (a == "-with-huber") { with_huber = true; flag = NONE; }
        else if (a == "-analytic") { analytic = true; flag = NONE; }
        else if (a == "-thresh") { set_threshold = true; flag = THRESH; }
        else if (a == "-consensus-max") flag = CMAX;
        else if (a == "-init") flag = INIT;
        else if (flag == EXCLUDE) excluded_cams.insert(std::stoi(a));
        else if (flag == THRESH) { threshold = std::stod(a); flag = NONE; }
        else if (flag == CMAX) { consensus_max = std::stoi(a); flag = NONE; }
        else if (flag == INIT) { init_path = a; flag = NONE; }
    }
    std::string suffix;                                         // find_solution.cpp:76-100
    if (use_subseqs) suffix += "_subseqs";
    if (with_huber) suffix += "_with_huber";
    if (!excluded_cams.empty()) { suffix += "_excluded_cams"; for (int c : excluded_cams) suffix += "_" + std::to_string(c); }
    if (set_threshold) { char d[16]; std::snprintf(d, sizeof d, "%.1f", threshold); suffix += "_thresh_" + std::string(d); }
    suffix += ".solution";
    const std::string initial_path = folder + "/initial" + suffix, final_path = folder + "/final" + suffix;
    try {
        std::chrono::duration<double> d(0);
        aar::MultiCamMapper mcm;
        if (!init_path.empty()) {
            if (!mcm.read_solution_file(init_path)) return 1;
            if (std::fabs(marker_size - mcm.get_marker_size()) > 1e-6) std::cout << "note: marker size of the solution file is " << mcm.get_marker_size() << std::endl;
        } else {
            std::vector<aar::CamConfig> cam_configs = aar::CamConfig::read_cam_configs(folder);
            if (cam_configs.empty()) { std::cerr << "find_solution: no <cam>/calib.yml under " << folder << std::endl; return 1; }
            std::vector<int> subseqs;
            if (use_subseqs) subseqs = aar::MultiCamMapper::read_subseqs(folder + "/subseqs.txt");
            aar::Initializer::Detections detections = aar::Initializer::read_detections_file(folder + "/aruco.detections", subseqs);
            auto start = std::chrono::system_clock::now();
            aar::Initializer initializer(marker_size, cam_configs, excluded_cams);
            initializer.threshold = threshold; initializer.consensus_max = consensus_max;
            initializer.set_detections(detections);
            initializer.obtain_pose_estimations();
            initializer.init_transforms();
            mcm.init((size_t)initializer.get_root_cam(), initializer.get_transforms_to_root_cam(), (size_t)initializer.get_root_marker(), initializer.get_transforms_to_root_marker(),
                     initializer.get_object_transforms(), initializer.get_frame_cam_markers(), (float)initializer.get_marker_size(), initializer.get_cam_configs());
            d += std::chrono::system_clock::now() - start;
            std::cout << "initialised " << initializer.get_transforms_to_root_cam().size() << " cameras, " << initializer.get_transforms_to_root_marker().size()
                      << " markers, " << initializer.get_object_transforms().size() << " frames" << std::endl;
            mcm.write_solution_file(initial_path);
            mcm.write_text_solution_file(initial_path + ".yaml");
        }
        mcm.set_optmize_flag_cam_poses(true); mcm.set_optmize_flag_marker_poses(true); mcm.set_optmize_flag_object_poses(true);
        mcm.set_optmize_flag_cam_intrinsics(false);              // find_solution.cpp:140
        if (with_huber) mcm.set_with_huber(true);
        if (analytic) mcm.set_analytic_jacobian(true);
        auto start = std::chrono::system_clock::now();
        mcm.solve();
        d += std::chrono::system_clock::now() - start;
        std::cout.precision(17); std::cout << "final_error: " << mcm.final_error << " iterations: " << mcm.iterations << std::endl;
        mcm.write_solution_file(final_path);
        mcm.write_text_solution_file(final_path + ".yaml");
        const int minutes = (int)(d.count() / 60); const long seconds = std::lround(d.count() - minutes * 60);
        std::cout << "The algorithm took: " << minutes << " minutes " << seconds << " seconds" << std::endl;
    } catch (const std::exception &e) { std::cerr << "find_solution: " << e.what() << std::endl; return 2; }
    return 0;
}
