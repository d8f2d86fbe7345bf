// find_solution — the caller of MultiCamMapper::solve() (/root/reference/apps/find_solution.cpp:138-177) on the CUDA path.
//   find_solution <folder> <marker_size> [-with-huber] [-init <initial.solution>]
// The reference builds its starting point with the Initializer (out of scope: O(n^2) host pose-graph code, SURVEY §2
// row 5) and writes it as <folder>/initial.solution before solving; this app starts from that file — the
// Initializer's output in the reference's own format — solves on the GPU and writes final.solution + .yaml.
#include <chrono>
#include <cmath>
#include <iostream>

#include "multicam_mapper.h"

int main(int argc, char **argv) {
    if (argc < 3) { std::cout << "Usage: find_solution <path_to_data_folder> <marker_size> [-with-huber] [-init <initial.solution>]" << std::endl; return -1; }
    const std::string folder = argv[1];
    bool with_huber = false; std::string init_path = folder + "/initial.solution";
    for (int i = 3; i < argc; i++) {
        const std::string a = argv[i];
        if (a == "-with-huber") with_huber = true;
        else if (a == "-init" && i + 1 < argc) init_path = argv[++i];
    }
    try {
        aar::MultiCamMapper mcm;
        if (!mcm.read_solution_file(init_path)) return 1;
        const double marker_size = std::atof(argv[2]);
        if (std::fabs(marker_size - mcm.get_marker_size()) > 1e-6) std::cout << "note: marker size of the solution file is " << mcm.get_marker_size() << std::endl;
        mcm.set_optmize_flag_cam_poses(true); mcm.set_optmize_flag_marker_poses(true); mcm.set_optmize_flag_object_poses(true);
        mcm.set_optmize_flag_cam_intrinsics(false);              // find_solution.cpp:140
        if (with_huber) mcm.set_with_huber(true);
        auto start = std::chrono::system_clock::now();
        mcm.solve();
        std::chrono::duration<double> d = std::chrono::system_clock::now() - start;
        std::cout << "final_error: " << mcm.final_error << " iterations: " << mcm.iterations << std::endl;
        mcm.write_solution_file(folder + "/final.solution");
        mcm.write_text_solution_file(folder + "/final.solution.yaml");
        const int minutes = (int)(d.count() / 60); const long seconds = std::lround(d.count() - minutes * 60);
        std::cout << "The algorithm took: " << minutes << " minutes " << seconds << " seconds" << std::endl;
    } catch (const std::exception &e) { std::cerr << "find_solution: " << e.what() << std::endl; return 2; }
    return 0;
}
