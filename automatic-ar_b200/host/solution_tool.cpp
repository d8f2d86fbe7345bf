// solution_tool — format round trips for the tests (no GPU needed):
//   solution_tool roundtrip <in.solution> <out.solution> [<out.yaml>]     read + write back (must be byte identical)
//   solution_tool detections <in.detections> <out.detections>             read + write back
//   solution_tool stereo <in.stereo_calib> <out.stereo_calib>             read + write back; prints root id and every 4x4 (incl. the inverses read_stereo_calib adds)
//   solution_tool truth <in.ground_truth> <out.ground_truth>              read + write back; prints every pose
//   solution_tool calib <data_folder> -                                   CamConfig::read_cam_configs: one line per camera, %.17g
//   solution_tool resolve_full <in.solution> <out.solution>               (GPU) the same with the default Config (camera intrinsics optimised too)
//   solution_tool resolve <in.solution> <out.solution>                    (GPU) re-create the mapper through the 8-argument
//                         init() — the Initializer-output path: raw corners, device undistortion — and solve()
#include <algorithm>
#include <cstdio>
#include <iostream>
#include "multicam_mapper.h"
int main(int argc, char **argv) {
    if (argc < 4) return -1;
    const std::string mode = argv[1];
    try {
        if (mode == "roundtrip") {
            aar::MultiCamMapper mcm;
            if (!mcm.read_solution_file(argv[2])) return 1;
            if (!mcm.write_solution_file(argv[3])) return 1;
            if (argc > 4) mcm.write_text_solution_file(argv[4]);
        } else if (mode == "detections") {
            auto d = aar::MultiCamMapper::read_detections_file(argv[2]);
            aar::MultiCamMapper::write_detections_file(argv[3], d);
            std::cout << d.size() << " frames" << std::endl;
        } else if (mode == "stereo") {
            std::map<int, std::map<int, aar::Mat44>> tr;
            const int root = aar::MultiCamMapper::read_stereo_calib(argv[2], tr);
            std::printf("root %d\n", root);
            for (auto &a : tr) for (auto &b : a.second) { std::printf("%d %d", a.first, b.first); for (int i = 0; i < 16; i++) std::printf(" %.17g", b.second.m[i]); std::printf("\n"); }
            aar::MultiCamMapper::write_stereo_calib(argv[3], tr, root);
        } else if (mode == "truth") {
            std::map<size_t, aar::Mat44> poses;
            aar::MultiCamMapper::read_ground_truth(argv[2], poses);
            for (auto &p : poses) { std::printf("%zu", p.first); for (int i = 0; i < 16; i++) std::printf(" %.17g", p.second.m[i]); std::printf("\n"); }
            aar::MultiCamMapper::write_ground_truth(argv[3], poses);
        } else if (mode == "calib") {
            const std::vector<aar::CamConfig> cc = aar::CamConfig::read_cam_configs(argv[2]);
            for (const aar::CamConfig &c : cc) {
                std::printf("%d %d", c.width, c.height);
                for (int i = 0; i < 9; i++) std::printf(" %.17g", c.K[i]);
                for (int i = 0; i < 5; i++) std::printf(" %.17g", c.dist[i]);
                std::printf("\n");
            }
        } else if (mode == "resolve" || mode == "resolve_full") {
            aar::MultiCamMapper in;
            if (!in.read_solution_file(argv[2])) return 1;
            int max_id = -1; for (auto &c : in.cam_configs) max_id = std::max(max_id, c.first);
            std::vector<aar::CamConfig> confs((size_t)max_id + 1);
            for (auto &c : in.cam_configs) confs[(size_t)c.first] = c.second;
            aar::MultiCamMapper mcm(in.get_root_cam(), in.transforms_to_root_cam, in.get_root_marker(), in.transforms_to_root_marker, in.object_to_global,
                                    in.frame_cam_markers, (float)in.get_marker_size(), confs);
            if (mode == "resolve") mcm.set_optmize_flag_cam_intrinsics(false);      // resolve_full: MultiCamMapper's default Config, intrinsics included
            mcm.solve();
            std::cout.precision(17); std::cout << "final_error: " << mcm.final_error << " iterations: " << mcm.iterations << std::endl;
            if (!mcm.write_solution_file(argv[3])) return 1;
        } else return -1;
    } catch (const std::exception &e) { std::cerr << e.what() << std::endl; return 2; }
    return 0;
}
