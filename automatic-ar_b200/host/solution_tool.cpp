// solution_tool — format round trips for the tests (no GPU needed):
//   solution_tool roundtrip <in.solution> <out.solution> [<out.yaml>]     read + write back (must be byte identical)
//   solution_tool detections <in.detections> <out.detections>             read + write back
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
        } else return -1;
    } catch (const std::exception &e) { std::cerr << e.what() << std::endl; return 2; }
    return 0;
}
