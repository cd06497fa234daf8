// ftb-render — the CLI of FuncTracer (FuncTracer/Program.fs) with the render loop replaced by libfunctracer_b200.
//
//   ftb-render <scene-file> [<output.png>]
//
// Same contract as Program.main / getInputStream / getOutputStream (Program.fs:71-100): argv[1] = scene file,
// exactly two arguments => PNG to that file, otherwise PNG bytes on stdout; log lines on stderr with the phases
// runTracer prints (Program.fs:53-67); exit code 1 with the message on a parse error (Program.fs:13-16).
// Parsing / PLY / BSP build / PNG encode are the host front end (libftb_frontend.so: the C++ stand-in for
// SceneParser.fs, PlyParser.fs, BspMesh.compile and ImageSharp, none of which can run in this image); the render
// loop (Program.fs:54-64) is ftb_scene_create + ftb_render.  Extra, optional environment: FTB_ASSETS (directory
// searched for mesh / texture files), FTB_SEED (RNG seed for soft shadows / depth of field), FTB_GPUS.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "../../../include/functracer_b200.h"
#include "../frontend/ftb_frontend.h"

int main(int argc, char** argv)
{
    const auto t0 = std::chrono::steady_clock::now();
    auto ms = [&]() { return (long long)std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::steady_clock::now() - t0).count(); };
    if (argc < 2) { std::fprintf(stderr, "usage: ftb-render <scene-file> [<output.png>]\n"); return 2; }
    std::fprintf(stderr, "Using input file: %s\n", argv[1]);
    std::ifstream in(argv[1]);
    if (!in) { std::fprintf(stderr, "cannot open %s\n", argv[1]); return 1; }
    std::stringstream ss;
    ss << in.rdbuf();
    const std::string text = ss.str();
    std::string assets;
    if (const char* a = std::getenv("FTB_ASSETS")) assets = a;
    else { assets = argv[1]; size_t k = assets.find_last_of('/'); assets = k == std::string::npos ? "." : assets.substr(0, k); }
    ftbf_scene* fs = nullptr;
    if (ftbf_parse(text.c_str(), assets.c_str(), &fs) != 0) {  // readScene (Program.fs:10-16)
        std::printf("%s\n", ftbf_last_error());
        return 1;
    }
    std::fprintf(stderr, "Parsed input %llims\n", ms());
    int W = 0, H = 0, spp = 0, sampling = 0;
    ftbf_options(fs, &W, &H, &spp, &sampling);
    std::vector<double> jitter(2 * (size_t)(spp > 0 ? spp : 1));
    const uint64_t seed = std::getenv("FTB_SEED") ? std::strtoull(std::getenv("FTB_SEED"), nullptr, 10) : (uint64_t)std::chrono::system_clock::now().time_since_epoch().count();
    ftbf_jitter_pattern(seed, spp > 0 ? spp : 1, jitter.data());  // Jitter.pattern random Jitter.circle spp (Image.fs:101-105)
    std::fprintf(stderr, "Generated rays: %llims\n", ms());
    ftb_scene* scene = nullptr;
    if (ftb_scene_create(ftbf_desc(fs), &scene) != 0) { std::fprintf(stderr, "%s\n", ftb_last_error()); return 1; }
    std::fprintf(stderr, "Geometry created\n");
    ftb_render_params p = {};
    p.width = W; p.height = H; p.spp = spp; p.sampling = sampling; p.jitter_xy = jitter.data();
    p.recursion_limit = 8;  // Shading.fs:142
    p.precision = FTB_PRECISION_FP32; p.seed = seed; p.out_format = FTB_OUT_RGBA8;  // Image.write's toByte on device
    p.n_gpus = std::getenv("FTB_GPUS") ? std::atoi(std::getenv("FTB_GPUS")) : 0;
    std::vector<uint8_t> rgba(4 * (size_t)W * H);
    if (ftb_render(scene, ftbf_camera(fs), &p, rgba.data(), nullptr, nullptr) != 0) { std::fprintf(stderr, "%s\n", ftb_last_error()); return 1; }
    std::fprintf(stderr, "Shaded scene %llims\n", ms());
    std::fprintf(stderr, "Writing output %llims\n", ms());
    int rc;
    if (argc == 3) {
        std::fprintf(stderr, "Using output file: %s\n", argv[2]);
        rc = ftbf_write_png(argv[2], W, H, rgba.data());
    } else {
        std::fprintf(stderr, "Using standard output\n");
        rc = ftbf_write_png("/dev/stdout", W, H, rgba.data());
    }
    if (rc != 0) { std::fprintf(stderr, "%s\n", ftbf_last_error()); return 1; }
    ftb_scene_destroy(scene);
    ftbf_destroy(fs);
    std::fprintf(stderr, "Elapsed Time: %llims\n", ms());
    return 0;
}
