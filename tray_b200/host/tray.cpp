// tray.cpp -- the reference's interactive driver (main.go:38-169) in its non-interactive form (`-exit`, or stdout not a
// terminal: render once, show, exit) on the CUDA backend: same flags (-s -r -d -w -seed -save -exit), RichScene(rand.New(seed)),
// RichSceneCamera, image size = round(s * W) x round(s * 2H) for a W x H terminal (main.go:92), SaveImage, downscale
// (BiLinear for s > 1, NearestNeighbor for s < 1) and the half-block frame -- the last three on the device.
// The key loop / ansipixels raw mode stays out of scope; -cols/-rows replace the terminal size query.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "tray.hpp"

int main(int argc, char** argv) {
    double supersample = 4;                       // main.go:39-47 defaults
    int rays = 64, depth = 12, workers = 0, cols = 80, rows = 24;
    uint64_t seed = 0;
    std::string save;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto val = [&](const char* name) -> const char* {
            std::string n = std::string("-") + name;
            if (a == n || a == "-" + n) { if (i + 1 < argc) return argv[++i]; fprintf(stderr, "missing value for %s\n", a.c_str()); exit(2); }
            for (std::string pre : {n + "=", "-" + n + "="}) if (a.rfind(pre, 0) == 0) return argv[i] + pre.size();
            return nullptr;
        };
        const char* v;
        if ((v = val("s"))) supersample = atof(v);
        else if ((v = val("r"))) rays = atoi(v);
        else if ((v = val("d"))) depth = atoi(v);
        else if ((v = val("w"))) workers = atoi(v);
        else if ((v = val("seed"))) seed = strtoull(v, nullptr, 10);
        else if ((v = val("save"))) save = v;
        else if ((v = val("cols"))) cols = atoi(v);
        else if ((v = val("rows"))) rows = atoi(v);
        else if (a == "-exit" || a == "--exit") {}
        else { fprintf(stderr, "unknown flag %s\n", a.c_str()); return 2; }
    }
    if (supersample <= 0) supersample = 1;       // main.go:60-63
    if (const char* c = getenv("COLUMNS")) if (cols == 80 && atoi(c) > 0) cols = atoi(c);
    if (const char* l = getenv("LINES")) if (rows == 24 && atoi(l) > 0) rows = atoi(l);
    try {
        fortio_rand::Rand rng = fortio_rand::New(seed);
        ray::Scene scene = ray::RichScene(rng);   // main.go:87-88: built once, outside OnResize
        const int w = (int)std::lround(supersample * cols), h = (int)std::lround(supersample * rows * 2);  // main.go:92
        auto rt = ray::New(w, h);
        rt->Seed = seed; rt->MaxDepth = depth; rt->NumRaysPerPixel = rays; rt->NumWorkers = workers;
        rt->SetCamera(ray::RichSceneCamera());
        auto t0 = std::chrono::steady_clock::now();
        rt->Render(&scene);
        if (!save.empty()) {
            rt->SaveImage(save);
            fprintf(stderr, "Saved rendered image to \"%s\"\n", save.c_str());
        }
        double present_ms = 0;
        std::string frame = rt->Present(cols, rows, &present_ms);
        double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        fwrite(frame.data(), 1, frame.size(), stdout);
        fprintf(stderr, "%d x %d image (%.1fx) Rays %d, Depth %d: %.1f ms (kernels %.1f ms, downscale + ANSI frame %.3f ms on the GPU)\n", w, h,
                supersample, rays, depth, ms, rt->Stats.kernel_ms, present_ms);
    } catch (const std::exception& e) {
        fprintf(stderr, "%s\n", e.what());
        return 1;
    }
    return 0;
}
