// benchmark.cpp -- the reference's batch driver (benchmark/benchmark.go:35-101) on the CUDA backend: same flags
// (-r -d -w -seed -width -height -save), RichScene(rand.New(seed)), RichSceneCamera, Render, PNG save.
// Extra flags: -gpus N (devices 0..N-1), -fma (fused discriminant), -ref (reference chunk streams, honours -w).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "tray.hpp"

int main(int argc, char** argv) {
    int rays = 10, depth = 20, workers = 1, width = 1200, height = 675, gpus = 1;  // benchmark.go:37-46 defaults
    uint64_t seed = 7;
    std::string save = "out.png";
    bool fma = false, ref = false;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto val = [&](const char* name) -> const char* {
            std::string n = std::string("-") + name;
            if (a == n || a == "-" + n) { if (i + 1 < argc) return argv[++i]; fprintf(stderr, "missing value for %s\n", a.c_str()); exit(2); }
            for (std::string pre : {n + "=", "-" + n + "="}) if (a.rfind(pre, 0) == 0) return argv[i] + pre.size();
            return nullptr;
        };
        const char* v;
        if ((v = val("r"))) rays = atoi(v);
        else if ((v = val("d"))) depth = atoi(v);
        else if ((v = val("w"))) workers = atoi(v);
        else if ((v = val("seed"))) seed = strtoull(v, nullptr, 10);
        else if ((v = val("width"))) width = atoi(v);
        else if ((v = val("height"))) height = atoi(v);
        else if ((v = val("save"))) save = v;
        else if ((v = val("gpus"))) gpus = atoi(v);
        else if (a == "-fma" || a == "--fma") fma = true;
        else if (a == "-ref" || a == "--ref") ref = true;
        else if (a.rfind("-progress", 0) == 0 || a.rfind("--progress", 0) == 0) {}
        else { fprintf(stderr, "unknown flag %s\n", a.c_str()); return 2; }
    }
    try {
        fortio_rand::Rand rng = fortio_rand::New(seed);
        ray::Scene scene = ray::RichScene(rng);
        fprintf(stderr, "Rendering image %dx%d with %d rays/pixel, max depth %d, %d workers, seed %llu: %zu objects\n", width, height,
                rays, depth, workers, (unsigned long long)seed, scene.Objects.size());
        auto rt = ray::New(width, height);
        rt->MaxDepth = depth; rt->NumRaysPerPixel = rays; rt->NumWorkers = workers; rt->Seed = seed;
        rt->SetCamera(ray::RichSceneCamera());
        rt->Precision = fma ? TRAY_FP64_FMA : TRAY_FP64_STRICT;
        rt->StreamMode = ref ? TRAY_STREAM_REFERENCE : TRAY_STREAM_PER_SAMPLE;
        for (int g = 0; g < gpus; g++) rt->Devices.push_back(g);
        auto t0 = std::chrono::steady_clock::now();
        rt->Render(&scene);
        double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        fprintf(stderr, "Rendered in %.3f s (kernels %.1f ms): %.1f Mpaths/s, %.1f Mrays/s on %d GPU(s)\n", s, rt->Stats.kernel_ms,
                rt->Stats.paths / rt->Stats.kernel_ms / 1e3, rt->Stats.segments / rt->Stats.kernel_ms / 1e3, rt->Stats.n_devices);
        if (!save.empty()) {
            double ms = rt->SaveImage(save);  // PNG encoded on the device (tray_encode_png)
            fprintf(stderr, "Saved rendered image to \"%s\" (PNG encoded on the GPU in %.2f ms)\n", save.c_str(), ms);
        }
    } catch (const std::exception& e) {
        fprintf(stderr, "%s\n", e.what());
        return 1;
    }
    return 0;
}
