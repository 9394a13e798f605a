// miro_cli.cpp — headless front end: render a ".miro" scene script on a B200 and write the PPM the reference's 'i' key
// would have written (src/MiroWindow.cpp:471-488 -> Image::writePPM, src/Image.cpp:132-154).  The reference has no
// headless mode (GLUT window only); SURVEY 8(f)-4.
//   miro_render scene.miro out.ppm [--assets DIR] [--device N | --devices A,B,.. [--shard-samples]] [--shard I N] [--stats]
#include "miro_host.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <chrono>
#include <vector>

int main(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "usage: %s scene.miro out.ppm [--assets DIR] [--device N] [--shard I N] [--stats]\n", argv[0]); return 2; }
    const char* assets = ".";
    int device = 0, shard_i = 0, shard_n = 1; bool stats = false, shard_samples = false;
    std::vector<int> devices;
    for (int i = 3; i < argc; ++i) {
        if (!strcmp(argv[i], "--assets") && i + 1 < argc) assets = argv[++i];
        else if (!strcmp(argv[i], "--device") && i + 1 < argc) device = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--devices") && i + 1 < argc) { for (char* t = strtok(argv[++i], ","); t; t = strtok(nullptr, ",")) devices.push_back(atoi(t)); }
        else if (!strcmp(argv[i], "--shard-samples")) shard_samples = true;
        else if (!strcmp(argv[i], "--shard") && i + 2 < argc) { shard_i = atoi(argv[++i]); shard_n = atoi(argv[++i]); }
        else if (!strcmp(argv[i], "--stats")) stats = true;
        else { fprintf(stderr, "unknown argument %s\n", argv[i]); return 2; }
    }
    miro::LoadedScene ls; std::string err;
    if (!miro::loadSceneScript(argv[1], assets, ls, err)) { fprintf(stderr, "miro_render: %s\n", err.c_str()); return 1; }
    auto t0 = std::chrono::steady_clock::now();
    if (!ls.scene->preCalc()) { fprintf(stderr, "miro_render: %s\n", ls.scene->lastError().c_str()); return 1; }
    auto t1 = std::chrono::steady_clock::now();
    ls.scene->setSampleSharding(shard_samples);
    if (!(devices.empty() ? ls.scene->attach(device) : ls.scene->attachDevices(devices.data(), (int)devices.size()))) { fprintf(stderr, "miro_render: %s\n", ls.scene->lastError().c_str()); return 1; }    // no GPU: fails here, loudly
    auto t2 = std::chrono::steady_clock::now();
    if (!ls.scene->raytraceImage(ls.camera.get(), ls.image.get(), shard_i, shard_n)) { fprintf(stderr, "miro_render: %s\n", ls.scene->lastError().c_str()); return 1; }
    auto t3 = std::chrono::steady_clock::now();
    ls.image->writePPM(argv[2]);
    if (stats) {
        miro_gpu_counters c;
        if (ls.scene->group()) miro_gpu_group_get_counters(ls.scene->group(), &c); else miro_gpu_get_counters(ls.scene->context(), &c);
        auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
        fprintf(stderr, "{\"build_ms\":%.2f,\"upload_ms\":%.2f,\"render_ms\":%.2f,\"rays\":%llu,\"Mrays_per_s\":%.1f,\"kernel_launches\":%llu}\n",
                ms(t0, t1), ms(t1, t2), ms(t2, t3), (unsigned long long)(c.rays_closest + c.rays_any),
                (c.rays_closest + c.rays_any) / ms(t2, t3) * 1e-3, (unsigned long long)c.kernel_launches);
    }
    return 0;
}
