// TEST-ONLY: prints the canonical dump of a host-built scene (compare with `mrt_ref dump-scene`).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "scene_graph.h"
int main(int argc, char **argv) {
    uint32_t scene = 0, W = 500, H = 500;
    std::string assets = "assets";
    const char *out = nullptr;
    for (int i = 1; i + 1 < argc; i++) {
        if (!strcmp(argv[i], "-scene")) scene = strtoul(argv[i + 1], 0, 0);
        if (!strcmp(argv[i], "-width")) W = strtoul(argv[i + 1], 0, 0);
        if (!strcmp(argv[i], "-height")) H = strtoul(argv[i + 1], 0, 0);
        if (!strcmp(argv[i], "-assets")) assets = argv[i + 1];
        if (!strcmp(argv[i], "-out")) out = argv[i + 1];
    }
    mrt::SceneGraph g;
    if (!mrt::build_scene(g, scene, float(W) / float(H), assets)) { fprintf(stderr, "scene: %s\n", g.error.c_str()); return 1; }
    FILE *f = out ? fopen(out, "w") : stdout;
    mrt::dump_scene(g, f);
    if (out) fclose(f);
    return 0;
}
