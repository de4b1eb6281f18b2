// The reference's nine hard-coded scenes (scene.cpp:51-529) rebuilt on the host
// scene graph, draw for draw: the same PCG32 stream (main.cpp:302) is consumed
// in the same order (arguments left -> right, the order of the reference's
// documented compiler; see oracle/patch_reference.py P8), so object placement,
// materials and BVH topology come out identical.
#include <cmath>
#include <cstdio>
#include <cstring>

#include "scene_graph.h"

namespace mrt {

static int load_ppm(SceneGraph &g, const std::string &path) {
    // Decoded copy of earthmap.jpg (binary PPM "P6").  JPEG decoding itself is
    // third-party load-time code in the reference (stb_image, scene.cpp:139) and is
    // out of scope here: the bytes are produced once by oracle/build_ref.sh.
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) { g.error = "cannot open " + path; return -1; }
    int w = 0, h = 0, maxv = 0;
    char magic[3] = {0, 0, 0};
    if (fscanf(f, "%2s %d %d %d", magic, &w, &h, &maxv) != 4 || strcmp(magic, "P6") != 0 || maxv != 255) {
        fclose(f);
        g.error = "bad ppm " + path;
        return -1;
    }
    fgetc(f);
    Image img;
    img.width = w;
    img.height = h;
    img.rgb.resize((size_t) w * h * 3);
    size_t got = fread(img.rgb.data(), 1, img.rgb.size(), f);
    fclose(f);
    if (got != img.rgb.size()) { g.error = "short ppm " + path; return -1; }
    g.images.push_back(std::move(img));
    return (int) g.images.size() - 1;
}

static Camera weekend_camera(float aspect) {   // scene.cpp:53-63 (shared by scenes 0-4)
    H3 cam_pos(11, 2.2f, 2.5f), lookat(2.8f, 0.5f, 1.2f), up(0, 1, 0);
    float focus_dist = hlength(cam_pos - lookat);
    return Camera(cam_pos, lookat, up, 27.0f, aspect, 0.09f, focus_dist, 0.0f, 1.0f);
}
static Camera cornell_camera(float aspect, float aperture) {   // scene.cpp:285-295
    H3 cam_pos(278, 278, -800), lookat(278, 278, 100), up(0, 1, 0);
    float focus_dist = hlength(cam_pos - lookat);
    return Camera(cam_pos, lookat, up, 40.0f, aspect, aperture, focus_dist, 0.0f, 1.0f);
}

// scene.cpp:51-119 (variant 2: scene.cpp:121-200)
static bool random_scene(SceneGraph &g, HostRng &rng, int n, float aspect, bool variant2, const std::string &assets) {
    g.camera = weekend_camera(aspect);
    std::vector<int> list;
    int earth = -1, checker_mat = -1, perlin = -1, perlin_small = -1;
    if (!variant2) {
        int checker = g.checker_tex(g.color_tex(H3(0.2f, 0.3f, 0.1f)), g.color_tex(H3(0.9f, 0.9f, 0.9f)), 10.0f);
        list.push_back(g.sphere(H3(0, -1000, 0), 1000, g.lambertian(checker)));
    } else {
        int img = load_ppm(g, assets + "/earthmap.ppm");
        if (img < 0) return false;
        earth = g.lambertian(g.image_tex(img));
        checker_mat = g.lambertian(g.checker_tex(g.color_tex(H3(0.2f, 0.3f, 0.1f)), g.color_tex(H3(0.9f, 0.9f, 0.9f)), 10.0f));
        perlin = g.lambertian(g.perlin_tex(1.0f));
        perlin_small = g.lambertian(g.perlin_tex(4.0f));
        list.push_back(g.sphere(H3(0, -1000, 0), 1000, perlin));
    }
    int half_sqrt_n = int(sqrtf(float(n)) * 0.5f);
    for (int a = -half_sqrt_n; a < half_sqrt_n; a++) {
        for (int b = -half_sqrt_n; b < half_sqrt_n; b++) {
            float choose_mat = rng.randf();
            float cx = rng.randf(), cz = rng.randf();
            H3 center(a + 0.9f * cx, 0.2f, b + 0.9f * cz);
            if (hlength(center - H3(4, 0.2f, 0)) > 0.9f) {
                int mat, sph;
                auto lambert_moving = [&]() {
                    float r0 = rng.randf(), r1 = rng.randf(), r2 = rng.randf(), r3 = rng.randf(), r4 = rng.randf(), r5 = rng.randf();
                    mat = g.lambertian(g.color_tex(H3(r0 * r1, r2 * r3, r4 * r5)));
                    sph = g.sphere(center, 0.2f, mat, center + H3(0, 0.5f * rng.randf(), 0), 0.0f, 1.0f);
                };
                auto metal_rand = [&]() {
                    float r0 = rng.randf(), r1 = rng.randf(), r2 = rng.randf(), r3 = rng.randf();
                    return g.metal(g.color_tex(0.5f * H3(1 + r0, 1 + r1, 1 + r2)), r3);
                };
                if (!variant2) {
                    if (choose_mat < 0.5f) {
                        lambert_moving();
                    } else if (choose_mat < 0.9f) {
                        mat = metal_rand();
                        sph = g.sphere(center, 0.2f, mat);
                    } else {
                        mat = g.dielectric(1.4f + rng.randf());
                        sph = g.sphere(center, 0.2f, mat);
                    }
                } else {
                    if (choose_mat < 0.3f) {
                        lambert_moving();
                    } else {
                        if (choose_mat < 0.6f) mat = metal_rand();
                        else if (choose_mat < 0.7f) mat = g.dielectric(1.4f + rng.randf());
                        else if (choose_mat < 0.75f) mat = earth;
                        else mat = perlin_small;
                        sph = g.sphere(center, 0.2f, mat);
                    }
                }
                list.push_back(sph);
            }
        }
    }
    list.push_back(g.sphere(H3(0, 1, 0), 1.0f, g.dielectric(1.5f)));
    if (!variant2) list.push_back(g.sphere(H3(-4, 1, 0), 1.0f, g.lambertian(g.color_tex(H3(0.4f, 0.2f, 0.1f)))));
    else list.push_back(g.sphere(H3(-4, 1, 0), 1.0f, checker_mat));
    list.push_back(g.sphere(H3(4, 1, 0), 1.0f, g.metal(g.color_tex(H3(0.7f, 0.6f, 0.5f)), 1.0f)));
    list.push_back(g.sphere(H3(4, 1, 3), 1.0f, g.dielectric(2.4f)));
    list.push_back(g.sphere(H3(4, 1, 3), -0.95f, g.dielectric(2.4f)));
    g.objects = g.bvh(list, 0, list.size(), 0.0f, 1.0f);
    g.biased = -1;
    return g.objects >= 0;
}

static bool two_spheres(SceneGraph &g, float aspect) {   // scene.cpp:203-228
    g.camera = weekend_camera(aspect);
    int checker = g.checker_tex(g.color_tex(H3(0.2f, 0.3f, 0.1f)), g.color_tex(H3(0.9f, 0.9f, 0.9f)), 10.0f);
    std::vector<int> l;
    l.push_back(g.sphere(H3(0, -10, 0), 10, g.lambertian(checker)));
    l.push_back(g.sphere(H3(0, 10, 0), 10, g.lambertian(checker)));
    g.objects = g.list(l, 0.0f, 1.0f);
    return true;
}

static bool spheres_perlin(SceneGraph &g, float aspect) {   // scene.cpp:230-254
    g.camera = weekend_camera(aspect);
    std::vector<int> l;
    l.push_back(g.sphere(H3(0, -1001, 0), 1000, g.lambertian(g.perlin_tex(1.0f))));
    l.push_back(g.sphere(H3(0, 1, 0), 2, g.lambertian(g.perlin_tex(4.0f))));
    l.push_back(g.sphere(H3(0.5f, -0.5f, 2), 0.5f, g.lambertian(g.perlin_tex(16.0f))));
    g.objects = g.list(l, 0.0f, 1.0f);
    return true;
}

static bool earth(SceneGraph &g, float aspect, const std::string &assets) {   // scene.cpp:256-285
    g.camera = weekend_camera(aspect);
    int img = load_ppm(g, assets + "/earthmap.ppm");
    if (img < 0) return false;
    int mat = g.lambertian(g.image_tex(img));
    std::vector<int> l;
    l.push_back(g.sphere(H3(0, -1001, 0), 1000, g.lambertian(g.perlin_tex(1.0f))));
    l.push_back(g.sphere(H3(0, 1, 0), 2, mat));
    l.push_back(g.sphere(H3(0.5f, -0.5f, 2), 0.5f, mat));
    g.objects = g.list(l, 0.0f, 1.0f);
    return true;
}

static bool cornell_box(SceneGraph &g, float aspect, bool all_lights, bool extra_triangles) {   // scene.cpp:283-332
    g.camera = cornell_camera(aspect, 0.0f);
    int red = g.lambertian(g.color_tex(H3(0.65f, 0.055f, 0.06f)));
    int white = g.lambertian(g.color_tex(H3(0.73f, 0.73f, 0.73f)));
    int green = g.lambertian(g.color_tex(H3(0.117f, 0.44f, 0.115f)));
    int light = g.diffuse_light(g.color_tex(H3(15.f, 15.f, 15.f)));
    int glass = g.dielectric(1.5f);
    std::vector<int> l;
    l.push_back(g.yz_rect(555, 0, 0, 555, 555, green));
    l.push_back(g.yz_rect(0, 555, 0, 555, 0, red));
    int lrect = g.xz_rect(343, 213, 227, 332, 554, light);
    l.push_back(lrect);
    l.push_back(g.xz_rect(555, 0, 0, 555, 555, white));
    l.push_back(g.xz_rect(0, 555, 0, 555, 0, white));
    l.push_back(g.xy_rect(555, 0, 0, 555, 555, white));
    l.push_back(g.translate(g.rotate_y(g.box(H3(0, 0, 0), H3(165, 330, 165), white), 15), H3(265, 0, 295)));
    int s = g.sphere(H3(190, 90, 190), 90, glass);
    l.push_back(s);
    if (extra_triangles) {
        // MRT_SCENE_EXTRA_TRIANGLES: two triangle_scene_objects (triangle.cpp:5-175; no stock scene uses the class) appended to the
        // object list -- one with a face normal, one with vertex normals; the oracle harness appends the same two (`-extra triangles`)
        l.push_back(g.triangle(H3(100, 300, 250), H3(400, 320, 300), H3(250, 520, 420), red));
        l.push_back(g.triangle(H3(420, 60, 120), H3(520, 60, 260), H3(470, 260, 180), H3(0, 0, -1), H3(-0.6f, 0, -0.8f), H3(0, 0.6f, -0.8f),
                               g.metal(g.color_tex(H3(0.8f, 0.85f, 0.88f)), 0.9f)));
    }
    g.objects = g.list(l, 0.0f, 1.0f);
    // the array holds {light, sphere} but the list is built with count 1 (scene.cpp:326-329); MRT_SCENE_ALL_LIGHTS uses both
    g.biased = all_lights ? g.list(std::vector<int>{lrect, s}, 0.0f, 1.0f) : g.list(std::vector<int>{lrect}, 0.0f, 1.0f);
    return true;
}

static bool cornell_smoke(SceneGraph &g, float aspect) {   // scene.cpp:334-378
    g.camera = cornell_camera(aspect, 0.0f);
    int red = g.lambertian(g.color_tex(H3(0.65f, 0.05f, 0.05f)));
    int white = g.lambertian(g.color_tex(H3(0.73f, 0.73f, 0.73f)));
    int green = g.lambertian(g.color_tex(H3(0.12f, 0.45f, 0.15f)));
    int light = g.diffuse_light(g.color_tex(H3(7.0f, 7.0f, 7.0f)));
    std::vector<int> l;
    l.push_back(g.yz_rect(555, 0, 0, 555, 555, green));
    l.push_back(g.yz_rect(0, 555, 0, 555, 0, red));
    int lrect = g.xz_rect(443, 113, 127, 432, 554, light);
    l.push_back(lrect);
    l.push_back(g.xz_rect(555, 0, 0, 555, 555, white));
    l.push_back(g.xz_rect(0, 555, 0, 555, 0, white));
    l.push_back(g.xy_rect(555, 0, 0, 555, 555, white));
    int smoke_box1 = g.translate(g.rotate_y(g.box(H3(0, 0, 0), H3(165, 165, 165), white), -18), H3(130, 0, 65));
    int smoke_box2 = g.translate(g.rotate_y(g.box(H3(0, 0, 0), H3(165, 330, 165), white), 15), H3(265, 0, 295));
    l.push_back(g.volume(smoke_box1, 0.01f, g.color_tex(H3(1.0f, 1.0f, 1.0f))));
    l.push_back(g.volume(smoke_box2, 0.01f, g.color_tex(H3(0.0f, 0.0f, 0.0f))));
    g.objects = g.list(l, 0.0f, 1.0f);
    g.biased = g.list(std::vector<int>{lrect}, 0.0f, 1.0f);
    return true;
}

static bool book2_final(SceneGraph &g, HostRng &rng, float aspect, const std::string &assets, bool all_lights) {   // scene.cpp:380-462
    H3 cam_pos(450, 278, -560), lookat(200, 278, 300), up(0, 1, 0);
    float focus_dist = hlength(cam_pos - lookat);
    g.camera = Camera(cam_pos, lookat, up, 40.0f, aspect, 0.0f, focus_dist, 0.0f, 1.0f);
    const int nb = 20, ns = 1000;
    int img = load_ppm(g, assets + "/earthmap.ppm");
    if (img < 0) return false;
    int earth = g.lambertian(g.image_tex(img));
    int white = g.lambertian(g.color_tex(H3(0.73f, 0.73f, 0.73f)));
    int green = g.lambertian(g.color_tex(H3(0.48f, 0.83f, 0.53f)));
    int light = g.diffuse_light(g.color_tex(H3(7.0f, 7.0f, 7.0f)));
    int orange = g.lambertian(g.color_tex(H3(0.7f, 0.3f, 0.1f)));
    int perlin = g.lambertian(g.perlin_tex(0.05f));

    std::vector<int> boxlist;
    for (int i = 0; i < nb; i++) {
        for (int j = 0; j < nb; j++) {
            float w = 100;
            float x0 = -1000 + i * w;
            float z0 = -1000 + j * w;
            float y0 = 0;
            float x1 = x0 + w;
            float y1 = 100 * (rng.randf() + 0.01f);
            float z1 = z0 + w;
            boxlist.push_back(g.box(H3(x0, y0, z0), H3(x1, y1, z1), green));
        }
    }
    std::vector<int> l;
    l.push_back(g.bvh(boxlist, 0, boxlist.size(), 0.0f, 1.0f));
    int lo = g.xz_rect(423, 123, 147, 412, 554, light);
    l.push_back(lo);
    H3 center(400, 400, 200);
    l.push_back(g.sphere(center, 50, orange, center + H3(30, 0, 0), 0, 1));
    const int gs = g.sphere(H3(260, 150, 45), 50, g.dielectric(1.5f));
    l.push_back(gs);
    l.push_back(g.sphere(H3(0, 150, 145), 50, g.metal(g.color_tex(H3(0.8f, 0.8f, 0.9f)), 0.1f)));
    l.push_back(g.sphere(H3(400, 200, 400), 100, earth));
    l.push_back(g.sphere(H3(220, 280, 300), 80, perlin));
    int boundary = g.sphere(H3(360, 150, 145), 70, g.dielectric(1.5f));
    l.push_back(boundary);
    l.push_back(g.volume(boundary, 0.2f, g.color_tex(H3(0.2f, 0.4f, 0.9f))));
    boundary = g.sphere(H3(0, 0, 0), 5000, g.dielectric((float) 1.5));
    l.push_back(g.volume(boundary, 0.0001f, g.color_tex(H3(1.0f, 1.0f, 1.0f))));
    std::vector<int> spherelist;
    for (int i = 0; i < ns; i++) {
        float r0 = rng.randf(), r1 = rng.randf(), r2 = rng.randf();
        spherelist.push_back(g.sphere(H3(165 * r0, 165 * r1, 165 * r2), 10, white));
    }
    l.push_back(g.translate(g.rotate_y(g.bvh(spherelist, 0, spherelist.size(), 0.0f, 1.0f), 15), H3(-100, 270, 395)));
    g.objects = g.list(l, 0.0f, 1.0f);
    // {light, glass sphere} with count 1 (scene.cpp:456-459); MRT_SCENE_ALL_LIGHTS uses both
    g.biased = all_lights ? g.list(std::vector<int>{lo, gs}, 0.0f, 1.0f) : g.list(std::vector<int>{lo}, 0.0f, 1.0f);
    return true;
}

static bool triangles(SceneGraph &g, float aspect, const std::string &assets) {   // scene.cpp:464-529
    g.camera = cornell_camera(aspect, 20.0f);
    int red = g.lambertian(g.color_tex(H3(0.65f, 0.05f, 0.05f)));
    int white = g.lambertian(g.color_tex(H3(0.73f, 0.73f, 0.73f)));
    int green = g.lambertian(g.color_tex(H3(0.12f, 0.45f, 0.15f)));
    int light = g.diffuse_light(g.color_tex(H3(4.0f, 4.0f, 4.0f)));
    int silver = g.metal(g.color_tex(H3(0.8f, 0.8f, 0.9f)), 0.9f);
    int dia = g.dielectric(2.4f);
    (void) red; (void) white; (void) green;
    std::vector<int> l;
    l.push_back(g.yz_rect(555, 0, 0, 555, 555, green));
    l.push_back(g.yz_rect(0, 555, 0, 555, 0, red));
    int lrect = g.xz_rect(443, 113, 127, 432, 554, light);
    l.push_back(lrect);
    l.push_back(g.xz_rect(555, 0, 0, 555, 555, white));
    l.push_back(g.xz_rect(0, 555, 0, 555, 0, white));
    l.push_back(g.xy_rect(555, 0, 0, 555, 555, silver));

    // a missing OBJ file is skipped silently, as in the reference (obj_loader.cpp:159-162)
    std::vector<Triangle> bunny;
    if (read_obj(assets + "/obj/bunny.obj", true, M4::scale(2000.0f), H3(195, -20, 280), M4::identity(), &bunny) && !bunny.empty())
        l.push_back(g.pod_bvh(std::move(bunny), dia));
    std::vector<Triangle> teapot;
    const float rad30 = 30 * (3.14159265358979323846f / 180.0f);
    if (read_obj(assets + "/obj/teapot3_no_vt.obj", false, M4::scale(250.0f), H3(393, 50, 108), M4::rotate_y(rad30), &teapot) && !teapot.empty())
        l.push_back(g.pod_bvh(std::move(teapot), dia));
    g.objects = g.list(l, 0.0f, 1.0f);
    g.biased = g.list(std::vector<int>{lrect}, 0.0f, 1.0f);
    return true;
}

bool build_scene(SceneGraph &g, uint32_t scene_and_flags, float aspect, const std::string &asset_dir) {
    const uint32_t scene = scene_and_flags & 0xFFu;
    const bool all_lights = (scene_and_flags & MRT_SCENE_ALL_LIGHTS) != 0;
    const bool extra_triangles = (scene_and_flags & MRT_SCENE_EXTRA_TRIANGLES) != 0;
    HostRng rng;
    rng.seed(11350390909718046443uLL, 6305599193148252115uLL);   // main.cpp:302
    g.sky = scene < 5;                                            // main.cpp:110
    switch (scene) {
    case 0: return random_scene(g, rng, 500, aspect, false, asset_dir);
    case 1: return random_scene(g, rng, 500, aspect, true, asset_dir);
    case 2: return two_spheres(g, aspect);
    case 3: return spheres_perlin(g, aspect);
    case 4: return earth(g, aspect, asset_dir);
    case 5: return cornell_box(g, aspect, all_lights, extra_triangles);
    case 6: return cornell_smoke(g, aspect);
    case 7: return book2_final(g, rng, aspect, asset_dir, all_lights);
    case 8: return triangles(g, aspect, asset_dir);
    default: g.error = "unknown scene"; return false;
    }
}

}  // namespace mrt
