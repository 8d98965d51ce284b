// Compiles include/rtb200_renderer.hpp the way the reference's harness would use its Renderer
// (utils/mainUtils.cpp:6-21) with stand-in value types shaped like the reference's Triangle / Materials / Image.
#include <cstdio>
#include <vector>
#include "rtb200_renderer.hpp"

struct P3 { float x, y, z; };
struct Tri { P3 _a, _b, _c, _tex_coords_u, _tex_coords_v; int _materialIndex; };
struct Rgb { float r, g, b; };
struct Mat { Rgb ambient_coeff, diffuse, specular, emission; float reflection, roughness, ns, specular_threshold; };
struct Mats { std::vector<Mat> materials; };
struct Xf { float m[4][4]; };

int main()
{
    try {
        rtb200::Renderer renderer(0);
        std::vector<Tri> tris;
        for (int i = 0; i < 64; i++) {
            float x = -2.0f + 0.06f * i, z = -4.0f - 0.01f * i;
            tris.push_back(Tri{{x, -1, z}, {x + 0.5f, -1, z}, {x + 0.25f, 1, z}, {0, 1, 0.5f}, {0, 0, 1}, 0});
        }
        RtSettings& s = renderer.render_settings();
        s.compute_shadows = 1;
        s.enable_ssaa = 1;
        s.ssaa_factor = 2;
        renderer.change_render_size(160, 90);
        renderer.change_camera_fov(80.0f);
        renderer.set_triangles(tris);
        Mats mats;
        mats.materials.push_back(Mat{{1, 1, 1}, {0.7f, 0.3f, 0.1f}, {0.5f, 0.5f, 0.5f}, {0, 0, 0}, 0.0f, 0.0f, 20.0f, 0.73f});
        renderer.set_materials(mats);
        renderer.set_light_position(P3{3, 3, 2});
        Xf cam = {{{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0.5f}, {0, 0, 0, 1}}};
        renderer.set_camera_transform(cam);
        renderer.ray_trace();
        renderer.post_process();
        const RtRenderStats& st = renderer.last_stats();
        size_t lit = 0;
        for (uint32_t c : renderer.get_image()) lit += (c != 0xff87ceebu);
        std::printf("rays %llu hits %llu launches %u lit %zu\n", (unsigned long long)st.primary_rays,
                    (unsigned long long)st.primary_hits, st.kernel_launches, lit);
        if (st.primary_rays != 160u * 90u * 4u || st.primary_hits == 0 || lit == 0) return 1;
        // the hybrid path: RenderThread::run calls raster_trace() when hybrid_rasterization_tracing is set (QT/mainWindowThreads.cpp:46-49)
        renderer.clear_z_buffer();
        renderer.clear_image();
        renderer.raster_trace();
        renderer.post_process();
        const RtRenderStats& rs = renderer.last_stats();
        size_t lit2 = 0;
        for (uint32_t c : renderer.get_image()) lit2 += (c != 0xff87ceebu);
        std::printf("raster: fragments %llu hits %llu launches %u lit %zu\n", (unsigned long long)rs.primary_rays,
                    (unsigned long long)rs.primary_hits, rs.kernel_launches, lit2);
        if (rs.primary_hits == 0 || lit2 == 0 || lit2 > lit + lit / 4 || lit2 + lit / 4 < lit) return 1;
        std::puts("adapter ok");
        return 0;
    } catch (const rtb200::Error& e) {
        std::printf("error %d: %s\n", e.code, e.what());
        return 2;
    }
}
