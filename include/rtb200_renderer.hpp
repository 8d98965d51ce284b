// rtb200_renderer.hpp -- header-only C++ adapter: the method names of the reference's `Renderer`
// (tp2/projets/renderer/renderer.h:38-169) for the ray-tracing path, forwarding to the C ABI of rtb200.h.
//
// It is duck-typed on the reference's own value types so that it compiles inside the reference tree without
// dragging those headers in here (nothing is copied from them):
//   TriangleT   : ._a ._b ._c (.x .y .z), ._tex_coords_u ._tex_coords_v (.x .y .z), ._materialIndex   (triangle.h:93-103)
//   MaterialsT  : .materials = std::vector<MaterialT>; MaterialT: .ambient_coeff .diffuse .specular .emission (.r .g .b),
//                 .reflection .roughness .ns .specular_threshold                                       (materials.h:14-38)
//   ImageT      : .width() .height() .data() -> const float* RGBA                                      (image.h:99-118)
//   TransformT  : .m[4][4] row-major                                                                   (mat.h)
//   PointT      : .x .y .z
// raster_trace() (the hybrid rasterizer) and the SSAO pass run on the device too; see INTEGRATION.md.
#pragma once

#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "rtb200.h"

namespace rtb200 {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& what) : std::runtime_error(what), code(c) {}
};

class Renderer {
public:
    explicit Renderer(int device = 0)
    {
        if (int rc = rt_create(device, &_ctx)) throw Error(rc, rt_last_error(nullptr));
        rt_default_settings(&_settings);
        for (int i = 0; i < 16; i++) _camera_to_world[i] = _previous_object_transform[i] = (i % 5 == 0) ? 1.0f : 0.0f;
        change_render_size(_settings.image_width, _settings.image_height);
    }
    ~Renderer() { rt_destroy(_ctx); }
    Renderer(const Renderer&) = delete;
    Renderer& operator=(const Renderer&) = delete;

    RtSettings& render_settings() { return _settings; }                                  // renderer.cpp:111-114

    void get_render_width_height(const RtSettings& s, int& w, int& h) const              // renderer.cpp:116-120
    {
        w = s.enable_ssaa ? s.image_width * s.ssaa_factor : s.image_width;
        h = s.enable_ssaa ? s.image_height * s.ssaa_factor : s.image_height;
    }

    template <class TriangleT>
    void set_triangles(const std::vector<TriangleT>& tris)                              // renderer.cpp:137-144
    {
        std::vector<float> xyz9(tris.size() * 9), uv6(tris.size() * 6);
        std::vector<int32_t> mat(tris.size());
        for (size_t i = 0; i < tris.size(); i++) {
            const TriangleT& t = tris[i];
            float* p = &xyz9[9 * i];
            p[0] = t._a.x; p[1] = t._a.y; p[2] = t._a.z; p[3] = t._b.x; p[4] = t._b.y; p[5] = t._b.z; p[6] = t._c.x; p[7] = t._c.y; p[8] = t._c.z;
            float* q = &uv6[6 * i];
            q[0] = t._tex_coords_u.x; q[1] = t._tex_coords_u.y; q[2] = t._tex_coords_u.z;
            q[3] = t._tex_coords_v.x; q[4] = t._tex_coords_v.y; q[5] = t._tex_coords_v.z;
            mat[i] = t._materialIndex;
        }
        check(rt_set_triangles(_ctx, xyz9.data(), uv6.data(), mat.data(), tris.size()));
        reconstruct_bvh_new();
    }

    void reconstruct_bvh_new() { check(rt_build_bvh(_ctx, _settings.bvh_max_depth, _settings.bvh_leaf_object_count)); }  // :243-246

    // MainWindow::load_obj (QT/mainwindow.cpp:251-282): read_meshio_data + MeshIOUtils::create_triangles with the native
    // loader (rt_obj_load), the GUI's overrides of material 0 (:258-261), Renderer::set_triangles, the mesh's materials
    // appended to the renderer's, precompute_materials, reset_previous_transform.  `transform`: row-major 4x4 or null.
    void load_obj(const char* filepath, const float* transform16 = nullptr)
    {
        RtObjMesh* mesh = nullptr;
        char err[512] = {0};
        const int rc = rt_obj_load(filepath, transform16, (int32_t)_materials.size(), &mesh, err, sizeof(err));
        if (rc != RT_OK) throw Error(rc, err);
        std::vector<RtMaterial> loaded(rt_obj_materials(mesh), rt_obj_materials(mesh) + rt_obj_material_count(mesh));
        if (!loaded.empty()) {
            loaded[0].roughness = 0.0f;
            loaded[0].reflection = 0.9f;
            for (int k = 0; k < 3; k++) { loaded[0].specular[k] = 0.2f; loaded[0].diffuse[k] = 0.5f; }
        }
        const int set = rt_set_triangles(_ctx, rt_obj_xyz9(mesh), rt_obj_uv6(mesh), rt_obj_material_indices(mesh), rt_obj_triangle_count(mesh));
        rt_obj_free(mesh);
        check(set);
        reconstruct_bvh_new();
        _materials.insert(_materials.end(), loaded.begin(), loaded.end());
        rt_precompute_materials(_materials.data(), _materials.size());
        check(rt_set_materials(_ctx, _materials.data(), _materials.size()));
        reset_previous_transform();
    }

    template <class MaterialsT>
    void set_materials(const MaterialsT& ms)                                             // renderer.cpp:150-152
    {
        std::vector<RtMaterial> out(ms.materials.size());
        for (size_t i = 0; i < out.size(); i++) {
            const auto& m = ms.materials[i];
            RtMaterial& o = out[i];
            o.ambient_coeff[0] = m.ambient_coeff.r; o.ambient_coeff[1] = m.ambient_coeff.g; o.ambient_coeff[2] = m.ambient_coeff.b;
            o.diffuse[0] = m.diffuse.r; o.diffuse[1] = m.diffuse.g; o.diffuse[2] = m.diffuse.b;
            o.specular[0] = m.specular.r; o.specular[1] = m.specular.g; o.specular[2] = m.specular.b;
            o.emission[0] = m.emission.r; o.emission[1] = m.emission.g; o.emission[2] = m.emission.b;
            o.reflection = m.reflection; o.roughness = m.roughness; o.ns = m.ns; o.specular_threshold = m.specular_threshold;
        }
        check(rt_set_materials(_ctx, out.data(), out.size()));
        _materials = out;
    }

    template <class ImageT> void set_ao_map(const ImageT& im) { set_map(RT_TEX_AO, im); }                // renderer.cpp:194-201
    template <class ImageT> void set_diffuse_map(const ImageT& im) { set_map(RT_TEX_DIFFUSE, im); }
    template <class ImageT> void set_normal_map(const ImageT& im) { set_map(RT_TEX_NORMAL, im); }
    template <class ImageT> void set_roughness_map(const ImageT& im) { set_map(RT_TEX_ROUGHNESS, im); }
    template <class ImageT> void set_displacement_map(const ImageT& im) { set_map(RT_TEX_DISPLACEMENT, im); }
    template <class ImageT> void set_skysphere(const ImageT& im) { set_map(RT_TEX_SKYSPHERE, im); }
    // Renderer::set_skybox(const Skybox&) -- renderer.cpp:199.  Skybox keeps its faces private (skybox.h:21), so the
    // adapter takes what Skybox's constructor takes: Image faces[6] = right, left, top, bottom, back, front.
    template <class ImageT> void set_skybox(const ImageT (&faces)[6])
    {
        for (int i = 0; i < 6; i++) set_map(RT_TEX_SKYBOX_RIGHT + i, faces[i]);
    }
    void clear_ao_map() { check(rt_clear_texture(_ctx, RT_TEX_AO)); }                                     // renderer.cpp:203-207
    void clear_diffuse_map() { check(rt_clear_texture(_ctx, RT_TEX_DIFFUSE)); }
    void clear_normal_map() { check(rt_clear_texture(_ctx, RT_TEX_NORMAL)); }
    void clear_roughness_map() { check(rt_clear_texture(_ctx, RT_TEX_ROUGHNESS)); }
    void clear_displacement_map() { check(rt_clear_texture(_ctx, RT_TEX_DISPLACEMENT)); }

    void change_camera_fov(float fov) { _fov = fov; push_camera(); }                                      // renderer.cpp:189
    void change_camera_aspect_ratio(float aspect) { _aspect = aspect; push_camera(); }                    // renderer.cpp:190
    // Renderer::add_analytic_shape(const AnalyticShapesTypes&) -- renderer.cpp:146.  Sphere and Plane keep their members
    // private (analyticShape.h:33-55), so the adapter takes what their constructors take.
    template <class PointT> void add_sphere(const PointT& center, float radius, int mat_index)
    {
        const float c[3] = {center.x, center.y, center.z};
        check(rt_add_sphere(_ctx, c, radius, mat_index));
    }
    template <class PointT, class VectorT> void add_plane(const PointT& point, const VectorT& normal, int mat_index)
    {
        const float p[3] = {point.x, point.y, point.z}, n[3] = {normal.x, normal.y, normal.z};
        check(rt_add_plane(_ctx, p, n, mat_index));
    }
    void clear_analytic_shapes() { check(rt_clear_analytic_shapes(_ctx)); }

    template <class PointT> void set_light_position(const PointT& p)                                      // renderer.cpp:191
    {
        float l[3] = {p.x, p.y, p.z};
        check(rt_set_light(_ctx, l));
    }
    template <class TransformT> void set_camera_transform(const TransformT& t)                            // renderer.cpp:226-233
    {
        for (int r = 0; r < 4; r++)
            for (int c = 0; c < 4; c++) _camera_to_world[4 * r + c] = t.m[r][c];
        const float origin[3] = {0, 0, 0};
        rt_transform_point(_camera_to_world, origin, _position);
        push_camera();
    }
    void reset_previous_transform()                                                                        // renderer.cpp:212
    {
        for (int i = 0; i < 16; i++) _previous_object_transform[i] = (i % 5 == 0) ? 1.0f : 0.0f;
    }
    template <class TransformT> void set_object_transform(const TransformT& t)                            // renderer.cpp:214-224
    {
        float prev_inv[16], now[16], composed[16];
        rt_invert_transform(_previous_object_transform, prev_inv);
        for (int r = 0; r < 4; r++)
            for (int c = 0; c < 4; c++) now[4 * r + c] = t.m[r][c];
        for (int r = 0; r < 4; r++)
            for (int c = 0; c < 4; c++) {
                float acc = 0;
                for (int k = 0; k < 4; k++) acc += now[4 * r + k] * prev_inv[4 * k + c];
                composed[4 * r + c] = acc;
            }
        check(rt_transform_triangles(_ctx, composed, _settings.bvh_max_depth, _settings.bvh_leaf_object_count));
        for (int i = 0; i < 16; i++) _previous_object_transform[i] = now[i];
    }

    void change_render_size(int width, int height)                                                         // renderer.cpp:250-261
    {
        _settings.image_width = width;
        _settings.image_height = height;
        int rw, rh;
        get_render_width_height(_settings, rw, rh);
        _aspect = (float)rw / rh;
        push_camera();
    }

    // Renderer::ray_trace() (renderer.cpp:1068-1116); the SSAA resolve of post_process() happens in the same call.
    void ray_trace() { render_frame(0); }
    // Renderer::raster_trace() (renderer.cpp:869-1006), what RenderThread::run calls when hybrid_rasterization_tracing is set
    // (QT/mainWindowThreads.cpp:46-49): clipping, rasterisation and the z-buffer on the device, every visible fragment shaded with
    // shadow rays and reflection fans as in ray_trace().  Uncovered pixels are the colour of clear_image() (renderer.cpp:175-180).
    void raster_trace() { render_frame(1); }
    void clear_image() {}
    // renderer.cpp:1118-1124: SSAO (enable_ssao) and the SSAA resolve have already run on the device, inside ray_trace()
    void post_process() {}
    // renderer.cpp:152-173 (called by QT/mainwindow.cpp:174-185): the z and normal buffers of the SSAO pass live on the device,
    // are sized by the frame and cleared at the start of every frame that uses them
    void prepare_ssao_buffers() {}
    void destroy_ssao_buffers() {}
    void clear_z_buffer() {}
    void clear_normal_buffer() {}

    // Renderer::get_image(): ARGB32, row 0 = bottom row (renderer.cpp:1086); copy_to() fills a QImage-like object.
    const std::vector<uint32_t>& get_image() const { return _image; }
    template <class QImageT> void copy_to(QImageT& img) const
    {
        for (int y = 0; y < _settings.image_height; y++)
            for (int x = 0; x < _settings.image_width; x++) img.setPixel(x, y, _image[(size_t)y * _settings.image_width + x]);
    }
    const RtRenderStats& last_stats() const { return _stats; }
    RtContext* context() { return _ctx; }

    // bool BVH::intersect(const Ray&, HitInfo&) const (bvh.h:307), batched.
    void intersect(const float* o3, const float* d3, size_t n, int32_t* tri_id, float* t, float* u, float* v)
    {
        check(rt_intersect(_ctx, o3, d3, n, tri_id, t, u, v));
    }

private:
    void check(int rc) { if (rc != RT_OK) throw Error(rc, rt_last_error(_ctx)); }
    template <class ImageT> void set_map(int slot, const ImageT& im) { check(rt_set_texture_f32(_ctx, slot, im.data(), im.width(), im.height())); }
    void push_camera()
    {
        float proj_inv[16];
        rt_perspective_inverse(_fov, _aspect, 0.1f, 1000.0f, proj_inv);                                    // Camera(), scene/camera.h:11
        check(rt_set_camera(_ctx, proj_inv, _camera_to_world, _position));
        check(rt_set_projection(_ctx, _fov, _aspect, 0.1f, 1000.0f));                                      // read by the SSAO pass and by raster_trace
    }

    void render_frame(int hybrid)
    {
        int rw, rh;
        get_render_width_height(_settings, rw, rh);
        _aspect = (float)rw / rh;
        push_camera();
        _image.resize((size_t)_settings.image_width * _settings.image_height);
        RtSettings s = _settings;
        s.hybrid_rasterization_tracing = hybrid;
        check(rt_render(_ctx, &s, _image.data(), &_stats));
    }

    RtContext* _ctx = nullptr;
    std::vector<RtMaterial> _materials;                                                  // Renderer::_materials (what load_obj appends to)
    RtSettings _settings;
    RtRenderStats _stats{};
    std::vector<uint32_t> _image;
    float _fov = 45.0f, _aspect = 1.0f;
    float _camera_to_world[16], _previous_object_transform[16];
    float _position[3] = {0, 0, 0};
};

} // namespace rtb200
