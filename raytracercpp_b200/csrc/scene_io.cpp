// scene_io.cpp -- the load half of the GUI's "load -> upload" path, host side: Wavefront OBJ + MTL into the flat arrays that
// rt_set_triangles / rt_set_materials upload.  SURVEY.md section 8(f)4.
//
// Replaces read_meshio_data (tp2/src/mesh_io.cpp:426-591) with read_materials_mtl (:213-304), followed by
// MeshIOUtils::create_triangles (tp2/projets/utils/meshIOUtils.cpp:4-33) -- what MainWindow::load_obj
// (tp2/projets/QT/mainwindow.cpp:251-282) does before Renderer::set_triangles -- and MainWindow::precompute_materials
// (:240-249).  Same results as the reference's loader, down to its quirks, because the parity tests compare triangle
// for triangle with the compiled reference (tests/test_scene_io.py):
//   * a face is a fan (0, k-1, k); indices are 1-based or negative (= from the end of what has been read so far);
//   * vertices are de-duplicated on (material, position, texcoord, normal) into an indexed mesh first, and a texture
//     coordinate is only appended when the vertex has one: a file that mixes faces with and without texture coordinates
//     ends up with the reference's misaligned texcoord array, reproduced here;
//   * faces before any `usemtl` get the material "default" (diffuse 0.8), appended on first use -- but only if the file has
//     loaded materials by then; without materials the index stays -1;
//   * `vt u v` keeps (u, v); normals are parsed and dropped (the renderer recomputes geometric normals);
//   * numbers go through strtof, like the reference's sscanf("%f").
// Image files are decoded by the caller (the reference uses the vendored stb_image, image_io.cpp:100-158); the decoded
// bytes go to rt_set_texture_u8, which applies read_image's `Color(u8) / 255` on the device.
#include "../../include/rtb200.h"

#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <tuple>
#include <vector>

#include "rt_math.h"

using namespace rtb;

struct RtObjMesh {
    std::vector<float> xyz9, uv6;
    std::vector<int32_t> mat;
    std::vector<RtMaterial> materials;
    std::vector<std::string> names;
    bool has_uv = false;
};

namespace {

void set_error(char* err, size_t cap, const std::string& msg)
{
    if (err && cap) snprintf(err, cap, "%s", msg.c_str());
}

std::string directory_of(const std::string& filename)                // pathname(), files.cpp:68-86 (non-Windows branch)
{
    std::string path = filename;
    for (char& c : path) if (c == '\\') c = '/';
    const size_t slash = path.find_last_of('/');
    return slash == std::string::npos ? std::string("./") : path.substr(0, slash + 1);
}

std::string forward_slashes(std::string s)                            // normalize_filename(), files.cpp:89-99
{
    for (char& c : s) if (c == '\\') c = '/';
    return s;
}

const char* skip_space(const char* p)
{
    while (*p && isspace((unsigned char)*p)) p++;
    return p;
}

// `count` floats after the keyword; false if fewer can be read (sscanf("%f ...") != count)
bool read_floats(const char* p, int count, float* out)
{
    for (int i = 0; i < count; i++) {
        char* end = nullptr;
        out[i] = strtof(p, &end);
        if (end == p) return false;
        p = end;
    }
    return true;
}

// the rest of the line up to CR / LF, as sscanf("%[^\r\n]") reads it (leading blanks skipped by the format's space)
bool read_name(const char* p, std::string& out)
{
    p = skip_space(p);
    const char* e = p;
    while (*e && *e != '\r' && *e != '\n') e++;
    if (e == p) return false;
    out.assign(p, e);
    return true;
}

bool keyword(const char* line, const char* word, const char** rest)
{
    const size_t n = strlen(word);
    if (strncmp(line, word, n) != 0) return false;
    *rest = line + n;
    return true;
}

struct MaterialTable {                                                 // Materials, materials.h:53-170 (what the path reads of it)
    std::vector<RtMaterial> materials;
    std::vector<std::string> names;
    int default_id = -1;

    static RtMaterial blank(float diffuse)                             // Material(Color) -- materials.h:37
    {
        RtMaterial m;
        memset(&m, 0, sizeof(m));
        m.ambient_coeff[0] = m.ambient_coeff[1] = m.ambient_coeff[2] = 1.0f;
        m.diffuse[0] = m.diffuse[1] = m.diffuse[2] = diffuse;
        return m;
    }
    int find(const std::string& name) const
    {
        if (name.empty()) return -1;
        for (size_t i = 0; i < names.size(); i++)
            if (names[i] == name) return (int)i;
        return -1;
    }
    int insert(const RtMaterial& m, const std::string& name)
    {
        int id = find(name);
        if (id == -1) {
            id = (int)materials.size();
            names.push_back(name);
            materials.push_back(m);
        }
        return id;
    }
    int default_index()                                                // Materials::default_material_index, materials.h:145-151
    {
        if (default_id == -1) default_id = insert(blank(0.8f), "default");
        return default_id;
    }
};

// read_materials_mtl -- mesh_io.cpp:213-304
bool read_mtl(const std::string& filename, MaterialTable& table)
{
    FILE* in = fopen(filename.c_str(), "rt");
    if (!in) return false;
    int current = -1;
    char buffer[1024];
    while (fgets(buffer, sizeof(buffer), in)) {
        const char* line = skip_space(buffer);
        const char* rest = nullptr;
        if (line[0] == 'n' && keyword(line, "newmtl", &rest)) {
            std::string name;
            if (read_name(rest, name)) current = table.insert(MaterialTable::blank(0.0f), name);     // Material(Black())
        }
        if (current < 0) continue;
        RtMaterial& m = table.materials[(size_t)current];
        float v[3];
        if (line[0] == 'K') {
            if (keyword(line, "Kd", &rest) && read_floats(rest, 3, v)) memcpy(m.diffuse, v, sizeof(v));
            else if (keyword(line, "Ks", &rest) && read_floats(rest, 3, v)) memcpy(m.specular, v, sizeof(v));
            else if (keyword(line, "Ke", &rest) && read_floats(rest, 3, v)) memcpy(m.emission, v, sizeof(v));
            else if (keyword(line, "Ka", &rest) && read_floats(rest, 3, v)) memcpy(m.ambient_coeff, v, sizeof(v));
        } else if (line[0] == 'N') {
            if (keyword(line, "Ns", &rest) && read_floats(rest, 1, v)) m.ns = v[0];
            // Ni (refraction index), Tf (transmission) and the map_* textures are not read by the renderer
        }
    }
    fclose(in);
    return true;
}

// one vertex of a face: p/t/n, p/t, p//n or p (0 = attribute absent, as in the reference); returns the number of characters
// consumed including the blanks after it, 0 at the end of the face
int read_face_vertex(const char* p, int& ip, int& it, int& in)
{
    const char* s = skip_space(p);
    char* end = nullptr;
    const long a = strtol(s, &end, 10);
    if (end == s) return 0;
    ip = (int)a; it = 0; in = 0;
    const char* q = end;
    auto number_at = [](const char* r, int& value, const char*& after) {
        if (!(isdigit((unsigned char)*r) || ((*r == '-' || *r == '+') && isdigit((unsigned char)r[1])))) return false;
        char* e = nullptr;
        value = (int)strtol(r, &e, 10);
        after = e;
        return true;
    };
    if (q[0] == '/' && q[1] == '/') {                                     // p//n
        const char* after = nullptr;
        if (number_at(q + 2, in, after)) q = after;
    } else if (q[0] == '/') {
        const char* after = nullptr;
        if (number_at(q + 1, it, after)) {                                // p/t
            q = after;
            if (q[0] == '/' && number_at(q + 1, in, after)) q = after;    // p/t/n
        }
    }
    while (*q && !isspace((unsigned char)*q)) q++;                        // whatever else clings to the vertex is skipped
    q = skip_space(q);
    return (int)(q - p);
}

} // namespace

extern "C" {

int rt_obj_load(const char* path, const float transform[16], int32_t current_material_count, RtObjMesh** out, char* err, size_t err_cap)
{
    if (!path || !out) { set_error(err, err_cap, "rt_obj_load: NULL argument"); return RT_ERR_INVALID; }
    *out = nullptr;
    FILE* in = fopen(path, "rt");
    if (!in) { set_error(err, err_cap, std::string("cannot open '") + path + "'"); return RT_ERR_INVALID; }

    // ---- read_meshio_data: the indexed mesh
    std::vector<V3> wpositions, wtexcoords, positions, texcoords;
    std::vector<int> indices, material_indices;
    MaterialTable table;
    std::map<std::tuple<int, int, int, int>, int> remap;
    int material_id = -1;
    bool failed = false;
    std::string why;
    char buffer[1024];
    std::vector<int> fp, ft, fn;
    int n_normals = 0;
    while (!failed && fgets(buffer, sizeof(buffer), in)) {
        const char* line = skip_space(buffer);
        const char* rest = nullptr;
        float v[3];
        if (line[0] == 'v') {
            if (line[1] == ' ') {
                if (!read_floats(line + 1, 3, v)) { failed = true; why = buffer; break; }
                wpositions.push_back(v3(v[0], v[1], v[2]));
            } else if (line[1] == 'n') {
                if (!read_floats(line + 2, 3, v)) { failed = true; why = buffer; break; }
                n_normals++;
            } else if (line[1] == 't') {
                if (!read_floats(line + 2, 2, v)) { failed = true; why = buffer; break; }
                wtexcoords.push_back(v3(v[0], v[1], 0.0f));
            }
        } else if (line[0] == 'f') {
            fp.clear(); ft.clear(); fn.clear();
            const char* p = line + 1;
            for (;;) {
                int a = 0, b = 0, c = 0;
                const int used = read_face_vertex(p, a, b, c);
                // the reference pushes a (0, 0, 0) entry before it tries to read, so the list ends with one invalid vertex
                fp.push_back(a); ft.push_back(b); fn.push_back(c);
                if (used == 0) break;
                p += used;
            }
            if (material_id == -1 && !table.materials.empty()) material_id = table.default_index();
            for (size_t k = 2; k + 1 < fp.size(); k++) {
                material_indices.push_back(material_id);
                const size_t corner[3] = {0, k - 1, k};
                for (int i = 0; i < 3; i++) {
                    const size_t c = corner[i];
                    const int pi = fp[c] < 0 ? (int)wpositions.size() + fp[c] : fp[c] - 1;
                    const int ti = ft[c] < 0 ? (int)wtexcoords.size() + ft[c] : ft[c] - 1;
                    const int ni = fn[c] < 0 ? n_normals + fn[c] : fn[c] - 1;     // normals only take part in the de-duplication key
                    if (pi < 0) break;
                    auto found = remap.insert(std::make_pair(std::make_tuple(material_id, pi, ti, ni), (int)remap.size()));
                    if (found.second) {
                        if (ti != -1) texcoords.push_back(ti >= 0 && ti < (int)wtexcoords.size() ? wtexcoords[(size_t)ti] : v3(0, 0, 0));
                        positions.push_back(pi < (int)wpositions.size() ? wpositions[(size_t)pi] : v3(0, 0, 0));
                    }
                    indices.push_back(found.first->second);
                }
            }
        } else if (line[0] == 'm') {
            std::string name;
            if (keyword(line, "mtllib", &rest) && read_name(rest, name)) {
                const std::string mtl = (name[0] != '/' && !(name.size() > 1 && name[1] == ':')) ? forward_slashes(directory_of(path) + name) : name;
                if (!read_mtl(mtl, table)) { failed = true; why = "cannot open materials '" + mtl + "'"; }
            }
        } else if (line[0] == 'u') {
            std::string name;
            if (keyword(line, "usemtl", &rest) && read_name(rest, name)) material_id = table.find(name);
        }
    }
    fclose(in);
    if (failed) { set_error(err, err_cap, std::string("rt_obj_load '") + path + "': " + why); return RT_ERR_INVALID; }
    if (positions.empty()) { set_error(err, err_cap, std::string("rt_obj_load '") + path + "': no geometry"); return RT_ERR_INVALID; }

    // ---- MeshIOUtils::create_triangles(data, current_material_count, transform)
    M4 m;
    if (transform) memcpy(m.m, transform, sizeof(m.m));
    else for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) m.m[i][j] = i == j ? 1.0f : 0.0f;
    RtObjMesh* mesh = new RtObjMesh();
    const size_t n_tris = indices.size() / 3;
    mesh->has_uv = !texcoords.empty();
    mesh->xyz9.resize(9 * n_tris);
    mesh->mat.resize(n_tris);
    if (mesh->has_uv) mesh->uv6.resize(6 * n_tris);
    for (size_t t = 0; t < n_tris; t++) {
        for (int k = 0; k < 3; k++) {
            const int idx = indices[3 * t + k];
            const V3 p = xform_point(m, positions[(size_t)idx]);          // Transform::operator()(Point), mat.cpp:83-100
            mesh->xyz9[9 * t + 3 * k + 0] = p.x; mesh->xyz9[9 * t + 3 * k + 1] = p.y; mesh->xyz9[9 * t + 3 * k + 2] = p.z;
            if (mesh->has_uv) {
                const V3 uv = (size_t)idx < texcoords.size() ? texcoords[(size_t)idx] : v3(0, 0, 0);   // (the reference reads past the array here)
                mesh->uv6[6 * t + k] = uv.x;
                mesh->uv6[6 * t + 3 + k] = uv.y;
            }
        }
        mesh->mat[t] = material_indices[t] + current_material_count;
    }
    mesh->materials = table.materials;
    mesh->names = table.names;
    *out = mesh;
    return RT_OK;
}

void rt_obj_free(RtObjMesh* mesh) { delete mesh; }
size_t rt_obj_triangle_count(const RtObjMesh* mesh) { return mesh ? mesh->mat.size() : 0; }
size_t rt_obj_material_count(const RtObjMesh* mesh) { return mesh ? mesh->materials.size() : 0; }
const float* rt_obj_xyz9(const RtObjMesh* mesh) { return mesh ? mesh->xyz9.data() : nullptr; }
const float* rt_obj_uv6(const RtObjMesh* mesh) { return mesh && mesh->has_uv ? mesh->uv6.data() : nullptr; }
const int32_t* rt_obj_material_indices(const RtObjMesh* mesh) { return mesh ? mesh->mat.data() : nullptr; }
const RtMaterial* rt_obj_materials(const RtObjMesh* mesh) { return mesh ? mesh->materials.data() : nullptr; }
const char* rt_obj_material_name(const RtObjMesh* mesh, size_t i) { return mesh && i < mesh->names.size() ? mesh->names[i].c_str() : nullptr; }

// MainWindow::precompute_materials -- QT/mainwindow.cpp:240-249: the third term of the luminance is a double product
// (0.0722 has no f suffix), the quotient and the exponent are float, std::pow is the float overload.
void rt_precompute_materials(RtMaterial* mats, size_t n)
{
    for (size_t i = 0; i < n; i++) {
        RtMaterial& m = mats[i];
        const float luminance = (float)((double)(0.2126f * m.specular[0] + 0.7152f * m.specular[1]) + 0.0722 * (double)m.specular[2]);
        m.specular_threshold = std::pow(1.0e-3f / luminance, 1 / m.ns);
    }
}

} // extern "C"
