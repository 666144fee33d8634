// host_loader.cpp -- .scn / .ply / .obj loaders of the host side of the drop-in.
//
// Re-implements the behaviour of the reference's hand-rolled tokenizers
// (code/parser.cpp) so that the scene structs come out BIT-IDENTICAL to the
// reference's: same whitespace rules (space, \n, \r -- tabs are not
// whitespace, parser.cpp:110-156), keywords matched by prefix (:13-32), the
// f64-accumulating number reader that is NOT strtof-equivalent (:158-250), fan
// triangulation of faces (:531-567, :845-979), `q w x y z` order (:1218-1225).
// Where the reference asserts (null-deref under MEKA_DEBUG) or loops forever,
// this code returns ORT_ERR_PARSE with a message instead.
#include "host_scene.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

namespace ort {

namespace {

struct Cursor
{
    const uint8_t *at;
    const uint8_t *end;
    bool done() const { return at >= end; }
};

inline bool is_ws(uint8_t c) { return c == ' ' || c == '\n' || c == '\r'; }
void skip_ws(Cursor &c) { while(c.at < c.end && is_ws(*c.at)) c.at++; }                    // eat_all_whitespaces
void skip_to_ws(Cursor &c) { while(c.at < c.end && !is_ws(*c.at)) c.at++; }               // eat_until_whitespace
void skip_to_eol(Cursor &c) { while(c.at < c.end && *c.at != '\n' && *c.at != '\r') c.at++; }   // eat_until_newline

// string_compare (parser.cpp:13-32): true when the text agrees with `kw` up to
// the end of the shorter of the two (the file buffer is NUL-padded here)
bool has_prefix(const Cursor &c, const char *kw)
{
    const uint8_t *a = c.at;
    while(a < c.end && *a != 0 && *kw != 0)
    {
        if(*a != (uint8_t)*kw) return false;
        a++; kw++;
    }
    return true;
}

struct Number
{
    bool is_float;
    union { int32_t i; float f; };
};

} // namespace

// eat_numeric, parser.cpp:158-250
static Number read_number(Cursor &c)
{
    Number r; r.is_float = false; r.i = 0;
    bool scientific = false;
    double adjust = 10.0f;
    double number = 0;
    while(c.at < c.end)
    {
        uint8_t ch = *c.at;
        if(ch >= '0' && ch <= '9') { number *= 10; number += ch - '0'; }
        else if(ch == '.') r.is_float = true;
        else if(ch == 'e') { scientific = true; break; }
        else break;
        if(r.is_float) adjust *= 0.1f;      // f64 *= f32 0.1: the step that makes this differ from strtof
        c.at++;
    }
    if(r.is_float) r.f = (float)(number * adjust);
    else r.i = (int32_t)number;
    if(scientific)
    {
        c.at++;                              // 'e'
        bool plus = c.at < c.end && *c.at == '+';
        c.at++;                              // the sign character (always eaten, parser.cpp:221)
        float sv = 0;
        while(c.at < c.end && *c.at >= '0' && *c.at <= '9') { sv *= 10; sv += (*c.at - '0'); c.at++; }
        r.f *= powf(10.0f, (plus ? 1 : -1) * sv);
    }
    return r;
}

int parse_numeric_text(const char *text, uint32_t *bits)
{
    Cursor c; c.at = (const uint8_t *)text; c.end = c.at + strlen(text);
    Number n = read_number(c);
    memcpy(bits, &n.f, 4);
    return n.is_float ? 1 : 0;
}

int read_file(const char *path, std::vector<uint8_t> *out, std::string *err)
{
    FILE *f = fopen(path, "rb");
    if(!f) { *err = std::string("cannot open ") + path; return ORT_ERR_IO; }
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    out->assign((size_t)(sz > 0 ? sz : 0) + 16, 0);      // NUL padding: prefix compares may look past the end
    if(sz > 0 && fread(out->data(), 1, (size_t)sz, f) != (size_t)sz) { fclose(f); *err = std::string("cannot read ") + path; return ORT_ERR_IO; }
    fclose(f);
    out->resize((size_t)(sz > 0 ? sz : 0));               // capacity keeps the padding
    return ORT_OK;
}

// ---------------------------------------------------------------------------
// PLY (ASCII), parser.cpp:269-570
// ---------------------------------------------------------------------------
namespace {

enum PlyTok { PLY_NULL, PLY_ELEMENT, PLY_VERTEX, PLY_FACE, PLY_END_HEADER, PLY_PROPERTY, PLY_F32, PLY_I32 };
struct PlyToken { PlyTok type; Number num; };

PlyToken ply_token(Cursor &c)                                             // eat_ply_token
{
    PlyToken t; t.type = PLY_NULL; t.num.is_float = false; t.num.i = 0;
    skip_ws(c);
    if(c.done()) return t;
    if(has_prefix(c, "element")) { t.type = PLY_ELEMENT; skip_to_ws(c); }
    else if(has_prefix(c, "vertex")) { t.type = PLY_VERTEX; skip_to_ws(c); }
    else if(has_prefix(c, "face")) { t.type = PLY_FACE; skip_to_ws(c); }
    else if(has_prefix(c, "end_header")) { t.type = PLY_END_HEADER; skip_to_ws(c); }
    else if(has_prefix(c, "property")) { t.type = PLY_PROPERTY; skip_to_eol(c); }
    else
    {
        bool neg = false;
        uint8_t ch = *c.at;
        if(ch == '-') { neg = true; c.at++; }
        if(neg || (ch >= '0' && ch <= '9'))
        {
            t.num = read_number(c);
            if(t.num.is_float) { t.type = PLY_F32; if(neg) t.num.f *= -1.0f; }
            else { t.type = PLY_I32; if(neg) t.num.i *= -1; }
        }
        else skip_to_eol(c);                                             // "ply", "format", "comment", ...
    }
    return t;
}

inline bool ply_is_number(const PlyToken &t) { return t.type == PLY_F32 || t.type == PLY_I32; }
inline float ply_value(const PlyToken &t) { return t.type == PLY_I32 ? (float)t.num.i : t.num.f; }

} // namespace

int load_ply(const uint8_t *mem, size_t size, std::vector<ort_v3> *vertices, std::vector<uint32_t> *indices, std::string *err)
{
    // header (parse_ply_header, parser.cpp:384-466)
    Cursor c; c.at = mem; c.end = mem + size;
    uint32_t vertex_count = 0, property_count = 0;
    bool header_done = false;
    while(!c.done() && !header_done)
    {
        PlyToken t = ply_token(c);
        if(t.type == PLY_ELEMENT)
        {
            PlyToken name = ply_token(c);
            if(name.type == PLY_VERTEX)
            {
                PlyToken n = ply_token(c);
                if(n.type != PLY_I32) { *err = "ply: element vertex without a count"; return ORT_ERR_PARSE; }
                // an ASCII vertex line is at least "0 0 0\n" = 6 bytes: a count the file cannot hold is a parse error,
                // not an allocation
                if(n.num.i < 0 || (uint64_t)n.num.i > (uint64_t)size / 6u + 1u) { *err = "ply: vertex count does not fit the file"; return ORT_ERR_PARSE; }
                vertex_count = (uint32_t)n.num.i;
            }
            else if(name.type != PLY_FACE) { *err = "ply: unsupported element"; return ORT_ERR_PARSE; }
        }
        else if(t.type == PLY_PROPERTY) property_count++;
        else if(t.type == PLY_END_HEADER) header_done = true;
    }
    if(!header_done) { *err = "ply: no end_header"; return ORT_ERR_PARSE; }
    if(property_count == 0) { *err = "ply: no properties"; return ORT_ERR_PARSE; }
    property_count--;        // the face list's own `property list ...` line (parser.cpp:431)
    if(property_count < 3) { *err = "ply: fewer than three vertex properties"; return ORT_ERR_PARSE; }

    // vertices (parse_ply, parser.cpp:486-529): position = the first three numbers of each line
    vertices->resize(vertex_count);
    for(uint32_t i = 0; i < vertex_count; ++i)
    {
        PlyToken x = ply_token(c), y = ply_token(c), z = ply_token(c);
        if(!ply_is_number(x) || !ply_is_number(y) || !ply_is_number(z)) { *err = "ply: malformed vertex line"; return ORT_ERR_PARSE; }
        (*vertices)[i].x = ply_value(x); (*vertices)[i].y = ply_value(y); (*vertices)[i].z = ply_value(z);
        skip_to_eol(c);
    }

    // faces, fan-triangulated (parser.cpp:533-567)
    indices->clear();
    for(;;)
    {
        if(c.done()) break;
        Cursor peek = c;
        if(ply_token(peek).type == PLY_NULL) break;
        PlyToken n = ply_token(c);
        if(n.type != PLY_I32 || n.num.i < 3) { *err = "ply: malformed face line"; return ORT_ERR_PARSE; }
        // every index of the face takes at least two bytes of the file ("0 ")
        if((uint64_t)n.num.i > (uint64_t)(c.end - c.at) / 2u + 1u) { *err = "ply: face arity does not fit the file"; return ORT_ERR_PARSE; }
        PlyToken a = ply_token(c), b = ply_token(c), d = ply_token(c);
        if(a.type != PLY_I32 || b.type != PLY_I32 || d.type != PLY_I32) { *err = "ply: malformed face indices"; return ORT_ERR_PARSE; }
        indices->push_back((uint32_t)a.num.i); indices->push_back((uint32_t)b.num.i); indices->push_back((uint32_t)d.num.i);
        for(int32_t k = 1; k < n.num.i - 2; ++k)
        {
            uint32_t second = indices->back();
            PlyToken e = ply_token(c);
            if(e.type != PLY_I32) { *err = "ply: malformed face indices"; return ORT_ERR_PARSE; }
            indices->push_back((uint32_t)a.num.i); indices->push_back(second); indices->push_back((uint32_t)e.num.i);
        }
    }
    return ORT_OK;
}

// ---------------------------------------------------------------------------
// OBJ, parser.cpp:571-982
// ---------------------------------------------------------------------------
namespace {

enum ObjTok { OBJ_NULL, OBJ_V, OBJ_VN, OBJ_VT, OBJ_F, OBJ_I32, OBJ_F32, OBJ_SLASH };
struct ObjToken { ObjTok type; Number num; };

ObjToken obj_token(Cursor &c)                                             // eat_obj_token
{
    ObjToken t; t.type = OBJ_NULL; t.num.is_float = false; t.num.i = 0;
    skip_ws(c);
    if(c.done()) return t;
    bool neg = false;
    if(*c.at == '-') { neg = true; c.at++; if(c.done()) return t; }
    if(has_prefix(c, "v ")) { t.type = OBJ_V; skip_to_ws(c); }
    else if(has_prefix(c, "vt ")) { t.type = OBJ_VT; skip_to_ws(c); }
    else if(has_prefix(c, "vn ")) { t.type = OBJ_VN; skip_to_ws(c); }
    else if(has_prefix(c, "f ")) { t.type = OBJ_F; skip_to_ws(c); }
    else if(has_prefix(c, "mtllib ") || has_prefix(c, "o ") || has_prefix(c, "usemtl ") ||
            has_prefix(c, "g ") || has_prefix(c, "body")) skip_to_eol(c);
    else if(*c.at == '/') { t.type = OBJ_SLASH; c.at++; }
    else if(*c.at == '#') skip_to_eol(c);
    else if(*c.at >= '0' && *c.at <= '9')
    {
        t.num = read_number(c);
        if(t.num.is_float) { t.type = OBJ_F32; if(neg) t.num.f *= -1.0f; }
        else { t.type = OBJ_I32; if(neg) t.num.i *= -1; }
    }
    else
    {
        // an unknown statement ("s off", "vp", ...): the reference's tokenizer does
        // not advance here and spins forever (parser.cpp:571-673); skip the line
        skip_to_eol(c);
    }
    return t;
}

inline bool obj_to_f32(const ObjToken &t, float *dst)                      // numeric_obj_token_to_f32
{
    if(t.type == OBJ_F32) { *dst = t.num.f; return true; }
    if(t.type == OBJ_I32) { *dst = (float)t.num.i; return true; }
    return false;
}

} // namespace

int load_obj(const uint8_t *mem, size_t size, std::vector<ort_v3> *positions, std::vector<uint32_t> *indices, std::string *err)
{
    if(size == 0) { *err = "obj: empty file"; return ORT_ERR_PARSE; }
    // pass 1 (pre_parse_obj, parser.cpp:687-781): which attributes appear decides the corner layout
    bool v_seen = false, vn_seen = false, vt_seen = false;
    {
        Cursor c; c.at = mem; c.end = mem + size;
        while(!c.done())
        {
            ObjToken t = obj_token(c);
            if(t.type == OBJ_V) v_seen = true;
            else if(t.type == OBJ_VN) vn_seen = true;
            else if(t.type == OBJ_VT) vt_seen = true;
        }
    }
    if(!v_seen) { *err = "obj: no vertex positions"; return ORT_ERR_PARSE; }
    // numbers per face corner: v | v//vn | v/vt | v/vt/vn
    const int per_corner = 1 + (vn_seen ? 1 : 0) + (vt_seen ? 1 : 0);
    if(vt_seen && !vn_seen) { *err = "obj: v/vt faces are not supported (nor by the reference, parser.cpp:930-933)"; return ORT_ERR_PARSE; }

    // pass 2 (parse_obj, parser.cpp:783-982)
    positions->clear(); indices->clear();
    Cursor c; c.at = mem; c.end = mem + size;
    while(!c.done())
    {
        ObjToken t = obj_token(c);
        if(t.type == OBJ_V)
        {
            ObjToken a = obj_token(c), b = obj_token(c), d = obj_token(c);
            ort_v3 p;
            if(!obj_to_f32(a, &p.x) || !obj_to_f32(b, &p.y) || !obj_to_f32(d, &p.z)) { *err = "obj: malformed v line"; return ORT_ERR_PARSE; }
            positions->push_back(p);
        }
        else if(t.type == OBJ_F)
        {
            // collect the integer tokens of the face; every per_corner-th one is a position index
            int number_count = 0;
            int32_t first = 0;
            int corners = 0;
            for(;;)
            {
                Cursor peek = c;
                ObjToken n = obj_token(peek);
                if(n.type == OBJ_I32)
                {
                    if(number_count % per_corner == 0)
                    {
                        int32_t idx = n.num.i - 1;                       // 1-based -> 0-based
                        if(corners == 0) first = idx;
                        if(corners < 3) indices->push_back((uint32_t)idx);
                        else
                        {
                            uint32_t prev = indices->back();             // fan: (first, previous, this)
                            indices->push_back((uint32_t)first); indices->push_back(prev); indices->push_back((uint32_t)idx);
                        }
                        corners++;
                    }
                    number_count++;
                    c = peek;
                }
                else if(n.type == OBJ_SLASH) c = peek;
                else break;
            }
            if(corners < 3) { *err = "obj: face with fewer than three corners"; return ORT_ERR_PARSE; }
        }
    }
    return ORT_OK;
}

int load_mesh_file(const char *path, std::vector<ort_v3> *vertices, std::vector<uint32_t> *indices, std::string *err)
{
    // get_extension (parser.cpp:91-108): everything after the FIRST '.' of the path
    const char *dot = strchr(path, '.');
    if(!dot) { *err = std::string("mesh path without extension: ") + path; return ORT_ERR_ARG; }
    const char *ext = dot + 1;
    std::vector<uint8_t> mem;
    int rc = read_file(path, &mem, err);
    if(rc != ORT_OK) return rc;
    // string_compare: prefix match in either direction
    auto ext_is = [&](const char *kw) { const char *a = ext; while(*a && *kw) { if(*a != *kw) return false; a++; kw++; } return true; };
    if(ext_is("ply")) return load_ply(mem.data(), mem.size(), vertices, indices, err);
    if(ext_is("obj")) return load_obj(mem.data(), mem.size(), vertices, indices, err);
    *err = std::string("unsupported mesh extension (the path's first '.' decides, parser.cpp:91): ") + path;
    return ORT_ERR_ARG;
}

// ---------------------------------------------------------------------------
// .scn, parser.cpp:984-1446
// ---------------------------------------------------------------------------
namespace {

enum ScnTok { SCN_NULL, SCN_SCREEN, SCN_CAMERA, SCN_AMBIENT, SCN_LIGHT, SCN_SPHERE, SCN_BRDF, SCN_BOX, SCN_CYLINDER,
              SCN_MESH, SCN_F32, SCN_I32, SCN_STRING, SCN_B, SCN_Q, SCN_Z };
struct ScnToken { ScnTok type; Number num; const uint8_t *start; int32_t length; };

ScnToken scn_token(Cursor &c)                                             // eat_scn_token
{
    ScnToken t; t.type = SCN_NULL; t.num.is_float = false; t.num.i = 0; t.start = 0; t.length = 0;
    skip_ws(c);
    if(c.done()) return t;
    static const struct { const char *kw; ScnTok type; } words[] = {
        { "screen", SCN_SCREEN }, { "camera", SCN_CAMERA }, { "ambient", SCN_AMBIENT }, { "light", SCN_LIGHT },
        { "sphere", SCN_SPHERE }, { "brdf", SCN_BRDF }, { "box", SCN_BOX }, { "cylinder", SCN_CYLINDER }, { "mesh", SCN_MESH } };
    bool matched = false;
    for(size_t i = 0; i < sizeof(words) / sizeof(words[0]) && !matched; ++i)
        if(has_prefix(c, words[i].kw)) { t.type = words[i].type; matched = true; }
    if(!matched)
    {
        uint8_t ch = *c.at;
        uint8_t nx = (c.at + 1 < c.end) ? c.at[1] : 0;
        if(ch == 'b' && nx == ' ') t.type = SCN_B;
        else if(ch == 'q' && nx == ' ') t.type = SCN_Q;
        else if(ch == 'z' && nx == ' ') t.type = SCN_Z;
        else if((ch >= 'a' && ch < 'z') || (ch >= 'A' && ch < 'Z'))       // sic: 'z'/'Z' excluded, parser.cpp:1059
        {
            t.type = SCN_STRING; t.start = c.at;
            skip_to_ws(c);
            t.length = (int32_t)(c.at - t.start);
        }
        else
        {
            bool neg = false;
            if(ch == '-') { neg = true; c.at++; }
            if(neg || (ch >= '0' && ch <= '9'))
            {
                t.num = read_number(c);
                if(t.num.is_float) { t.type = SCN_F32; if(neg) t.num.f *= -1.0f; }
                else { t.type = SCN_I32; if(neg) t.num.i *= -1; }
            }
        }
    }
    skip_to_ws(c);                                                       // parser.cpp:1120: always
    return t;
}

struct ScnReader
{
    Cursor c;
    std::string *err;
    bool ok;
    float f32(const char *what)
    {
        ScnToken t = scn_token(c);
        if(t.type != SCN_F32) { if(ok) *err = std::string("scn: expected a float (write 1.0, not 1) for ") + what; ok = false; return 0.f; }
        return t.num.f;
    }
    int32_t i32(const char *what)
    {
        ScnToken t = scn_token(c);
        if(t.type != SCN_I32) { if(ok) *err = std::string("scn: expected an integer for ") + what; ok = false; return 0; }
        return t.num.i;
    }
    void expect(ScnTok type, const char *what)
    {
        ScnToken t = scn_token(c);
        if(t.type != type) { if(ok) *err = std::string("scn: expected ") + what; ok = false; }
    }
    float number(const char *what)     // f32 or i32 accepted (mesh quaternion / degree, parser.cpp:1374-1423)
    {
        ScnToken t = scn_token(c);
        if(t.type == SCN_F32) return t.num.f;
        if(t.type == SCN_I32) return (float)t.num.i;
        (void)what;
        return 0.f;                    // the reference leaves the (zeroed) field untouched
    }
};

} // namespace

int parse_scene_text(const uint8_t *mem, size_t size, const char *base_dir, ParsedScene *scene, std::string *err)
{
    scene->materials.clear();
    OrtMaterial zero_mat; memset(&zero_mat, 0, sizeof(zero_mat));
    scene->materials.push_back(zero_mat);                                 // material 0 = "miss", parser.cpp:1187
    ScnReader r; r.c.at = mem; r.c.end = mem + size; r.err = err; r.ok = true;
    while(r.c.at != r.c.end && r.ok)
    {
        ScnToken t = scn_token(r.c);
        switch(t.type)
        {
            case SCN_SCREEN:
                scene->output_width = r.i32("screen width");
                scene->output_height = r.i32("screen height");
                break;
            case SCN_CAMERA:
            {
                float x = r.f32("camera x"), y = r.f32("camera y"), z = r.f32("camera z");
                r.expect(SCN_B, "'b' after the camera position");
                float ratio = r.f32("camera height ratio");
                r.expect(SCN_Q, "'q' before the camera orientation");
                float qw = r.f32("camera q.w"), qx = r.f32("camera q.x"), qy = r.f32("camera q.y"), qz = r.f32("camera q.z");
                scene->camera_p.x = x; scene->camera_p.y = y; scene->camera_p.z = z;
                scene->camera_height_ratio = ratio;
                scene->camera_quaternion.x = qx; scene->camera_quaternion.y = qy; scene->camera_quaternion.z = qz; scene->camera_quaternion.w = qw;
            } break;
            case SCN_AMBIENT:
                scene->ambient.x = r.f32("ambient r"); scene->ambient.y = r.f32("ambient g"); scene->ambient.z = r.f32("ambient b");
                break;
            case SCN_LIGHT:
            {
                int32_t cr = r.i32("light r"), cg = r.i32("light g"), cb = r.i32("light b");
                OrtMaterial m; memset(&m, 0, sizeof(m));
                m.is_light = 1;
                m.emit_color.x = (float)cr; m.emit_color.y = (float)cg; m.emit_color.z = (float)cb;
                scene->materials.push_back(m);
            } break;
            case SCN_SPHERE:
            {
                OrtSphere s;
                s.center.x = r.f32("sphere x"); s.center.y = r.f32("sphere y"); s.center.z = r.f32("sphere z");
                s.r = r.f32("sphere r");
                s.mat_index = (uint32_t)scene->materials.size() - 1;
                scene->spheres.push_back(s);
                if(scene->materials[s.mat_index].is_light)               // parser.cpp:1262-1266
                    scene->lights.push_back(ParsedLight{ ORT_SHAPE_SPHERE, (uint32_t)scene->spheres.size() - 1 });
            } break;
            case SCN_BRDF:
            {
                OrtMaterial m; memset(&m, 0, sizeof(m));
                m.diffuse.x = r.f32("brdf diffuse r"); m.diffuse.y = r.f32("brdf diffuse g"); m.diffuse.z = r.f32("brdf diffuse b");
                m.specular.x = r.f32("brdf specular r"); m.specular.y = r.f32("brdf specular g"); m.specular.z = r.f32("brdf specular b");
                m.specular.w = (float)r.i32("brdf alpha");
                Cursor peek = r.c;
                ScnToken nx = scn_token(peek);
                if(nx.type == SCN_F32 || nx.type == SCN_I32)
                {
                    m.transmission.x = r.f32("brdf transmission r"); m.transmission.y = r.f32("brdf transmission g");
                    m.transmission.z = r.f32("brdf transmission b");
                    m.ior = r.f32("brdf ior");
                }
                scene->materials.push_back(m);
            } break;
            case SCN_BOX:
            {
                OrtAAB b;
                b.min.x = r.f32("box x"); b.min.y = r.f32("box y"); b.min.z = r.f32("box z");
                float dx = r.f32("box dx"), dy = r.f32("box dy"), dz = r.f32("box dz");
                b.max.x = b.min.x + dx; b.max.y = b.min.y + dy; b.max.z = b.min.z + dz;
                b.mat_index = (uint32_t)scene->materials.size() - 1;
                scene->boxes.push_back(b);
            } break;
            case SCN_CYLINDER:
            {
                OrtCylinder cy;
                cy.base.x = r.f32("cylinder base x"); cy.base.y = r.f32("cylinder base y"); cy.base.z = r.f32("cylinder base z");
                cy.axis.x = r.f32("cylinder axis x"); cy.axis.y = r.f32("cylinder axis y"); cy.axis.z = r.f32("cylinder axis z");
                cy.r = r.f32("cylinder r");
                cy.mat_index = (uint32_t)scene->materials.size() - 1;
                scene->cylinders.push_back(cy);
                if(cy.mat_index)                                         // sic: every cylinder with a material, parser.cpp:1345-1348
                    scene->lights.push_back(ParsedLight{ ORT_SHAPE_CYLINDER, (uint32_t)scene->cylinders.size() - 1 });
            } break;
            case SCN_MESH:
            {
                ParsedMesh mi; memset(&mi.translate, 0, sizeof(mi.translate)); memset(&mi.quaternion, 0, sizeof(mi.quaternion));
                mi.scale = 0.f; mi.degree = 0.f;
                ScnToken name = scn_token(r.c);
                if(name.type != SCN_STRING) { if(r.ok) *err = "scn: expected a mesh file name"; r.ok = false; break; }
                mi.translate.x = r.f32("mesh tx"); mi.translate.y = r.f32("mesh ty"); mi.translate.z = r.f32("mesh tz");
                mi.scale = r.f32("mesh scale");
                ScnToken det = scn_token(r.c);
                if(det.type == SCN_Z)
                {
                    mi.degree = r.number("mesh degree");
                    r.expect(SCN_Q, "'q' after the mesh axis rotation");
                }
                else if(det.type != SCN_Q) { if(r.ok) *err = "scn: expected 'z' or 'q' after the mesh scale"; r.ok = false; break; }
                float qw = r.number("mesh q.w"), qx = r.number("mesh q.x"), qy = r.number("mesh q.y"), qz = r.number("mesh q.z");
                mi.quaternion.x = qx; mi.quaternion.y = qy; mi.quaternion.z = qz; mi.quaternion.w = qw;
                mi.file_path = std::string(base_dir) + std::string((const char *)name.start, (size_t)name.length);
                mi.mat_index = (uint32_t)scene->materials.size() - 1;
                scene->meshes.push_back(mi);
            } break;
            default:
                break;
        }
    }
    return r.ok ? ORT_OK : ORT_ERR_PARSE;
}

} // namespace ort
