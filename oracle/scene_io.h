/*
 * scene_io.h -- TEST INFRASTRUCTURE (oracle side).  Reader for the "LTRSCN01"
 * scene files written by lighter_b200/scenes.py, plus the glue that feeds one
 * through the public ltr_* API.  The same translation unit is compiled twice:
 * against the reference sources under /root/reference (-> oracle/_ref/) and
 * against liblighter_b200.so (-> the C++ drop-in smoke driver).  It only uses
 * the public API so it works for both.
 *
 * Scene file layout (little endian, no padding):
 *   char  magic[8] = "LTRSCN01"
 *   CfgBlock cfg                       (see struct below)
 *   u32 n_meshes;  per mesh:  u32 ident_len; bytes; u32 n_parts;
 *        per part: u32 vcount, icount; i32 shadow;
 *                  f32 pos[3*vc], nrm[3*vc], uv1[2*vc], uv2[2*vc]; u32 idx[ic]
 *   u32 n_inst;    per inst:  u32 mesh; f32 matrix[16]; f32 importance;
 *                             i32 shadow; u32 ident_len; bytes; u32 force_w, force_h
 *   u32 n_lights;  per light: 20 x 4 bytes in ltr_LightInfo order
 *   u32 n_probes;  per probe: u32 id; f32 pos[3]; f32 nrm[3]
 */
#pragma once
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <string.h>
#include <stdint.h>
#include <string>
#include <vector>
#include <map>

#include "lighter.h"

namespace sio {

struct CfgBlock {
    uint32_t max_lightmap_size, default_width, default_height;
    float    global_size_factor, max_correct_dist, max_correct_angle;
    float    clear_color[3], ambient_color[3];
    int32_t  bounce_count;
    int32_t  sample_fn_kind;      /* 0 none, 1 = "red wall" rule of the rad1 scenario */
    float    ao_distance, ao_multiplier, ao_falloff, ao_effect, ao_divergence;
    float    ao_color[3];
    int32_t  ao_num_samples;
    float    blur_size;
    int32_t  ds2x;
    int32_t  generate_normalmap_data;
    int32_t  size_fn_kind;        /* 0 default size function, 1 = per-instance forced size */
};

struct Part { std::vector<float> pos, nrm, uv1, uv2; std::vector<uint32_t> idx; int32_t shadow; uint32_t vc, ic; };
struct Mesh { std::string ident; std::vector<Part> parts; };
struct Inst { uint32_t mesh; float matrix[16]; float importance; int32_t shadow; std::string ident; uint32_t fw, fh; };
struct Probe { uint32_t id; float pos[3], nrm[3]; };

struct SceneFile {
    CfgBlock cfg;
    std::vector<Mesh> meshes;
    std::vector<Inst> insts;
    std::vector<ltr_LightInfo> lights;
    std::vector<Probe> probes;
    std::map<std::string, std::pair<uint32_t, uint32_t> > forced;   /* inst ident -> size */
};

static inline bool rd(FILE *f, void *p, size_t n) { return n == 0 || fread(p, 1, n, f) == n; }

static bool load(const char *path, SceneFile &S)
{
    FILE *f = fopen(path, "rb");
    if (!f) { fprintf(stderr, "scene_io: cannot open %s\n", path); return false; }
    char magic[8];
    bool ok = rd(f, magic, 8) && memcmp(magic, "LTRSCN01", 8) == 0;
    ok = ok && rd(f, &S.cfg, sizeof(S.cfg));
    uint32_t n = 0;
    ok = ok && rd(f, &n, 4);
    S.meshes.resize(ok ? n : 0);
    for (size_t m = 0; ok && m < S.meshes.size(); ++m) {
        uint32_t len = 0, np = 0;
        ok = rd(f, &len, 4);
        S.meshes[m].ident.resize(len);
        ok = ok && rd(f, &S.meshes[m].ident[0], len) && rd(f, &np, 4);
        S.meshes[m].parts.resize(ok ? np : 0);
        for (size_t p = 0; ok && p < S.meshes[m].parts.size(); ++p) {
            Part &P = S.meshes[m].parts[p];
            ok = rd(f, &P.vc, 4) && rd(f, &P.ic, 4) && rd(f, &P.shadow, 4);
            if (!ok) break;
            P.pos.resize(3 * (size_t)P.vc); P.nrm.resize(3 * (size_t)P.vc);
            P.uv1.resize(2 * (size_t)P.vc); P.uv2.resize(2 * (size_t)P.vc);
            P.idx.resize(P.ic);
            ok = rd(f, P.pos.data(), P.pos.size() * 4) && rd(f, P.nrm.data(), P.nrm.size() * 4) &&
                 rd(f, P.uv1.data(), P.uv1.size() * 4) && rd(f, P.uv2.data(), P.uv2.size() * 4) &&
                 rd(f, P.idx.data(), P.idx.size() * 4);
        }
    }
    ok = ok && rd(f, &n, 4);
    S.insts.resize(ok ? n : 0);
    for (size_t i = 0; ok && i < S.insts.size(); ++i) {
        Inst &I = S.insts[i];
        uint32_t len = 0;
        ok = rd(f, &I.mesh, 4) && rd(f, I.matrix, 64) && rd(f, &I.importance, 4) &&
             rd(f, &I.shadow, 4) && rd(f, &len, 4);
        if (!ok) break;
        I.ident.resize(len);
        ok = rd(f, &I.ident[0], len) && rd(f, &I.fw, 4) && rd(f, &I.fh, 4);
        if (ok && I.fw && I.fh) S.forced[I.ident] = std::make_pair(I.fw, I.fh);
    }
    ok = ok && rd(f, &n, 4);
    S.lights.resize(ok ? n : 0);
    for (size_t l = 0; ok && l < S.lights.size(); ++l) {
        static_assert(sizeof(ltr_LightInfo) == 80, "light record is 20 words");
        ok = rd(f, &S.lights[l], sizeof(ltr_LightInfo));
    }
    ok = ok && rd(f, &n, 4);
    S.probes.resize(ok ? n : 0);
    for (size_t s = 0; ok && s < S.probes.size(); ++s)
        ok = rd(f, &S.probes[s].id, 4) && rd(f, S.probes[s].pos, 12) && rd(f, S.probes[s].nrm, 12);
    fclose(f);
    if (!ok) fprintf(stderr, "scene_io: %s is truncated or not a scene file\n", path);
    return ok;
}

/* size hook used when cfg.size_fn_kind == 1: userdata = SceneFile* */
static LTRBOOL forced_size_fn(ltr_Config *config, const char *, size_t, const char *inst_ident,
                              size_t inst_ident_size, float, float, u32 out_size[2])
{
    SceneFile *S = (SceneFile *)config->userdata;
    std::map<std::string, std::pair<uint32_t, uint32_t> >::iterator it =
        S->forced.find(std::string(inst_ident ? inst_ident : "", inst_ident_size));
    if (it == S->forced.end()) return 0;
    out_size[0] = it->second.first;
    out_size[1] = it->second.second;
    return 1;
}

/* material hook used when cfg.sample_fn_kind == 1 (the rule the reference's rad1 scenario installs,
 * lighter_test.cpp:470-480: surfaces at x <= -3 facing +x get a dark red albedo) */
static LTRBOOL redwall_sample_fn(ltr_Config *, ltr_SampleRequest *req)
{
    if (req->position[0] <= -3 && req->normal[0] > 0.1f) {
        req->out_diffuse_color[0] = 0.5f;
        req->out_diffuse_color[1] = 0.05f;
        req->out_diffuse_color[2] = 0.02f;
    }
    return 1;
}

/* material hook for cfg.sample_fn_kind == 2: a position / normal / part dependent rule for the synthetic workloads (4-unit
 * checker of two albedos, downward-facing surfaces glow, every 64th call is declined) -- the same rule the library exports as
 * ltrx_SampleFnChecker so that large bakes can be timed with a native callback; stated twice on purpose (the oracle links
 * nothing of the product). */
static LTRBOOL checker_sample_fn(ltr_Config *, ltr_SampleRequest *req)
{
    const int s = (int)floorf(req->position[0] * 0.25f) + (int)floorf(req->position[1] * 0.25f);
    if (((s ^ (int)req->part_id) & 63) == 63) return 0;
    if (s & 1) { req->out_diffuse_color[0] = 0.5f; req->out_diffuse_color[1] = 0.05f; req->out_diffuse_color[2] = 0.02f; }
    else { req->out_diffuse_color[0] = 0.7f; req->out_diffuse_color[1] = 0.7f * req->tex1u; req->out_diffuse_color[2] = 0.6f; }
    if (req->normal[2] < -0.5f) { req->out_emissive_color[0] = 0.1f; req->out_emissive_color[1] = 0.1f; req->out_emissive_color[2] = 0.3f; }
    return 1;
}

/* Feed the scene through the public API.  `threads` <= 0 keeps the library default. */
static ltr_Scene *build(SceneFile &S, int threads)
{
    ltr_Scene *scene = ltr_CreateScene();
    ltr_Config cfg;
    ltr_GetConfig(&cfg, scene);
    const CfgBlock &c = S.cfg;
    cfg.max_lightmap_size = c.max_lightmap_size;
    cfg.default_width = c.default_width;
    cfg.default_height = c.default_height;
    cfg.global_size_factor = c.global_size_factor;
    cfg.max_correct_dist = c.max_correct_dist;
    cfg.max_correct_angle = c.max_correct_angle;
    memcpy(cfg.clear_color, c.clear_color, 12);
    memcpy(cfg.ambient_color, c.ambient_color, 12);
    cfg.bounce_count = c.bounce_count;
    cfg.sample_fn = c.sample_fn_kind == 1 ? redwall_sample_fn : c.sample_fn_kind == 2 ? checker_sample_fn : NULL;
    cfg.ao_distance = c.ao_distance;
    cfg.ao_multiplier = c.ao_multiplier;
    cfg.ao_falloff = c.ao_falloff;
    cfg.ao_effect = c.ao_effect;
    cfg.ao_divergence = c.ao_divergence;
    memcpy(cfg.ao_color_rgb, c.ao_color, 12);
    cfg.ao_num_samples = c.ao_num_samples;
    cfg.blur_size = c.blur_size;
    cfg.ds2x = c.ds2x;
    cfg.generate_normalmap_data = c.generate_normalmap_data;
    if (c.size_fn_kind == 1) { cfg.size_fn = forced_size_fn; cfg.userdata = &S; }
    if (threads > 0) cfg.max_num_threads = threads;
    ltr_SetConfig(scene, &cfg);

    std::vector<ltr_Mesh *> handles;
    for (size_t m = 0; m < S.meshes.size(); ++m) {
        Mesh &M = S.meshes[m];
        ltr_Mesh *h = ltr_CreateMesh(scene, M.ident.c_str(), M.ident.size());
        for (size_t p = 0; p < M.parts.size(); ++p) {
            Part &P = M.parts[p];
            ltr_MeshPartInfo pi;
            memset(&pi, 0, sizeof(pi));
            pi.positions_f3 = P.pos.data();  pi.stride_positions = 12;
            pi.normals_f3 = P.nrm.data();    pi.stride_normals = 12;
            pi.texcoords1_f2 = P.uv1.data(); pi.stride_texcoords1 = 8;
            pi.texcoords2_f2 = P.uv2.data(); pi.stride_texcoords2 = 8;
            pi.indices = P.idx.data();
            pi.vertex_count = P.vc;
            pi.index_count = P.ic;
            pi.shadow = P.shadow;
            if (!ltr_MeshAddPart(h, &pi)) fprintf(stderr, "scene_io: ltr_MeshAddPart failed\n");
        }
        handles.push_back(h);
    }
    for (size_t i = 0; i < S.insts.size(); ++i) {
        Inst &I = S.insts[i];
        ltr_MeshInstanceInfo ii;
        memset(&ii, 0, sizeof(ii));
        memcpy(ii.matrix, I.matrix, 64);
        ii.importance = I.importance;
        ii.shadow = I.shadow;
        ii.ident = I.ident.c_str();
        ii.ident_size = I.ident.size();
        ltr_MeshAddInstance(handles[I.mesh], &ii);
    }
    for (size_t l = 0; l < S.lights.size(); ++l) ltr_LightAdd(scene, &S.lights[l]);
    for (size_t s = 0; s < S.probes.size(); ++s) {
        ltr_SampleInfo si;
        memset(&si, 0, sizeof(si));
        si.id = S.probes[s].id;
        memcpy(si.position, S.probes[s].pos, 12);
        memcpy(si.normal, S.probes[s].nrm, 12);
        ltr_SampleAdd(scene, &si);
    }
    return scene;
}

} // namespace sio
