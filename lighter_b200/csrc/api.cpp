/*
 * api.cpp -- the ltr_* C API (drop-in boundary) over the GPU bake pipeline.
 *
 * Behavioural contract restated from the reference (file:line into /root/reference):
 *   - defaults of ltr_GetConfig(cfg, NULL)                       lighter.cpp:1179-1207
 *   - ltr_SetConfig returns 1                                    lighter.cpp:1210-1214
 *   - inputs are copied at add time, strided                     lighter.cpp:1226-1280
 *   - ltr_MeshAddPart only rejects index_count<3 && %3!=0        lighter.cpp:1237
 *   - one libc rand() is consumed per ltr_LightAdd               lighter.cpp:1300
 *   - ltr_Start is asynchronous; ltr_GetStatus is true while a stage string is set, and also
 *     before the start ("not started")                           lighter.cpp:1147-1164, lighter_int.hpp:971
 *   - outputs are owned by the scene until ltr_DestroyScene      lighter_int.hpp:982-994
 * Differences, on purpose: ltr_DestroyScene joins the bake thread first (the reference frees under
 * a running worker); errors from CUDA/NCCL are never thrown across the C ABI: after a failed bake
 * ltr_GetStatus returns 0 (done) with the stage string "failed: <reason>" instead of "finished" and
 * lightmap_count stays 0; ltrx_GetError returns the same reason.
 */
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

#include "gpu.h"
#include "scene.h"

ltr_Scene::ltr_Scene() : stage("not started"), completion(0.f)
{
    ltr_GetConfig(&config, nullptr);
    memset(&stats, 0, sizeof(stats));
    memset(nccl_id, 0, sizeof(nccl_id));
    MeshInstance *probe_container = new MeshInstance;
    memset(probe_container->matrix, 0, sizeof(probe_container->matrix));
    instances.push_back(probe_container);
}

ltr_Scene::~ltr_Scene()
{
    if (worker.joinable()) worker.join();
    bake_free(this);
    for (MeshInstance *mi : instances) delete mi;
    for (ltr_Mesh *m : meshes) delete m;
    for (ltr_WorkOutput &wo : outputs) free(wo.normals_xyzf);     /* lightmap_rgb points into output_arena */
    ltrgpu_host_free(output_arena);
}

extern "C" {

LTRBOOL ltr_DefaultSizeFunc(ltr_Config *config, const char *, size_t, const char *, size_t,
                            float computed_surface_area, float inst_importance, u32 out_size[2])
{
    /* ref: lighter.cpp:7-28 -- next pow2 of factor*sqrt(area)*importance, refuse above the cap */
    float side = config->global_size_factor * sqrtf(computed_surface_area) * inst_importance;
    if (side < 1) side = 1;
    u32 p2 = ltr_NextPowerOfTwo((u32)side);
    if (p2 > config->max_lightmap_size) return 0;
    out_size[0] = out_size[1] = p2;
    return 1;
}

ltr_Scene *ltr_CreateScene(void) { return new ltr_Scene; }

void ltr_DestroyScene(ltr_Scene *scene) { delete scene; }

void ltr_Start(ltr_Scene *scene)
{
    if (scene->started) return;          /* one bake per scene, as in the reference */
    scene->started = true;
    scene->stage.store("starting");
    scene->worker = std::thread(bake_main, scene);
}

void ltr_Abort(ltr_Scene *) { /* no-op, as in the reference (lighter.cpp:1154-1157) */ }

LTRBOOL ltr_GetStatus(ltr_Scene *scene, ltr_WorkStatus *wsout)
{
    const char *st = scene->stage.load(std::memory_order_acquire);
    wsout->completion = scene->completion.load(std::memory_order_relaxed);
    wsout->stage = st ? st : (scene->failed_stage.empty() ? "finished" : scene->failed_stage.c_str());
    return st != nullptr;
}

void ltr_Sleep(int ms)
{
    if (ms <= 0) return;
    struct timespec ts = { ms / 1000, (long)(ms % 1000) * 1000000L };
    nanosleep(&ts, nullptr);
}

void ltr_GetConfig(ltr_Config *cfg, ltr_Scene *opt_scene)
{
    if (opt_scene) { memcpy(cfg, &opt_scene->config, sizeof(*cfg)); return; }
    memset(cfg, 0, sizeof(*cfg));
    cfg->size_fn = ltr_DefaultSizeFunc;
    cfg->max_num_threads = 0x7fff;
    cfg->max_tree_memory = 128u * 1024u * 1024u;
    cfg->max_lightmap_size = 1024;
    cfg->default_width = cfg->default_height = 64;
    cfg->global_size_factor = 4;
    cfg->max_correct_dist = 0.1f;
    cfg->max_correct_angle = 60;
    cfg->ao_multiplier = 1.2f;
    cfg->ao_falloff = 1;
    cfg->ao_num_samples = 17;
    cfg->blur_size = 0.5f;
}

LTRCODE ltr_SetConfig(ltr_Scene *scene, ltr_Config *cfg)
{
    memcpy(&scene->config, cfg, sizeof(*cfg));
    return 1;
}

ltr_Mesh *ltr_CreateMesh(ltr_Scene *scene, const char *ident, size_t ident_size)
{
    ltr_Mesh *m = new ltr_Mesh;
    m->scene = scene;
    if (ident) m->ident.assign(ident, ident_size);
    scene->meshes.push_back(m);
    return m;
}

LTRBOOL ltr_MeshAddPart(ltr_Mesh *mesh, ltr_MeshPartInfo *pi)
{
    if (pi->index_count < 3 && pi->index_count % 3 != 0) return 0;
    MeshPart mp = { pi->vertex_count, (u32)mesh->vpos.size(), pi->index_count, (u32)mesh->indices.size(), pi->shadow };
    size_t nv = mesh->vpos.size() + pi->vertex_count;
    mesh->vpos.resize(nv); mesh->vnrm.resize(nv); mesh->vtex1.resize(nv); mesh->vtex2.resize(nv);
    const char *p = (const char *)pi->positions_f3, *n = (const char *)pi->normals_f3;
    const char *t1 = (const char *)pi->texcoords1_f2, *t2 = (const char *)pi->texcoords2_f2;
    for (u32 i = 0; i < pi->vertex_count; ++i) {
        memcpy(&mesh->vpos[mp.vertex_offset + i], p + (size_t)pi->stride_positions * i, 12);
        memcpy(&mesh->vnrm[mp.vertex_offset + i], n + (size_t)pi->stride_normals * i, 12);
        memcpy(&mesh->vtex1[mp.vertex_offset + i], t1 + (size_t)pi->stride_texcoords1 * i, 8);
        memcpy(&mesh->vtex2[mp.vertex_offset + i], t2 + (size_t)pi->stride_texcoords2 * i, 8);
    }
    mesh->indices.insert(mesh->indices.end(), pi->indices, pi->indices + pi->index_count);
    mesh->parts.push_back(mp);
    return 1;
}

LTRBOOL ltr_MeshAddInstance(ltr_Mesh *mesh, ltr_MeshInstanceInfo *mii)
{
    MeshInstance *mi = new MeshInstance;
    mi->mesh = mesh;
    mi->importance = mii->importance;
    mi->shadow = mii->shadow != 0;
    if (mii->ident) mi->ident.assign(mii->ident, mii->ident_size);
    memcpy(mi->matrix, mii->matrix, sizeof(mi->matrix));
    mi->lm_width = mi->lm_height = 128;
    mesh->scene->instances.push_back(mi);
    return 1;
}

void ltr_LightAdd(ltr_Scene *scene, ltr_LightInfo *li)
{
    Light L;
    L.type = li->type;
    L.position = mk3(li->position[0], li->position[1], li->position[2]);
    L.direction = norm3(mk3(li->direction[0], li->direction[1], li->direction[2]));
    L.up_direction = norm3(mk3(li->up_direction[0], li->up_direction[1], li->up_direction[2]));
    L.color = mk3(li->color_rgb[0], li->color_rgb[1], li->color_rgb[2]);
    L.range = li->range;
    L.power = li->power;
    L.light_radius = li->light_radius;
    L.shadow_sample_count = li->shadow_sample_count > 1 ? li->shadow_sample_count : 1;
    L.spot_angle_out = li->spot_angle_out;
    L.spot_angle_in = li->spot_angle_in;
    L.spot_curve = li->spot_curve;
    /* The reference draws one randf() here for its (unread) per-light sample table
     * (lighter.cpp:1300).  Consume it so the AO random offsets that follow line up; the value
     * rotates the sample spiral of the sampled-shadow extension mode. */
    L.randoff = (float)rand() / (float)RAND_MAX;
    scene->lights.push_back(L);
}

void ltr_SampleAdd(ltr_Scene *scene, ltr_SampleInfo *si)
{
    ltr_SampleInfo s = *si;
    s.out_color[0] = s.out_color[1] = s.out_color[2] = 0;
    scene->probes.push_back(s);
}

void ltr_GetWorkOutputInfo(ltr_Scene *scene, ltr_WorkOutputInfo *woutinfo)
{
    woutinfo->lightmap_count = (u32)scene->outputs.size();
    woutinfo->sample_count = (u32)scene->probes.size();
    woutinfo->samples = scene->probes.empty() ? nullptr : scene->probes.data();
}

LTRBOOL ltr_GetWorkOutput(ltr_Scene *scene, u32 which, ltr_WorkOutput *wout)
{
    if (which >= scene->outputs.size()) return 0;
    *wout = scene->outputs[which];
    return 1;
}

u32 ltr_NextPowerOfTwo(u32 x)
{
    if (x == 0) return 0;                       /* (0-1 | ...) + 1 wraps to 0 in the reference too */
    u32 p = 1;
    while (p < x && p) p <<= 1;
    return p;                                   /* x > 2^31 wraps to 0 like the bit-smear form */
}

} /* extern "C" */
