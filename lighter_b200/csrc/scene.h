/*
 * scene.h -- host-side scene object behind the opaque ltr_Scene / ltr_Mesh handles.
 *
 * Mirrors what the reference keeps per scene (ref: lighter_int.hpp:837-1051) only as far as the
 * C API needs: copied-in mesh parts, instances, lights, probes, config, outputs and the
 * stage/completion pair polled by ltr_GetStatus.  All bake state lives on the GPU (gpu.h).
 */
#pragma once
#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include "lighter.h"
#include "lighter_b200.h"
#include "vmath.h"

struct V2 { float x, y; };

struct MeshPart {
    u32 vertex_count, vertex_offset, index_count, index_offset;
    int shadow;
};

struct ltr_Mesh {
    ltr_Scene *scene;
    std::string ident;
    std::vector<V3> vpos, vnrm;
    std::vector<V2> vtex1, vtex2;
    std::vector<u32> indices;
    std::vector<MeshPart> parts;
};

struct MeshInstance {
    ltr_Mesh *mesh = nullptr;       /* null for instance 0, the probe container */
    std::string ident;
    float importance = 0;
    float matrix[16];
    bool shadow = false;
    u32 lm_width = 0, lm_height = 0;
};

struct Light {
    u32 type;
    V3 position, direction, up_direction, color;
    float range, power, light_radius;
    int shadow_sample_count;
    float spot_angle_out, spot_angle_in, spot_curve;
    float randoff;             /* the randf() the reference draws per ltr_LightAdd (lighter.cpp:1300) */
};

struct ltr_Scene {
    ltr_Scene();
    ~ltr_Scene();

    ltr_Config config;
    std::vector<ltr_Mesh *> meshes;
    std::vector<MeshInstance *> instances;     /* [0] = probe container (ref: lighter_int.hpp:971-981) */
    std::vector<Light> lights;
    std::vector<ltr_SampleInfo> probes;
    std::vector<ltr_WorkOutput> outputs;
    float *output_arena = nullptr;             /* page-locked block all lightmap_rgb pointers point into (one D2H copy) */

    /* status (ref: lighter_int.hpp:1045-1046, lighter.cpp:1159-1164) */
    std::atomic<const char *> stage;
    std::atomic<float> completion;
    std::thread worker;
    bool started = false;

    /* sharding: this process bakes lumels [rank/world) of the global lumel array */
    int rank = 0, world = 1;
    unsigned char nccl_id[128];
    bool have_nccl_id = false;
    int device = -1;                          /* CUDA device ordinal; -1 = current */

    /* diagnostics */
    ltrx_Stats stats;
    std::string error;                        /* first fatal error of the bake, "" if none */
    std::string failed_stage;                 /* "failed: <error>": the stage string ltr_GetStatus hands out after a failed bake */
    int shadow_mode = 0;                      /* 0 = reference distance march, 1 = sampled any-hit shadow rays (ltrx_SetShadowMode) */
    int output_root_only = 0;                 /* sharded bake: only rank 0 reads the lightmaps back (ltrx_SetOutputRoot); the others report none */
    int keep_debug = 0;                       /* keep stage arrays for ltrx_Get* */
    struct Bake *bake = nullptr;              /* pipeline state incl. device buffers (bake.cpp) */
};

void bake_main(ltr_Scene *S);                 /* bake.cpp: the stage pipeline, runs on S->worker */
void bake_free(ltr_Scene *S);
