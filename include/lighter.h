/*
 * lighter.h -- public C ABI of the B200-native lightmap baker.
 *
 * This header is the drop-in boundary.  Every type and entry point below is
 * binary-compatible with the interface the reference baker publishes in its
 * own lighter.h (cited per item as ref:<first>-<last> = line range in
 * /root/reference/lighter.h).  A caller compiled against the reference header
 * links against liblighter_b200.so unchanged.  Layout equality is asserted at
 * compile time in lighter_b200/csrc/abi_check.cpp and at test time in
 * tests/test_abi.py.
 *
 * Nothing here exposes CUDA, torch or C++ types: plain pointers, sizes, ints
 * and floats only.  GPU-specific extensions (sharding, stats, stage dumps)
 * live in lighter_b200.h so this file stays a pure mirror.
 */
#ifndef LIGHTER_H_B200
#define LIGHTER_H_B200

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ref:13-21 -- export decoration; empty outside Windows DLL builds */
#ifndef LTRAPI
#  if defined(_WIN32) && !defined(LTR_STATIC)
#    ifdef LTRBUILD
#      define LTRAPI __declspec(dllexport)
#    else
#      define LTRAPI __declspec(dllimport)
#    endif
#  else
#    define LTRAPI __attribute__((visibility("default")))
#  endif
#endif

/* ref:24-25 -- both are plain int */
#define LTRBOOL int
#define LTRCODE int

/* ref:45-50 -- scalar aliases and the row-vector (v * M) matrix convention */
typedef int32_t  i32;
typedef uint32_t u32;
typedef float    ltr_VEC3[3];
typedef float    ltr_VEC4[4];
typedef ltr_VEC4 ltr_MAT4[4];

/* ref:53 */
#define LTR_VEC3_SET(v, x, y, z) do { (v)[0] = (x); (v)[1] = (y); (v)[2] = (z); } while (0)

/* ref:56-57 -- return codes (declared by the reference, unused by its code) */
#define LTRC_SUCCESS 0
#define LTRC_ATEND   1

/* ref:59-61 -- light kinds */
#define LTR_LT_POINT  1
#define LTR_LT_SPOT   2
#define LTR_LT_DIRECT 3

/* ref:64-79 -- one indexed triangle list of a mesh; arrays are copied at add time */
typedef struct ltr_MeshPartInfo {
    const void *positions_f3;
    const void *normals_f3;
    const void *texcoords1_f2;
    const void *texcoords2_f2;      /* lightmap UV set */
    u32         stride_positions;   /* byte strides */
    u32         stride_normals;
    u32         stride_texcoords1;
    u32         stride_texcoords2;
    const u32  *indices;
    u32         vertex_count;
    u32         index_count;
    int         shadow;             /* part casts shadows / occludes */
} ltr_MeshPartInfo;

/* ref:81-89 */
typedef struct ltr_MeshInstanceInfo {
    ltr_MAT4    matrix;
    float       importance;
    int         shadow;
    const char *ident;
    size_t      ident_size;
} ltr_MeshInstanceInfo;

/* ref:91-106 */
typedef struct ltr_LightInfo {
    u32      type;
    ltr_VEC3 position;
    ltr_VEC3 direction;
    ltr_VEC3 up_direction;
    ltr_VEC3 color_rgb;
    float    range;
    float    power;
    float    light_radius;
    int      shadow_sample_count;
    float    spot_angle_out;
    float    spot_angle_in;
    float    spot_curve;
} ltr_LightInfo;

/* ref:108-115 -- free-standing probe */
typedef struct ltr_SampleInfo {
    u32      id;
    ltr_VEC3 position;
    ltr_VEC3 normal;
    ltr_VEC3 out_color;
} ltr_SampleInfo;

/* ref:117-133 -- argument of the radiosity material callback */
typedef struct ltr_SampleRequest {
    ltr_VEC3    position;
    ltr_VEC3    normal;
    float       tex0u, tex0v;
    float       tex1u, tex1v;
    uint32_t    part_id;

    const char *mesh_ident;
    size_t      mesh_ident_size;
    const char *inst_ident;
    size_t      inst_ident_size;

    ltr_VEC3    out_diffuse_color;
    ltr_VEC3    out_emissive_color;
} ltr_SampleRequest;

/* ref:135-189 -- copied whole by ltr_GetConfig / ltr_SetConfig */
typedef struct ltr_Config ltr_Config;
struct ltr_Config {
    void *userdata;
    /* lightmap sizing hook; return 0 to fall back to default_width/height */
    LTRBOOL (*size_fn)(ltr_Config *config,
                       const char *mesh_ident, size_t mesh_ident_size,
                       const char *inst_ident, size_t inst_ident_size,
                       float computed_surface_area, float inst_importance,
                       u32 out_size[2]);
    int      max_num_threads;
    size_t   max_tree_memory;
    u32      max_lightmap_size;
    u32      default_width;
    u32      default_height;
    float    global_size_factor;
    float    max_correct_dist;
    float    max_correct_angle;
    ltr_VEC3 clear_color;
    ltr_VEC3 ambient_color;
    /* radiosity */
    int      bounce_count;
    LTRBOOL (*sample_fn)(ltr_Config *config, ltr_SampleRequest *req);
    /* ambient occlusion */
    float    ao_distance;
    float    ao_multiplier;
    float    ao_falloff;
    float    ao_effect;
    float    ao_divergence;
    ltr_VEC3 ao_color_rgb;
    int      ao_num_samples;
    /* post-process */
    float    blur_size;
    int      ds2x;
    int      generate_normalmap_data;
};

/* ref:191-201 */
LTRAPI LTRBOOL ltr_DefaultSizeFunc(ltr_Config *config,
                                   const char *mesh_ident, size_t mesh_ident_size,
                                   const char *inst_ident, size_t inst_ident_size,
                                   float computed_surface_area, float inst_importance,
                                   u32 out_size[2]);

/* ref:203-209 */
typedef struct ltr_WorkOutputInfo {
    u32             lightmap_count;
    u32             sample_count;
    ltr_SampleInfo *samples;
} ltr_WorkOutputInfo;

/* ref:211-223 -- buffers are owned by the scene until ltr_DestroyScene */
typedef struct ltr_WorkOutput {
    u32         uid;
    const char *mesh_ident;
    size_t      mesh_ident_size;
    const char *inst_ident;
    size_t      inst_ident_size;
    float      *lightmap_rgb;   /* width*height*3 */
    float      *normals_xyzf;   /* width*height*4 or NULL */
    u32         width;
    u32         height;
} ltr_WorkOutput;

/* ref:225-226 */
typedef struct ltr_Mesh  ltr_Mesh;
typedef struct ltr_Scene ltr_Scene;

/* ref:228-233 */
typedef struct ltr_WorkStatus {
    float       completion;
    const char *stage;
} ltr_WorkStatus;

/* ref:237-244 -- scene lifetime, asynchronous bake, polling */
LTRAPI ltr_Scene *ltr_CreateScene(void);
LTRAPI void       ltr_DestroyScene(ltr_Scene *scene);
LTRAPI void       ltr_Start(ltr_Scene *scene);
LTRAPI void       ltr_Abort(ltr_Scene *scene);
LTRAPI LTRBOOL    ltr_GetStatus(ltr_Scene *scene, ltr_WorkStatus *wsout);
LTRAPI void       ltr_Sleep(int ms);
LTRAPI void       ltr_GetConfig(ltr_Config *cfg, ltr_Scene *opt_scene);
LTRAPI LTRCODE    ltr_SetConfig(ltr_Scene *scene, ltr_Config *cfg);

/* ref:247-251 -- scene input */
LTRAPI ltr_Mesh  *ltr_CreateMesh(ltr_Scene *scene, const char *ident, size_t ident_size);
LTRAPI LTRBOOL    ltr_MeshAddPart(ltr_Mesh *mesh, ltr_MeshPartInfo *mpinfo);
LTRAPI LTRBOOL    ltr_MeshAddInstance(ltr_Mesh *mesh, ltr_MeshInstanceInfo *mii);
LTRAPI void       ltr_LightAdd(ltr_Scene *scene, ltr_LightInfo *li);
LTRAPI void       ltr_SampleAdd(ltr_Scene *scene, ltr_SampleInfo *si);

/* ref:254-255 -- readback */
LTRAPI void       ltr_GetWorkOutputInfo(ltr_Scene *scene, ltr_WorkOutputInfo *woutinfo);
LTRAPI LTRBOOL    ltr_GetWorkOutput(ltr_Scene *scene, u32 which, ltr_WorkOutput *wout);

/* ref:258 */
LTRAPI u32        ltr_NextPowerOfTwo(u32 x);

#ifdef __cplusplus
}
#endif

#endif /* LIGHTER_H_B200 */
