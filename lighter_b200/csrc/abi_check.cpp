/*
 * abi_check.cpp -- compile-time proof that include/lighter.h lays its structs out exactly as the
 * reference's lighter.h does on x86-64 (sizes/offsets below were printed by a probe compiled
 * against /root/reference/lighter.h:64-233; tests/test_abi.py re-checks them through ctypes).
 */
#include <stddef.h>
#include "lighter.h"

#define SZ(T, n) static_assert(sizeof(T) == n, "sizeof(" #T ") differs from the reference ABI")
#define OFF(T, f, n) static_assert(offsetof(T, f) == n, "offsetof(" #T "," #f ") differs from the reference ABI")

SZ(ltr_MeshPartInfo, 72);
OFF(ltr_MeshPartInfo, positions_f3, 0);
OFF(ltr_MeshPartInfo, texcoords2_f2, 24);
OFF(ltr_MeshPartInfo, stride_positions, 32);
OFF(ltr_MeshPartInfo, indices, 48);
OFF(ltr_MeshPartInfo, vertex_count, 56);
OFF(ltr_MeshPartInfo, index_count, 60);
OFF(ltr_MeshPartInfo, shadow, 64);
SZ(ltr_MeshInstanceInfo, 88);
OFF(ltr_MeshInstanceInfo, importance, 64);
OFF(ltr_MeshInstanceInfo, shadow, 68);
OFF(ltr_MeshInstanceInfo, ident, 72);
OFF(ltr_MeshInstanceInfo, ident_size, 80);
SZ(ltr_LightInfo, 80);
OFF(ltr_LightInfo, position, 4);
OFF(ltr_LightInfo, color_rgb, 40);
OFF(ltr_LightInfo, range, 52);
OFF(ltr_LightInfo, shadow_sample_count, 64);
OFF(ltr_LightInfo, spot_curve, 76);
SZ(ltr_SampleInfo, 40);
OFF(ltr_SampleInfo, out_color, 28);
SZ(ltr_SampleRequest, 104);
OFF(ltr_SampleRequest, tex0u, 24);
OFF(ltr_SampleRequest, part_id, 40);
OFF(ltr_SampleRequest, mesh_ident, 48);
OFF(ltr_SampleRequest, inst_ident_size, 72);
OFF(ltr_SampleRequest, out_diffuse_color, 80);
OFF(ltr_SampleRequest, out_emissive_color, 92);
SZ(ltr_Config, 144);
OFF(ltr_Config, size_fn, 8);
OFF(ltr_Config, max_num_threads, 16);
OFF(ltr_Config, max_tree_memory, 24);
OFF(ltr_Config, max_lightmap_size, 32);
OFF(ltr_Config, global_size_factor, 44);
OFF(ltr_Config, clear_color, 56);
OFF(ltr_Config, ambient_color, 68);
OFF(ltr_Config, bounce_count, 80);
OFF(ltr_Config, sample_fn, 88);
OFF(ltr_Config, ao_distance, 96);
OFF(ltr_Config, ao_color_rgb, 116);
OFF(ltr_Config, ao_num_samples, 128);
OFF(ltr_Config, blur_size, 132);
OFF(ltr_Config, ds2x, 136);
OFF(ltr_Config, generate_normalmap_data, 140);
SZ(ltr_WorkOutputInfo, 16);
OFF(ltr_WorkOutputInfo, samples, 8);
SZ(ltr_WorkOutput, 64);
OFF(ltr_WorkOutput, mesh_ident, 8);
OFF(ltr_WorkOutput, lightmap_rgb, 40);
OFF(ltr_WorkOutput, normals_xyzf, 48);
OFF(ltr_WorkOutput, width, 56);
OFF(ltr_WorkOutput, height, 60);
SZ(ltr_WorkStatus, 16);
OFF(ltr_WorkStatus, stage, 8);

extern "C" LTRAPI int ltrx_abi_checked(void) { return 1; }
