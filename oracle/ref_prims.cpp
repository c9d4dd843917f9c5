/*
 * ref_prims.cpp -- TEST INFRASTRUCTURE.  extern "C" shims onto the reference's OWN primitive
 * functions so tests can call them one at a time (several are file-static in the reference, so
 * the translation unit is pulled in by #include from where it lies; no reference source is copied
 * into this repository).  Built by oracle/Makefile into oracle/_ref/libref_prims.so with
 * -I/root/reference.  Used only by tests/ to pin the C restatement (oracle/lighter_oracle.c) and to
 * check device kernels primitive-by-primitive.
 */
#include "lighter_math.cpp"   /* resolved through -I$(REF): /root/reference/lighter_math.cpp */

static inline Vec3 v3(const float *p) { return V3(p[0], p[1], p[2]); }
static inline void st3(float *o, const Vec3 &v) { o[0] = v.x; o[1] = v.y; o[2] = v.z; }

extern "C" {

/* lighter_math.cpp:875-913 */
float refp_point_tri_distance(const float *p, const float *t)
{ return PointTriangleDistance(v3(p), v3(t), v3(t + 3), v3(t + 6)); }

/* lighter_math.cpp:915-952 */
int refp_point_proj_on_tri(const float *p, const float *t)
{ return PointProjOnTriangle(v3(p), v3(t), v3(t + 3), v3(t + 6)) ? 1 : 0; }

/* lighter_math.cpp:315-354 */
float refp_seg_tri(const float *l1, const float *l2, const float *t)
{ return IntersectLineSegmentTriangle(v3(l1), v3(l2), v3(t), v3(t + 3), v3(t + 6)); }

/* lighter_math.cpp:618-650 with the ray set-up of lighter_int.hpp:680-702 */
int refp_ray_aabb(const float *r0, const float *r1, const float *bbmin, const float *bbmax)
{
    BaseRayQuery q;
    q.SetRay(v3(r0), v3(r1));
    return RayAABBTest(q.ray_origin, q._ray_inv_dir, q.ray_len, v3(bbmin), v3(bbmax)) ? 1 : 0;
}

/* lighter_int.hpp:456-471 */
void refp_spiral_dir(const float *dir, float randoff, int i, int n, float *out)
{ st3(out, Vec3::CreateSpiralDirVector(v3(dir), randoff, i, n)); }

/* lighter_math.cpp:150-181 */
float refp_triangle_area3(const float *t) { return TriangleArea(v3(t), v3(t + 3), v3(t + 6)); }
float refp_sample_area(const float *uv, const float *t)
{ return CalculateSampleArea(V2(uv[0], uv[1]), V2(uv[2], uv[3]), V2(uv[4], uv[5]), v3(t), v3(t + 3), v3(t + 6)); }

/* lighter_math.cpp:187-198 */
void refp_transform(const float *m16, const float *pos, const float *nrm, int n, float *opos, float *onrm)
{
    Mat4 M;
    memcpy(M.a, m16, 64);
    std::vector<Vec3> ip(n), in(n), op(n), on(n);
    for (int i = 0; i < n; ++i) { ip[i] = v3(pos + 3 * i); in[i] = v3(nrm + 3 * i); }
    TransformPositions(VDATA(op), VDATA(ip), n, M);
    TransformNormals(VDATA(on), VDATA(in), n, M);
    for (int i = 0; i < n; ++i) { st3(opos + 3 * i, op[i]); st3(onrm + 3 * i, on[i]); }
}

/* lighter_math.cpp:245-302; images are w*h arrays of 3,3,4 floats */
void refp_raster_tri(float *img1, float *img2, float *img3, int w, int h, float margin,
                     const float *p, const float *va, const float *vb, const float *vc)
{
    RasterizeTriangle2D_x2_ex((Vec3 *)img1, (Vec3 *)img2, (Vec4 *)img3, w, h, margin,
                              V2(p[0], p[1]), V2(p[2], p[3]), V2(p[4], p[5]),
                              v3(va), v3(va + 3), v3(va + 6), v3(vb), v3(vb + 3), v3(vb + 6),
                              V4(vc[0], vc[1], vc[2], vc[3]), V4(vc[4], vc[5], vc[6], vc[7]), V4(vc[8], vc[9], vc[10], vc[11]));
}

/* lighter_math.cpp:357-424 */
void refp_gauss_kernel(float *out, int ext, float radius) { Generate_Gaussian_Kernel(out, ext, radius); }
void refp_convolve_transpose(float *src, float *dst, unsigned w, unsigned h, int ext, float *kernel)
{
    std::vector<float> tmp(((w > h ? w : h) + 2 * ext) * 3);
    Convolve_Transpose(src, dst, w, h, ext, kernel, tmp.data());
}
void refp_downsample2x(float *dst, unsigned dw, unsigned dh, float *src, unsigned sw, unsigned sh)
{ Downsample2X(dst, dw, dh, src, sw, sh); }

/* TriTree (lighter_math.cpp:785-801 build, :827-871 rays, :984-989 distance, :1040-1044 offset) */
void *refp_tritree_create(const float *tris, int count)
{
    std::vector<Triangle> T(count);
    for (int i = 0; i < count; ++i) { T[i].P1 = v3(tris + 9 * i); T[i].P2 = v3(tris + 9 * i + 3); T[i].P3 = v3(tris + 9 * i + 6); }
    TriTree *tt = new TriTree;
    tt->SetTris(VDATA(T), T.size());
    return tt;
}
void refp_tritree_destroy(void *h) { delete (TriTree *)h; }
int refp_tritree_tri_count(void *h) { return (int)((TriTree *)h)->m_tris.size(); }
float refp_tritree_distance(void *h, const float *p, float dist) { return ((TriTree *)h)->GetDistance(v3(p), dist); }
int refp_tritree_anyhit(void *h, const float *a, const float *b) { return ((TriTree *)h)->IntersectRay(v3(a), v3(b)) ? 1 : 0; }
float refp_tritree_closest(void *h, const float *a, const float *b, int *tid)
{ int32_t t = -1; float r = ((TriTree *)h)->IntersectRayDist(v3(a), v3(b), &t); if (tid) *tid = t; return r; }
void refp_tritree_offset(void *h, float *P, const float *N, float dist)
{ Vec3 p = v3(P); ((TriTree *)h)->OffsetSample(p, v3(N), dist); st3(P, p); }
/* node dump: 8 words per node (min3,max3,ch,ido) then the item index stream */
int refp_tritree_node_count(void *h) { return (int)((TriTree *)h)->m_bbTree.m_nodes.size(); }
int refp_tritree_item_count(void *h) { return (int)((TriTree *)h)->m_bbTree.m_itemidx.size(); }
void refp_tritree_dump(void *h, void *nodes, int32_t *items)
{
    AABBTree &bt = ((TriTree *)h)->m_bbTree;
    memcpy(nodes, VDATA(bt.m_nodes), bt.m_nodes.size() * sizeof(AABBTree::Node));
    if (bt.m_itemidx.size()) memcpy(items, VDATA(bt.m_itemidx), bt.m_itemidx.size() * 4);
}

/* BSP closest hit + normal (lighter_math.cpp:490-533), used by the overlap correction */
void *refp_bsp_create(const float *tris, int count)
{
    std::vector<Triangle> T;
    for (int i = 0; i < count; ++i) {
        Triangle t = { v3(tris + 9 * i), v3(tris + 9 * i + 3), v3(tris + 9 * i + 6) };
        if (t.CheckIsUseful()) T.push_back(t);
    }
    BSPTree *b = new BSPTree;
    if (T.size()) b->SetTriangles(VDATA(T), T.size());
    return b;
}
void refp_bsp_destroy(void *h) { delete (BSPTree *)h; }
float refp_bsp_closest(void *h, const float *a, const float *b, float *nrm)
{ Vec3 n = V3(0); float r = ((BSPTree *)h)->IntersectRay(v3(a), v3(b), &n); st3(nrm, n); return r; }

} /* extern "C" */
