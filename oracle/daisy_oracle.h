/*
 * daisy_oracle.h -- CPU restatement of the DaisyRiot form-factor + radiosity hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * Every function cites the reference file:line it restates ("VS/" = "visual studio/").
 * Arithmetic follows the reference's FP32 operation order with NO fused multiply-add
 * (compile with -ffp-contract=off), so that results are reproducible bit for bit.
 *
 * PARITY STATUS
 *   - unoccluded form factors, ray generation, matrix assembly, gather: pinned against the
 *     reference's own sources compiled from /root/reference (oracle/ref_build -> oracle/_ref).
 *   - ray/triangle closest hit (what OptiX Prime 4.1.1 did, closed source, not in the
 *     checkout): PARITY UNPINNED.  The oracle defines it as the watertight test of
 *     Woop/Benthin/Wald (JCGT 2013) in FP32 with a (t, triangleId) lexicographic minimum;
 *     pinned only by analytic known-answer tests.
 */
#ifndef DAISY_ORACLE_H
#define DAISY_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    const float *vertices; /* nv x 3  (glm::vec3 packed)                 VS/MeshS.h:15 */
    int nv;
    const float *normals;  /* nn x 3                                      VS/MeshS.h:16 */
    int nn;
    const int *tri;        /* ntri x 6 {v0,v1,v2,n0,n1,n2}                VS/Vertex.h:11-14 */
    int ntri;
} orc_mesh;

typedef struct { float t; int triangleId; float u, v; } orc_hit; /* VS/optix_functionality.h:10-14 */

/* variant: 0 = device path (double pi, VS/parallellism.cu:197-207), 1 = host path (float pi, VS/triangle_math.cpp:49-58) */
enum { ORC_FF_DEVICE = 0, ORC_FF_HOST = 1 };

float orc_surface3(const float *a, const float *b, const float *c);          /* VS/triangle_math.cpp:31-35 */
float orc_surface_tri(const orc_mesh *m, int tri);                            /* VS/triangle_math.cpp:41-46 */
void  orc_centre_tri(const orc_mesh *m, int tri, float out[3]);               /* VS/triangle_math.cpp:16-21 */
void  orc_avg_normal(const orc_mesh *m, int tri, float out[3]);               /* VS/triangle_math.cpp:23-29 */
void  orc_divide4(const orc_mesh *m, int tri, float out[4][3][3]);            /* VS/triangle_math.cpp:60-74 */
void  orc_uv2xyz(const orc_mesh *m, int tri, float u, float v, float out[3]); /* VS/triangle_math.cpp:3-9 */
float orc_point_ff(const float opos[3], const float onrm[3], const float dpos[3], const float dnrm[3],
                   float surface, int variant);                               /* VS/parallellism.cu:197-207 */
float orc_p2p_ff(const orc_mesh *m, int origin, int dest, int variant);       /* VS/parallellism.cu:113-151 */

/* dense unoccluded rows: out[(r-row0)*ntri + c], diagonal included (VS/parallellism.cu:91-111) */
void orc_unoccluded_rows(const orc_mesh *m, int row0, int row1, int variant, float *out);

/* ray for sample i of pair (origin->dest): ray6 = {ox,oy,oz,dx,dy,dz}   VS/OptixPrimeFunctionality.cpp:191-196 */
void orc_pair_ray(const orc_mesh *m, int origin, int dest, float u, float v, float ray6[6]);

/* closest hit ------------------------------------------------------------------------------- */
/* single ray vs single triangle (watertight); returns 1 on hit with t>0 */
int orc_ray_tri(const float ray6[6], const float *a, const float *b, const float *c, float *t, float *u, float *v);

typedef struct orc_bvh orc_bvh;
orc_bvh *orc_bvh_build(const orc_mesh *m);
void orc_bvh_free(orc_bvh *b);
/* VS/OptixPrimeFunctionality.cpp:66-81 (RTP_QUERY_TYPE_CLOSEST, HIT_T_TRIID_U_V; miss => t=-1,id=-1) */
void orc_query_closest(const orc_mesh *m, const orc_bvh *b, int n, const float *rays6, orc_hit *hits);
void orc_query_closest_brute(const orc_mesh *m, int n, const float *rays6, orc_hit *hits);

/* visibility + matrix assembly ------------------------------------------------------------- */
/* Rows [row0,row1) of RadMat as a dense float matrix (VS/OptixPrimeFunctionality.cpp:6-34,169-242).
 * uv: S x 2 sample pattern (S<=64).  F_out: (row1-row0) x ntri or NULL.  masks_out: same shape,
 * bit i = sample i saw the destination, for the pair oriented min(r,c)->max(r,c), or NULL.
 * brute != 0 uses orc_query_closest_brute.  Returns number of rays cast (pairs counted once per row).
 * reciprocity != 0 restates calculateRadiosityMatrix (VS/OptixPrimeFunctionality.cpp:311-366). */
int64_t orc_radmat_rows(const orc_mesh *m, const orc_bvh *b, const float *uv, int S, int row0, int row1,
                        int variant, int reciprocity, int brute, float *F_out, uint64_t *masks_out,
                        int nthreads);

int64_t orc_radmat_rowlist(const orc_mesh *m, const orc_bvh *b, const float *uv, int S, const int *rows, int nrows,
                           int variant, int reciprocity, int brute, float *F_out, uint64_t *masks_out, int nthreads);

/* rays of the listed rows whose closest hit distance is shared bit for bit by two different triangles (parity report) */
int64_t orc_count_ties(const orc_mesh *m, const orc_bvh *b, const float *uv, int S, const int *rows, int nrows, int variant, int nthreads);

/* upper-triangle form (c > r only): F_rc = RadMat(r,c), F_cr = RadMat(c,r), masks; direct variant */
int64_t orc_radmat_upper(const orc_mesh *m, const orc_bvh *b, const float *uv, int S, int row0, int row1,
                         int variant, float *F_rc, float *F_cr, uint64_t *masks_out, int nthreads);

/* gather ----------------------------------------------------------------------------------- */
/* one pass of VS/Lightning.h:196-226 on a dense row-major F (N x ldF):
 *   bounced_k = F * res_k ; res'[:,p] = M[mat[p]] * bounced[:,p] ; B_k += res'_k
 * res,B: K x N band-major.  M: nmat x K x K column-major (Eigen::MatrixXf).
 * accum: 0 = FP32 sequential ascending column (Eigen/src/SparseCore/SparseDenseProduct.h:197-208),
 *        1 = FP64 accumulate, rounded to FP32 once per dot product.
 * band_sums[K] receives sum_p res'_k[p] (double accumulation). */
void orc_gather_pass(const float *F, int64_t ldF, int N, int K, float *res, float *B,
                     const float *M, const int *mat_idx, int accum, double *band_sums, int nthreads);

int orc_num_procs(void);

#ifdef __cplusplus
}
#endif
#endif
