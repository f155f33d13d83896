/*
 * oracle/pmvs_oracle.h -- TEST INFRASTRUCTURE.  CPU restatement (plain C) of the PMVS patch-optimisation
 * hot path of /root/reference, used ONLY as the checker by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg.  The product path (cmvs-pmvs_b200/csrc) never links or calls it.
 *
 * Pinning: checked against the reference's own objects (oracle/_ref/libpmvs_ref.so, built from
 * /root/reference by oracle/Makefile) and against tests/golden/ vectors generated from them
 * (tests/golden/make_golden.py).  The optimiser is oracle/nm3.h on both sides: PARITY UNPINNED for
 * optimiser iterates (nlopt 2.6.1 LN_BOBYQA is not vendored by the reference and absent here).
 */
#ifndef PMVS_ORACLE_H
#define PMVS_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pmvso_ctx pmvso_ctx;

/* option values as in /root/reference/source/pmvs/option.cpp:10-28 and findMatch.cpp:30-106 */
pmvso_ctx* pmvso_create(int num, int tnum, int level, int csize, int wsize, int min_image_num,
                        float threshold, float max_angle_deg);
void pmvso_destroy(pmvso_ctx* c);
/* P: 3x4 row-major float32 as parsed from txt/%08d.txt (CONTOUR) */
void pmvso_set_camera(pmvso_ctx* c, int index, const float* P);
/* rgb: interleaved uint8, builds level+3 pyramid levels (image.cpp:228-325) */
void pmvso_set_image(pmvso_ctx* c, int index, int w, int h, const unsigned char* rgb);
/* visdata2[index] = list; default (never set) = all other images */
void pmvso_set_visdata2(pmvso_ctx* c, int index, const int* list, int n);
void pmvso_set_thresholds(pmvso_ctx* c, float ncc, float ncc_before);
void pmvso_set_xtol(pmvso_ctx* c, double xtol, double step, int maxeval);

void pmvso_image_dims(const pmvso_ctx* c, int index, int level, int* w, int* h);
void pmvso_image_bytes(const pmvso_ctx* c, int index, int level, unsigned char* out);
void pmvso_camera(const pmvso_ctx* c, int index, int level, float* P, float* centre, float* oaxis,
                  float* xaxis, float* yaxis, float* zaxis, float* ipscale);

void pmvso_project(const pmvso_ctx* c, int index, const float* coord, int level, float* out3);
float pmvso_get_unit(const pmvso_ctx* c, int index, const float* coord);
void pmvso_get_color(const pmvso_ctx* c, int index, float x, float y, int level, float* rgb);
void pmvso_get_paxes(const pmvso_ctx* c, int index, const float* coord, const float* normal, float* px, float* py);
/* returns 0 ok / 1 rejected; *newlevel gets the pyramid level that was sampled (or -1) */
int pmvso_grab_tex(const pmvso_ctx* c, const float* coord, const float* normal, int ref, int index,
                   float* tex, int* newlevel);
void pmvso_normalize(float* tex, int n);
float pmvso_dot(const float* a, const float* b, int n);

void pmvso_set_scales(const pmvso_ctx* c, const float* coord, const int* images, int n, float* dscale, float* ascale);
void pmvso_encode(const pmvso_ctx* c, const float* coord, const float* normal, const int* images, int n, float dscale, double* x);
void pmvso_decode(const pmvso_ctx* c, const float* coord, const float* normal, const int* images, int n, float dscale,
                  const double* x, float* ocoord, float* onormal);
double pmvso_my_f(const pmvso_ctx* c, const float* coord, const float* normal, const int* images, int n,
                  float dscale, const double* x);
double pmvso_compute_incc(const pmvso_ctx* c, const float* coord, const float* normal, const int* images, int n, int robust);
void pmvso_set_inccs(const pmvso_ctx* c, const float* coord, const float* normal, const int* images, int n, int robust, float* out);
void pmvso_set_inccs_matrix(const pmvso_ctx* c, const float* coord, const float* normal, const int* images, int n, int robust, float* out);
int pmvso_refine(const pmvso_ctx* c, float* coord, float* normal, const int* images, int n, float dscale, float* ncc, int* evals);
/* OpenMP-free batched loop, `threads` pthreads; returns seconds */
double pmvso_refine_batch(const pmvso_ctx* c, int P, int V, float* coords, float* normals, const int* images,
                          const float* dscales, float* nccs, int* evals, unsigned char* ok, int threads);

int pmvso_pre_process(const pmvso_ctx* c, const float* coord, const float* normal, int* images, int* n, int cap,
                      float* dscale, float* ascale);
/* postProcess at _depth == 0 (no depth maps yet): image-set selection, grids, timages, score2 */
int pmvso_post_process(const pmvso_ctx* c, const float* coord, const float* normal, float ncc, int* images, int* n,
                       int cap, int* grids, int* timages, float* tmp);

/* ---- filter stage: depth maps, visibility, gains (source/pmvs/filter.cpp, patchOrganizerS.cpp) ----------
 * The patch table is handed over as arrays (CSR for the per-patch image lists); indexes into it are what the
 * reference calls CPatch::_id after collectPatches(). */
void pmvso_set_depth(pmvso_ctx* c, int depth);                       /* CFindMatch::_depth */
void pmvso_store_set(pmvso_ctx* c, int P, const float* coords, const float* normals, const float* ncc, const float* dscale,
                     const int* img_off, const int* images, const int* grids,
                     const int* vimg_off, const int* vimages, const int* vgrids, const int* timages);
void pmvso_grid_dims(const pmvso_ctx* c, int image, int* gw, int* gh);
void pmvso_build_depth_maps(pmvso_ctx* c);                           /* CFilter::setDepthMaps, filter.cpp:667-732 */
void pmvso_get_depth_map(const pmvso_ctx* c, int image, int* out);   /* patch id per cell, -1 = empty */
int pmvso_is_visible(const pmvso_ctx* c, const float* coord, const float* normal, int image, int ix, int iy, float strict);
/* CPatchOrganizerS::setVImagesVGrids (patchOrganizerS.cpp:420-450) for store patch k with an empty _vimages */
int pmvso_set_vimages(const pmvso_ctx* c, int k, int* vimages, int* vgrids, int cap);
/* CFilter::filterExactThread's test (filter.cpp:315-343) for store patch k in (image, ix, iy): 1 = keep */
int pmvso_filter_exact_safe(const pmvso_ctx* c, int k, int image, int ix, int iy);
int pmvso_is_neighbor(const pmvso_ctx* c, int a, int b, float thr);  /* CFindMatch::isNeighbor, findMatch.cpp:120-149 */
float pmvso_compute_gain(const pmvso_ctx* c, int k);
/* CExpand::computeRadius (expand.cpp:182-198), CPatchOrganizerS::findNeighbors (patchOrganizerS.cpp:528-651),
 * CExpand::findEmptyBlocks (expand.cpp:108-180), CFilter::filterNeighborThread + filterQuad (filter.cpp:357-462) */
/* CDetectFeatures (detectFeatures.cpp:50-125): CHarris::run + CDifferenceOfGaussians::run on the working-level image;
 * xy float[2*cap], resp float[cap], types int[cap] (0 Harris, 1 DoG); returns the number of features */
int pmvso_detect_features(const pmvso_ctx* c, int index, int gspeedup, float* xy, float* resp, int* types, int cap);
float pmvso_compute_radius(const pmvso_ctx* c, int k);
int pmvso_find_neighbors(const pmvso_ctx* c, int k, float scale, int margin, int skipvis, int* out, int cap);
int pmvso_find_empty_blocks(const pmvso_ctx* c, int k, float* radius_out);
int pmvso_filter_neighbor(const pmvso_ctx* c, int k, float quad, float* residual_out, int* ncount_out);
int pmvso_check(const pmvso_ctx* c, int k, float quad, float* gain_out);   /* COptim::check, optim.cpp:363-383 */                 /* CFilter::computeGain, filter.cpp:88-146 */

#ifdef __cplusplus
}
#endif
#endif
