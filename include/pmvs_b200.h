/*
 * include/pmvs_b200.h -- C ABI of the B200-native PMVS patch-optimisation path.
 *
 * The reference has no FFI layer: its hot path is the in-process C++ contract
 *     if (preProcess(patch,id,seed)) fail; refinePatch(patch,id,100); if (postProcess(patch,id,seed)) fail;
 * (/root/reference/source/pmvs/seed.cpp:397-409, source/pmvs/expand.cpp:225-237) on top of
 * CPhotoSetS (images + cameras).  This header is the boundary a maintainer binds instead
 * (INTEGRATION.md shows the stub): one opaque context per GPU owns the image pyramids and camera
 * tables in HBM; every hot call is BATCHED over patches (structure-of-arrays, plain pointers).
 *
 * Conventions
 *   - every function returns 0 on success, a negative PMVSB_E* code otherwise; pmvsb_last_error()
 *     returns the message of the last failure on that context (never throws across the boundary);
 *   - the caller owns all host buffers, the library owns all device memory;
 *   - one context is used by one host thread at a time; multi-GPU = one context per device;
 *   - calls return after the work is complete (stream-synchronised) unless suffixed _dev, which take
 *     DEVICE pointers, enqueue on the context's stream and return immediately (pmvsb_sync to wait);
 *   - there is NO CPU fallback: a context cannot be created without a CUDA device.
 *
 * Patch batch layout (P patches):
 *   coords   float[4*P]   x y z 1            (Patch::CPatch::_coord, include/pmvs/patch.hpp:32)
 *   normals  float[4*P]   nx ny nz 0         (CPatch::_normal, patch.hpp:34)
 *   images   int32[stride*P]  image indexes, reference image first   (CPatch::_images, patch.hpp:38)
 *   nimages  int32[P] or NULL (= stride for every patch)
 *   dscales  float[P]     depth step per pixel (CPatch::_dscale, patch.hpp:72; set by setScales)
 */
#ifndef PMVS_B200_H
#define PMVS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PMVSB_OK 0
#define PMVSB_EINVAL (-1)   /* bad argument */
#define PMVSB_ECUDA (-2)    /* CUDA runtime error */
#define PMVSB_ESTATE (-3)   /* call out of order (e.g. scene not finalised) */
#define PMVSB_ENOMEM (-4)
#define PMVSB_ERANGE (-5)   /* a patch's image list outgrew its capacity (PMVSB_MAX_VIEWS): results would differ from the reference's */

#define PMVSB_MAX_TAU 8     /* views used by refinePatch / computeINCC (tau = min(2*minImageNum, num)) */
#define PMVSB_MAX_VIEWS 64  /* views per patch accepted by set_inccs / pre- and post-process kernels */
#define PMVSB_EGROW (-6)    /* a wave's message is larger than the peer mailbox slots: re-export with pmvsb_peer_needed() bytes and call again */
#define PMVSB_MAX_RANKS 16  /* ranks of one peer-memory exchange (GPUs of one NVLink domain) */
#define PMVSB_PEER_HANDLE_BYTES 64

typedef struct pmvsb_ctx pmvsb_ctx;

/* ---- context ------------------------------------------------------------------------------------
 * Replaces CFindMatch::init's option plumbing (/root/reference/source/pmvs/findMatch.cpp:30-106):
 * level, csize, wsize, minImageNum, threshold, maxAngle as in the option file
 * (source/pmvs/option.cpp:47-109); tau, nccThresholdBefore, the 60-degree angle thresholds and the
 * level+3 pyramid depth are derived exactly as the reference derives them. */
int pmvsb_create(pmvsb_ctx** out, int device, int num_images, int num_target, int level, int csize,
                 int wsize, int min_image_num, float threshold, float max_angle_deg);
int pmvsb_destroy(pmvsb_ctx* ctx);
const char* pmvsb_last_error(const pmvsb_ctx* ctx);
const char* pmvsb_version(void);
int pmvsb_device_count(void);   /* CUDA devices visible to this process (0 when there is none) */

/* Image::CCamera::init + updateCamera (source/image/camera.cpp:13-54,109-136) and
 * COptim::setAxesScales (source/pmvs/optim.cpp:43-64): P = 3x4 row-major float32 of txt/%08d.txt. */
int pmvsb_upload_camera(pmvsb_ctx* ctx, int index, const float* P);
/* Image::CImage::alloc + buildImage (source/image/image.cpp:113-180,228-325): rgb = interleaved
 * uint8 w*h*3 (what readAnyImage produces); the level+3 pyramid is built on the device. */
int pmvsb_upload_image(pmvsb_ctx* ctx, int index, int width, int height, const uint8_t* rgb);
/* Masks and edge maps (optional).  CImage::alloc reads <prefix>masks/%08d.{pgm,pbm} and <prefix>edges/%08d.{pgm,pbm}
 * (source/image/image.cpp:146-176; file names source/image/photoSetS.cpp:43-49): gray = uint8[width*height] as stored in the
 * file, at the size of the image, uploaded AFTER its image.  which = 0: mask, a pixel is inside when 127 < v;
 * which = 1: edge map, inside when 1 < v.  The "any of the 2x2 parents" pyramid of buildMask / buildEdge
 * (image.cpp:326-393) is built on the device down to the working level, the only level the path reads
 * (getMask / getEdge are always called with CFindMatch::_level). */
int pmvsb_upload_mask(pmvsb_ctx* ctx, int index, int which, int width, int height, const uint8_t* gray);
/* Option `setEdge` (CPhotoSetS::setEdge, source/image/photoSetS.cpp:91-95 -> CImage::setEdge, image.cpp:407-471): the edge map
 * of EVERY image recomputed from its level-0 colour gradients (replaces uploaded edge files, as in findMatch.cpp:74-76). */
int pmvsb_set_edge(pmvsb_ctx* ctx, float threshold);
/* Option `useBound`: SOption::_bindexes (source/pmvs/option.cpp:301-325), image indexes whose frames bound the reconstruction
 * (CFindMatch::insideBimages, source/pmvs/findMatch.cpp:109-118). */
int pmvsb_set_bimages(pmvsb_ctx* ctx, const int32_t* list, int n);
/* parity hook: the working-level map of an image (CImage::getMask(level) / getEdge(level), 255 in / 0 out);
 * *present = 0 when the image has none (out untouched; out may be NULL to query) */
int pmvsb_download_mask(pmvsb_ctx* ctx, int index, int which, uint8_t* out, int* present);
/* SOption::_visdata2 (source/pmvs/option.cpp:202-299): candidate images for addImages. Default: all. */
int pmvsb_set_visdata2(pmvsb_ctx* ctx, int index, const int32_t* list, int n);
/* After all cameras and images are uploaded: builds the device tables.  Required before any batch call. */
int pmvsb_finalize_scene(pmvsb_ctx* ctx);

/* CFindMatch::_nccThreshold / _nccThresholdBefore / _depth (updateThreshold, findMatch.cpp:23-28,196,216) */
int pmvsb_set_thresholds(pmvsb_ctx* ctx, float ncc_threshold, float ncc_threshold_before);
/* Optimiser knobs replacing nlopt's set_xtol_rel / set_maxeval (optim.cpp:623-624). Defaults 1e-3, 1.0, 1000. */
int pmvsb_set_optimizer(pmvsb_ctx* ctx, double xtol, double step, int maxeval);

/* ---- parity hooks (small, synchronous, host pointers) ------------------------------------------- */
int pmvsb_image_dims(pmvsb_ctx* ctx, int index, int level, int* width, int* height);
/* pyramid level back as interleaved RGB (CImage::getImage, include/image/image.hpp:197-209) */
int pmvsb_download_image(pmvsb_ctx* ctx, int index, int level, uint8_t* rgb);
/* derived camera constants: P at `level` (12), centre (4), oaxis (4), x/y/z axes (3 each), ipscale */
int pmvsb_get_camera(pmvsb_ctx* ctx, int index, int level, float* P, float* centre, float* oaxis,
                     float* xaxis, float* yaxis, float* zaxis, float* ipscale);
/* CPhotoSetS::project (include/image/photoSetS.hpp:81-83): N points, image index per point -> float[3*N] */
int pmvsb_project_batch(pmvsb_ctx* ctx, int n, const float* coords, const int32_t* image, int level, float* out);
/* COptim::grabTex as my_f calls it (optim.cpp:815-863): for patch p and its k-th image
 * (k < min(stride, nimages[p])) writes flag[p*stride+k] (0 grabbed / 1 rejected),
 * newlevel[p*stride+k] and the raw wsize*wsize*3 texture. */
int pmvsb_grab_tex_batch(pmvsb_ctx* ctx, int P, int stride, const float* coords, const float* normals,
                         const int32_t* images, const int32_t* nimages, float* tex, int32_t* flag,
                         int32_t* newlevel);
/* COptim::my_f (optim.cpp:507-578): x = double[3*P] (depth/dscale, angle1/ascale, angle2/ascale),
 * coords/normals = the START patch (centre, ray, reference axes).  f = double[P]. */
int pmvsb_eval_objective_batch(pmvsb_ctx* ctx, int P, int stride, const float* coords, const float* normals,
                               const int32_t* images, const int32_t* nimages, const float* dscales,
                               const double* x, double* f);
/* COptim::computeINCC (optim.cpp:865-938) with setWeightsT weights: out = double[P] */
int pmvsb_compute_incc_batch(pmvsb_ctx* ctx, int P, int stride, const float* coords, const float* normals,
                             const int32_t* images, const int32_t* nimages, int robust, double* out);
/* COptim::setINCCs vector form (optim.cpp:709-744): out = float[stride*P] */
int pmvsb_set_inccs_batch(pmvsb_ctx* ctx, int P, int stride, const float* coords, const float* normals,
                          const int32_t* images, const int32_t* nimages, int robust, float* out);
/* CPatchOrganizerS::setScales (source/pmvs/patchOrganizerS.cpp:663-684): dscale, ascale = float[P] */
int pmvsb_set_scales_batch(pmvsb_ctx* ctx, int P, int stride, const float* coords, const int32_t* images,
                           const int32_t* nimages, float* dscale, float* ascale);

/* `_pss.getMask(coord, _level) == 0 || insideBimages(coord) == 0` as expandSub, collectCandidates and postProcess test it
 * (source/pmvs/expand.cpp:212, seed.cpp:314, optim.cpp:153) for n points (coords float[4n]): inside[k] = 1 when the point passes
 * (CPhotoSetS::getMask walks every image of the scene, include/image/photoSetS.hpp:109-116). */
int pmvsb_mask_gate_batch(pmvsb_ctx* ctx, int n, const float* coords, uint8_t* inside);
/* COptim::removeImagesEdge (source/pmvs/optim.cpp:385-396), called by expandSub before preProcess (expand.cpp:219): images whose
 * edge map rejects the patch centre leave the list (order kept); images / nimages updated in place. */
int pmvsb_remove_images_edge_batch(pmvsb_ctx* ctx, int P, int stride, const float* coords, int32_t* images, int32_t* nimages);

/* ---- visible-image-set selection around the hot call (integer results: bit-exact) -----------------------
 * COptim::preProcess (optim.cpp:95-122) for a batch of candidates: addImages -> constraintImages(nccThresholdBefore)
 * -> sortImages -> setScales -> minImageNum check -> checkAngles.  images/nimages are updated in place
 * (stride = capacity per patch, at most PMVSB_MAX_VIEWS); verdict 0 = keep, 1 = reject (as the reference returns).
 * The reference's lists are unbounded: when addImages finds more images than the capacity holds, the call fails with
 * PMVSB_ERANGE instead of truncating. */
int pmvsb_pre_process_batch(pmvsb_ctx* ctx, int P, int stride, const float* coords, const float* normals,
                            int32_t* images, int32_t* nimages, float* dscale, float* ascale, int32_t* verdict);
/* COptim::postProcess (optim.cpp:150-190) at CFindMatch::_depth == 0 (seed round): addImages ->
 * constraintImages(nccThreshold) -> filterImagesByAngle -> setRefImage -> constraintImages -> setGrids, then
 * _timages and _tmp = score2.  grids = int32[2*stride*P] (cell x,y per image), ncc = refinePatch's _ncc. */
int pmvsb_post_process_batch(pmvsb_ctx* ctx, int P, int stride, const float* coords, const float* normals,
                             const float* ncc, int32_t* images, int32_t* nimages, int32_t* grids,
                             int32_t* timages, float* tmp, int32_t* verdict);

/* The reference's in-process contract for one candidate patch,
 *     if (preProcess(patch, id, seed)) fail;  refinePatch(patch, id, 100);  if (postProcess(patch, id, seed)) fail;
 * (source/pmvs/seed.cpp:397-409, expand.cpp:225-237), for a whole wave in ONE call: the stages chain on the device (survivors are
 * compacted between them, nothing returns to the host in between).  Candidates: coords / normals float[4P], image lists as CSR
 * (img_off int32[P+1], images).  postProcess includes setVImagesVGrids at _depth >= 1 and COptim::check at _depth >= 2
 * (pmvsb_set_depth; both read the resident table and its depth maps).  Results stay on the device until pmvsb_evaluate_fetch:
 * *accepted candidates with *entries image entries and *ventries visible-image entries in total; *refined = refinePatch calls
 * (candidates that passed preProcess). */
int pmvsb_evaluate_batch(pmvsb_ctx* ctx, int P, const float* coords, const float* normals, const int32_t* img_off, const int32_t* images, float quad,
                         int32_t* accepted, int32_t* entries, int32_t* ventries, int32_t* refined);
/* Results of the last pmvsb_evaluate_batch (any pointer may be NULL): verdict int32[P] (0 accepted, 1 rejected by preProcess,
 * 2 by postProcess); for the accepted candidates, in candidate order: index (candidate number), coords / normals float[4A],
 * scal float[4A] = (_ncc, _dscale, _ascale, _tmp), timages, _images / _grids as CSR (img_off int32[A+1], images int32[E],
 * grids int32[2E]) and _vimages / _vgrids likewise. */
int pmvsb_evaluate_fetch(pmvsb_ctx* ctx, int32_t* verdict, int32_t* index, float* coords, float* normals, float* scal, int32_t* timages,
                         int32_t* img_off, int32_t* images, int32_t* grids, int32_t* vimg_off, int32_t* vimages, int32_t* vgrids);

/* ---- filter stage: depth maps, visibility, gains ----------------------------------------------------------
 * The host keeps the cell bookkeeping (CPatchOrganizerS) and hands the current patch table over as arrays
 * (index = CPatch::_id after collectPatches(), source/pmvs/patchOrganizerS.cpp:207-236):
 *   coords, normals float[4P]; ncc, dscale float[P]; img_off int32[P+1] + images int32[E] + grids int32[2E]
 *   (CPatch::_images/_grids as CSR); vimg_off/vimages/vgrids likewise (CPatch::_vimages/_vgrids); timages int32[P]. */
int pmvsb_set_depth(pmvsb_ctx* ctx, int depth);   /* CFindMatch::_depth: 0 makes isVisible trivially true (patchOrganizerS.cpp:493) */
int pmvsb_grid_dims(pmvsb_ctx* ctx, int image, int* gwidth, int* gheight);   /* patchOrganizerS.cpp:72-77 */
int pmvsb_store_upload(pmvsb_ctx* ctx, int P, const float* coords, const float* normals, const float* ncc, const float* dscale,
                       const int32_t* img_off, const int32_t* images, const int32_t* grids,
                       const int32_t* vimg_off, const int32_t* vimages, const int32_t* vgrids, const int32_t* timages);
/* CPatchOrganizerS::addPatch for n patches committed since the last upload (patchOrganizerS.cpp:312-349): full
 * records are appended (table ids P, P+1, ...; offsets relative to img_off[0] / vimg_off[0]), the cell lists are
 * rebuilt on the device and, once depth maps exist, the new patches are merged into them (updateDepthMaps,
 * patchOrganizerS.cpp:351-381).  Every store call stays valid on the extended table. */
int pmvsb_store_append(pmvsb_ctx* ctx, int n, const float* coords, const float* normals, const float* ncc, const float* dscale,
                       const int32_t* img_off, const int32_t* images, const int32_t* grids,
                       const int32_t* vimg_off, const int32_t* vimages, const int32_t* vgrids, const int32_t* timages);
/* CFilter::setDepthMaps (source/pmvs/filter.cpp:667-732): nearest patch per cell of every target image */
int pmvsb_build_depth_maps(pmvsb_ctx* ctx);
int pmvsb_download_depth_map(pmvsb_ctx* ctx, int image, int32_t* patch_id);   /* gw*gh ids, -1 = empty */
/* CPatchOrganizerS::updateDepthMaps (patchOrganizerS.cpp:351-381) for n patches committed since the last upload:
 * their coords are appended to the table (ids P, P+1, ...) and merged into the depth maps.  Only the visibility
 * calls may be used on an appended table; gains / filterExact need a fresh pmvsb_store_upload.  Returns
 * PMVSB_ENOMEM when the table's spare capacity is exhausted (re-upload instead). */
int pmvsb_depth_maps_add(pmvsb_ctx* ctx, int n, const float* coords);
/* CPatchOrganizerS::setVImagesVGrids (patchOrganizerS.cpp:420-450) for every table patch, from an empty _vimages:
 * vimages int32[vcap*P], vgrids int32[2*vcap*P], nv int32[P] */
int pmvsb_set_vimages_store(pmvsb_ctx* ctx, int vcap, int32_t* vimages, int32_t* vgrids, int32_t* nv);
/* CFilter::setDepthMapsVGridsVPGridsAddPatchV (filter.cpp:734-783): setVImagesVGrids for every table patch written
 * into the table's own _vimages/_vgrids (additive = 0: from empty lists, 1: keep and append), _vpgrids rebuilt.
 * total = number of (patch, visible image) entries afterwards; pmvsb_store_download_vimages returns them as CSR
 * (vimg_off int32[P+1], vimages int32[total], vgrids int32[2*total]). */
int pmvsb_store_update_vimages(pmvsb_ctx* ctx, int additive, int32_t* total);
int pmvsb_store_download_vimages(pmvsb_ctx* ctx, int32_t* vimg_off, int32_t* vimages, int32_t* vgrids);
/* ---- the table reorganised on the device: a filter round without marshalling patches through the host ----------------------
 * Creation sequence numbers of table patches [first, first + n): the order of a cell's patch vector in the reference (addPatch
 * appends, patchOrganizerS.cpp:312-349).  Default after pmvsb_store_upload: the table order; appended patches are younger than
 * every patch before them. */
int pmvsb_store_set_seq(pmvsb_ctx* ctx, int first, int n, const int32_t* seq);
int pmvsb_store_counts(pmvsb_ctx* ctx, int32_t* P, int32_t* entries, int32_t* ventries);
/* CPatchOrganizerS::removePatch for the patches with keep[k] == 0 (keep = NULL: none), then
 * CFilter::setDepthMapsVGridsVPGridsAddPatchV(additive) (source/pmvs/filter.cpp:734-783): the survivors renumbered in
 * collectPatches order (patchOrganizerS.cpp:207-236: first appearance in (image, cell) order, creation order inside a cell),
 * _pgrids rebuilt, depth maps (setDepthMaps), setVImagesVGrids for every patch (additive = 0 from empty _vimages), _vpgrids.
 * additive = 2: renumbering and depth maps only, _vimages kept as they are (collectPatches + setDepthMaps, what the table needs
 * when an expansion round starts, expand.cpp:42-47).
 * new_count = patches left; perm (optional, int32[new_count]) = the OLD table index of each patch of the new table. */
int pmvsb_store_rebuild(pmvsb_ctx* ctx, const uint8_t* keep, int additive, int32_t* new_count, int32_t* perm);
/* CFilter::filterExact on the table (filter.cpp:203-355): the visibility re-test of every image entry (filterExactThread), the
 * lists pruned (surviving target images in ascending image order, then the non-target ones), _timages, then setRefImage +
 * setGrids for the survivors; keep[k] = 0 for a patch left with fewer than minImageNum images or without a target image
 * (its lists are emptied; pmvsb_store_rebuild removes it).  _pgrids is rebuilt; depth maps and _vimages are untouched. */
int pmvsb_filter_exact_apply_store(pmvsb_ctx* ctx, uint8_t* keep);
/* The neighbour tests of CFilter::filterSmallGroupsSub (filter.cpp:602-665) for every table patch p: the patches q listed in
 * _pgrids / _vpgrids of the 3x3 cells around p's cell in its reference image with isNeighbor(p, q, neighbor_threshold), as CSR
 * adjacency (adj_off int32[P+1]; adj int32[*total], written when *total <= cap: call once with cap = 0 to size it). */
int pmvsb_small_group_edges_store(pmvsb_ctx* ctx, float neighbor_threshold, int32_t* adj_off, int32_t* adj, int cap, int32_t* total);
/* CFilter::filterSmallGroups (filter.cpp:524-600): keep[k] = 0 for the patches of a connected group smaller than
 * *group_threshold = max(20, P / 10000).  The neighbour tests run on the device (above); the labelling walk is sequential by
 * definition (a patch takes the label of the first walk that reaches it) and runs over that adjacency on the calling thread. */
int pmvsb_filter_small_groups_store(pmvsb_ctx* ctx, float neighbor_threshold, uint8_t* keep, int32_t* group_threshold);
/* table lists back to the host (any pointer may be NULL): seq, timages int32[P]; img_off int32[P+1]; images int32[E];
 * grids int32[2E] (P, E from pmvsb_store_counts); _vimages through pmvsb_store_download_vimages */
int pmvsb_store_download_lists(pmvsb_ctx* ctx, int32_t* seq, int32_t* timages, int32_t* img_off, int32_t* images, int32_t* grids);
/* _pgrids (visible = 0) / _vpgrids (1) of the table as the device built them: cell_off int32[cells+1] over the
 * flattened cells of the target images in image order, cell_patch int32[cell_off[cells]] (may be NULL) -- parity hook */
int pmvsb_download_cell_lists(pmvsb_ctx* ctx, int visible, int32_t* cell_off, int32_t* cell_patch);
/* CExpand::findEmptyBlocks (source/pmvs/expand.cpp:108-180) for n table patches: bit i of mask[k] is set when
 * direction i (angle 2 pi i / 6 in the ortho(normal) frame) already has a neighbour from
 * findNeighbors(patch, scale 4, margin 1) at in-plane distance in [r/6, 2.5 r]; radius[k] = r = computeRadius
 * (expand.cpp:182-198).  The caller places candidates in the clear directions. */
int pmvsb_find_empty_blocks_store(pmvsb_ctx* ctx, int n, const int32_t* ids, uint8_t* mask, float* radius);
/* CFilter::filterNeighborThread + filterQuad (filter.cpp:357-462) for every table patch: reject[k] = 1 when the
 * patch has fewer than 6 neighbours (findNeighbors scale 4, margin 2, skipvis 1) or its quadric-fit residual is
 * >= quad.  Optional outputs: residual float[P] (-1 = too few neighbours), ncount int32[P] unique neighbours,
 * overflow = patches whose neighbour set did not fit on chip (kept, never silently rejected). */
int pmvsb_filter_neighbor_store(pmvsb_ctx* ctx, float quad, uint8_t* reject, float* residual, int32_t* ncount, int32_t* overflow);
/* COptim::check (source/pmvs/optim.cpp:363-383), the last step of postProcess at _depth >= 2, for a batch of candidates that
 * are not in the table: gain[p] = CFilter::computeGain (filter.cpp:88-146, stored to CPatch::_tmp) against the table's
 * cells; reject[p] = 1 when gain < 0, or when findNeighbors(scale 4, margin 2) has more than 6 neighbours and the
 * quadric residual (filterQuad) is >= quad.  Lists as in pre/post-process (stride / vstride = capacity per patch). */
int pmvsb_check_batch(pmvsb_ctx* ctx, int P, int stride, const float* coords, const float* normals, const float* ncc, const float* dscale,
                      const int32_t* timages, const int32_t* images, const int32_t* nimages, const int32_t* grids, int vstride,
                      const int32_t* vimages, const int32_t* nv, const int32_t* vgrids, float quad, float* gain, uint8_t* reject, int32_t* overflow);
/* CFilter::filterExactThread's visibility re-test (filter.cpp:315-343): safe uint8[E], one flag per image entry */
int pmvsb_filter_exact_store(pmvsb_ctx* ctx, uint8_t* safe);
/* CFilter::filterOutsideThread / computeGain (filter.cpp:88-201): gains float[P] */
int pmvsb_compute_gains_store(pmvsb_ctx* ctx, float* gains);

/* CPatchOrganizerS::setVImagesVGrids (patchOrganizerS.cpp:420-450) for a batch of patches that are not (yet) in the
 * table -- postProcess calls it on every candidate once _depth >= 1 (optim.cpp:184-186).  Images already in
 * images[] or vimages[] count as used; newly visible target images are appended to vimages/vgrids (capacity
 * vstride per patch), nv is updated.  Uses the depth maps of the last pmvsb_build_depth_maps. */
int pmvsb_set_vimages_batch(pmvsb_ctx* ctx, int P, int stride, const float* coords, const float* normals,
                            const int32_t* images, const int32_t* nimages, int vstride, int32_t* vimages,
                            int32_t* nv, int32_t* vgrids);
/* COptim::setRefImage (optim.cpp:208-254) + setGrids for a batch (what filterExact does after pruning image
 * lists, filter.cpp:277-280): images reordered in place, grids = int32[2*stride*P]; nimages becomes 0 for a patch
 * without target images. */
int pmvsb_set_ref_image_batch(pmvsb_ctx* ctx, int P, int stride, const float* coords, const float* normals,
                              int32_t* images, int32_t* nimages, int32_t* grids);
/* vertex colour of writePLY (patchOrganizerS.cpp:713-727): mean over the patch's images of the bilinear colour at
 * its projection (level = option level), rounded; rgb = uint8[3*P] */
int pmvsb_patch_colors_batch(pmvsb_ctx* ctx, int P, int stride, const float* coords, const int32_t* images,
                             const int32_t* nimages, uint8_t* rgb);

/* ---- the hot call -------------------------------------------------------------------------------
 * COptim::refinePatch for a whole seed / expansion frontier in one launch (optim.cpp:496-502,580-658):
 * in-kernel bounded Nelder-Mead over (depth, angle1, angle2) around my_f, then the final
 * computeINCC.  coords/normals are updated in place for patches with ok=1 and left untouched otherwise
 * (optim.cpp:649-655).  ncc = 1 - unrobustincc(computeINCC(robust)), evals = objective evaluations,
 * ok = 1 iff the optimiser stopped on its x-tolerance (MAXEVAL = failure, optim.cpp:644). */
int pmvsb_refine_batch(pmvsb_ctx* ctx, int P, int stride, float* coords, float* normals,
                       const int32_t* images, const int32_t* nimages, const float* dscales,
                       float* ncc, int32_t* evals, uint8_t* ok);
/* Same with DEVICE pointers; asynchronous on the context stream. */
int pmvsb_refine_batch_dev(pmvsb_ctx* ctx, int P, int stride, float* d_coords, float* d_normals,
                           const int32_t* d_images, const int32_t* d_nimages, const float* d_dscales,
                           float* d_ncc, int32_t* d_evals, uint8_t* d_ok);
/* ---- feature detection ---------------------------------------------------------------------
 * CDetectFeatures::runThread for one image (source/pmvs/detectFeatures.cpp:77-118): CHarris::run (sigma 4,
 * source/pmvs/harris.cpp:174-240) then CDifferenceOfGaussians::run (scales 1..3, source/pmvs/dog.cpp:96-198) on the
 * working-level image of the device pyramid, at most 4 points per (2*gspeedup)^2-pixel block and detector.  Output in the
 * reference's order (Harris strongest first, then DoG strongest first): xy float[2*cap] pixel coordinates, response
 * float[cap], type int32[cap] (0 Harris, 1 DoG); *count = number of features found (may exceed cap). */
int pmvsb_detect_features(pmvsb_ctx* ctx, int index, int gspeedup, int cap, float* xy, float* response, int32_t* type, int32_t* count);

/* ---- seed candidate enumeration ------------------------------------------------------------------
 * CSeed::readPoints (source/pmvs/seed.cpp:23-36): the features of image `index` (what pmvsb_detect_features returned, or the
 * caller's own), binned by cell on upload in the order given. */
int pmvsb_set_features(pmvsb_ctx* ctx, int index, int n, const float* xy, const int32_t* type);
/* CSeed::collectCells + collectCandidates + unproject (seed.cpp:207-384) for EVERY feature of reference image `index` that sits in
 * a cell open for a patch, in one launch.  views[nviews] = COptim::collectImages' list (optim.cpp:66-93; nviews <= PMVSB_MAX_TAU);
 * blocked = uint8 per cell of every image's grid, images back to back (targets and others; cell = y * gwidth + x of
 * pmvsb_grid_dims), non-zero where CSeed::canAdd(image, x, y) is 0 (seed.cpp:325-338: mask, occupied, trial count reached).
 * Outputs, per reference feature r < *nref in (cell, feature-in-cell) order: ref_feature[r] (index into the image's feature
 * list), ref_cell[r], and its candidates [ref_start[r], ref_start[r] + ref_count[r]) in coords float[4 * ..] (the triangulated
 * point, w = 1), other_image, other_feature (index into that image's list), resp = |dist to camera `index` - dist to the other
 * camera|; a feature's candidates are in ascending resp order, ties in the reference's enumeration order.  The candidate set and
 * all values equal the reference's; see pmvs_seed.cuh on the order.  Sizing: when *nref > cap_ref or *total > cap nothing (or
 * only the per-feature arrays) is written -- call again with room. */
int pmvsb_seed_candidates(pmvsb_ctx* ctx, int index, int nviews, const int32_t* views, const uint8_t* blocked, int cap_ref, int32_t* nref,
                          int32_t* ref_feature, int32_t* ref_cell, int32_t* ref_start, int32_t* ref_count, int cap, int32_t* total,
                          float* coords, int32_t* other_image, int32_t* other_feature, float* resp);

/* ---- multi-GPU -------------------------------------------------------------------------------------
 * One process and one context per GPU, images and cameras replicated (the reference shares one CPhotoSetS between its
 * worker threads, findMatch.hpp).  The candidates of a wave are independent given the grid snapshot, so each rank
 * evaluates a contiguous shard and the ONLY exchange is the all-gather of the per-candidate result records, after
 * which every rank applies the same deterministic commit (the role of the per-image locks around _pgrids in
 * source/pmvs/expand.cpp:225-237).  NCCL is loaded at run time (libnccl.so.2).
 * rank 0 creates the 128-byte id, the caller hands it to the other ranks (any channel), everybody calls comm_init. */
int pmvsb_comm_unique_id(pmvsb_ctx* ctx, uint8_t* id128);
int pmvsb_comm_init(pmvsb_ctx* ctx, int rank, int world, const uint8_t* id128);
/* host buffers: send = bytes, recv = world * bytes in rank order; identity copy when no communicator (world = 1) */
int pmvsb_allgather(pmvsb_ctx* ctx, const void* send, size_t bytes, void* recv);

/* Peer-memory exchange (the default of a multi-GPU pmvs2 run; NCCL stays as the fallback).  Every rank owns a mailbox in its
 * GPU's memory -- a header of flags plus one slot per rank -- which the other ranks map into their address space through CUDA
 * IPC.  A wave's message is then written by ONE kernel straight into every rank's mailbox (peer stores over NVLink / NVSwitch,
 * the message being packed from the result arrays on the fly), followed by a system-scope fence and a flag; the receiver
 * waits for the flags of all ranks and unpacks from its own memory.  No communicator bring-up, no staging copies, no
 * collective launch.
 *   pmvsb_peer_export : (re)allocates this rank's mailbox with `slot_bytes` per rank and returns its 64-byte IPC handle;
 *   pmvsb_peer_open   : handles = world * 64 bytes in rank order (any channel carries them); maps the peers' mailboxes.  After it
 *                       pmvsb_evaluate_allgather uses the mailboxes.  Fails (PMVSB_ECUDA) where the GPUs cannot reach each other;
 *   pmvsb_peer_close  : unmaps the peers' mailboxes (before a re-export: every rank closes, then every rank exports again);
 *   pmvsb_peer_needed : slot size in bytes that the wave which returned PMVSB_EGROW needs.
 * Ranks may share one GPU (the handles of another PROCESS on the same device open like any other). */
int pmvsb_peer_export(pmvsb_ctx* ctx, int rank, int world, size_t slot_bytes, uint8_t* handle64);
int pmvsb_peer_open(pmvsb_ctx* ctx, const uint8_t* handles);
int pmvsb_peer_close(pmvsb_ctx* ctx);
size_t pmvsb_peer_needed(const pmvsb_ctx* ctx);

/* pmvsb_refine_batch_dev followed by the all-gather of the refined records over peer memory (what the reference's worker threads
 * get for free from shared memory, source/pmvs/expand.cpp:225-237: every worker sees every accepted patch).  One kernel after
 * the refine kernel packs each patch's 48-byte record -- coord[4], normal[4], ncc, ok, evaluations, 0 as floats -- from the
 * result arrays and stores it into this rank's slot of EVERY rank's mailbox (coalesced 16-byte peer stores over NVLink), its
 * last block raises a flag in every mailbox, and a stream-ordered wait for all ranks' flags is queued (cuStreamWaitValue32: the
 * host does not block, no collective is launched).  PMVSB_GATHER_IN_KERNEL=1 moves the stores into the refine kernel itself
 * (each finished patch's leader lane posts its record underneath the evaluations still running) -- measured 1 % slower, kept
 * for A/B.  Work queued on the context's stream after this call sees the records of all ranks: rank k's record p at
 * *records + k * *rank_stride_floats + 12 * p.  Slots are used as two halves by call parity, so the records of one call stay
 * valid until the call after the next.  Needs pmvsb_peer_export (slot_bytes >= 2 * 48 * n; world 1 allowed) and, for
 * world > 1, pmvsb_peer_open; every rank must call it the same number of times.  PMVSB_EGROW as above. */
int pmvsb_refine_batch_dev_gather(pmvsb_ctx* ctx, int n, int stride, float* coords, float* normals, const int32_t* images,
                                  const int32_t* nimages, const float* dscales, float* ncc, int32_t* evals, uint8_t* ok,
                                  const float** records, size_t* rank_stride_floats);

/* The wave exchange, device to device: after pmvsb_evaluate_batch on this rank's contiguous shard [shard_lo, shard_lo + P) of a
 * wave of total_candidates, every rank calls this; the accepted candidates' records of all ranks (and all verdicts) are
 * all-gathered -- through the peer mailboxes when pmvsb_peer_open succeeded, else by ONE ncclAllGather straight from device
 * memory -- and pmvsb_evaluate_fetch then returns the WHOLE wave, `index` being wave-wide candidate numbers.  No-op without a
 * communicator of more than one rank.  PMVSB_EGROW: see pmvsb_peer_needed. */
int pmvsb_evaluate_allgather(pmvsb_ctx* ctx, int shard_lo, int total_candidates);
/* sizes of what pmvsb_evaluate_fetch will return (after pmvsb_evaluate_batch or pmvsb_evaluate_allgather) */
int pmvsb_evaluate_counts(pmvsb_ctx* ctx, int32_t* candidates, int32_t* accepted, int32_t* entries, int32_t* ventries);
double pmvsb_exchanged_bytes(const pmvsb_ctx* ctx);   /* bytes received through pmvsb_evaluate_allgather so far */

int pmvsb_sync(pmvsb_ctx* ctx);
/* the CUDA stream (cudaStream_t) the context launches on, for event timing by the caller */
void* pmvsb_stream(pmvsb_ctx* ctx);
/* launch on a caller-owned stream instead (e.g. the stream a collective will be ordered after);
 * NULL restores the context's own stream */
int pmvsb_set_stream(pmvsb_ctx* ctx, void* cuda_stream);
/* launches issued by this context since creation (bench.py's gpu_launches) */
uint64_t pmvsb_launch_count(const pmvsb_ctx* ctx);
/* duration in ms of the most recent pmvsb_refine_batch[_dev] kernel, measured with CUDA events on the
 * context stream (valid after pmvsb_sync) */
float pmvsb_last_refine_ms(pmvsb_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif
