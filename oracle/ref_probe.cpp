// oracle/ref_probe.cpp -- TEST INFRASTRUCTURE.
//
// C entry points (ctypes) around the REFERENCE'S OWN OBJECTS, compiled from /root/reference by
// oracle/Makefile into oracle/_ref/libpmvs_ref.so.  Nothing here re-implements the path: every call
// lands in the reference's COptim / CPatchOrganizerS / CPhotoSetS code.  It is used to
//   * pin oracle/pmvs_oracle.c (the C restatement) and generate tests/golden/ vectors,
//   * time the reference's CPU path (bench.py --impl reference, cpu_baseline kind "reference").
// The optimiser behind refinePatch is oracle/nm3.h through oracle/shim/nlopt.hpp (nlopt is absent).
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <list>
#include <map>
#include <memory>
#include <mutex>
#include <numeric>
#include <queue>
#include <set>
#include <shared_mutex>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include "nlopt.hpp"

// Reach the protected members of COptim (my_f, grabTex, per-thread scratch).  Layout is unaffected.
#define protected public
#define private public
#include "pmvs/findMatch.hpp"
#include "pmvs/option.hpp"
#include "pmvs/harris.hpp"
#include "pmvs/dog.hpp"
#include "pmvs/point.hpp"
#undef protected
#undef private

using namespace PMVS3;

namespace {
std::unique_ptr<CFindMatch> g_fm;
std::unique_ptr<SOption> g_opt;

inline Vec4f v4(const float* p) { return Vec4f(p[0], p[1], p[2], p[3]); }
inline void out4(const Vec4f& v, float* p) { for (int i = 0; i < 4; ++i) p[i] = v[i]; }

void fill_patch(Patch::CPatch& patch, const float* coord, const float* normal, const int* images, int n) {
  patch._coord = v4(coord);
  patch._normal = v4(normal);
  patch._images.assign(images, images + n);
}

// what refinePatchBFGS does before calling the optimiser (optim.cpp:584-596)
void setup_thread_context(const Patch::CPatch& patch, int id) {
  COptim& o = g_fm->_optim;
  o._centersT[id] = patch._coord;
  o._raysT[id] = patch._coord - g_fm->_pss._photos[patch._images[0]].OpticalCenter();
  unitize(o._raysT[id]);
  o._indexesT[id] = patch._images;
  o._dscalesT[id] = patch._dscale;
  o._ascalesT[id] = M_PI / 48.0f;
  o.setWeightsT(patch, id);
}
}  // namespace

extern "C" {

// Parse <prefix><option>, load images, build pyramids, detect features (skipped per image when
// <prefix>models/%08d.affin<level> exists, detectFeatures.cpp:65-73), init all stages.
int ref_open(const char* prefix, const char* option) {
  g_fm.reset();
  g_opt.reset(new SOption());
  g_opt->init(prefix, option);
  g_fm.reset(new CFindMatch());
  g_fm->init(*g_opt);
  return 0;
}

void ref_close(void) { g_fm.reset(); g_opt.reset(); }

// [num, tnum, level, csize, wsize, minImageNum, tau, CPU, depth]
void ref_config(int* out) {
  out[0] = g_fm->_num; out[1] = g_fm->_tnum; out[2] = g_fm->_level; out[3] = g_fm->_csize;
  out[4] = g_fm->_wsize; out[5] = g_fm->_minImageNumThreshold; out[6] = g_fm->_tau;
  out[7] = g_fm->_CPU; out[8] = g_fm->_depth;
}
// [nccThreshold, nccThresholdBefore, angleThreshold0, angleThreshold1, maxAngleThreshold, quad]
void ref_thresholds(float* out) {
  out[0] = g_fm->_nccThreshold; out[1] = g_fm->_nccThresholdBefore; out[2] = g_fm->_angleThreshold0;
  out[3] = g_fm->_angleThreshold1; out[4] = g_fm->_maxAngleThreshold; out[5] = g_fm->_quadThreshold;
}
void ref_set_depth(int depth) { g_fm->_depth = depth; }
void ref_set_thresholds(float ncc, float nccBefore) { g_fm->_nccThreshold = ncc; g_fm->_nccThresholdBefore = nccBefore; }

int ref_num_levels(void) { return g_fm->_pss._maxLevel; }
void ref_image_dims(int index, int level, int* w, int* h) {
  *w = g_fm->_pss.getWidth(index, level);
  *h = g_fm->_pss.getHeight(index, level);
}
void ref_image_bytes(int index, int level, unsigned char* out) {
  const std::vector<unsigned char>& im = g_fm->_pss._photos[index].getImage(level);
  std::memcpy(out, im.data(), im.size());
}
// P: 12 floats at `level`; centre 4; oaxis 4; xaxis/yaxis/zaxis 3 each (COptim's), ipscale (COptim's)
void ref_camera(int index, int level, float* P, float* centre, float* oaxis, float* xaxis, float* yaxis,
                float* zaxis, float* ipscale) {
  const Image::CPhoto& ph = g_fm->_pss._photos[index];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 4; ++c) P[4 * r + c] = ph._projection[level][r][c];
  out4(ph._center, centre);
  out4(ph._oaxis, oaxis);
  for (int i = 0; i < 3; ++i) {
    xaxis[i] = g_fm->_optim._xaxes[index][i];
    yaxis[i] = g_fm->_optim._yaxes[index][i];
    zaxis[i] = g_fm->_optim._zaxes[index][i];
  }
  *ipscale = g_fm->_optim._ipscales[index];
}
int ref_visdata2(int index, int* out, int cap) {
  const std::vector<int>& v = g_fm->_visdata2[index];
  int n = std::min((int)v.size(), cap);
  for (int i = 0; i < n; ++i) out[i] = v[i];
  return (int)v.size();
}

void ref_project(int index, const float* coord, int level, float* out) {
  Vec3f p = g_fm->_pss.project(index, v4(coord), level);
  out[0] = p[0]; out[1] = p[1]; out[2] = p[2];
}
float ref_get_unit(int index, const float* coord) { return g_fm->_optim.getUnit(index, v4(coord)); }
void ref_get_color(int index, float x, float y, int level, float* rgb) {
  Vec3f c = g_fm->_pss.getColor(index, x, y, level);
  rgb[0] = c[0]; rgb[1] = c[1]; rgb[2] = c[2];
}
void ref_get_paxes(int index, const float* coord, const float* normal, float* px, float* py) {
  Vec4f a, b;
  g_fm->_optim.getPAxes(index, v4(coord), v4(normal), a, b);
  out4(a, px); out4(b, py);
}
// grabTex as my_f calls it: axes from the reference image `ref`, sample image `index`.
// Returns the reference's flag (0 = texture grabbed, 1 = rejected); tex gets wsize*wsize*3 floats.
int ref_grab_tex(const float* coord, const float* normal, int ref, int index, float* tex) {
  Vec4f px, py;
  g_fm->_optim.getPAxes(ref, v4(coord), v4(normal), px, py);
  std::vector<float> t;
  int flag = g_fm->_optim.grabTex(v4(coord), px, py, v4(normal), index, g_fm->_wsize, t);
  if (flag == 0) std::memcpy(tex, t.data(), t.size() * sizeof(float));
  return flag;
}
void ref_normalize(float* tex, int n) {
  std::vector<float> t(tex, tex + n);
  COptim::normalize(t);
  std::memcpy(tex, t.data(), n * sizeof(float));
}
float ref_dot(const float* a, const float* b, int n) {
  std::vector<float> t0(a, a + n), t1(b, b + n);
  return g_fm->_optim.dot(t0, t1);
}

void ref_set_scales(const float* coord, const int* images, int n, float* dscale, float* ascale) {
  Patch::CPatch patch;
  float nrm[4] = {0, 0, 0, 0};
  fill_patch(patch, coord, nrm, images, n);
  g_fm->_pos.setScales(patch);
  *dscale = patch._dscale;
  *ascale = patch._ascale;
}

void ref_encode(const float* coord, const float* normal, const int* images, int n, float dscale, double* x) {
  Patch::CPatch patch;
  fill_patch(patch, coord, normal, images, n);
  patch._dscale = dscale;
  setup_thread_context(patch, 0);
  g_fm->_optim.encode(patch._coord, patch._normal, x, 0);
}
// coord/normal = the START patch (defines centre, ray, images); x = parameter vector to evaluate/decode.
void ref_decode(const float* coord, const float* normal, const int* images, int n, float dscale,
                const double* x, float* ocoord, float* onormal) {
  Patch::CPatch patch;
  fill_patch(patch, coord, normal, images, n);
  patch._dscale = dscale;
  setup_thread_context(patch, 0);
  Vec4f c, nn;
  g_fm->_optim.decode(c, nn, x, 0);
  out4(c, ocoord); out4(nn, onormal);
}
double ref_my_f(const float* coord, const float* normal, const int* images, int n, float dscale, const double* x) {
  Patch::CPatch patch;
  fill_patch(patch, coord, normal, images, n);
  patch._dscale = dscale;
  setup_thread_context(patch, 0);
  int id = 0;
  return COptim::my_f(3, x, nullptr, &id);
}
double ref_compute_incc(const float* coord, const float* normal, const int* images, int n, int robust) {
  Patch::CPatch patch;
  fill_patch(patch, coord, normal, images, n);
  g_fm->_optim.setWeightsT(patch, 0);
  return g_fm->_optim.computeINCC(patch._coord, patch._normal, patch._images, 0, robust);
}
// ref-vs-all (vector form) of setINCCs; out has n floats
void ref_set_inccs(const float* coord, const float* normal, const int* images, int n, int robust, float* out) {
  Patch::CPatch patch;
  fill_patch(patch, coord, normal, images, n);
  std::vector<float> inccs;
  g_fm->_optim.setINCCs(patch, inccs, patch._images, 0, robust);
  std::memcpy(out, inccs.data(), n * sizeof(float));
}
// all-pairs (matrix form); out has n*n floats
void ref_set_inccs_matrix(const float* coord, const float* normal, const int* images, int n, int robust, float* out) {
  Patch::CPatch patch;
  fill_patch(patch, coord, normal, images, n);
  std::vector<std::vector<float> > inccs;
  g_fm->_optim.setINCCs(patch, inccs, patch._images, 0, robust);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) out[i * n + j] = inccs[i][j];
}

// refinePatch.  In/out coord, normal; out ncc, evals.  Returns 1 when the optimiser reported success.
int ref_refine(float* coord, float* normal, const int* images, int n, float dscale, float* ncc, int* evals) {
  Patch::CPatch patch;
  fill_patch(patch, coord, normal, images, n);
  patch._dscale = dscale;
  const bool ok = g_fm->_optim.refinePatchBFGS(patch, 0, 1000, 1);
  out4(patch._coord, coord); out4(patch._normal, normal);
  *ncc = patch._ncc;
  *evals = nlopt::last_evals();
  return ok ? 1 : 0;
}

// Batched refinePatch over `threads` host threads (one COptim scratch slot each; needs CPU >= threads
// in the option file).  images: P x V row-major.  Returns seconds spent inside the loop.
double ref_refine_batch(int P, int V, float* coords, float* normals, const int* images, const float* dscales,
                        float* nccs, int* evals, unsigned char* ok, int threads) {
  if (threads < 1) threads = 1;
  if (threads > g_fm->_CPU) threads = g_fm->_CPU;
  std::atomic<int> next(0);
  auto t0 = std::chrono::steady_clock::now();
  auto work = [&](int id) {
    for (;;) {
      int b = next.fetch_add(64);
      if (b >= P) break;
      int e = std::min(P, b + 64);
      for (int p = b; p < e; ++p) {
        Patch::CPatch patch;
        fill_patch(patch, coords + 4 * p, normals + 4 * p, images + (size_t)V * p, V);
        patch._dscale = dscales[p];
        const bool s = g_fm->_optim.refinePatchBFGS(patch, id, 1000, 1);
        out4(patch._coord, coords + 4 * p); out4(patch._normal, normals + 4 * p);
        nccs[p] = patch._ncc;
        evals[p] = nlopt::last_evals();
        ok[p] = s ? 1 : 0;
      }
    }
  };
  std::vector<std::thread> th;
  for (int i = 1; i < threads; ++i) th.emplace_back(work, i);
  work(0);
  for (auto& t : th) t.join();
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// preProcess / postProcess on one patch.  images in/out (cap entries), returns the reference's verdict
// (0 keep, 1 reject).  *n is updated.  grids (2 ints per image) are filled by postProcess's setGrids.
int ref_pre_process(const float* coord, const float* normal, int* images, int* n, int cap, float* dscale, float* ascale) {
  Patch::CPatch patch;
  fill_patch(patch, coord, normal, images, *n);
  int r = g_fm->_optim.preProcess(patch, 0, 0);
  *n = std::min((int)patch._images.size(), cap);
  for (int i = 0; i < *n; ++i) images[i] = patch._images[i];
  *dscale = patch._dscale; *ascale = patch._ascale;
  return r;
}
int ref_post_process(const float* coord, const float* normal, float ncc, int* images, int* n, int cap, int* grids,
                     int* vimages, int* nv, int* vgrids, int* timages, float* tmp) {
  Patch::CPatch patch;
  fill_patch(patch, coord, normal, images, *n);
  patch._ncc = ncc;
  int r = g_fm->_optim.postProcess(patch, 0, 0);
  *n = std::min((int)patch._images.size(), cap);
  for (int i = 0; i < *n; ++i) images[i] = patch._images[i];
  for (int i = 0; i < *n && i < (int)patch._grids.size(); ++i) { grids[2 * i] = patch._grids[i][0]; grids[2 * i + 1] = patch._grids[i][1]; }
  *nv = std::min((int)patch._vimages.size(), cap);
  for (int i = 0; i < *nv; ++i) { vimages[i] = patch._vimages[i]; vgrids[2 * i] = patch._vgrids[i][0]; vgrids[2 * i + 1] = patch._vgrids[i][1]; }
  *timages = patch._timages;
  *tmp = patch._tmp;
  return r;
}

// ---- masks / edges / bounding images --------------------------------------------------------------------------
// CImage::getMask(level) (which = 0) / getEdge(level) (1) bytes of one image; returns the byte count (0 = no map)
int ref_map_bytes(int index, int which, int level, unsigned char* out) {
  const Image::CPhoto& ph = g_fm->_pss._photos[index];
  const std::vector<unsigned char>& m = which == 0 ? ph.CImage::getMask(level) : ph.CImage::getEdge(level);
  if (out && !m.empty()) std::memcpy(out, m.data(), m.size());
  return (int)m.size();
}
// the gate of expandSub / collectCandidates / postProcess (expand.cpp:212, seed.cpp:314, optim.cpp:153): 1 = passes
int ref_mask_gate(const float* coord) {
  return !(g_fm->_pss.getMask(v4(coord), g_fm->_level) == 0 || g_fm->insideBimages(v4(coord)) == 0);
}
int ref_get_mask_image(const float* coord, int index) { return g_fm->_pss.getMask(v4(coord), index, g_fm->_level); }
int ref_get_edge(const float* coord, int index) { return g_fm->_pss.getEdge(v4(coord), index, g_fm->_level); }
int ref_num_bimages(void) { return (int)g_fm->_bindexes.size(); }
// COptim::removeImagesEdge; returns the new length
int ref_remove_images_edge(const float* coord, int* images, int n) {
  Patch::CPatch patch;
  float nrm[4] = {0, 0, 0, 0};
  fill_patch(patch, coord, nrm, images, n);
  g_fm->_optim.removeImagesEdge(patch);
  for (int i = 0; i < (int)patch._images.size(); ++i) images[i] = patch._images[i];
  return (int)patch._images.size();
}

// Whole reference run (seed + 3 x expand/filter), CFindMatch::run.
void ref_run(void) { g_fm->run(); }
int ref_num_patches(void) { g_fm->_pos.collectPatches(1); return (int)g_fm->_pos._ppatches.size(); }
// after ref_num_patches: copy patch i
void ref_get_patch(int i, float* coord, float* normal, float* ncc_dscale_ascale, int* images, int* n, int cap,
                   int* vimages, int* nv) {
  const Patch::CPatch& p = *g_fm->_pos._ppatches[i];
  out4(p._coord, coord); out4(p._normal, normal);
  ncc_dscale_ascale[0] = p._ncc; ncc_dscale_ascale[1] = p._dscale; ncc_dscale_ascale[2] = p._ascale;
  *n = std::min((int)p._images.size(), cap);
  for (int k = 0; k < *n; ++k) images[k] = p._images[k];
  *nv = std::min((int)p._vimages.size(), cap);
  for (int k = 0; k < *nv; ++k) vimages[k] = p._vimages[k];
}

// ---- filter-stage state (after ref_run): the reference's own depth maps, visibility test, gains -------------
// collectPatches(0) numbers the patches: _ppatches[k]->_id == k
int ref_collect_patches(void) { g_fm->_pos.collectPatches(0); return (int)g_fm->_pos._ppatches.size(); }
void ref_grid_dims(int image, int* gw, int* gh) { *gw = g_fm->_pos._gwidths[image]; *gh = g_fm->_pos._gheights[image]; }
// full patch record k: coord4, normal4, [ncc, dscale, ascale, tmp], images+grids, vimages+vgrids, [timages, fix, flag]
void ref_get_patch_full(int k, float* coord, float* normal, float* scal4, int* images, int* grids, int* n, int* vimages, int* vgrids,
                        int* nv, int* misc3, int cap) {
  const Patch::CPatch& p = *g_fm->_pos._ppatches[k];
  out4(p._coord, coord); out4(p._normal, normal);
  scal4[0] = p._ncc; scal4[1] = p._dscale; scal4[2] = p._ascale; scal4[3] = p._tmp;
  *n = std::min((int)p._images.size(), cap);
  for (int i = 0; i < *n; ++i) { images[i] = p._images[i]; grids[2 * i] = p._grids[i][0]; grids[2 * i + 1] = p._grids[i][1]; }
  *nv = std::min((int)p._vimages.size(), cap);
  for (int i = 0; i < *nv; ++i) { vimages[i] = p._vimages[i]; vgrids[2 * i] = p._vgrids[i][0]; vgrids[2 * i + 1] = p._vgrids[i][1]; }
  misc3[0] = p._timages; misc3[1] = p._fix; misc3[2] = p._flag;
}
// CFilter::setDepthMaps over the collected patches; out = patch id per cell of `image` (-1 = empty)
void ref_set_depth_maps(void) { g_fm->_filter.setDepthMaps(); }
void ref_get_depth_map(int image, int* out) {
  const int n = g_fm->_pos._gwidths[image] * g_fm->_pos._gheights[image];
  for (int i = 0; i < n; ++i) {
    const Patch::PPatch& q = g_fm->_pos._dpgrids[image][i];
    out[i] = (q == CPatchOrganizerS::_MAXDEPTH) ? -1 : q->_id;
  }
}
int ref_is_visible(int k, int image, int ix, int iy, float strict) {
  return g_fm->_pos.isVisible(*g_fm->_pos._ppatches[k], image, ix, iy, strict, 0);
}
int ref_is_neighbor(int a, int b, float thr) { return g_fm->isNeighbor(*g_fm->_pos._ppatches[a], *g_fm->_pos._ppatches[b], thr); }
float ref_compute_gain(int k) { return g_fm->_filter.computeGain(*g_fm->_pos._ppatches[k], 0); }
// CPatchOrganizerS::setVImagesVGrids on a COPY of patch k whose _vimages were cleared first
int ref_set_vimages(int k, int* vimages, int* vgrids, int cap) {
  Patch::CPatch p = *g_fm->_pos._ppatches[k];
  p._vimages.clear(); p._vgrids.clear();
  g_fm->_pos.setVImagesVGrids(p);
  const int n = std::min((int)p._vimages.size(), cap);
  for (int i = 0; i < n; ++i) { vimages[i] = p._vimages[i]; vgrids[2 * i] = p._vgrids[i][0]; vgrids[2 * i + 1] = p._vgrids[i][1]; }
  return n;
}
// patches of one cell of _pgrids, in list order
int ref_get_cell(int image, int cell, int* out, int cap) {
  const std::vector<Patch::PPatch>& v = g_fm->_pos._pgrids[image][cell];
  const int n = std::min((int)v.size(), cap);
  for (int i = 0; i < n; ++i) out[i] = v[i]->_id;
  return (int)v.size();
}
// CPatchOrganizerS::findNeighbors of table patch k: unique neighbours as table ids, ascending
int ref_find_neighbors(int k, float scale, int margin, int skipvis, int* out, int cap) {
  std::vector<Patch::PPatch> nb;
  g_fm->_pos.findNeighbors(*g_fm->_pos._ppatches[k], nb, 0, scale, margin, skipvis);
  std::vector<int> ids;
  for (const auto& q : nb) ids.push_back(q->_id);
  std::sort(ids.begin(), ids.end());
  for (int i = 0; i < (int)ids.size() && i < cap; ++i) out[i] = ids[i];
  return (int)ids.size();
}
float ref_compute_radius(int k) { return g_fm->_expand.computeRadius(*g_fm->_pos._ppatches[k]); }
// CExpand::findEmptyBlocks on a copy of patch k with _dflag cleared: bit i = direction i is NOT offered, i.e. fill[i] > 0
int ref_find_empty_blocks(int k) {
  Patch::PPatch pp(new Patch::CPatch(*g_fm->_pos._ppatches[k]));
  pp->_dflag = 0;
  std::vector<std::vector<Vec4f> > can;
  g_fm->_expand.findEmptyBlocks(pp, can);
  int mask = 0;
  for (int i = 0; i < (int)can.size(); ++i) if (can[i].empty()) mask |= 1 << i;
  return mask;
}
// CFilter::filterNeighborThread's test for table patch k (filter.cpp:375-388): 1 = reject
int ref_filter_neighbor(int k, float quad, int* ncount) {
  g_fm->_quadThreshold = quad;   // option `quad` (source/pmvs/option.cpp), default 2.5
  std::vector<Patch::PPatch> nb;
  g_fm->_pos.findNeighbors(*g_fm->_pos._ppatches[k], nb, 0, 4, 2, 1);
  if (ncount) *ncount = (int)nb.size();
  if ((int)nb.size() < 6) return 1;
  return g_fm->_filter.filterQuad(*g_fm->_pos._ppatches[k], nb);
}
// CDetectFeatures::runThread's body for one image (detectFeatures.cpp:77-118): CHarris then CDifferenceOfGaussians on the
// working-level image, each result multiset read in reverse (strongest first)
int ref_detect_features(int index, int gspeedup, float* xy, float* resp, int* types, int cap) {
  Image::CPhotoSetS& pss = g_fm->_pss;
  const int level = g_fm->_level;
  int n = 0;
  auto emit = [&](std::multiset<CPoint>& result) {
    for (auto it = result.rbegin(); it != result.rend(); ++it) {
      if (n < cap) { xy[2 * n] = it->_icoord[0]; xy[2 * n + 1] = it->_icoord[1]; resp[n] = it->_response; types[n] = it->_type; }
      ++n;
    }
  };
  {
    CHarris harris;
    std::multiset<CPoint> result;
    harris.run(pss._photos[index].getImage(level), pss._photos[index].CImage::getMask(level), pss._photos[index].CImage::getEdge(level),
               pss._photos[index].getWidth(level), pss._photos[index].getHeight(level), gspeedup, 4.0f, result);
    emit(result);
  }
  {
    CDifferenceOfGaussians dog;
    std::multiset<CPoint> result;
    dog.run(pss._photos[index].getImage(level), pss._photos[index].CImage::getMask(level), pss._photos[index].CImage::getEdge(level),
            pss._photos[index].getWidth(level), pss._photos[index].getHeight(level), gspeedup, 1.0f, 3.0f, result);
    emit(result);
  }
  return n;
}
// COptim::check on a COPY of table patch k (check clears the image list of a rejected patch): 1 = reject, *gain = _tmp
int ref_check(int k, float quad, float* gain) {
  g_fm->_quadThreshold = quad;
  Patch::CPatch p = *g_fm->_pos._ppatches[k];
  const int r = g_fm->_optim.check(p);
  *gain = p._tmp;
  return r;
}
// ---- seed stage (before ref_run: CFindMatch::run frees CSeed::_ppoints when the seeds are done) ----------------------
// features of image `index` as CSeed holds them (cells row-major, a cell's features in detection order)
int ref_features(int index, float* xy, int* type, int cap) {
  int n = 0;
  for (const auto& cell : g_fm->_seed._ppoints[index])
    for (const auto& pp : cell) {
      if (n < cap) { xy[2 * n] = pp->_icoord[0]; xy[2 * n + 1] = pp->_icoord[1]; type[n] = pp->_type; }
      ++n;
    }
  return n;
}
int ref_collect_images(int index, int* out, int cap) {
  std::vector<int> indexes;
  g_fm->_optim.collectImages(index, indexes);
  if (g_fm->_tau < (int)indexes.size()) indexes.resize(g_fm->_tau);   // seed.cpp:124-126
  for (int i = 0; i < (int)indexes.size() && i < cap; ++i) out[i] = indexes[i];
  return (int)indexes.size();
}
// CPatchOrganizerS::_counts of one target image (what CSeed::canAdd reads beside _pgrids)
void ref_set_counts(int image, const unsigned char* counts) {
  std::vector<unsigned char>& c = g_fm->_pos._counts[image];
  std::memcpy(c.data(), counts, c.size());
}
int ref_can_add(int image, int x, int y) { return g_fm->_seed.canAdd(image, x, y); }
// CSeed::collectCandidates for feature p of cell `cell` of image `index`: per candidate the other image, the other
// feature's pixel, the triangulated point and _response.  (The order is the reference's sort of shared_ptr values.)
int ref_collect_candidates(int index, int cell, int p, int* other_image, float* other_xy, float* coords, float* resp, int cap) {
  std::vector<int> indexes;
  g_fm->_optim.collectImages(index, indexes);
  if (g_fm->_tau < (int)indexes.size()) indexes.resize(g_fm->_tau);
  std::vector<PPoint> vcp;
  g_fm->_seed.collectCandidates(index, indexes, *g_fm->_seed._ppoints[index][cell][p], vcp);
  for (int i = 0; i < (int)vcp.size() && i < cap; ++i) {
    other_image[i] = vcp[i]->_itmp;
    other_xy[2 * i] = vcp[i]->_icoord[0]; other_xy[2 * i + 1] = vcp[i]->_icoord[1];
    out4(vcp[i]->_coord, coords + 4 * i);
    resp[i] = vcp[i]->_response;
  }
  return (int)vcp.size();
}

// ---- expansion cell rules on the reference's own grids ----------------------------------------------------------------
void ref_get_occupancy(int image, int* out) {
  const auto& g = g_fm->_pos._pgrids[image];
  for (size_t c = 0; c < g.size(); ++c) out[c] = (int)g[c].size();
}
void ref_get_counts(int image, unsigned char* out) {
  const std::vector<unsigned char>& c = g_fm->_pos._counts[image];
  std::memcpy(out, c.data(), c.size());
}
void ref_set_count_threshold1(int v) { g_fm->_countThreshold1 = v; }
static void raw_patch(Patch::CPatch& p, const int* images, const int* grids, int n, const int* vimages, const int* vgrids, int nv) {
  p._images.assign(images, images + n);
  for (int i = 0; i < n; ++i) p._grids.push_back(TVec2<int>(grids[2 * i], grids[2 * i + 1]));
  p._vimages.assign(vimages, vimages + nv);
  for (int i = 0; i < nv; ++i) p._vgrids.push_back(TVec2<int>(vgrids[2 * i], vgrids[2 * i + 1]));
}
// CExpand::checkCounts / updateCounts on a patch made of the given lists
int ref_check_counts_raw(const int* images, const int* grids, int n) {
  Patch::CPatch p;
  raw_patch(p, images, grids, n, nullptr, nullptr, 0);
  return g_fm->_expand.checkCounts(p);
}
int ref_update_counts_raw(const int* images, const int* grids, int n, const int* vimages, const int* vgrids, int nv) {
  Patch::CPatch p;
  raw_patch(p, images, grids, n, vimages, vgrids, nv);
  return g_fm->_expand.updateCounts(p);
}

// ---- filter-round stages on the reference's own state --------------------------------------------------------------
// removePatch for the table patches (numbering of ref_collect_patches) with keep[k] == 0, then
// CFilter::setDepthMapsVGridsVPGridsAddPatchV(additive); old_index[i] = former table index of new table patch i.
static std::map<const Patch::CPatch*, int> g_old_index;
static int report_old_indexes(int* old_index, int cap) {
  g_fm->_pos.collectPatches(0);
  const int n = (int)g_fm->_pos._ppatches.size();
  for (int i = 0; i < n && i < cap; ++i) old_index[i] = g_old_index[g_fm->_pos._ppatches[i].get()];
  return n;
}
int ref_remove_and_rebuild(const unsigned char* keep, int additive, int* old_index, int cap) {
  g_fm->_pos.collectPatches(0);
  const std::vector<Patch::PPatch> pp = g_fm->_pos._ppatches;
  g_old_index.clear();
  for (int k = 0; k < (int)pp.size(); ++k) g_old_index[pp[k].get()] = k;
  for (int k = 0; k < (int)pp.size(); ++k)
    if (!keep[k]) g_fm->_pos.removePatch(pp[k]);
  g_fm->_filter.setDepthMapsVGridsVPGridsAddPatchV(additive);
  return report_old_indexes(old_index, cap);
}
// CFilter::filterSmallGroups on the current state; survivors as former table indexes (numbering of the last ref_remove_and_rebuild)
int ref_filter_small_groups(int* old_index, int cap) {
  g_fm->_filter.filterSmallGroups();
  return report_old_indexes(old_index, cap);
}
// Moves the flagged table patches a fraction t of the way towards the camera of their reference image (removePatch, shift,
// setGrids, addPatch): they now hide what lies behind them, which gives CFilter::filterExact something to prune.
void ref_shift_patches(const unsigned char* flags, float t) {
  g_fm->_pos.collectPatches(0);
  std::vector<Patch::PPatch> pp = g_fm->_pos._ppatches;
  for (int k = 0; k < (int)pp.size(); ++k) {
    if (!flags[k]) continue;
    g_fm->_pos.removePatch(pp[k]);
    const Vec4f c = g_fm->_pss._photos[pp[k]->_images[0]].OpticalCenter();
    pp[k]->_coord = pp[k]->_coord + t * (c - pp[k]->_coord);
    pp[k]->_coord[3] = 1.0f;
    g_fm->_pos.setGrids(*pp[k]);
    g_fm->_pos.addPatch(pp[k]);
  }
}
// CFilter::filterExact on the current state (depth maps of the last rebuild); survivors as former table indexes
int ref_filter_exact(int* old_index, int cap) {
  g_fm->_filter.filterExact();
  return report_old_indexes(old_index, cap);
}
int ref_get_depth_flag(void) { return g_fm->_depth; }
float ref_neighbor_threshold(int which) { return which == 0 ? g_fm->_neighborThreshold : (which == 1 ? g_fm->_neighborThreshold1 : g_fm->_neighborThreshold2); }

// x-tolerance floor of the nm3 stand-in (oracle/shim/nlopt.hpp): 1e-3 by default, 0 = the reference's own xtol_rel 1e-7
void ref_set_xtol_floor(double v) { nlopt::xtol_floor() = v; }
double ref_get_xtol_floor(void) { return nlopt::xtol_floor(); }

unsigned long long ref_total_evals(void) { return nlopt::total_evals(); }
unsigned long long ref_total_calls(void) { return nlopt::total_calls(); }
void ref_reset_counters(void) { nlopt::total_evals() = 0; nlopt::total_calls() = 0; }

}  // extern "C"
