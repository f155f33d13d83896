// cmvs-pmvs_b200/csrc/pmvs_select.cuh
//
// Visible-image-set selection around the refinement: COptim::preProcess / postProcess and the helpers they
// call (addImages, constraintImages, sortImages, setScales, checkAngles, filterImagesByAngle, setRefImage,
// setGrids; /root/reference/source/pmvs/optim.cpp:95-254, 284-321, 398-444).
//
// These produce INTEGER results (image lists, grid cells, verdicts) that must equal the reference's bit for
// bit, and they are decided by comparing NCC values with thresholds.  So, unlike the refine kernel, the
// texture arithmetic here keeps the reference's exact association: ONE LANE OWNS ONE VIEW and walks its 49
// texels, its normalisation and its dot products sequentially in the reference's order (textures live in
// shared memory, one column per view).  One warp handles one patch; up to kSelMaxViews views per patch.
#pragma once
#include "pmvs_device.cuh"

namespace pmvsb {

constexpr int kSelMaxViews = 64;

struct SelectParams {          // host-tabulated constants (same libm as the reference)
  const int32_t* vis_off;      // visdata2 as CSR: candidates of image i are vis_idx[vis_off[i] .. vis_off[i+1])
  const int32_t* vis_idx;
  float cos_angle0_f;          // (float)cos(angleThreshold0): addImages compares in float (optim.cpp:416,438)
  float sort_threshold;        // (float)(1.0f - cos(10 deg)) (optim.cpp:287)
  float angle_dot_lo, angle_dot_hi;  // checkAngles: minAngle < (float)acos(d) < maxAngle  <=>  lo <= d <= hi
  float ncc_threshold, ncc_threshold_before;
  int* overflow;               // set when addImages found more images than the list capacity holds (reported, never silent)
};

// per-warp scratch in shared memory
// MAXV = view capacity: 37.6 KB at 64 views (6 resident warps per SM), 9.7 KB at 16 (23 warps); kernels whose list length is
// known up front (setRefImage) are launched once per capacity class
template <int WSIZE, int MAXV = kSelMaxViews>
struct SelScratch {
  static constexpr int TS = 3 * WSIZE * WSIZE;
  float tex[TS][MAXV];      // column v = texture of view v (bank-conflict free per lane, broadcast for v fixed)
  int images[MAXV];
  int tmp_images[MAXV];
  int aux[MAXV];            // the full image list while `images` holds the target subset (setRefImage)
  float val[MAXV];          // inccs / units
  float rays[MAXV][4];
  unsigned char valid[MAXV];
};

// grabTex (optim.cpp:815-863) by ONE lane, sequential, into column `slot`; then normalize (optim.cpp:1031-1067).
template <int WSIZE, int MAXV>
__device__ __forceinline__ bool lane_grab_normalize(const SceneDev& s, SelScratch<WSIZE, MAXV>& sc, int slot, int index, const float* coord,
                                                    const float* px, const float* py, const float* pz) {
  constexpr int TS = SelScratch<WSIZE, MAXV>::TS;
  CamDev cam;
  load_cam(s, index, cam);
  const ViewWin w = view_window<WSIZE>(s, cam, index, coord, px, py, pz);
  if (w.newlevel < 0) return false;
  const LevelDev lv = s.levels[index * s.nlevels + w.newlevel];
  float lx = w.lx, ly = w.ly;
  int k = 0;
  for (int y = 0; y < WSIZE; ++y) {
    float vx = lx, vy = ly;
    lx += w.dyx; ly += w.dyy;
    for (int x = 0; x < WSIZE; ++x) {
      float rgb[3];
      get_color(lv, vx, vy, rgb);
      sc.tex[k][slot] = rgb[0]; sc.tex[k + 1][slot] = rgb[1]; sc.tex[k + 2][slot] = rgb[2];
      k += 3;
      vx += w.dxx; vy += w.dxy;
    }
  }
  float ave0 = 0.f, ave1 = 0.f, ave2 = 0.f;
  for (int i = 0; i < TS; i += 3) { ave0 += sc.tex[i][slot]; ave1 += sc.tex[i + 1][slot]; ave2 += sc.tex[i + 2][slot]; }
  const float n = (float)(WSIZE * WSIZE);
  ave0 /= n; ave1 /= n; ave2 /= n;
  float sq = 0.0f;
  for (int i = 0; i < TS; i += 3) {
    const float f0 = ave0 - sc.tex[i][slot], f1 = ave1 - sc.tex[i + 1][slot], f2 = ave2 - sc.tex[i + 2][slot];
    sq += f0 * f0 + f1 * f1 + f2 * f2;
  }
  float sd = sqrtf(sq / (float)TS);
  if (sd == 0.0f) sd = 1.0f;
  for (int i = 0; i < TS; i += 3) {
    sc.tex[i][slot] = (sc.tex[i][slot] - ave0) / sd;
    sc.tex[i + 1][slot] = (sc.tex[i + 1][slot] - ave1) / sd;
    sc.tex[i + 2][slot] = (sc.tex[i + 2][slot] - ave2) / sd;
  }
  return true;
}

// COptim::dot (optim.cpp:1069-1077), sequential
template <int WSIZE, int MAXV>
__device__ __forceinline__ float lane_dot(const SelScratch<WSIZE, MAXV>& sc, int a, int b) {
  constexpr int TS = SelScratch<WSIZE, MAXV>::TS;
  float ans = 0.0f;
  for (int i = 0; i < TS; ++i) ans += sc.tex[i][a] * sc.tex[i][b];
  return ans / (float)TS;
}

// order-preserving compaction of sc.images[0..n) by keep(i); returns the new length (warp-collective)
template <class Scratch, class Pred>
__device__ __forceinline__ int warp_compact(Scratch& sc, int n, int lane, Pred keep) {
  int out = 0;
  for (int base = 0; base < n; base += 32) {
    const int i = base + lane;
    const bool k = i < n && keep(i);
    const int img = i < n ? sc.images[i] : -1;
    const unsigned m = __ballot_sync(kFull, k);
    __syncwarp();
    if (k) sc.tmp_images[out + __popc(m & ((1u << lane) - 1u))] = img;
    out += __popc(m);
  }
  __syncwarp();
  for (int i = lane; i < out; i += 32) sc.images[i] = sc.tmp_images[i];
  __syncwarp();
  return out;
}

// COptim::addImages (optim.cpp:398-444)
template <class Scratch>
__device__ __forceinline__ int sel_add_images(const SceneDev& s, const SelectParams& sp, Scratch& sc, int n, int cap, int lane,
                                              const float* coord, const float* normal) {
  const int ref = sc.images[0];
  const int beg = sp.vis_off[ref], end = sp.vis_off[ref + 1];
  for (int base = beg; base < end; base += 32) {
    const int c = base + lane;
    bool ok = false;
    int im = -1;
    if (c < end) {
      im = sp.vis_idx[c];
      bool used = false;
      for (int j = 0; j < n; ++j) used |= (sc.images[j] == im);   // lists are short; the reference's `used` flags
      if (!used) {
        CamDev cam;
        load_cam(s, im, cam);
        float ic[3];
        project(cam, coord, ic);
        const LevelDev lv = s.levels[im * s.nlevels + s.level];
        const bool outside = ic[0] < 0.0f || (float)(lv.w - 1) <= ic[0] || ic[1] < 0.0f || (float)(lv.h - 1) <= ic[1];
        if (!outside && get_edge_img(s, cam, im, coord)) {
          float ray[4] = {cam.centre[0] - coord[0], cam.centre[1] - coord[1], cam.centre[2] - coord[2], cam.centre[3] - coord[3]};
          unitize4(ray);
          ok = sp.cos_angle0_f <= dot4(ray, normal);
        }
      }
    }
    // the `used` test above only looks at images present BEFORE this 32-candidate chunk; candidates are distinct
    // image ids (visdata2 has no duplicates), so appending inside the chunk cannot create a duplicate
    const unsigned m = __ballot_sync(kFull, ok);
    const int pos = n + __popc(m & ((1u << lane) - 1u));
    if (ok && pos < cap) sc.images[pos] = im;
    if (ok && pos >= cap) *sp.overflow = 1;   // the reference has no cap (optim.cpp:439): the entry point turns this into an error
    n = min(cap, n + __popc(m));
    __syncwarp();
  }
  return n;
}

// textures + validity of views [0, n) with axes from images[0] (what both setINCCs forms do first, optim.cpp:709-722)
template <int WSIZE, int MAXV>
__device__ __forceinline__ void sel_grab_all(const SceneDev& s, SelScratch<WSIZE, MAXV>& sc, int n, int lane, const float* coord, const float* normal) {
  CamDev refcam;
  load_cam(s, sc.images[0], refcam);
  float px[4], py[4];
  get_paxes(refcam, s.level, coord, normal, px, py);
  for (int i = lane; i < n; i += 32) sc.valid[i] = lane_grab_normalize<WSIZE>(s, sc, i, sc.images[i], coord, px, py, normal) ? 1 : 0;
  __syncwarp();
}

// COptim::constraintImages (optim.cpp:192-206): keep image 0 and every i with 1 - NCC(0, i) < 1 - threshold
template <int WSIZE, int MAXV>
__device__ __forceinline__ int sel_constraint_images(const SceneDev& s, SelScratch<WSIZE, MAXV>& sc, int n, int lane, const float* coord,
                                                     const float* normal, float ncc_threshold) {
  sel_grab_all<WSIZE>(s, sc, n, lane, coord, normal);
  const bool ref_ok = sc.valid[0] != 0;
  for (int i = lane; i < n; i += 32) {
    float v = 2.0f;
    if (ref_ok) {
      if (i == 0) v = 0.0f;
      else if (sc.valid[i]) v = 1.0f - lane_dot<WSIZE>(sc, 0, i);
    }
    sc.val[i] = v;
  }
  __syncwarp();
  const float lim = 1.0f - ncc_threshold;
  return warp_compact(sc, n, lane, [&](int i) { return i == 0 || sc.val[i] < lim; });
}

// COptim::sortImages, newm == 1 (optim.cpp:284-321) with computeUnits (optim.cpp:473-494)
template <class Scratch>
__device__ __forceinline__ int sel_sort_images(const SceneDev& s, const SelectParams& sp, Scratch& sc, int n, int lane, const float* coord,
                                               const float* normal) {
  // computeUnits: drop views with ray . normal <= 0, unit = getUnit / dot
  for (int i = lane; i < n; i += 32) {
    CamDev cam;
    load_cam(s, sc.images[i], cam);
    float ray[4] = {cam.centre[0] - coord[0], cam.centre[1] - coord[1], cam.centre[2] - coord[2], cam.centre[3] - coord[3]};
    unitize4(ray);
    const float d = dot4(ray, normal);
    sc.valid[i] = d > 0.0f ? 1 : 0;
    sc.val[i] = d > 0.0f ? get_unit(cam, s.level, coord) / d : 0.0f;
    sc.rays[i][0] = ray[0]; sc.rays[i][1] = ray[1]; sc.rays[i][2] = ray[2]; sc.rays[i][3] = ray[3];
  }
  __syncwarp();
  // compact images, units and rays together (order preserving); tmp_images reused as index map
  int m = 0;
  for (int base = 0; base < n; base += 32) {
    const int i = base + lane;
    const bool k = i < n && sc.valid[i];
    const unsigned b = __ballot_sync(kFull, k);
    if (k) sc.tmp_images[m + __popc(b & ((1u << lane) - 1u))] = i;
    m += __popc(b);
  }
  __syncwarp();
  // gather through registers (two elements per lane cover 64 views)
  int gi[2]; float gu[2], gr[2][4]; int gim[2];
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const int i = lane + 32 * t;
    gi[t] = i < m ? sc.tmp_images[i] : 0;
    gim[t] = sc.images[gi[t]]; gu[t] = sc.val[gi[t]];
#pragma unroll
    for (int k = 0; k < 4; ++k) gr[t][k] = sc.rays[gi[t]][k];
  }
  __syncwarp();
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const int i = lane + 32 * t;
    if (i < m) {
      sc.tmp_images[i] = gim[t]; sc.val[i] = gu[t];
#pragma unroll
      for (int k = 0; k < 4; ++k) sc.rays[i][k] = gr[t][k];
      sc.valid[i] = 1;   // = still in the candidate pool
    }
  }
  __syncwarp();
  if (m < 2) return 0;
  if (lane == 0) sc.val[0] = 0.0f;
  __syncwarp();
  // greedy: repeatedly take the first minimum, then penalise views within 10 degrees of it
  const float threshold = sp.sort_threshold;
  for (int out = 0; out < m; ++out) {
    float best = 3.0e38f; int bi = 1 << 30;
    for (int i = lane; i < m; i += 32)
      if (sc.valid[i] && (sc.val[i] < best)) { best = sc.val[i]; bi = i; }   // ascending i: keeps the first minimum per lane
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(kFull, best, o);
      const int oi = __shfl_xor_sync(kFull, bi, o);
      if (ob < best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    if (lane == 0) { sc.images[out] = sc.tmp_images[bi]; sc.valid[bi] = 0; }
    __syncwarp();
    const float r0 = sc.rays[bi][0], r1 = sc.rays[bi][1], r2 = sc.rays[bi][2], r3 = sc.rays[bi][3];
    for (int j = lane; j < m; j += 32) {
      if (!sc.valid[j]) continue;
      const float dt = r0 * sc.rays[j][0] + r1 * sc.rays[j][1] + r2 * sc.rays[j][2] + r3 * sc.rays[j][3];
      const float ftmp = smin(threshold, smax(threshold / 2.0f, 1.0f - dt));
      sc.val[j] = sc.val[j] * (threshold / ftmp);
    }
    __syncwarp();
  }
  return m;
}

// CPatchOrganizerS::setScales (patchOrganizerS.cpp:663-684); uniform across the warp
template <class Scratch>
__device__ __forceinline__ void sel_set_scales(const SceneDev& s, const Scratch& sc, int n, const float* coord, float& dscale, float& ascale) {
  CamDev cam;
  load_cam(s, sc.images[0], cam);
  const float unit = get_unit(cam, s.level, coord);
  const float unit2 = 2.0f * unit;
  float ray[4] = {coord[0] - cam.centre[0], coord[1] - cam.centre[1], coord[2] - cam.centre[2], coord[3] - cam.centre[3]};
  unitize4(ray);
  const int inum = s.tau < n ? s.tau : n;
  float ds = 0.0f;
  for (int i = 1; i < inum; ++i) {
    load_cam(s, sc.images[i], cam);
    float a[3], b[3];
    project(cam, coord, a);
    const float t[4] = {coord[0] - ray[0] * unit2, coord[1] - ray[1] * unit2, coord[2] - ray[2] * unit2, coord[3] - ray[3] * unit2};
    project(cam, t, b);
    const float d[3] = {a[0] - b[0], a[1] - b[1], a[2] - b[2]};
    ds += sqrtf(dot3(d, d));
  }
  ds /= (float)(inum - 1);
  ds = unit2 / ds;
  dscale = ds;
  ascale = (float)atan((double)(ds / (unit * (float)s.wsize / 2.0f)));
}

// CPhotoSetS::checkAngles (source/image/photoSetS.cpp:164-189): 1 = reject (no view pair inside (minAngle, maxAngle))
template <class Scratch>
__device__ __forceinline__ int sel_check_angles(const SceneDev& s, const SelectParams& sp, Scratch& sc, int n, int lane, const float* coord) {
  for (int i = lane; i < n; i += 32) {
    CamDev cam;
    load_cam(s, sc.images[i], cam);
    float ray[4] = {cam.centre[0] - coord[0], cam.centre[1] - coord[1], cam.centre[2] - coord[2], cam.centre[3] - coord[3]};
    unitize4(ray);
    sc.rays[i][0] = ray[0]; sc.rays[i][1] = ray[1]; sc.rays[i][2] = ray[2]; sc.rays[i][3] = ray[3];
  }
  __syncwarp();
  bool found = false;
  for (int pair = lane; pair < n * n; pair += 32) {
    const int i = pair / n, j = pair % n;
    if (j <= i) continue;
    float d = sc.rays[i][0] * sc.rays[j][0] + sc.rays[i][1] * sc.rays[j][1] + sc.rays[i][2] * sc.rays[j][2] + sc.rays[i][3] * sc.rays[j][3];
    d = smax(-1.0f, smin(1.0f, d));
    found |= (d >= sp.angle_dot_lo && d <= sp.angle_dot_hi);
  }
  return __any_sync(kFull, found) ? 0 : 1;
}

// COptim::filterImagesByAngle (optim.cpp:124-148)
template <class Scratch>
__device__ __forceinline__ int sel_filter_by_angle(const SceneDev& s, Scratch& sc, int n, int lane, const float* coord, const float* normal) {
  for (int i = lane; i < n; i += 32) {
    CamDev cam;
    load_cam(s, sc.images[i], cam);
    float ray[4] = {cam.centre[0] - coord[0], cam.centre[1] - coord[1], cam.centre[2] - coord[2], cam.centre[3] - coord[3]};
    unitize4(ray);
    sc.valid[i] = dot4(ray, normal) < s.cos_angle1 ? 0 : 1;   // same float-vs-double compare as grabTex's gate
  }
  __syncwarp();
  if (!sc.valid[0]) return 0;   // reference image dropped: the whole list goes
  return warp_compact(sc, n, lane, [&](int i) { return sc.valid[i] != 0; });
}

// COptim::setRefImage (optim.cpp:208-254): robust all-pairs matrix (setINCCs matrix form, optim.cpp:746-781) over
// the TARGET images of the patch; the reference image becomes the one with the smallest row sum and is swapped
// into slot 0.  Returns the list length (0 when the patch has no target image: the reference clears the list).
template <int WSIZE, int MAXV>
__device__ __forceinline__ int sel_set_ref_image(const SceneDev& s, SelScratch<WSIZE, MAXV>& sc, int n, int lane, const float* coord,
                                                 const float* normal) {
  for (int i = lane; i < n; i += 32) sc.aux[i] = sc.images[i];
  __syncwarp();
  const int m = warp_compact(sc, n, lane, [&](int i) { return sc.images[i] < s.tnum; });   // images := indexes
  int refimg = -1;
  if (m > 0) {
    sel_grab_all<WSIZE>(s, sc, m, lane, coord, normal);
    for (int i = lane; i < m; i += 32) {
      float sum = 0.0f;
      for (int j = 0; j < m; ++j) {
        float v = 0.0f;
        if (j != i) v = (sc.valid[i] && sc.valid[j]) ? robustincc(1.0f - lane_dot<WSIZE>(sc, i < j ? i : j, i < j ? j : i)) : 2.0f;
        sum += v;
      }
      sc.val[i] = sum;
    }
    __syncwarp();
    float best = 1073741824.0f;  // INT_MAX / 2 as float (optim.cpp:236)
    int bi = 1 << 30;
    for (int i = lane; i < m; i += 32)
      if (sc.val[i] < best) { best = sc.val[i]; bi = i; }
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(kFull, best, o);
      const int oi = __shfl_xor_sync(kFull, bi, o);
      if (ob < best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    // no row below INT_MAX/2 cannot happen for finite sums; mirror the reference's refindex = -1 hazard by keeping slot 0
    refimg = bi < m ? sc.images[bi] : sc.aux[0];
  }
  __syncwarp();
  for (int i = lane; i < n; i += 32) sc.images[i] = sc.aux[i];
  __syncwarp();
  if (m == 0) return 0;
  if (lane == 0) {
    for (int i = 0; i < n; ++i)
      if (sc.images[i] == refimg) { const int t = sc.images[0]; sc.images[0] = refimg; sc.images[i] = t; break; }
  }
  __syncwarp();
  return n;
}

// CPatchOrganizerS::setGrids (patchOrganizerS.cpp:400-414): cell of the patch in each of its images
template <class Scratch>
__device__ __forceinline__ void sel_set_grids(const SceneDev& s, const Scratch& sc, int n, int lane, const float* coord, int32_t* grids) {
  for (int i = lane; i < n; i += 32) {
    CamDev cam;
    load_cam(s, sc.images[i], cam);
    float ic[3];
    project(cam, coord, ic);
    grids[2 * i] = ((int)floorf(ic[0] + 0.5f)) / s.csize;
    grids[2 * i + 1] = ((int)floorf(ic[1] + 0.5f)) / s.csize;
  }
}

// COptim::preProcess (optim.cpp:95-122).  Returns the verdict (0 keep / 1 reject); n, dscale, ascale updated.
template <int WSIZE, int MAXV>
__device__ __forceinline__ int sel_pre_process(const SceneDev& s, const SelectParams& sp, SelScratch<WSIZE, MAXV>& sc, int& n, int cap, int lane,
                                               const float* coord, const float* normal, float& dscale, float& ascale) {
  dscale = 0.0f; ascale = 0.0f;
  n = sel_add_images(s, sp, sc, n, cap, lane, coord, normal);
  n = sel_constraint_images<WSIZE>(s, sc, n, lane, coord, normal, sp.ncc_threshold_before);
  n = sel_sort_images(s, sp, sc, n, lane, coord, normal);
  if (n > 0) sel_set_scales(s, sc, n, coord, dscale, ascale);
  if (n < s.min_image_num) return 1;
  if (sel_check_angles(s, sp, sc, n, lane, coord)) { n = 0; return 1; }
  return 0;
}

// COptim::postProcess (optim.cpp:150-190) up to _tmp = score2 (setVImagesVGrids and check are separate calls).
template <int WSIZE, int MAXV>
__device__ __forceinline__ int sel_post_process(const SceneDev& s, const SelectParams& sp, SelScratch<WSIZE, MAXV>& sc, int& n, int cap, int lane,
                                                const float* coord, const float* normal, float ncc, int32_t* grids, int& timages, float& tmp) {
  timages = 0; tmp = 0.0f;
  if (n < s.min_image_num) return 1;
  if (!mask_gate_warp(s, coord, lane)) return 1;   // optim.cpp:153
  n = sel_add_images(s, sp, sc, n, cap, lane, coord, normal);
  n = sel_constraint_images<WSIZE>(s, sc, n, lane, coord, normal, sp.ncc_threshold);
  n = sel_filter_by_angle(s, sc, n, lane, coord, normal);
  if (n < s.min_image_num) return 1;
  n = sel_set_ref_image<WSIZE>(s, sc, n, lane, coord, normal);
  if (n == 0) return 1;
  n = sel_constraint_images<WSIZE>(s, sc, n, lane, coord, normal, sp.ncc_threshold);
  if (n < s.min_image_num) return 1;
  sel_set_grids(s, sc, n, lane, coord, grids);
  int t = 0;
  for (int i = lane; i < n; i += 32) t += sc.images[i] < s.tnum ? 1 : 0;
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(kFull, t, o);
  timages = t;
  tmp = smax(0.0f, ncc - sp.ncc_threshold) * (float)t;   // CPatch::score2 (include/pmvs/patch.hpp:48-50)
  return 0;
}

}  // namespace pmvsb
