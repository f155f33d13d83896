// cmvs-pmvs_b200/csrc/pmvs_seed.cuh
//
// Seed candidate enumeration: CSeed::collectCells + collectCandidates + unproject
// (/root/reference/source/pmvs/seed.cpp:207-384) for every feature of one reference image in one launch.
//
// One warp owns one feature p0 of the reference image.  For each of the <= tau images picked by collectImages the lanes walk
// the epipolar line of p0 through that image's cell grid (three cells per step), test the features binned there -- same
// detector type, point-to-epipolar-line distance below _epThreshold -- and triangulate the survivors.  The arithmetic follows
// the reference operation by operation: the epipolar geometry and the 3x3 normal equations in double (numeric/mat3.hpp,
// mat4.hpp orderings), the entries of the 4x3 system and everything after the triangulation in float, so the candidate SET and
// every candidate's coordinates equal the reference's bit for bit.  The ORDER of a feature's candidates is the one the
// reference documents ("from the closest": ascending |dist to camera 0 - dist to camera 1|, ties in enumeration order); the
// reference itself sorts shared_ptr values, i.e. by allocation address (seed.cpp:322), which no restatement can follow.
#pragma once
#include "pmvs_device.cuh"

namespace pmvsb {

constexpr int kSeedCap = 512;         // candidates kept per feature (more would be an error reported to the caller)
constexpr int kSeedWarps = 2;         // warps per CTA (shared memory: kSeedWarps * kSeedCap * 32 B)

struct SeedDev {                      // features of every image, binned by cell (CSeed::_ppoints, seed.cpp:25-36)
  const float* fxy;                   // feature coordinates (x, y) at the working level, all images back to back
  const int32_t* ftype;               // 0 Harris / 1 DoG
  const int32_t* fcell_base;          // per image: first cell of its grid in fcell_off (all images, targets and others)
  const int32_t* fcell_off;           // CSR over the flattened cells: features of a cell in detection order
  const int32_t* flist;               // feature ids (global)
  const int32_t* gw;                  // grid width / height per image
  const int32_t* gh;
};

struct SeedView {                     // one image of collectImages' list, relative to the reference image
  double F[9];                        // Image::setF(reference, other, level) (include/image/camera.hpp:129-151), row-major
  int image;
  int pad;
};
struct SeedParams {
  SeedView view[kMaxTau];
  int index, nviews;
  float ep_threshold;                 // CFindMatch::_epThreshold (findMatch.cpp:106)
};

struct SeedHit {                      // 32 B
  float x, y, z, resp;
  int other_image, other_feature;
  unsigned key_hi, key_lo;            // enumeration order: (view, step along the line, which of the three cells, feature in cell)
};

// the point gate of collectCandidates (seed.cpp:314) evaluated by one thread
__device__ __forceinline__ bool mask_gate_thread(const SceneDev& s, const float* X) {
  if (s.mask_lv)
    for (int i = 0; i < s.num; ++i)
      if (get_mask_img(s, i, X) == 0) return false;
  for (int i = 0; i < s.n_bimages; ++i)
    if (inside_bimage(s, s.bimages[i], X) == 0) return false;
  return true;
}

// CSeed::unproject (seed.cpp:340-384): rows of the 4x3 system in FLOAT (Vec4f / Vec3f operands), normal equations and the
// adjoint inverse in double (mat4.hpp:364-376, mat3.hpp:275-292)
__device__ __forceinline__ void seed_unproject(const CamDev& c0, const CamDev& c1, float x0, float y0, float x1, float y1, float* coord) {
  double A[4][3], b[4];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    A[0][k] = (double)(c0.P[0][k] - x0 * c0.P[2][k]);
    A[1][k] = (double)(c0.P[1][k] - y0 * c0.P[2][k]);
    A[2][k] = (double)(c1.P[0][k] - x1 * c1.P[2][k]);
    A[3][k] = (double)(c1.P[1][k] - y1 * c1.P[2][k]);
  }
  b[0] = (double)(x0 * c0.P[2][3] - c0.P[0][3]);
  b[1] = (double)(y0 * c0.P[2][3] - c0.P[1][3]);
  b[2] = (double)(x1 * c1.P[2][3] - c1.P[0][3]);
  b[3] = (double)(y1 * c1.P[2][3] - c1.P[1][3]);
  double M[3][3], v[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int j = 0; j < 3; ++j) M[i][j] = A[0][i] * A[0][j] + A[1][i] * A[1][j] + A[2][i] * A[2][j] + A[3][i] * A[3][j];
    v[i] = A[0][i] * b[0] + A[1][i] * b[1] + A[2][i] * b[2] + A[3][i] * b[3];
  }
  // adjoint rows: m1 ^ m2, m2 ^ m0, m0 ^ m1 with u ^ v = (u1 v2 - v1 u2, -u0 v2 + v0 u2, u0 v1 - v0 u1)
  double adj[3][3];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const double* u = M[(r + 1) % 3];
    const double* w = M[(r + 2) % 3];
    adj[r][0] = u[1] * w[2] - w[1] * u[2];
    adj[r][1] = -u[0] * w[2] + w[0] * u[2];
    adj[r][2] = u[0] * w[1] - w[0] * u[1];
  }
  const double d = adj[0][0] * M[0][0] + adj[0][1] * M[0][1] + adj[0][2] * M[0][2];
  // d == 0: the reference leaves iATA3 as constructed (zeros), ans = 0
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    double ans = 0.0;
    if (d != 0.0) ans = (adj[0][i] / d) * v[0] + (adj[1][i] / d) * v[1] + (adj[2][i] / d) * v[2];
    coord[i] = (float)ans;
  }
  coord[3] = 1.0f;
}

// blocked[cell] != 0  <=>  CSeed::canAdd(image, x, y) == 0 (seed.cpp:325-338), flattened like fcell_off.
// For reference feature r (global id ref_feats[r]): up to kSeedCap hits, sorted, written to out[start[r] ..), count[r] hits.
__global__ void __launch_bounds__(kSeedWarps * 32) k_seed_candidates(SceneDev s, SeedDev sd, SeedParams sp, const uint8_t* __restrict__ blocked, int nref,
                                                                      const int32_t* __restrict__ ref_feats, int32_t* __restrict__ count,
                                                                      int32_t* __restrict__ start, int32_t* __restrict__ cursor, int capacity,
                                                                      SeedHit* __restrict__ out, int32_t* __restrict__ overflow) {
  __shared__ SeedHit hits[kSeedWarps][kSeedCap];
  __shared__ int nhits[kSeedWarps];
  __shared__ int base[kSeedWarps];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * kSeedWarps + wib;
  if (r >= nref) return;
  if (lane == 0) nhits[wib] = 0;
  __syncwarp();
  const int f0 = ref_feats[r];
  const float x0 = sd.fxy[2 * f0], y0 = sd.fxy[2 * f0 + 1];
  const int type0 = sd.ftype[f0];
  CamDev c0;
  load_cam(s, sp.index, c0);
  const double px = (double)x0, py = (double)y0;
  for (int k = 0; k < sp.nviews; ++k) {
    const double* F = sp.view[k].F;
    const int other = sp.view[k].image;
    // collectCells (seed.cpp:207-268): line = transpose(F) * (x0, y0, 1)
    const double l0 = F[0] * px + F[3] * py + F[6] * 1.0;
    const double l1 = F[1] * px + F[4] * py + F[7] * 1.0;
    const double l2 = F[2] * px + F[5] * py + F[8] * 1.0;
    if (l0 == 0.0 && l1 == 0.0) continue;
    const bool vertical = fabs(l0) > fabs(l1);
    const int gw = sd.gw[other], gh = sd.gh[other];
    const int steps = vertical ? gh : gw;
    CamDev c1;
    load_cam(s, other, c1);
    for (int t = lane; t < steps; t += 32) {
      const float fa = (float)(((double)t + 0.5) * (double)s.csize - (double)0.5f);
      float fb = vertical ? (float)((-l1 * (double)fa - l2) / l0) : (float)((-l0 * (double)fa - l2) / l1);
      fb = smax(-2147483648.0f, smin(2147483648.0f, fb));
      const int ib = ((int)floor((double)(fb + 0.5f))) / s.csize;
#pragma unroll
      for (int sub = 0; sub < 3; ++sub) {
        const int cb = sub == 0 ? ib : (sub == 1 ? ib - 1 : ib + 1);
        const int cx = vertical ? cb : t, cy = vertical ? t : cb;
        if (cx < 0 || gw <= cx || cy < 0 || gh <= cy) continue;
        const int cell = sd.fcell_base[other] + cy * gw + cx;
        if (blocked[cell]) continue;
        for (int j = sd.fcell_off[cell]; j < sd.fcell_off[cell + 1]; ++j) {
          const int f1 = sd.flist[j];
          if (sd.ftype[f1] != type0) continue;
          const float x1 = sd.fxy[2 * f1], y1 = sd.fxy[2 * f1 + 1];
          // computeEPD(F, p0, p1) (camera.hpp:118-127): distance of p0 to the epipolar line F * p1, returned as float
          const double qx = (double)x1, qy = (double)y1;
          double e0 = F[0] * qx + F[1] * qy + F[2] * 1.0;
          double e1 = F[3] * qx + F[4] * qy + F[5] * 1.0;
          double e2 = F[6] * qx + F[7] * qy + F[8] * 1.0;
          const double nn = sqrt(e0 * e0 + e1 * e1);
          float epd = 0.0f;
          if (nn != 0.0) { e0 /= nn; e1 /= nn; e2 /= nn; epd = (float)fabs(e0 * px + e1 * py + e2 * 1.0); }
          if (sp.ep_threshold <= epd) continue;
          float X[4];
          seed_unproject(c0, c1, x0, y0, x1, y1, X);
          if (dot4(c0.P[2], X) <= 0.0f) continue;          // behind the reference camera (seed.cpp:313)
          if (!mask_gate_thread(s, X)) continue;            // seed.cpp:314
          const float d0[4] = {X[0] - c0.centre[0], X[1] - c0.centre[1], X[2] - c0.centre[2], X[3] - c0.centre[3]};
          const float d1[4] = {X[0] - c1.centre[0], X[1] - c1.centre[1], X[2] - c1.centre[2], X[3] - c1.centre[3]};
          const float resp = fabsf(fsqrt(dot4(d0, d0)) - fsqrt(dot4(d1, d1)));
          const int slot = atomicAdd(&nhits[wib], 1);
          if (slot < kSeedCap) {
            SeedHit h;
            h.x = X[0]; h.y = X[1]; h.z = X[2]; h.resp = resp;
            h.other_image = other; h.other_feature = f1;
            h.key_hi = ((unsigned)k << 20) | (unsigned)t;
            h.key_lo = ((unsigned)sub << 24) | (unsigned)(j - sd.fcell_off[cell]);
            hits[wib][slot] = h;
          }
        }
      }
    }
    __syncwarp();
  }
  __syncwarp();
  int n = nhits[wib];
  if (n > kSeedCap) { if (lane == 0) atomicAdd(overflow, 1); n = kSeedCap; }
  if (lane == 0) base[wib] = n > 0 ? atomicAdd(cursor, n) : 0;
  __syncwarp();
  const int b = base[wib];
  if (lane == 0) { count[r] = (b + n <= capacity) ? n : 0; start[r] = b; if (b + n > capacity) atomicAdd(overflow + 1, 1); }
  if (b + n > capacity) return;
  // rank sort by (resp, enumeration key): rank = number of hits that come before
  for (int i = lane; i < n; i += 32) {
    const SeedHit h = hits[wib][i];
    int rank = 0;
    for (int j = 0; j < n; ++j) {
      const SeedHit& o = hits[wib][j];
      const bool before = o.resp < h.resp || (o.resp == h.resp && (o.key_hi < h.key_hi || (o.key_hi == h.key_hi && o.key_lo < h.key_lo)));
      rank += before ? 1 : 0;
    }
    out[b + rank] = h;
  }
}

}  // namespace pmvsb
