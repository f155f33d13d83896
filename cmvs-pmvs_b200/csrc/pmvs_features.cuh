// cmvs-pmvs_b200/csrc/pmvs_features.cuh
//
// Feature detection on the working-level image: CHarris::run (source/pmvs/harris.cpp:174-240) and
// CDifferenceOfGaussians::run (source/pmvs/dog.cpp:96-198) with the separable convolutions of CDetector
// (include/pmvs/detector.hpp:23-94, unmasked images: coordinates clamp to the image).  One thread per output pixel
// walks the filter taps in the reference's order with plain FMUL + FADD (-fmad=false), so every plane is bit-exact;
// one warp per 2*gspeedup-pixel block replays the reference's per-block multiset<CPoint> (4 strongest points).
// All paths relative to /root/reference.
#pragma once
#include "pmvs_device.cuh"

namespace pmvsb {

// RGBA8 level -> three float planes, ((int)byte) / 255.0f (harris.cpp:17-19)
__global__ void k_feat_planes(LevelDev lv, float* __restrict__ planes) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = lv.w * lv.h;
  if (i >= n) return;
  const uchar4 p = lv.pix[i];
  planes[i] = fdiv((float)(int)p.x, 255.0f);
  planes[n + i] = fdiv((float)(int)p.y, 255.0f);
  planes[2 * n + i] = fdiv((float)(int)p.z, 255.0f);
}

// CDetector::convolveX / convolveY, the overloads that take a mask (detector.hpp:26-94: coordinates clamp to the image;
// a masked-out pixel yields 0 and masked-out taps are skipped); mask = null for an image without mask and edge map.
// blockIdx.z selects the plane.
template <bool VERTICAL>
__global__ void k_feat_conv(const float* __restrict__ src, float* __restrict__ dst, int w, int h, const float* __restrict__ taps, int ntaps,
                            const unsigned char* __restrict__ mask) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= w) return;
  const size_t plane = (size_t)blockIdx.z * w * h;
  const float* s = src + plane;
  const int margin = ntaps / 2;
  float acc = 0.0f;
  if (!mask || mask[(size_t)y * w + x] != 0)
    for (int j = 0; j < ntaps; ++j) {
      int xt = x, yt = y;
      if (VERTICAL) { yt = y + j - margin; yt = yt < 0 ? 0 : (h <= yt ? h - 1 : yt); }
      else { xt = x + j - margin; xt = xt < 0 ? 0 : (w <= xt ? w - 1 : xt); }
      if (mask && mask[(size_t)yt * w + xt] == 0) continue;
      acc += __ldg(taps + j) * __ldg(s + (size_t)yt * w + xt);
    }
  dst[plane + (size_t)y * w + x] = acc;
}
// CHarris::init / CDifferenceOfGaussians::init (harris.cpp:22-43, dog.cpp:240-261): mask AND edge where both exist
__global__ void k_feat_mask(const unsigned char* __restrict__ a, const unsigned char* __restrict__ b, int n, unsigned char* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = !a ? b[i] : (!b ? a[i] : ((a[i] && b[i]) ? 255 : 0));
}

// CHarris::preprocess2's products (harris.cpp:60-79): Vec3f * Vec3f summed over the channels left to right
__global__ void k_feat_products(const float* __restrict__ dx, const float* __restrict__ dy, int n, float* __restrict__ out,
                                const unsigned char* __restrict__ mask) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (mask && mask[i] == 0) { out[i] = 0.0f; out[n + i] = 0.0f; out[2 * n + i] = 0.0f; return; }   // harris.cpp:74
  const float a0 = dx[i], a1 = dx[n + i], a2 = dx[2 * n + i];
  const float b0 = dy[i], b1 = dy[n + i], b2 = dy[2 * n + i];
  out[i] = 0.0f + ((a0 * a0 + a1 * a1) + a2 * a2);
  out[n + i] = 0.0f + ((b0 * b0 + b1 * b1) + b2 * b2);
  out[2 * n + i] = 0.0f + ((a0 * b0 + a1 * b1) + a2 * b2);
}

// CHarris::setResponse (harris.cpp:139-172): D - 0.06 tr^2 in double, then the 4-neighbour non-maximum suppression
__global__ void k_feat_harris_response(const float* __restrict__ m, int n, float* __restrict__ resp, const unsigned char* __restrict__ mask) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (mask && mask[i] == 0) { resp[i] = 0.0f; return; }   // harris.cpp:147
  const float xx = m[i], yy = m[n + i], xy = m[2 * n + i];
  const float D = xx * yy - xy * xy;
  const float tr = xx + yy;
  resp[i] = (float)((double)D - 0.06 * (double)tr * (double)tr);
}
__global__ void k_feat_nms(const float* __restrict__ resp, int w, int h, float* __restrict__ out) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= w) return;
  const size_t o = (size_t)y * w + x;
  float v = resp[o];
  if (y >= 1 && y < h - 1 && x >= 1 && x < w - 1)
    if (v < resp[o + 1] || v < resp[o - 1] || v < resp[o + w] || v < resp[o - w]) v = 0.0f;
  out[o] = v;
}

// CDifferenceOfGaussians::setRes' norm (dog.cpp:200-222) and setDOG (82-94)
__global__ void k_feat_norm(const float* __restrict__ c, int n, float* __restrict__ res) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float a = c[i], b = c[n + i], d = c[2 * n + i];
  res[i] = fsqrt((a * a + b * b) + d * d);
}
__global__ void k_feat_sub(const float* __restrict__ next, const float* __restrict__ cur, int n, float* __restrict__ dog) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  dog[i] = next[i] - cur[i];
}

// the per-block multiset<CPoint>: ascending by response, equal keys in insertion order, begin() dropped beyond four
struct FeatTop {
  float r[5];
  int x[5], y[5];
  int n;
  __device__ __forceinline__ void insert(float v, int px, int py) {   // n <= 4 on entry
    int pos = n;
    while (pos > 0 && v < r[pos - 1]) { r[pos] = r[pos - 1]; x[pos] = x[pos - 1]; y[pos] = y[pos - 1]; --pos; }
    r[pos] = v; x[pos] = px; y[pos] = py;
    ++n;
    if (n > 4) {
      for (int k = 1; k < 5; ++k) { r[k - 1] = r[k]; x[k - 1] = x[k]; y[k - 1] = y[k]; }
      n = 4;
    }
  }
};

// One warp per feature block.  MODE 0: Harris (harris.cpp:196-222: candidates are the non-zero suppressed responses,
// inserted only while the block has room or beats its weakest point).  MODE 1: DoG (dog.cpp:150-183: scales i = 2, 3 in
// turn; a pixel is an extremum of cdog against its 8 neighbours and against pdog / ndog; detected pixels are skipped at
// the next scale; every extremum is inserted, the weakest dropped).
template <int MODE>
__global__ void k_feat_select(const float* __restrict__ p0, const float* __restrict__ p1, const float* __restrict__ p2,
                              const float* __restrict__ p3, int w, int h, int gridsize, int gw, int gh, int margin_a, int margin_b,
                              unsigned char* __restrict__ seen, float* __restrict__ out_r, int* __restrict__ out_xy, int* __restrict__ out_n) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= gw * gh) return;
  const int bx = warp % gw, by = warp / gw;
  FeatTop top;
  top.n = 0;
#pragma unroll
  for (int k = 0; k < 5; ++k) { top.r[k] = 0.0f; top.x[k] = 0; top.y[k] = 0; }
  const int x_end = min((bx + 1) * gridsize, w), y_end = min((by + 1) * gridsize, h);
  const int passes = MODE == 0 ? 1 : 2;
  for (int pass = 0; pass < passes; ++pass) {
    const int margin = pass == 0 ? margin_a : margin_b;
    const float* pd = pass == 0 ? p0 : p1;
    const float* cd = pass == 0 ? p1 : p2;
    const float* nd = pass == 0 ? p2 : p3;
    for (int y = by * gridsize; y < y_end; ++y) {
      if (y < margin || y >= h - margin) continue;
      for (int xb = bx * gridsize; xb < x_end; xb += 32) {
        const int x = xb + lane;
        const bool inside = x < x_end && x >= margin && x < w - margin;
        float v = 0.0f;
        bool cand = false;
        if (inside) {
          const size_t o = (size_t)y * w + x;
          if (MODE == 0) {
            v = p0[o];
            cand = v != 0.0f;
          } else {
            const float c = cd[o];
            if (!seen[o] && c != 0.0f) {
              bool ext;
              if (0.0f < c)
                ext = cd[o - w - 1] < c && cd[o - 1] < c && cd[o + w - 1] < c && cd[o - w] < c && cd[o + w] < c && cd[o - w + 1] < c &&
                      cd[o + 1] < c && cd[o + w + 1] < c && pd[o] < c && nd[o] < c;
              else
                ext = cd[o - w - 1] > c && cd[o - 1] > c && cd[o + w - 1] > c && cd[o - w] > c && cd[o + w] > c && cd[o - w + 1] > c &&
                      cd[o + 1] > c && cd[o + w + 1] > c && c < pd[o] && c < nd[o];
              if (ext) { seen[o] = 1; v = fabsf(c); cand = true; }
            }
          }
        }
        unsigned m = __ballot_sync(kFull, cand);
        while (m) {   // ascending x: the reference's scan order
          const int l = __ffs(m) - 1;
          m &= m - 1;
          const float vv = __shfl_sync(kFull, v, l);
          if (MODE == 1 || top.n < 4 || top.r[0] < vv) top.insert(vv, xb + l, y);
        }
      }
    }
    __syncwarp();
  }
  if (lane == 0) {
    out_n[warp] = top.n;
    for (int k = 0; k < top.n; ++k) { out_r[warp * 4 + k] = top.r[k]; out_xy[(warp * 4 + k) * 2] = top.x[k]; out_xy[(warp * 4 + k) * 2 + 1] = top.y[k]; }
  }
}

}  // namespace pmvsb
