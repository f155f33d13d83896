// What bounds the texture pipe of k_refine_g on B200: does a TLD4 cost per ACTIVE lane, per active quad or per instruction,
// and what do the alternatives to three one-channel gathers per sample cost (one 16-byte footprint texel, one hardware
// bilinear fetch, gathers for some rows + global loads for the others)?  Window-like locality as in tex_rate_probe.cu.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/probe/tex_lane_probe tools/probe/tex_lane_probe.cu
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s failed: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

enum { ALL32, SKIP_LANE7, SKIP_QUAD7, HALF_WARP, FOOTPRINT16, HW_BILINEAR, MIX_TLD4_LDG, LDG_ONLY, ONE_WINDOW, LANE_WINDOWS, NEAR4, NEAR8, NEAR16, NEAR32, NMODES };
static const char* kNames[NMODES] = {
    "TLD4 x3, 32 lanes            ", "TLD4 x3, lanes 7/15/23/31 off", "TLD4 x3, lanes 28..31 off    ", "TLD4 x3, lanes 16..31 off    ",
    "1 point fetch of a 16 B texel", "1 hardware-bilinear fetch    ", "rows 0-3 TLD4 x3, 4-6 LDG x4 ", "LDG x4                       ",
    "TLD4 x3, ONE window per warp ", "TLD4 x3, a window per LANE   ",
    "TLD4 x3, 4 windows 4 px apart", "TLD4 x3, 4 windows 8 px apart", "TLD4 x3, 4 windows 16 px apart", "TLD4 x3, 4 windows 32 px apart"};

template <int MODE>
__global__ void __launch_bounds__(128, 8) k(cudaTextureObject_t gather, cudaTextureObject_t foot, cudaTextureObject_t lin, const uint32_t* pix,
                                            int W, int H, int iters, float* out) {
  const int lane = threadIdx.x & 31;
  // SKIP_QUAD7 packs four 7-lane groups into lanes 0..27; every other mode uses 8-lane groups
  const int g = MODE == SKIP_QUAD7 ? lane / 7 : lane >> 3, gl = MODE == SKIP_QUAD7 ? lane % 7 : lane & 7;
  const bool active = MODE == SKIP_LANE7 ? gl != 7 : MODE == SKIP_QUAD7 ? lane < 28 : MODE == HALF_WARP ? lane < 16 : true;
  // locality of a warp's 32 footprints: four windows (the kernel's layout), ONE window (4 rows x 8 columns of it at a time), or 32
  unsigned seed = (blockIdx.x * 4 + (threadIdx.x >> 5)) * 4 + (MODE == ONE_WINDOW ? 0 : (g & 3));
  if (MODE == LANE_WINDOWS) seed = seed * 32 + lane;
  // NEARd: the warp's four windows sit on a 2 x 2 grid d pixels apart (what a warp-local, locality-ordered hand-out would give)
  constexpr bool kNear = MODE == NEAR4 || MODE == NEAR8 || MODE == NEAR16 || MODE == NEAR32;
  constexpr float kD = MODE == NEAR4 ? 4.f : MODE == NEAR8 ? 8.f : MODE == NEAR16 ? 16.f : 32.f;
  if (kNear) seed = (blockIdx.x * 4 + (threadIdx.x >> 5)) * 4;
  float acc = 0.f;
  for (int it = 0; it < iters; ++it) {
    if (it % 150 == 0) seed = seed * 1664525u + 1013904223u;   // a window is re-sampled ~150 times (one Nelder-Mead run)
    float bx = 8.0f + (float)((seed >> 8) % (W - 96)), by = 8.0f + (float)((seed >> 20) % (H - 96));
    if (kNear) { bx += (float)(g & 1) * kD; by += (float)(g >> 1) * kD; }
    for (int row = 0; row < 7; ++row) {
      const float jit = (float)(it & 7) * 0.05f;
      // ONE_WINDOW: the warp's four groups take four rows of one window.  The sub-pixel term keeps the seven iterations' coordinates
      // distinct: without it (first run, profiles/r1_tex_lane_probe.txt: 7.4 clk) rows 0/2/4/6 and 1/3/5 were identical requests and the
      // compiler merged them, so that figure is 2/7 of the four-window cost and says nothing about the hardware
      const float ry = MODE == ONE_WINDOW ? (float)((row & 1) * 4 + g) + (float)(row >> 1) * 0.13f : (float)row;
      const float x = bx + gl * 0.93f + ry * 0.11f + jit, y = by + ry * 0.97f + gl * 0.07f + jit;
      const float lxf = truncf(x), lyf = truncf(y);
      if (!active) continue;
      if (MODE == FOOTPRINT16) {
        const uint4 q = tex2D<uint4>(foot, lxf + 0.5f, lyf + 0.5f);
        acc += __uint_as_float((q.x ^ q.y ^ q.z ^ q.w) & 0x3fffffffu);
      } else if (MODE == HW_BILINEAR) {
        const float4 q = tex2D<float4>(lin, x + 0.5f, y + 0.5f);
        acc += q.x + q.y + q.z;
      } else if (MODE == LDG_ONLY || (MODE == MIX_TLD4_LDG && row >= 4)) {
        const uint32_t* p = pix + ((int)lyf * W + (int)lxf);
        const uint32_t a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + W), d = __ldg(p + W + 1);
        acc += __uint_as_float((a ^ b ^ c ^ d) & 0x3fffffffu);
      } else {
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
          const float4 q = tex2Dgather<float4>(gather, lxf + 1.0f, lyf + 1.0f, ch);
          acc += (q.x + q.y) + (q.z + q.w);
        }
      }
    }
  }
  if (acc == 1234.5f) out[0] = acc;
}

template <int MODE>
void run(cudaTextureObject_t gather, cudaTextureObject_t foot, cudaTextureObject_t lin, const uint32_t* pix, int W, int H, float* d, int iters) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<148 * 8, 128>>>(gather, foot, lin, pix, W, H, 8, d);
  cudaEventRecord(e0);
  k<MODE><<<148 * 8, 128>>>(gather, foot, lin, pix, W, H, iters, d);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double warp_rows = 8.0 * 4 * iters * 7;   // per SM
  printf("%s: %7.2f ms, %5.1f clk per warp-row per SM (1.965 GHz)\n", kNames[MODE], ms, ms * 1e-3 * 1.965e9 / warp_rows);
}

int main(int argc, char**) {
  const int W = 6400, H = 4800;   // atlas-sized
  std::vector<uchar4> img((size_t)W * H);
  for (size_t i = 0; i < img.size(); ++i) img[i] = make_uchar4(i * 7, i * 13, i * 3, 0);
  cudaChannelFormatDesc cd = cudaCreateChannelDesc<uchar4>();
  cudaArray_t arr;
  CK(cudaMallocArray(&arr, &cd, W, H, cudaArrayTextureGather));
  CK(cudaMemcpy2DToArray(arr, 0, 0, img.data(), W * 4, W * 4, H, cudaMemcpyHostToDevice));
  uint32_t* pix; CK(cudaMalloc(&pix, img.size() * 4)); CK(cudaMemcpy(pix, img.data(), img.size() * 4, cudaMemcpyHostToDevice));
  cudaResourceDesc rd = {}; rd.resType = cudaResourceTypeArray; rd.res.array.array = arr;
  cudaTextureDesc td = {}; td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp; td.filterMode = cudaFilterModePoint;
  td.readMode = cudaReadModeNormalizedFloat; td.normalizedCoords = 0;
  cudaTextureObject_t gather; CK(cudaCreateTextureObject(&gather, &rd, &td, nullptr));
  td.filterMode = cudaFilterModeLinear;
  cudaTextureObject_t lin; CK(cudaCreateTextureObject(&lin, &rd, &td, nullptr));
  // footprint atlas: texel (x, y) holds the RGBA words of (x,y), (x+1,y), (x,y+1), (x+1,y+1)
  std::vector<uint4> fimg((size_t)W * H);
  const uint32_t* w32 = reinterpret_cast<const uint32_t*>(img.data());
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      const int x1 = x + 1 < W ? x + 1 : x, y1 = y + 1 < H ? y + 1 : y;
      fimg[(size_t)y * W + x] = make_uint4(w32[(size_t)y * W + x], w32[(size_t)y * W + x1], w32[(size_t)y1 * W + x], w32[(size_t)y1 * W + x1]);
    }
  cudaChannelFormatDesc cf = cudaCreateChannelDesc<uint4>();
  cudaArray_t farr;
  CK(cudaMallocArray(&farr, &cf, W, H));
  CK(cudaMemcpy2DToArray(farr, 0, 0, fimg.data(), (size_t)W * 16, (size_t)W * 16, H, cudaMemcpyHostToDevice));
  cudaResourceDesc frd = {}; frd.resType = cudaResourceTypeArray; frd.res.array.array = farr;
  cudaTextureDesc ftd = {}; ftd.addressMode[0] = ftd.addressMode[1] = cudaAddressModeClamp; ftd.filterMode = cudaFilterModePoint;
  ftd.readMode = cudaReadModeElementType; ftd.normalizedCoords = 0;
  cudaTextureObject_t foot; CK(cudaCreateTextureObject(&foot, &frd, &ftd, nullptr));
  float* d; CK(cudaMalloc(&d, 4));
  const int iters = 2000;
  if (argc > 1) {   // "near": only the locality sweep
    run<ALL32>(gather, foot, lin, pix, W, H, d, iters);
    run<NEAR4>(gather, foot, lin, pix, W, H, d, iters); run<NEAR8>(gather, foot, lin, pix, W, H, d, iters);
    run<NEAR16>(gather, foot, lin, pix, W, H, d, iters); run<NEAR32>(gather, foot, lin, pix, W, H, d, iters);
    run<ONE_WINDOW>(gather, foot, lin, pix, W, H, d, iters);
    CK(cudaDeviceSynchronize());
    return 0;
  }
  run<ALL32>(gather, foot, lin, pix, W, H, d, iters);
  run<SKIP_LANE7>(gather, foot, lin, pix, W, H, d, iters);
  run<SKIP_QUAD7>(gather, foot, lin, pix, W, H, d, iters);
  run<HALF_WARP>(gather, foot, lin, pix, W, H, d, iters);
  run<FOOTPRINT16>(gather, foot, lin, pix, W, H, d, iters);
  run<HW_BILINEAR>(gather, foot, lin, pix, W, H, d, iters);
  run<MIX_TLD4_LDG>(gather, foot, lin, pix, W, H, d, iters);
  run<LDG_ONLY>(gather, foot, lin, pix, W, H, d, iters);
  run<ONE_WINDOW>(gather, foot, lin, pix, W, H, d, iters);
  run<LANE_WINDOWS>(gather, foot, lin, pix, W, H, d, iters);
  CK(cudaDeviceSynchronize());
  return 0;
}
