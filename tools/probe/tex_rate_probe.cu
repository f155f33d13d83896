// TLD4 (tex2Dgather) issue rate on B200 with window-like locality, alone and mixed with FMA work.
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s failed: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
template <int NFMA, bool LDG>
__global__ void __launch_bounds__(128, 8) k(cudaTextureObject_t t, const uint32_t* pix, int W, int H, int iters, float* out) {
  const int lane = threadIdx.x & 31, g = lane >> 3, gl = lane & 7;
  unsigned seed = (blockIdx.x * 4 + (threadIdx.x >> 5)) * 4 + g;
  float acc = 0.f;
  for (int it = 0; it < iters; ++it) {
    if (it % 150 == 0) seed = seed * 1664525u + 1013904223u;   // a window is re-sampled ~150 times (one Nelder-Mead run)
    const float bx = 8.0f + (float)((seed >> 8) % (W - 32)), by = 8.0f + (float)((seed >> 20) % (H - 32));   // one window per group
    for (int row = 0; row < 7; ++row) {
      const float jit = (float)(it & 7) * 0.05f;
      const float x = bx + gl * 0.93f + row * 0.11f + jit, y = by + row * 0.97f + gl * 0.07f + jit;
      const float lxf = truncf(x), lyf = truncf(y);
      float v[12];
      if (LDG) {
        const uint32_t* p = pix + ((int)lyf * W + (int)lxf);
        uint32_t a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + W), d = __ldg(p + W + 1);
        for (int ch = 0; ch < 3; ++ch) {
          v[4*ch] = (float)((a >> (8*ch)) & 255); v[4*ch+1] = (float)((b >> (8*ch)) & 255); v[4*ch+2] = (float)((c >> (8*ch)) & 255); v[4*ch+3] = (float)((d >> (8*ch)) & 255);
        }
      } else {
        for (int ch = 0; ch < 3; ++ch) { float4 q = tex2Dgather<float4>(t, lxf + 1.0f, lyf + 1.0f, ch); v[4*ch] = q.x; v[4*ch+1] = q.y; v[4*ch+2] = q.z; v[4*ch+3] = q.w; }
      }
      float s = x - lxf;
#pragma unroll
      for (int f = 0; f < NFMA; ++f) s = fmaf(s, v[f % 12], y);
#pragma unroll
      for (int f = 0; f < 12; ++f) acc += v[f];
      acc += s;
    }
  }
  if (acc == 1234.5f) out[0] = acc;
}
template <int NFMA, bool LDG>
float run(cudaTextureObject_t tex, const uint32_t* pix, int W, int H, float* d, int iters) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<NFMA, LDG><<<148 * 8, 128>>>(tex, pix, W, H, 8, d);
  cudaEventRecord(e0);
  k<NFMA, LDG><<<148 * 8, 128>>>(tex, pix, W, H, iters, d);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double warp_trips = 148.0 * 8 * 4 * iters * 7;
  printf("NFMA %3d %s: %.2f ms, %.1f clk per warp-trip per SM (1.965 GHz)\n", NFMA, LDG ? "LDG x4 + cvt" : "TLD4 x3     ", ms, ms * 1e-3 * 1.965e9 / (warp_trips / 148.0));
  return ms;
}
int main() {
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
  cudaTextureObject_t tex; CK(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
  float* d; CK(cudaMalloc(&d, 4));
  cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
  printf("maxTexture2DGather %d x %d, maxTexture2D %d x %d\n", pr.maxTexture2DGather[0], pr.maxTexture2DGather[1], pr.maxTexture2D[0], pr.maxTexture2D[1]);
  const int iters = 2000;
  run<0, false>(tex, pix, W, H, d, iters); run<24, false>(tex, pix, W, H, d, iters); run<48, false>(tex, pix, W, H, d, iters); run<72, false>(tex, pix, W, H, d, iters);
  run<0, true>(tex, pix, W, H, d, iters); run<24, true>(tex, pix, W, H, d, iters); run<48, true>(tex, pix, W, H, d, iters);
  CK(cudaDeviceSynchronize());
  return 0;
}
