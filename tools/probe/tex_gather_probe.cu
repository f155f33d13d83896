// tex2Dgather on an RGBA8 CUDA array with normalized-float reads and unnormalized coordinates:
// which component is which texel, and are the values exactly byte/255?   nvcc -arch=sm_100a tex_gather_probe.cu
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s failed: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
__global__ void k(cudaTextureObject_t t, int lx, int ly, float4* out) {
  for (int c = 0; c < 3; ++c) out[c] = tex2Dgather<float4>(t, (float)lx + 1.0f, (float)ly + 1.0f, c);
  out[3] = tex2D<float4>(t, (float)lx + 0.5f, (float)ly + 0.5f);
}
__global__ void kall(cudaTextureObject_t t, int w, int h, int* bad) {
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= w - 1 || y >= h - 1) return;
  float4 g = tex2Dgather<float4>(t, (float)x + 1.0f, (float)y + 1.0f, 1);   // green = (x*7+y*13)&255 pattern
  int e00 = (x * 7 + y * 13 + 1) & 255, e10 = ((x + 1) * 7 + y * 13 + 1) & 255, e01 = (x * 7 + (y + 1) * 13 + 1) & 255, e11 = ((x + 1) * 7 + (y + 1) * 13 + 1) & 255;
  if (g.w != (float)e00 / 255.0f || g.z != (float)e10 / 255.0f || g.x != (float)e01 / 255.0f || g.y != (float)e11 / 255.0f) atomicAdd(bad, 1);
}
int main() {
  const int W = 203, H = 77;
  std::vector<uchar4> img(W * H);
  for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) img[y * W + x] = make_uchar4((x * 7 + y * 13) & 255, (x * 7 + y * 13 + 1) & 255, (x * 3 + y * 5) & 255, 0);
  cudaChannelFormatDesc cd = cudaCreateChannelDesc<uchar4>();
  cudaArray_t arr;
  CK(cudaMallocArray(&arr, &cd, W, H, cudaArrayTextureGather));
  CK(cudaMemcpy2DToArray(arr, 0, 0, img.data(), W * 4, W * 4, H, cudaMemcpyHostToDevice));
  cudaResourceDesc rd = {}; rd.resType = cudaResourceTypeArray; rd.res.array.array = arr;
  cudaTextureDesc td = {}; td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp; td.filterMode = cudaFilterModePoint;
  td.readMode = cudaReadModeNormalizedFloat; td.normalizedCoords = 0;
  cudaTextureObject_t tex;
  CK(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
  float4* d; CK(cudaMalloc(&d, 4 * sizeof(float4)));
  const int lx = 10, ly = 20;
  k<<<1, 1>>>(tex, lx, ly, d);
  float4 h[4]; CK(cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost));
  for (int c = 0; c < 3; ++c) printf("comp %d: x=%.1f y=%.1f z=%.1f w=%.1f\n", c, h[c].x * 255, h[c].y * 255, h[c].z * 255, h[c].w * 255);
  auto px = [&](int x, int y) { return img[y * W + x]; };
  printf("expect R: (x0,y0)=%d (x1,y0)=%d (x0,y1)=%d (x1,y1)=%d\n", px(lx, ly).x, px(lx + 1, ly).x, px(lx, ly + 1).x, px(lx + 1, ly + 1).x);
  printf("point: %.1f %.1f %.1f\n", h[3].x * 255, h[3].y * 255, h[3].z * 255);
  int* bad; CK(cudaMalloc(&bad, 4)); CK(cudaMemset(bad, 0, 4));
  kall<<<dim3((W + 63) / 64, H), 64>>>(tex, W, H, bad);
  int hb; CK(cudaMemcpy(&hb, bad, 4, cudaMemcpyDeviceToHost));
  printf("exhaustive gather check (order w=x0y0 z=x1y0 x=x0y1 y=x1y1, value == byte/255.0f exactly): %d mismatches\n", hb);
  return 0;
}
