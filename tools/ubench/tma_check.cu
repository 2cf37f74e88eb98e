// tma_check.cu - isolates the TMA tile load used by photo.cu (tools/ubench, not part of the library).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tma_check tma_check.cu
//   ./tma_check <variant>   0: param-space descriptor (__grid_constant__), 1: descriptor in global memory
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../mal_b200/csrc/mal_tma.cuh"
using namespace mal;

struct Maps { TileMap m[2]; };

__global__ void k_param(const __grid_constant__ Maps maps, float* out, int ox, int oy, int n, int elems) {
  extern __shared__ __align__(128) unsigned char sm[];
  float* tile = reinterpret_cast<float*>(sm);
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(sm + ((elems * 4 + 127) / 128 * 128));
  if (threadIdx.x == 0) mbar_init(bar, 1);
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, elems * 4);
    tma_load_3d(tile, &maps.m[0], ox, oy, n, bar);
  }
  mbar_wait(bar, 0);
  for (int i = threadIdx.x; i < elems; i += blockDim.x) out[i] = tile[i];
}
__global__ void k_global(const TileMap* map, float* out, int ox, int oy, int n, int elems) {
  extern __shared__ __align__(128) unsigned char sm[];
  float* tile = reinterpret_cast<float*>(sm);
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(sm + ((elems * 4 + 127) / 128 * 128));
  if (threadIdx.x == 0) mbar_init(bar, 1);
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, elems * 4);
    tma_load_3d(tile, map, ox, oy, n, bar);
  }
  mbar_wait(bar, 0);
  for (int i = threadIdx.x; i < elems; i += blockDim.x) out[i] = tile[i];
}

int main(int argc, char** argv) {
  const int variant = argc > 1 ? atoi(argv[1]) : 0;
  const int W = 640, H = 192, N = 6, bw = argc > 6 ? atoi(argv[6]) : 36, bh = argc > 2 ? atoi(argv[2]) : 18, bn = argc > 3 ? atoi(argv[3]) : 3;
  const int ox = argc > 4 ? atoi(argv[4]) : 31, oy = argc > 5 ? atoi(argv[5]) : 15, n0 = 3;
  std::vector<float> h((size_t)W * H * N);
  for (size_t i = 0; i < h.size(); i++) h[i] = (float)(i % 100003) * 0.25f;
  float *d, *out;
  cudaMalloc(&d, h.size() * 4);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  const int elems = bw * bh * bn;
  cudaMalloc(&out, elems * 4);
  Maps maps;
  memset(&maps, 0, sizeof(maps));
  if (!tile_map_encode(&maps.m[0], d, W, H, N, bw, bh, bn)) { printf("encode failed\n"); return 2; }
  const size_t smem = (elems * 4 + 127) / 128 * 128 + 64;
  if (variant == 0) {
    cudaFuncSetAttribute(k_param, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_param<<<1, 128, smem>>>(maps, out, ox, oy, n0, elems);
  } else {
    TileMap* dm;
    cudaMalloc(&dm, sizeof(TileMap));
    cudaMemcpy(dm, &maps.m[0], sizeof(TileMap), cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(k_global, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_global<<<1, 128, smem>>>(dm, out, ox, oy, n0, elems);
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("variant %d box %dx%dx%d at (%d,%d): CUDA error %s\n", variant, bw, bh, bn, ox, oy, cudaGetErrorString(e)); return 1; }
  std::vector<float> o(elems);
  cudaMemcpy(o.data(), out, elems * 4, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int c = 0; c < bn; c++)
    for (int j = 0; j < bh; j++)
      for (int i = 0; i < bw; i++) {
        const int gx = ox + i, gy = oy + j, gn = n0 + c;
        const bool in = gx >= 0 && gx < W && gy >= 0 && gy < H && gn < N;
        const float want = in ? h[((size_t)gn * H + gy) * W + gx] : 0.0f;
        if (o[(c * bh + j) * bw + i] != want) bad++;
      }
  printf("variant %d box %dx%dx%d at (%d,%d): %s (%d mismatches)\n", variant, bw, bh, bn, ox, oy, bad ? "MISMATCH" : "ok", bad);
  return bad != 0;
}
