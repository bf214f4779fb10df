// Which SM does block b of a persistent grid land on?  (decides how k_sweep_group_tma maps blocks to rows so that the
// CTAs sharing an SM work on adjacent rows.)  Build: nvcc -arch=sm_100a -o smid_map smid_map.cu ; run: ./smid_map [blocks_per_sm=6]
#include <cstdio>
#include <cstdlib>
#include <vector>
__global__ void k(int *smid, long long spin)
{
  extern __shared__ unsigned char sm[];
  unsigned s;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(s));
  if (threadIdx.x == 0) smid[blockIdx.x] = (int)s;
  const long long t0 = clock64();
  while (clock64() - t0 < spin) { sm[threadIdx.x] = (unsigned char)threadIdx.x; }
}
int main(int argc, char **argv)
{
  const int per = argc > 1 ? atoi(argv[1]) : 6;
  int dev = 0, nsm = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
  const int grid = per * nsm, smem = 27648 + 96;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  int *d;
  cudaMalloc(&d, grid * sizeof(int));
  k<<<grid, 128, smem>>>(d, 200000);
  std::vector<int> h(grid);
  cudaMemcpy(h.data(), d, grid * sizeof(int), cudaMemcpyDeviceToHost);
  printf("nsm %d grid %d err %s\n", nsm, grid, cudaGetErrorString(cudaGetLastError()));
  int same = 0, pair = 0;
  for (int b = 0; b + nsm < grid; b++) same += h[b] == h[b + nsm];
  for (int b = 0; b + 1 < grid; b++) pair += h[b] == h[b + 1];
  printf("blocks b and b+nsm on the same SM: %d of %d; blocks b and b+1 on the same SM: %d of %d\n", same, grid - nsm, pair, grid - 1);
  printf("first 40 blocks -> SM:");
  for (int b = 0; b < 40; b++) printf(" %d", h[b]);
  printf("\nblocks nsm..nsm+19 -> SM:");
  for (int b = nsm; b < nsm + 20; b++) printf(" %d", h[b]);
  printf("\n");
  return 0;
}
