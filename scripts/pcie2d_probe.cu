// Developer tool (GPU box): host-link probe - flat vs padded-row 2-D copies, one direction and both at once
// (profiles/r2_pcie2d_probe.txt).  nvcc -O2 -gencode arch=compute_100a,code=sm_100a -o scratch/pcie2d scripts/pcie2d_probe.cu
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
int main()
{
  const size_t Ntx = 8196, pitch = 8208, rows = 8196, planes = 4;
  double *h_in, *h_out, *d_a, *d_b;
  cudaMallocHost(&h_in, planes * rows * Ntx * 8);
  cudaMallocHost(&h_out, planes * rows * Ntx * 8);
  cudaMalloc(&d_a, planes * rows * pitch * 8);
  cudaMalloc(&d_b, planes * rows * pitch * 8);
  cudaStream_t s1, s2;
  cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking);
  cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking);
  const double gb = planes * rows * Ntx * 8 / 1e9;
  for (int mode = 0; mode < 3; ++mode) // 0 flat, 1 2-D per plane and block, 2 2-D with 128-byte aligned host rows
    for (int dir = 0; dir < 3; ++dir) // 0 up, 1 down, 2 both
      for (int B : {256, 2048})
      {
        double best = 0;
        for (int rep = 0; rep < 3; ++rep)
        {
          cudaDeviceSynchronize();
          auto t0 = std::chrono::steady_clock::now();
          for (size_t r0 = 0; r0 < rows; r0 += B)
          {
            const size_t nr = (r0 + B <= rows) ? B : rows - r0;
            for (size_t f = 0; f < planes; ++f)
            {
              double *h1 = h_in + (f * rows + r0) * Ntx, *h2 = h_out + (f * rows + r0) * Ntx;
              double *d1 = d_a + (f * rows + r0) * pitch, *d2 = d_b + (f * rows + r0) * pitch;
              if (mode == 0)
              {
                if (dir != 1) cudaMemcpyAsync(d1, h1, nr * Ntx * 8, cudaMemcpyHostToDevice, s1);
                if (dir != 0) cudaMemcpyAsync(h2, d2, nr * Ntx * 8, cudaMemcpyDeviceToHost, s2);
              }
              else
              {
                const size_t w = (mode == 1) ? Ntx * 8 : 8192 * 8;
                if (dir != 1) cudaMemcpy2DAsync(d1, pitch * 8, h1, Ntx * 8, w, nr, cudaMemcpyHostToDevice, s1);
                if (dir != 0) cudaMemcpy2DAsync(h2, Ntx * 8, d2, pitch * 8, w, nr, cudaMemcpyDeviceToHost, s2);
              }
            }
          }
          cudaDeviceSynchronize();
          const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
          if (gb / s > best) best = gb / s;
        }
        printf("mode %d (%s) dir %d (%s) block %4d rows: %.1f GB/s per direction\n", mode,
               mode == 0 ? "flat" : (mode == 1 ? "2-D, 65568-byte rows" : "2-D, 65536 of 65568 bytes"), dir,
               dir == 0 ? "up" : (dir == 1 ? "down" : "both"), B, best);
      }
  return 0;
}
