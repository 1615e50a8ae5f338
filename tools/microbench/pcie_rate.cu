// Microbenchmark (B200 box): what the host <-> device path can carry, so that the end-to-end number of bench.py has a measured
// ceiling beside it: pinned contiguous cudaMemcpyAsync H2D, D2H and both at once (duplex, two streams), the same through
// cudaMemcpy2DAsync with the row sizes the sliced host path of the library uses (sl_capi.cu: process_host_sliced), and
// write-combined input buffers.
//   nvcc -O2 -o /tmp/pcie_rate tools/microbench/pcie_rate.cu && /tmp/pcie_rate [device]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf ("CUDA error %s at %s:%d\n", cudaGetErrorString (e_), __FILE__, __LINE__); exit (1); } } while (0)
static double now () { return std::chrono::duration<double> (std::chrono::steady_clock::now ().time_since_epoch ()).count (); }

int main (int argc, char **argv)
{
  const int dev = argc > 1 ? atoi (argv[1]) : 0;
  CK (cudaSetDevice (dev));
  const size_t bytes = (size_t) 1 << 30;                  // 1 GiB each way per repetition
  char *h_in, *h_out, *h_wc, *d_in, *d_out;
  CK (cudaHostAlloc (&h_in, bytes, cudaHostAllocDefault)); CK (cudaHostAlloc (&h_out, bytes, cudaHostAllocDefault));
  CK (cudaHostAlloc (&h_wc, bytes, cudaHostAllocWriteCombined));
  memset (h_in, 1, bytes); memset (h_out, 2, bytes); memset (h_wc, 3, bytes);
  CK (cudaMalloc (&d_in, bytes)); CK (cudaMalloc (&d_out, bytes));
  cudaStream_t s0, s1; CK (cudaStreamCreateWithFlags (&s0, cudaStreamNonBlocking)); CK (cudaStreamCreateWithFlags (&s1, cudaStreamNonBlocking));
  const int reps = 4;
  auto timeit = [&] (const char *name, auto fn, double bytes_each_way, bool duplex)
  {
    fn (); CK (cudaDeviceSynchronize ());
    const double t0 = now ();
    for (int r = 0; r < reps; r++) fn ();
    CK (cudaDeviceSynchronize ());
    const double dt = (now () - t0) / reps;
    printf ("%-72s %7.2f GB/s %s\n", name, bytes_each_way / dt / 1e9, duplex ? "each way (duplex)" : "");
  };
  timeit ("contiguous pinned H2D", [&] { CK (cudaMemcpyAsync (d_in, h_in, bytes, cudaMemcpyHostToDevice, s0)); }, (double) bytes, false);
  timeit ("contiguous pinned D2H", [&] { CK (cudaMemcpyAsync (h_out, d_out, bytes, cudaMemcpyDeviceToHost, s1)); }, (double) bytes, false);
  timeit ("contiguous pinned H2D + D2H at once", [&] { CK (cudaMemcpyAsync (d_in, h_in, bytes, cudaMemcpyHostToDevice, s0)); CK (cudaMemcpyAsync (h_out, d_out, bytes, cudaMemcpyDeviceToHost, s1)); }, (double) bytes, true);
  timeit ("write-combined input H2D + D2H at once", [&] { CK (cudaMemcpyAsync (d_in, h_wc, bytes, cudaMemcpyHostToDevice, s0)); CK (cudaMemcpyAsync (h_out, d_out, bytes, cudaMemcpyDeviceToHost, s1)); }, (double) bytes, true);
  // 64 MiB pieces (the library's slice size), contiguous, both directions at once
  timeit ("contiguous, 16 x 64 MiB pieces per direction, duplex", [&] {
    for (size_t o = 0; o < bytes; o += (size_t) 64 << 20)
    { CK (cudaMemcpyAsync (d_in + o, h_in + o, (size_t) 64 << 20, cudaMemcpyHostToDevice, s0)); CK (cudaMemcpyAsync (h_out + o, d_out + o, (size_t) 64 << 20, cudaMemcpyDeviceToHost, s1)); } }, (double) bytes, true);
  // strided 2-D copies: `rows` rows of `row` bytes, host pitch = bytes / rows (a time slice of all channels), device compact
  for (size_t row : { (size_t) 6144, (size_t) 15360, (size_t) 30720, (size_t) 61440, (size_t) 245760 })
    for (size_t rows : { (size_t) 1024, (size_t) 8192 })
    {
      const size_t pitch = bytes / rows;
      if (row > pitch) continue;
      const size_t slices = pitch / row;                 // slices that cover the whole buffer
      const size_t use = slices > 64 ? 64 : slices;      // bounded work per repetition
      char name[128]; snprintf (name, sizeof name, "2-D strided, %zu rows x %zu B per slice (%zu slices), duplex", rows, row, use);
      timeit (name, [&] {
        for (size_t s = 0; s < use; s++)
        {
          CK (cudaMemcpy2DAsync (d_in + s * rows * row, row, h_in + s * row, pitch, row, rows, cudaMemcpyHostToDevice, s0));
          CK (cudaMemcpy2DAsync (h_out + s * row, pitch, d_out + s * rows * row, row, row, rows, cudaMemcpyDeviceToHost, s1));
        } }, (double) (use * rows * row), true);
    }
  return 0;
}
