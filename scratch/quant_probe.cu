// Streaming-quantizer design probe: fp32 [rows, cols] -> fp16 sign codes + per-chunk partial row sums (the one-pass XnorNet
// quantizer), warp-per-task with U 16-byte loads in flight per lane.  Prints achieved GB/s for each (U, CHUNK, blocks/SM).
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdio.h>
#include <stdint.h>

template <int U, int CHUNK, int MINB>
__global__ void __launch_bounds__(256, MINB) probe(const float* __restrict__ x, __half* __restrict__ codes, float* __restrict__ part,
                                                   int rows, int cols) {
  const int lane = threadIdx.x & 31;
  const int nchunks = cols / CHUNK;
  const long task = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (task >= (long)rows * nchunks) return;
  const long row = task / nchunks;
  const int ch = (int)(task - row * nchunks);
  const float* xr = x + row * cols + (long)ch * CHUNK;
  __half* cr = codes + row * cols + (long)ch * CHUNK;
  float sum = 0.f;
  for (int base0 = 0; base0 < CHUNK; base0 += 128 * U) {
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = __ldcs(reinterpret_cast<const float4*>(xr + base0 + u * 128 + 4 * lane));
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const float4 t = v[u];
      sum += (t.x + t.y) + (t.z + t.w);
      __half2 a = __floats2half2_rn((float)((t.x > 0.f) - (t.x < 0.f)), (float)((t.y > 0.f) - (t.y < 0.f)));
      __half2 b = __floats2half2_rn((float)((t.z > 0.f) - (t.z < 0.f)), (float)((t.w > 0.f) - (t.w < 0.f)));
      uint2 o;
      o.x = *reinterpret_cast<uint32_t*>(&a);
      o.y = *reinterpret_cast<uint32_t*>(&b);
      *reinterpret_cast<uint2*>(cr + base0 + u * 128 + 4 * lane) = o;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if (lane == 0) part[(long)ch * rows + row] = sum;
}

// thread-per-16-bytes copy-like variant: no row structure (upper bound for this access pattern)
template <int U>
__global__ void __launch_bounds__(256) flat(const float* __restrict__ x, __half* __restrict__ codes, long n4) {
  long i = ((long)blockIdx.x * 256 * U) + threadIdx.x;
  float4 v[U];
#pragma unroll
  for (int u = 0; u < U; ++u) v[u] = (i + u * 256 < n4) ? __ldcs(reinterpret_cast<const float4*>(x) + i + u * 256) : make_float4(0, 0, 0, 0);
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const float4 t = v[u];
    __half2 a = __floats2half2_rn((float)((t.x > 0.f) - (t.x < 0.f)), (float)((t.y > 0.f) - (t.y < 0.f)));
    __half2 b = __floats2half2_rn((float)((t.z > 0.f) - (t.z < 0.f)), (float)((t.w > 0.f) - (t.w < 0.f)));
    uint2 o;
    o.x = *reinterpret_cast<uint32_t*>(&a);
    o.y = *reinterpret_cast<uint32_t*>(&b);
    if (i + u * 256 < n4) reinterpret_cast<uint2*>(codes)[i + u * 256] = o;
  }
}

static float* xs[3];
static __half* cs[3];
static float* part;
static const int rows = 8192, cols = 4096;

template <typename F>
static void run(const char* name, F launch) {
  cudaEvent_t s, e;
  cudaEventCreate(&s); cudaEventCreate(&e);
  for (int i = 0; i < 3; ++i) launch(i % 3);
  cudaDeviceSynchronize();
  cudaEventRecord(s);
  const int n = 30;
  for (int i = 0; i < n; ++i) launch(i % 3);
  cudaEventRecord(e);
  cudaDeviceSynchronize();
  float ms;
  cudaEventElapsedTime(&ms, s, e);
  const double bytes = (double)rows * cols * 6;
  printf("%-28s %7.1f us  %6.0f GB/s  (%s)\n", name, ms / n * 1e3, bytes / (ms / n * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
}

#define PROBE(U, CH, MB) run("U=" #U " chunk=" #CH " minb=" #MB, [](int b) { \
  probe<U, CH, MB><<<(rows * (cols / CH) + 7) / 8, 256>>>(xs[b], cs[b], part, rows, cols); })

int main() {
  for (int i = 0; i < 3; ++i) {
    cudaMalloc(&xs[i], (size_t)rows * cols * 4);
    cudaMalloc(&cs[i], (size_t)rows * cols * 2);
    cudaMemset(xs[i], 0x3c, (size_t)rows * cols * 4);
  }
  cudaMalloc(&part, 8 * rows * 4);
  PROBE(4, 4096, 4); PROBE(4, 2048, 4); PROBE(4, 1024, 4); PROBE(4, 512, 4);
  PROBE(8, 4096, 4); PROBE(8, 2048, 4); PROBE(8, 1024, 4);
  PROBE(4, 1024, 6); PROBE(4, 2048, 6); PROBE(8, 1024, 6); PROBE(2, 1024, 8); PROBE(4, 1024, 8); PROBE(2, 512, 8);
  const long n4 = (long)rows * cols / 4;
  run("flat U=4", [=](int b) { flat<4><<<(unsigned)((n4 + 1023) / 1024), 256>>>(xs[b], cs[b], n4); });
  run("flat U=8", [=](int b) { flat<8><<<(unsigned)((n4 + 2047) / 2048), 256>>>(xs[b], cs[b], n4); });
  run("flat U=2", [=](int b) { flat<2><<<(unsigned)((n4 + 511) / 512), 256>>>(xs[b], cs[b], n4); });
  run("cudaMemcpy d2d 134MB (r+w 268)", [](int b) { cudaMemcpyAsync(xs[(b + 1) % 3], xs[b], (size_t)rows * cols * 4, cudaMemcpyDeviceToDevice); });
  return 0;
}
