// lat_probe.cu -- dependent-issue latency of the DPX/IADD instructions (one warp, one chain).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
constexpr int ITER = 8192;
template <int V>
__global__ void probe(uint32_t* out, uint32_t k1, uint32_t k2, long long* cyc) {
  uint32_t a = threadIdx.x + k1, b = threadIdx.x * 3 + k2, e = k1 * 5, f = k2 * 7;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITER; it++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
      if (V == 0) a = __viaddmax_u16x2(a, 0x00010001u, b);                       // ALU -> ALU
      else if (V == 1) a = __vimax3_u16x2(a, b, e);
      else if (V == 2) a = a + b;                                                 // IADD
      else if (V == 3) { uint32_t h = __vimax3_u16x2(a, e, f); uint32_t hg = h - k1; e = __viaddmax_u16x2(e, 0x00010001u, hg); a = h ^ 0; }  // real chain h->hg->E
      else if (V == 4) { uint32_t h = __vimax3_u16x2(a, e, f); e = __viaddmax_u16x2(e, 0x00010001u, h); a = h; }      // 2-op chain
      else if (V == 5) { uint32_t h = __vimax3_u16x2(b, e, f); uint32_t hg = h - k1; e = __viaddmax_u16x2(e, 0x00010001u, hg); f = __viaddmax_u16x2(f, 0x00010001u, hg); b += k2; }
    }
  }
  long long t1 = clock64();
  if ((a ^ b ^ e ^ f) == 0x12345678u) out[threadIdx.x] = a;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
template <int V> void run(const char* name, double ops_per_u, uint32_t* out, long long* dcyc) {
  probe<V><<<1, 32>>>(out, 3, 5, dcyc); cudaDeviceSynchronize();
  probe<V><<<1, 32>>>(out, 3, 5, dcyc); cudaDeviceSynchronize();
  long long h; cudaMemcpy(&h, dcyc, 8, cudaMemcpyDeviceToHost);
  printf("{\"chain\": \"%s\", \"cycles_per_step\": %.2f}\n", name, (double)h / (ITER * 8.0));
}
int main() {
  uint32_t* out; long long* dcyc; cudaMalloc(&out, 4096); cudaMalloc(&dcyc, 64);
  run<0>("viaddmax->viaddmax", 1, out, dcyc);
  run<1>("vimax3->vimax3", 1, out, dcyc);
  run<2>("iadd->iadd", 1, out, dcyc);
  run<3>("vimax3->sub->viaddmax (E chain of one cell)", 3, out, dcyc);
  run<4>("vimax3->viaddmax (2-op chain)", 2, out, dcyc);
  run<5>("cell with E and F updates, E chain", 3, out, dcyc);
  return 0;
}
