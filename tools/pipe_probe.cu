// pipe_probe.cu -- issue-rate microbenchmark for the integer/DPX pipes on sm_100a.
// Measures thread-level results per clock per SM for the instructions the Gotoh
// recurrence is built from ("measure, don't assume").
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o pipe_probe pipe_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); return 1;}}while(0)

constexpr int NCH = 8;      // independent chains per thread
constexpr int ITER = 4096;  // loop trips

enum Op { VIADDMAX_U16X2, VIMAX3_U16X2, VMAXU2, VIADDMAX_S32, VIMAX3_S32, IADD, IMAD, LOP, PRMT_,
          MIX_3DPX_2IMAD, MIX_3DPX_2IADD, MIX_5DPX, MIX_3DPX_2IMAD_LDS, SHFL, LDS32, MIX_CELL, VIADD_U, MIX_VIADD, MIX_LDSIMM, LDSIMM };

__device__ __forceinline__ uint32_t imad(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d; asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d;
}

template <int OP>
__global__ void __launch_bounds__(1024, 1) probe(uint32_t* out, const uint32_t k1, const uint32_t k2, uint32_t one, long long* cyc) {
  __shared__ uint32_t sm[1024 + 64];
  uint32_t a[NCH], b[NCH];
#pragma unroll
  for (int i = 0; i < NCH; i++) { a[i] = threadIdx.x * 7 + i + k1; b[i] = threadIdx.x * 3 + i * 5 + k2; }
  sm[threadIdx.x] = threadIdx.x + k1; 
  if (threadIdx.x < 64) sm[1024 + threadIdx.x] = k2;
  __syncthreads();
  const uint32_t* sp = sm + (threadIdx.x & 31) + (k1 & 31);
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITER; it++) {
#pragma unroll
    for (int i = 0; i < NCH; i++) {
      if (OP == VIADDMAX_U16X2) a[i] = __viaddmax_u16x2(a[i], k1, b[i]);
      else if (OP == VIMAX3_U16X2) a[i] = __vimax3_u16x2(a[i], k1, b[i]);
      else if (OP == VMAXU2) a[i] = __vmaxu2(a[i], b[i]) , b[i] = __vmaxu2(b[i], k1);
      else if (OP == VIADDMAX_S32) a[i] = __viaddmax_s32(a[i], k1, b[i]);
      else if (OP == VIMAX3_S32) a[i] = __vimax3_s32(a[i], k1, b[i]);
      else if (OP == IADD) a[i] = a[i] + b[i];
      else if (OP == IMAD) a[i] = imad(a[i], one, b[i]);
      else if (OP == LOP) a[i] = (a[i] & b[i]) ^ k1;
      else if (OP == PRMT_) a[i] = __byte_perm(a[i], b[i], 0x4240);
      else if (OP == SHFL) a[i] = __shfl_up_sync(0xffffffffu, a[i], 1);
      else if (OP == LDS32) a[i] = sp[(a[i] & 31) * 32 + i];   // conflict-free (same bank group per lane)
      else if (OP == VIADD_U) a[i] = a[i] - k1;
      else if (OP == LDSIMM) { a[i] ^= sp[i * 33 + (it & 7) * 64]; }
      else if (OP == MIX_VIADD) {
        uint32_t t = a[i] + sp[0];
        uint32_t h = __vimax3_u16x2(t, b[i], k1);
        uint32_t hg = h - k1;
        a[i] = __viaddmax_u16x2(a[i], k2, hg);
        b[i] = __viaddmax_u16x2(b[i], k2, hg);
      } else if (OP == MIX_LDSIMM) {
        uint32_t s = sp[i * 33 + (it & 7) * 64];
        uint32_t t = imad(a[i], one, s);
        uint32_t h = __vimax3_u16x2(t, b[i], k1);
        uint32_t hg = imad(h, one, k1);
        a[i] = __viaddmax_u16x2(a[i], k2, hg);
        b[i] = __viaddmax_u16x2(b[i], k2, hg);
      }
      else if (OP == MIX_3DPX_2IMAD) {
        // 3 DPX + 2 IMAD per group, the recurrence's instruction mix
        uint32_t t = imad(a[i], one, k2);
        uint32_t h = __vimax3_u16x2(t, b[i], k1);
        uint32_t hg = imad(h, one, k1);
        a[i] = __viaddmax_u16x2(a[i], k2, hg);
        b[i] = __viaddmax_u16x2(b[i], k2, hg);
      } else if (OP == MIX_3DPX_2IADD) {
        uint32_t t = a[i] + k2;
        uint32_t h = __vimax3_u16x2(t, b[i], k1);
        uint32_t hg = h + k1;
        a[i] = __viaddmax_u16x2(a[i], k2, hg);
        b[i] = __viaddmax_u16x2(b[i], k2, hg);
      } else if (OP == MIX_5DPX) {
        uint32_t m = __viaddmax_u16x2(a[i], k2, b[i]);
        uint32_t h = __vmaxu2(m, k1);
        uint32_t hg = __viaddmax_u16x2(h, k1, 0u);
        a[i] = __viaddmax_u16x2(a[i], k2, hg);
        b[i] = __viaddmax_u16x2(b[i], k2, hg);
      } else if (OP == MIX_3DPX_2IMAD_LDS || OP == MIX_CELL) {
        uint32_t s = sp[((it + i) & 31) * 32];
        uint32_t t = imad(a[i], one, s);
        uint32_t h = __vimax3_u16x2(t, b[i], k1);
        uint32_t hg = imad(h, one, k1);
        a[i] = __viaddmax_u16x2(a[i], k2, hg);
        b[i] = __viaddmax_u16x2(b[i], k2, hg);
      }
    }
  }
  long long t1 = clock64();
  uint32_t r = 0;
#pragma unroll
  for (int i = 0; i < NCH; i++) r ^= a[i] ^ b[i];
  if (r == 0x12345678u) out[threadIdx.x] = r;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
int run(const char* name, double ops_per_inner, int threads, int nsm, uint32_t* out, long long* dcyc) {
  probe<OP><<<nsm, threads>>>(out, 3, 5, 1, dcyc);   // warm
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  probe<OP><<<nsm, threads>>>(out, 3, 5, 1, dcyc);
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  static long long h[1024];
  CK(cudaMemcpy(h, dcyc, sizeof(long long) * nsm, cudaMemcpyDeviceToHost));
  long long mx = 0; double avg = 0;
  for (int i = 0; i < nsm; i++) { if (h[i] > mx) mx = h[i]; avg += h[i]; }
  avg /= nsm;
  double ops = (double)threads * NCH * ITER * ops_per_inner;
  printf("{\"op\": \"%s\", \"threads_per_sm\": %d, \"ops_per_clk_per_sm\": %.2f, \"cycles_avg\": %.0f, \"cycles_max\": %lld, \"ms\": %.4f, \"eff_mhz\": %.0f}\n",
         name, threads, ops / avg, avg, mx, ms, mx / (ms * 1e3));
  return 0;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int nsm = p.multiProcessorCount;
  printf("{\"device\": \"%s\", \"sms\": %d, \"cc\": \"%d.%d\", \"clock_khz\": %d}\n", p.name, nsm, p.major, p.minor, p.clockRate);
  uint32_t* out; long long* dcyc;
  CK(cudaMalloc(&out, 4096 * 4)); CK(cudaMalloc(&dcyc, 1024 * 8));
  for (int threads : {1024, 512, 384, 256}) {
    run<VIADDMAX_U16X2>("viaddmax_u16x2", 1, threads, nsm, out, dcyc);
    run<VIMAX3_U16X2>("vimax3_u16x2", 1, threads, nsm, out, dcyc);
    run<VMAXU2>("vmaxu2", 2, threads, nsm, out, dcyc);
    run<VIADDMAX_S32>("viaddmax_s32", 1, threads, nsm, out, dcyc);
    run<VIMAX3_S32>("vimax3_s32", 1, threads, nsm, out, dcyc);
    run<IADD>("iadd", 1, threads, nsm, out, dcyc);
    run<IMAD>("imad", 1, threads, nsm, out, dcyc);
    run<LOP>("lop3", 1, threads, nsm, out, dcyc);
    run<PRMT_>("prmt", 1, threads, nsm, out, dcyc);
    run<SHFL>("shfl", 1, threads, nsm, out, dcyc);
    run<LDS32>("lds32", 1, threads, nsm, out, dcyc);
    run<MIX_3DPX_2IMAD>("mix_3dpx_2imad(5 instr)", 5, threads, nsm, out, dcyc);
    run<MIX_3DPX_2IADD>("mix_3dpx_2iadd(5 instr)", 5, threads, nsm, out, dcyc);
    run<MIX_5DPX>("mix_5dpx(5 instr)", 5, threads, nsm, out, dcyc);
    run<MIX_3DPX_2IMAD_LDS>("mix_3dpx_2imad_1lds(6 instr)", 6, threads, nsm, out, dcyc);
    run<VIADD_U>("viadd_uniform", 1, threads, nsm, out, dcyc);
    run<MIX_VIADD>("mix_3dpx_viadd_iadd(5 instr)", 5, threads, nsm, out, dcyc);
    run<LDSIMM>("lds_imm(+1 lop)", 1, threads, nsm, out, dcyc);
    run<MIX_LDSIMM>("mix_3dpx_2imad_1ldsimm(6 instr)", 6, threads, nsm, out, dcyc);
  }
  return 0;
}
