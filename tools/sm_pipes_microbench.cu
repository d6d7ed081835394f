// FP32 pipe microbenchmark: scalar vs packed (f32x2) add/mul/fma issue rates on sm_100a, and co-issue with LDS/SHFL
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
typedef unsigned long long u64;
#define ITER 8192
template<int MODE>
__global__ void __launch_bounds__(1024) kern(float* out, float a, float b, long long* cyc) {
  __shared__ float sm[2048];
  for (int i = threadIdx.x; i < 2048; i += 1024) sm[i] = a * i;
  __syncthreads();
  float x[16];
  #pragma unroll
  for (int i = 0; i < 16; i++) x[i] = a * (threadIdx.x + i);
  u64 X[8];
  #pragma unroll
  for (int i = 0; i < 8; i++) { float2 t = make_float2(x[2*i], x[2*i+1]); X[i] = *reinterpret_cast<u64*>(&t); }
  float2 ab = make_float2(a, b), ba = make_float2(b, a);
  u64 AB = *reinterpret_cast<u64*>(&ab), BA = *reinterpret_cast<u64*>(&ba);
  float ls = 0.f;
  long long t0 = clock64();
  #pragma unroll 1
  for (int it = 0; it < ITER; ++it) {
    if (MODE == 0) {  // 16 FFMA
      #pragma unroll
      for (int i = 0; i < 16; i++) x[i] = fmaf(x[i], a, b);
    } else if (MODE == 1) {  // 16 FADD
      #pragma unroll
      for (int i = 0; i < 16; i++) x[i] = x[i] + x[(i+5)&15];
    } else if (MODE == 2) {  // 16 FMUL
      #pragma unroll
      for (int i = 0; i < 16; i++) x[i] = x[i] * x[(i+5)&15];
    } else if (MODE == 3) {  // 8 FFMA2 (=16 fma lanes)
      #pragma unroll
      for (int i = 0; i < 8; i++) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(X[i]) : "l"(AB), "l"(BA));
    } else if (MODE == 4) {  // 8 FADD2
      #pragma unroll
      for (int i = 0; i < 8; i++) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(X[i]) : "l"(X[(i+3)&7]));
    } else if (MODE == 5) {  // 8 FMUL2
      #pragma unroll
      for (int i = 0; i < 8; i++) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(X[i]) : "l"(X[(i+3)&7]));
    } else if (MODE == 6) {  // 16 FFMA 3-reg distinct operands
      #pragma unroll
      for (int i = 0; i < 16; i++) x[i] = fmaf(x[i], x[(i+5)&15], x[(i+9)&15]);
    } else if (MODE == 7) {  // 8 FFMA2 3-reg
      #pragma unroll
      for (int i = 0; i < 8; i++) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(X[i]) : "l"(X[(i+3)&7]), "l"(X[(i+5)&7]));
    } else if (MODE == 8) {  // 8 FADD2 + 8 LDS
      #pragma unroll
      for (int i = 0; i < 8; i++) {
        asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(X[i]) : "l"(X[(i+3)&7]));
        ls += sm[(threadIdx.x + i * 32 + it) & 2047];
      }
    } else if (MODE == 9) {  // 16 FADD + 8 LDS
      #pragma unroll
      for (int i = 0; i < 8; i++) {
        x[2*i] = x[2*i] + x[(2*i+5)&15]; x[2*i+1] = x[2*i+1] + x[(2*i+6)&15];
        ls += sm[(threadIdx.x + i * 32 + it) & 2047];
      }
    } else if (MODE == 10) {  // 8 FADD2 + 8 SHFL
      #pragma unroll
      for (int i = 0; i < 8; i++) {
        asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(X[i]) : "l"(X[(i+3)&7]));
        ls += __shfl_xor_sync(0xffffffffu, ls, 1 + i);
      }
    } else if (MODE == 11) {  // 8 FADD2 + 8 IADD/LOP (alu pipe)
      #pragma unroll
      for (int i = 0; i < 8; i++) {
        asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(X[i]) : "l"(X[(i+3)&7]));
        x[i] = __int_as_float((__float_as_int(x[i]) ^ it) + i);
      }
    } else if (MODE == 12) {  // 16 FADD + 8 alu
      #pragma unroll
      for (int i = 0; i < 8; i++) {
        x[8+i] = x[8+i] + x[8+((i+3)&7)];
        asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(X[i]) : "l"(X[(i+3)&7]));
        x[i] = __int_as_float((__float_as_int(x[i]) ^ it) + i);
      }
    } else if (MODE == 13) {  // 8 FADD2 + 8 FADD (mixed)
      #pragma unroll
      for (int i = 0; i < 8; i++) {
        asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(X[i]) : "l"(X[(i+3)&7]));
        x[i] = x[i] + x[(i+5)&7];
      }
    } else if (MODE == 14) {  // 8 LDS.128 only
      #pragma unroll
      for (int i = 0; i < 8; i++) {
        float4 v = reinterpret_cast<float4*>(sm)[(threadIdx.x + i * 32 + it) & 511];
        x[i] += v.x; x[i+8] += v.y + v.z * v.w;
      }
    }
  }
  long long t1 = clock64();
  float s = ls;
  #pragma unroll
  for (int i = 0; i < 16; i++) s += x[i];
  #pragma unroll
  for (int i = 0; i < 8; i++) { float2 t = *reinterpret_cast<float2*>(&X[i]); s += t.x + t.y; }
  out[blockIdx.x * 1024 + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template<int MODE> void run(const char* name, int flop_lanes_per_iter, int instr_per_iter) {
  float* out; long long* cyc; int nb = 148;  // 4 CTAs x 8 warps = 32 warps/SM = 8 per SMSP
  cudaMalloc(&out, nb * 1024 * 4); cudaMalloc(&cyc, nb * 8);
  kern<MODE><<<nb, 1024>>>(out, 1.0001f, 0.5f, cyc); cudaDeviceSynchronize();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0); kern<MODE><<<nb, 1024>>>(out, 1.0001f, 0.5f, cyc); cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long h[592]; cudaMemcpy(h, cyc, nb * 8, cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < nb; i++) avg += h[i]; avg /= nb;
  // per SMSP: 8 warps each ITER * instr_per_iter instrs over avg cycles
  double ipc = 8.0 * ITER * instr_per_iter / avg;
  double lanes = 8.0 * ITER * flop_lanes_per_iter * 32 / avg;  // fp ops per clk per SMSP
  double clk = ms * 1e-3 * 1.965e9; printf("%-34s cyc %9.0f  instr/clk/SMSP %.3f  fp-lanes/clk/SMSP %.1f  ms %.4f  (by event@1965MHz: ipc %.3f)\n", name, avg, ipc, lanes, ms, 8.0 * ITER * instr_per_iter / clk);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run<0>("16 FFMA (imm/const operands)", 16, 16);
  run<6>("16 FFMA 3-reg", 16, 16);
  run<1>("16 FADD", 16, 16);
  run<2>("16 FMUL", 16, 16);
  run<3>("8 FFMA2 (2 shared operands)", 16, 8);
  run<7>("8 FFMA2 3-reg", 16, 8);
  run<4>("8 FADD2", 16, 8);
  run<5>("8 FMUL2", 16, 8);
  run<13>("8 FADD2 + 8 FADD", 24, 16);
  run<8>("8 FADD2 + 8 LDS (+8 FADD,alu)", 16, 8);
  run<9>("16 FADD + 8 LDS (+8 FADD,alu)", 16, 16);
  run<10>("8 FADD2 + 8 SHFL(+FADD)", 16, 8);
  run<11>("8 FADD2 + 16 ALU", 16, 24);
  run<12>("8 FADD2 + 8 FADD + 16 ALU", 24, 32);
  run<14>("8 LDS.128 (+fp)", 0, 8);
  return 0;
}
