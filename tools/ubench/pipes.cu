// pipes.cu -- issue/pipe rates of the SASS instructions the raycast march is made of, on one B200.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu && ./pipes
// Each kernel runs N_IT iterations of 8 independent chains of one instruction per thread, with
// 148*4 CTAs of 256 threads (8 warps per SMSP).  Reported: warp-instructions per cycle per SM.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
#define N_IT 4096
#define CHAINS 8
#define DEF_KERNEL(NAME, DECL, INIT, BODY, SINK)                           \
  __global__ void NAME(float* out, int n_it, float seed) {                 \
    DECL;                                                                   \
    _Pragma("unroll") for (int c = 0; c < CHAINS; ++c) { INIT; }           \
    for (int it = 0; it < n_it; ++it) {                                     \
      _Pragma("unroll") for (int c = 0; c < CHAINS; ++c) { BODY; }         \
    }                                                                       \
    float acc = 0;                                                          \
    _Pragma("unroll") for (int c = 0; c < CHAINS; ++c) { SINK; }           \
    if (acc == 123.456f) out[threadIdx.x] = acc;                            \
  }

DEF_KERNEL(k_ffma, float v[CHAINS], v[c] = seed + c, asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(v[c]) : "f"(seed)), acc += v[c])
DEF_KERNEL(k_fadd, float v[CHAINS], v[c] = seed + c, asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(v[c]) : "f"(seed)), acc += v[c])
DEF_KERNEL(k_fadd_rm, float v[CHAINS], v[c] = seed + c, asm volatile("add.rm.f32 %0, %0, %1;" : "+f"(v[c]) : "f"(seed)), acc += v[c])
DEF_KERNEL(k_fmul, float v[CHAINS], v[c] = seed + c, asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(v[c]) : "f"(seed)), acc += v[c])
DEF_KERNEL(k_ffma2, u64 v[CHAINS]; u64 s2 = ((u64)__float_as_uint(seed) << 32) | __float_as_uint(seed), v[c] = s2 + c,
           asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(v[c]) : "l"(s2)), acc += __uint_as_float((unsigned)v[c]))
DEF_KERNEL(k_fadd2_rm, u64 v[CHAINS]; u64 s2 = ((u64)__float_as_uint(seed) << 32) | __float_as_uint(seed), v[c] = s2 + c,
           asm volatile("add.rm.f32x2 %0, %0, %1;" : "+l"(v[c]) : "l"(s2)), acc += __uint_as_float((unsigned)v[c]))
DEF_KERNEL(k_fmul2, u64 v[CHAINS]; u64 s2 = ((u64)__float_as_uint(seed) << 32) | __float_as_uint(seed), v[c] = s2 + c,
           asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(v[c]) : "l"(s2)), acc += __uint_as_float((unsigned)v[c]))
DEF_KERNEL(k_lop3, unsigned v[CHAINS]; unsigned s = __float_as_uint(seed), v[c] = s + c,
           asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v[c]) : "r"(s), "r"(c)), acc += v[c])
DEF_KERNEL(k_iadd3, unsigned v[CHAINS]; unsigned s = __float_as_uint(seed), v[c] = s + c,
           asm volatile("add.u32 %0, %0, %1;" : "+r"(v[c]) : "r"(s)), acc += v[c])
DEF_KERNEL(k_shf, unsigned v[CHAINS]; unsigned s = __float_as_uint(seed), v[c] = s + c,
           asm volatile("shf.r.wrap.b32 %0, %0, %1, 7;" : "+r"(v[c]) : "r"(s)), acc += v[c])
DEF_KERNEL(k_imad, unsigned v[CHAINS]; unsigned s = __float_as_uint(seed), v[c] = s + c,
           asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(v[c]) : "r"(s)), acc += v[c])
DEF_KERNEL(k_imad_wide, u64 v[CHAINS]; unsigned s = __float_as_uint(seed), v[c] = s + c,
           asm volatile("{.reg .u32 lo; cvt.u32.u64 lo, %0; mad.wide.u32 %0, lo, %1, %0;}" : "+l"(v[c]) : "r"(s)), acc += (unsigned)v[c])
DEF_KERNEL(k_fmnmx, float v[CHAINS], v[c] = seed + c, asm volatile("min.f32 %0, %0, %1;" : "+f"(v[c]) : "f"(seed)), acc += v[c])
DEF_KERNEL(k_fsetp_sel, float v[CHAINS], v[c] = seed + c,
           asm volatile("{.reg .pred p; setp.eq.f32 p, %0, %1; selp.f32 %0, %1, %0, p;}" : "+f"(v[c]) : "f"(seed)), acc += v[c])
DEF_KERNEL(k_padd, float v[CHAINS], v[c] = seed + c,
           asm volatile("{.reg .pred p; setp.gt.f32 p, %0, %1; @p add.rn.f32 %0, %0, %1;}" : "+f"(v[c]) : "f"(seed)), acc += v[c])
DEF_KERNEL(k_mufu_rcp, float v[CHAINS], v[c] = seed + c, asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(v[c])), acc += v[c])
DEF_KERNEL(k_i2f, float v[CHAINS], v[c] = seed + c,
           asm volatile("{.reg .u32 t; mov.b32 t, %0; cvt.rn.f32.u32 %0, t;}" : "+f"(v[c])), acc += v[c])
DEF_KERNEL(k_i2f_u8, float v[CHAINS], v[c] = seed + c,
           asm volatile("{.reg .u32 t; .reg .u16 h; mov.b32 t, %0; cvt.u16.u32 h, t; cvt.rn.f32.u16 %0, h;}" : "+f"(v[c])), acc += v[c])
DEF_KERNEL(k_f2i_rd, float v[CHAINS], v[c] = seed + c,
           asm volatile("{.reg .s32 t; cvt.rmi.s32.f32 t, %0; mov.b32 %0, t;}" : "+f"(v[c])), acc += v[c])
DEF_KERNEL(k_frnd_floor, float v[CHAINS], v[c] = seed + c, asm volatile("cvt.rmi.f32.f32 %0, %0;" : "+f"(v[c])), acc += v[c])
DEF_KERNEL(k_fdiv, float v[CHAINS], v[c] = seed + c, asm volatile("div.rn.f32 %0, %0, %1;" : "+f"(v[c]) : "f"(seed)), acc += v[c])
// 50/50 mixes: do the fma pipe and the alu pipe overlap?
DEF_KERNEL(k_mix_ffma_lop3, float v[CHAINS]; unsigned w[CHAINS]; unsigned s = __float_as_uint(seed), v[c] = seed + c; w[c] = s + c,
           asm volatile("fma.rn.f32 %0, %0, %2, %2;\n\tlop3.b32 %1, %1, %3, %4, 0x96;" : "+f"(v[c]), "+r"(w[c]) : "f"(seed), "r"(s), "r"(c)),
           acc += v[c] + w[c])
DEF_KERNEL(k_mix_ffma2_lop3, u64 v[CHAINS]; unsigned w[CHAINS]; unsigned s = __float_as_uint(seed); u64 s2 = ((u64)s << 32) | s, v[c] = s2 + c; w[c] = s + c,
           asm volatile("fma.rn.f32x2 %0, %0, %2, %2;\n\tlop3.b32 %1, %1, %3, %4, 0x96;" : "+l"(v[c]), "+r"(w[c]) : "l"(s2), "r"(s), "r"(c)),
           acc += (unsigned)v[c] + w[c])

template <class K>
static void run(const char* name, K kern, double inst_per_iter) {
  float* out;
  cudaMalloc(&out, 4096);
  int dev_clock_khz = 0, sms = 0;
  cudaDeviceGetAttribute(&dev_clock_khz, cudaDevAttrClockRate, 0);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  dim3 grid(sms * 4), block(256);
  cudaEvent_t a, b;
  cudaEventCreate(&a), cudaEventCreate(&b);
  kern<<<grid, block>>>(out, N_IT, 1.0001f);
  kern<<<grid, block>>>(out, N_IT, 1.0001f);
  cudaEventRecord(a);
  for (int r = 0; r < 5; ++r) kern<<<grid, block>>>(out, N_IT, 1.0001f);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  ms /= 5;
  const double warp_inst = (double)grid.x * (block.x / 32) * N_IT * CHAINS * inst_per_iter;
  // cycles at the nominal max clock; the real clock is printed by nvidia-smi beside this run
  const double cyc = ms * 1e-3 * dev_clock_khz * 1e3;
  printf("%-18s %8.3f ms  %6.3f warp-inst/clk/SM (at %d MHz nominal)  = %6.2f thread-ops/clk/SM\n", name, ms, warp_inst / cyc / sms,
         dev_clock_khz / 1000, warp_inst * 32 / cyc / sms);
  cudaFree(out);
}

int main() {
  run("FFMA", k_ffma, 1);
  run("FADD", k_fadd, 1);
  run("FADD.RM", k_fadd_rm, 1);
  run("FMUL", k_fmul, 1);
  run("FFMA2", k_ffma2, 1);
  run("FADD2.RM", k_fadd2_rm, 1);
  run("FMUL2", k_fmul2, 1);
  run("LOP3", k_lop3, 1);
  run("IADD", k_iadd3, 1);
  run("SHF", k_shf, 1);
  run("IMAD", k_imad, 1);
  run("IMAD.WIDE", k_imad_wide, 1);
  run("FMNMX", k_fmnmx, 1);
  run("FSETP+FSEL", k_fsetp_sel, 2);
  run("FSETP+@P FADD", k_padd, 2);
  run("MUFU.RCP", k_mufu_rcp, 1);
  run("I2F.U32", k_i2f, 1);
  run("I2F.U16", k_i2f_u8, 1);
  run("F2I.FLOOR", k_f2i_rd, 1);
  run("FRND.FLOOR", k_frnd_floor, 1);
  run("FDIV(ieee)", k_fdiv, 1);
  run("FFMA+LOP3", k_mix_ffma_lop3, 2);
  run("FFMA2+LOP3", k_mix_ffma2_lop3, 2);
  return 0;
}
