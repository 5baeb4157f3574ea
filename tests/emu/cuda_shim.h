// cuda_shim.h -- TEST INFRASTRUCTURE.  Host definitions of the CUDA device intrinsics wx_device.cuh uses, so that the
// header can be compiled by g++ (-ffp-contract=off) and the very code the kernels run can be executed on the CPU, lane by
// lane, and compared with the oracle without a GPU.  Every shim is the IEEE-754 binary32 operation the PTX ISA specifies for
// the intrinsic; MUFU.RCP (rcp.approx, <= 1 ulp) is 1/v bumped by a test-selected number of ulps, because the march's
// result must not depend on which of the allowed values the hardware returns.
#pragma once
#define WX_HOST_EMU 1
#define WX_NO_F32X2 1  // the packed forms are two independent IEEE operations: use the scalar definitions
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#ifndef __noinline__
#define __noinline__ __attribute__((noinline))
#endif

static inline float __uint_as_float(uint32_t u) {
  float f;
  memcpy(&f, &u, 4);
  return f;
}
static inline uint32_t __float_as_uint(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  return u;
}
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
// add.rm.f32: the exact sum rounded toward -inf.  s = RN(a + b) and the exact rounding error (TwoSum, Knuth):
// a negative error means RN rounded up.
static inline float __fadd_rd(float a, float b) {
  volatile float s = a + b;
  if (!isfinite(s)) {
    if (isinf(s) && isfinite(a) && isfinite(b) && s > 0.f) return 3.40282346638528859812e+38f;  // overflow rounds down to FLT_MAX
    return s;
  }
  volatile float bb = s - a;
  volatile float err = (a - (s - bb)) + (b - bb);
  if (err < 0.f) return nextafterf(s, -INFINITY);
  if (s == 0.f && err == 0.f) {  // exact zero: -0 in round-down unless both operands are +0
    const bool both_pos_zero = a == 0.f && b == 0.f && !signbit(a) && !signbit(b);
    return both_pos_zero ? 0.f : -0.f;
  }
  return s;
}
static inline int __float2int_rd(float v) {  // cvt.rmi.s32.f32: saturating, NaN -> 0
  if (v != v) return 0;
  const float f = floorf(v);
  if (f >= 2147483648.f) return 2147483647;
  if (f < -2147483648.f) return (int)-2147483648LL;
  return (int)f;
}
static inline int __float2int_rn(float v) {  // cvt.rni.s32.f32: round half to even
  if (v != v) return 0;
  const float f = nearbyintf(v);  // default rounding mode
  if (f >= 2147483648.f) return 2147483647;
  if (f < -2147483648.f) return (int)-2147483648LL;
  return (int)f;
}
template <class T>
static inline T __ldg(const T* p) { return *p; }
// atomics of the work queue (sequentially consistent here; the protocol relies on atomicity only)
static inline uint32_t atomicAdd(uint32_t* p, uint32_t v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline uint32_t atomicCAS(uint32_t* p, uint32_t compare, uint32_t val) {
  __atomic_compare_exchange_n(p, &compare, val, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
  return compare;  // the old value, as CUDA returns it
}

// rcp.approx.ftz.f32 stand-in
extern thread_local int wx_emu_rcp_bump;  // ulps added to RN(1/v): -1, 0, +1
static inline float wx_emu_rcp(float v) {
  float r = 1.0f / v;
  for (int k = 0; k < wx_emu_rcp_bump; ++k) r = nextafterf(r, INFINITY);
  for (int k = 0; k > wx_emu_rcp_bump; --k) r = nextafterf(r, -INFINITY);
  return r;
}

// per-step trace of the fast march (see wx_emu.cpp): dv0 = cursor test result at the start of the lookup, dbits = cursor depth after it
struct WxEmuTrace {
  uint8_t* steps;  // one byte per step: (start level class << 4) | end level class
  uint32_t n, cap;
};
extern thread_local WxEmuTrace* wx_emu_trace;
static inline void wx_emu_step(uint32_t dv0, uint32_t dbits) {
  WxEmuTrace* t = wx_emu_trace;
  if (!t || t->n >= t->cap) return;
  // start: 0 root (dv >= 4096), 1 N5 table, 2 N4 table, 3 leaf brick ; end: level the lookup ended on (0 none, 1 N5 slot, 2 N4 slot, 3 leaf)
  const uint32_t start = dv0 >= 4096u ? 0u : (dv0 >= 128u ? 1u : (dv0 >= 8u ? 2u : 3u));
  const uint32_t end = (dbits & 0x80000000u) ? ((dbits >> 28) & 7u) : (dbits == 0u ? 3u : (dbits == 8u ? 2u : 1u));
  t->steps[t->n++] = (uint8_t)((start << 4) | end);
}
#define WX_EMU_STEP(dv0, dbits) wx_emu_step(dv0, dbits)
