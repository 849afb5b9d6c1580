// Carry-chain primitives for 256-bit integer arithmetic on sm_100a.
//
// On the device each wrapper is ONE PTX instruction (add.cc / addc.cc /
// mad.lo.cc / madc.hi.cc ...).  ptxas fuses a (mad.lo.cc, madc.hi.cc) pair on
// the same multiplicands into a single IMAD.WIDE.U32[.X] with a predicate
// carry, which is what the field multiplier in pa_fe.cuh is built from
// (32 x 32 -> 64 multiply-accumulate chains, north_star bullet 1).
//
// When the header is compiled by a host compiler (tests/hostcheck only — the
// product library is device code and has no CPU path) the same wrappers are
// emulated with an explicit carry flag so the limb algorithms above them can be
// checked on a machine without a GPU.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define PA_HD __host__ __device__ __forceinline__
#define PA_D __device__ __forceinline__
#else
#define PA_HD inline
#define PA_D inline
#endif

typedef uint32_t u32;
typedef uint64_t u64;

#if defined(__CUDACC__)
__constant__ u32 pa_opq_z;  // always 0, never written: an operand ptxas cannot fold (pa_fe.cuh, PA_OPQ_MODE)
#endif

#if defined(__CUDA_ARCH__)

PA_D u32 add_cc(u32 a, u32 b) { u32 r; asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
PA_D u32 addc_cc(u32 a, u32 b) { u32 r; asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
PA_D u32 addc(u32 a, u32 b) { u32 r; asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
PA_D u32 sub_cc(u32 a, u32 b) { u32 r; asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
PA_D u32 subc_cc(u32 a, u32 b) { u32 r; asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
PA_D u32 subc(u32 a, u32 b) { u32 r; asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
PA_D u32 mul_lo(u32 a, u32 b) { u32 r; asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
PA_D u32 mul_hi(u32 a, u32 b) { u32 r; asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
PA_D u32 mad_lo_cc(u32 a, u32 b, u32 c) { u32 r; asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
PA_D u32 mad_hi_cc(u32 a, u32 b, u32 c) { u32 r; asm volatile("mad.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
PA_D u32 madc_lo_cc(u32 a, u32 b, u32 c) { u32 r; asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
PA_D u32 madc_hi_cc(u32 a, u32 b, u32 c) { u32 r; asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
PA_D u32 madc_hi(u32 a, u32 b, u32 c) { u32 r; asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
PA_D u32 madc_lo(u32 a, u32 b, u32 c) { u32 r; asm volatile("madc.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }

#else  // host emulation of the PTX condition-code register (tests/hostcheck only)

static thread_local u32 pa_cc_ = 0;
inline u32 add_cc(u32 a, u32 b) { u64 s = (u64)a + b; pa_cc_ = (u32)(s >> 32); return (u32)s; }
inline u32 addc_cc(u32 a, u32 b) { u64 s = (u64)a + b + pa_cc_; pa_cc_ = (u32)(s >> 32); return (u32)s; }
inline u32 addc(u32 a, u32 b) { u64 s = (u64)a + b + pa_cc_; return (u32)s; }
inline u32 sub_cc(u32 a, u32 b) { u64 s = (u64)a - b; pa_cc_ = (u32)(s >> 63); return (u32)s; }
inline u32 subc_cc(u32 a, u32 b) { u64 s = (u64)a - b - pa_cc_; pa_cc_ = (u32)(s >> 63); return (u32)s; }
inline u32 subc(u32 a, u32 b) { u64 s = (u64)a - b - pa_cc_; return (u32)s; }
inline u32 mul_lo(u32 a, u32 b) { return (u32)((u64)a * b); }
inline u32 mul_hi(u32 a, u32 b) { return (u32)(((u64)a * b) >> 32); }
inline u32 mad_lo_cc(u32 a, u32 b, u32 c) { u64 s = (u64)mul_lo(a, b) + c; pa_cc_ = (u32)(s >> 32); return (u32)s; }
inline u32 mad_hi_cc(u32 a, u32 b, u32 c) { u64 s = (u64)mul_hi(a, b) + c; pa_cc_ = (u32)(s >> 32); return (u32)s; }
inline u32 madc_lo_cc(u32 a, u32 b, u32 c) { u64 s = (u64)mul_lo(a, b) + c + pa_cc_; pa_cc_ = (u32)(s >> 32); return (u32)s; }
inline u32 madc_hi_cc(u32 a, u32 b, u32 c) { u64 s = (u64)mul_hi(a, b) + c + pa_cc_; pa_cc_ = (u32)(s >> 32); return (u32)s; }
inline u32 madc_hi(u32 a, u32 b, u32 c) { u64 s = (u64)mul_hi(a, b) + c + pa_cc_; return (u32)s; }
inline u32 madc_lo(u32 a, u32 b, u32 c) { u64 s = (u64)mul_lo(a, b) + c + pa_cc_; return (u32)s; }

#endif
