// f2.cuh -- two fp32 values processed by ONE instruction (sm_100a FADD2 / FMUL2 / FFMA2, PTX *.f32x2).
//
// The effect kernels are bound by instruction issue, not by the FP32 pipe (profiles/): a packed
// instruction costs one issue slot for two IEEE-754 round-to-nearest results, each bit-identical to
// the scalar instruction's.  Wherever the reference does the same arithmetic on independent values
// (the reverb's four delay lines, the equalizer's wet channels, the per-output-channel pan adds) the
// kernels hold them as F2 pairs.
//
// No contraction: the reference's products and sums are rounded separately (SURVEY.md section 0,
// fact 5).  ptxas contracts `mul.rn.f32x2` + `add.rn.f32x2` into FFMA2 even with --fmad=false
// (measured, CUDA 12.9), so a product is computed as fma(a, b, -0.0) with the -0.0 pair read from
// constant memory (opaque to the compiler): x*y + (-0) rounds to exactly x*y, signed zeros included,
// and an FFMA2 cannot be merged with the add that consumes it.
//
// The host build (tests/emu) implements F2 as two floats with scalar operations in the same order.
#ifndef OALSFX_F2_CUH
#define OALSFX_F2_CUH

namespace oalsfx {

#if defined(__CUDACC__)
static __constant__ unsigned long long kF2NegZero = 0x8000000080000000ULL; // (-0.0f, -0.0f)
#endif

#if defined(__CUDA_ARCH__)

struct F2 { unsigned long long v; };

OALSFX_HD F2 f2(float lo, float hi)
{
	F2 r;
	asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi));
	return r;
}
OALSFX_HD float f2_lo(F2 a)
{
	[[maybe_unused]] float lo, hi;
	asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a.v));
	return lo;
}
OALSFX_HD float f2_hi(F2 a)
{
	[[maybe_unused]] float lo, hi;
	asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a.v));
	return hi;
}
OALSFX_HD F2 f2_add(F2 a, F2 b)
{
	F2 r;
	asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
	return r;
}
OALSFX_HD F2 f2_sub(F2 a, F2 b)
{
	F2 r;
	asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
	return r;
}
OALSFX_HD F2 f2_mul(F2 a, F2 b)
{
	F2 r;
	asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(kF2NegZero));
	return r;
}

#else

struct F2 { float lo, hi; };

OALSFX_HD F2 f2(float lo, float hi) { return F2{lo, hi}; }
OALSFX_HD float f2_lo(F2 a) { return a.lo; }
OALSFX_HD float f2_hi(F2 a) { return a.hi; }
OALSFX_HD F2 f2_add(F2 a, F2 b) { return F2{a.lo + b.lo, a.hi + b.hi}; }
OALSFX_HD F2 f2_sub(F2 a, F2 b) { return F2{a.lo - b.lo, a.hi - b.hi}; }
OALSFX_HD F2 f2_mul(F2 a, F2 b) { return F2{a.lo * b.lo, a.hi * b.hi}; }

#endif

OALSFX_HD F2 f2_bcast(float s) { return f2(s, s); }
OALSFX_HD F2 operator+(F2 a, F2 b) { return f2_add(a, b); }
OALSFX_HD F2 operator-(F2 a, F2 b) { return f2_sub(a, b); }
OALSFX_HD F2 operator*(F2 a, F2 b) { return f2_mul(a, b); }
OALSFX_HD F2 operator*(F2 a, float s) { return f2_mul(a, f2_bcast(s)); }
OALSFX_HD F2 operator*(float s, F2 a) { return f2_mul(f2_bcast(s), a); }

} // namespace oalsfx

#endif
