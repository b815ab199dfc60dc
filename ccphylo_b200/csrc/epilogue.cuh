/*
 * epilogue.cuh -- the fused epilogue: integer counts -> one cell of D (and N).
 *
 * Pair mode restates reference fsacmpthrd.c:419-475, shared-mask mode
 * fsacmpthrd.c:247-255; fixed-point cells follow bytescale.h:22 (dtouc).
 * Every floating-point step is an explicit IEEE round-to-nearest intrinsic so
 * nvcc cannot contract a multiply-add into an FMA: the results must be
 * bit-identical to the reference's x86-64 SSE2 doubles / floats.
 */
#ifndef CCG_EPILOGUE_CUH
#define CCG_EPILOGUE_CUH

#include "ccg_internal.h"

/* double -> int32 as x86-64 cvttsd2si: truncate; out of range / NaN gives
 * INT32_MIN ("integer indefinite").  The reference stores the low 16 / 8 bits
 * of that into its unsigned short / unsigned char cells. */
__device__ __forceinline__ int ccg_cvttsd2si(double x) {
	if(!(x > -2147483649.0 && x < 2147483648.0)) return (int) 0x80000000;
	return __double2int_rz(x);
}

__device__ __forceinline__ void ccg_store_fixed(void *base, long long cell, int elem_size, double v) {
	int t = ccg_cvttsd2si(v);
	if(elem_size == 2) ((unsigned short *) base)[cell] = (unsigned short) t;
	else ((unsigned char *) base)[cell] = (unsigned char) t;
}

/* slot_i > slot_j are sample slots; excluded samples have rank < 0. */
__device__ __forceinline__ void ccg_write_cell(const EpilogueParams &ep, int slot_i, int slot_j,
                                               unsigned mism, unsigned inc) {
	int r = ep.rank[slot_i];
	int c = ep.rank[slot_j];
	if(r < 0 || c < 0) return;
	long long cell = ep.row_base ? ep.row_base[r] + c : (long long) r * (r - 1) / 2 + c;
	if(ep.row_plus1) {
		if(r + 1 != ep.row_plus1) return;
		cell = c;
	}

	if(ep.mode == 0) {
		bool ok = ep.minLength <= inc;
		if(ep.row_plus1 && !ok) inc = 0;
		unsigned long long scaled = (unsigned long long) mism * ep.norm;
		if(ep.elem_size == 8) {
			double d;
			if(!ok) d = -1.0;
			else if(ep.norm) d = __ddiv_rn(__ull2double_rn(scaled), __uint2double_rn(inc));
			else d = __uint2double_rn(mism);
			((double *) ep.D)[cell] = d;
			if(ep.N) ((double *) ep.N)[cell] = __uint2double_rn(inc);
		} else if(ep.elem_size == 4) {
			float f;
			if(!ok) f = -1.0f;
			else if(ep.norm) {
				f = __fdiv_rn(__ull2float_rn(scaled), __uint2float_rn(inc));
				/* 0 / 0 (no included position and a zero threshold): the reference's SSE division gives the default NaN
				 * with the sign bit set, which prints as "-nan"; CUDA's canonical float NaN is 0x7fffffff ("nan") */
				if(f != f) f = __int_as_float((int) 0xFFC00000u);
			} else f = __uint2float_rn(mism);
			((float *) ep.D)[cell] = f;
			if(ep.N) ((float *) ep.N)[cell] = __uint2float_rn(inc);
		} else {
			double d;
			if(!ok) d = __dadd_rn(__dmul_rn(-1.0, ep.byteScale), 0.0);
			else if(ep.norm) d = __ddiv_rn(__dadd_rn(__dmul_rn(__ull2double_rn(scaled), ep.byteScale), 0.5), __uint2double_rn(inc));
			else d = __dadd_rn(__dmul_rn(__uint2double_rn(mism), ep.byteScale), 0.5);
			ccg_store_fixed(ep.D, cell, ep.elem_size, d);
			if(ep.N) ccg_store_fixed(ep.N, cell, ep.elem_size, __dadd_rn(__dmul_rn(__uint2double_rn(inc), ep.byteScale), 0.5));
		}
	} else {
		double v = __dmul_rn(ep.nFactor, __uint2double_rn(mism));
		if(ep.elem_size == 8) ((double *) ep.D)[cell] = v;
		else if(ep.elem_size == 4) ((float *) ep.D)[cell] = __double2float_rn(v);
		else ccg_store_fixed(ep.D, cell, ep.elem_size, __dadd_rn(__dmul_rn(v, ep.byteScale), 0.5));
	}
}

#endif
