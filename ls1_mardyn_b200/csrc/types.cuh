// types.cuh -- calc-precision vector records (float4 in the mixed build, double4 in the double build)
#pragma once
#include "common.cuh"

namespace ls1b200 {

template <class C> struct Vec4;
template <> struct Vec4<float> { typedef float4 type; };
template <> struct Vec4<double> { typedef double4 type; };

template <class C> __device__ __forceinline__ typename Vec4<C>::type make_mrec(double x, double y, double z, uint32_t bits);
template <> __device__ __forceinline__ float4 make_mrec<float>(double x, double y, double z, uint32_t bits) {
	return make_float4((float) x, (float) y, (float) z, __uint_as_float(bits));
}
template <> __device__ __forceinline__ double4 make_mrec<double>(double x, double y, double z, uint32_t bits) {
	return make_double4(x, y, z, __longlong_as_double((long long) bits));
}

template <class C> __device__ __forceinline__ typename Vec4<C>::type make_c4(C x, C y, C z, C w);
template <> __device__ __forceinline__ float4 make_c4<float>(float x, float y, float z, float w) { return make_float4(x, y, z, w); }
template <> __device__ __forceinline__ double4 make_c4<double>(double x, double y, double z, double w) { return make_double4(x, y, z, w); }

template <class C> __device__ __forceinline__ uint32_t wbits(C w);
template <> __device__ __forceinline__ uint32_t wbits<float>(float w) { return __float_as_uint(w); }
template <> __device__ __forceinline__ uint32_t wbits<double>(double w) { return (uint32_t) __double_as_longlong(w); }

}  // namespace ls1b200
