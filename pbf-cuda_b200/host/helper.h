// helper.h — the few non-GL declarations of the reference's fluids/helper.h that host code above
// the solver relies on (typedef uint, float3/int3, ceilDiv), so that sources written against the
// reference's headers compile unchanged against this shim. Plain C++: no CUDA headers needed.
#pragma once
#include <math.h>
#include <stdint.h>

#ifndef __VECTOR_TYPES_H__   // CUDA's vector_types.h not included: supply the three types we use
struct float3 { float x, y, z; };
struct int3 { int x, y, z; };
static inline float3 make_float3(float x, float y, float z) { float3 r = {x, y, z}; return r; }
static inline int3 make_int3(int x, int y, int z) { int3 r = {x, y, z}; return r; }
static inline float3 operator+(float3 a, float3 b) { return make_float3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline float3 operator-(float3 a, float3 b) { return make_float3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline float3 operator*(float3 a, float b) { return make_float3(a.x * b, a.y * b, a.z * b); }
#endif

typedef unsigned int uint;

// reference helper.h:10 hard-codes 130000; here it is only the DEFAULT capacity of a Simulator
#ifndef MAX_PARTICLE_NUM
#define MAX_PARTICLE_NUM 130000
#endif

inline int ceilDiv(int a, int b) { return (int)((a + b - 1) / b); }
