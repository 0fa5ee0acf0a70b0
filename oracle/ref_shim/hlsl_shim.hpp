// hlsl_shim.hpp -- TEST INFRASTRUCTURE ONLY (part of the oracle/_ref recipe, see oracle/build_ref.sh).
//
// A minimal C++ environment in which the reference's OWN HLSL text compiles with g++:
// /root/reference/Assets/_Shaders/BVH/BVH.compute, Raytracing/Raytracing.compute and Sorting/*.compute are fed
// through a purely syntactic sed pass (strip `#pragma`, `[numthreads(..)]`, `: SV_*` semantics, give float literals
// an `f` suffix -- HLSL literals are fp32 --, spell `.xyz` / `.xy` swizzles as calls) and then #included below the
// declarations in this header. Constants.cginc is included unmodified (-I /root/reference). No reference text lives
// in this repository: the generated translation units go to oracle/_ref/ (git-ignored).
//
// What this header DEFINES is only what HLSL leaves to the implementation:
//   dot(a,b)      = a.x*b.x + a.y*b.y + a.z*b.z, left to right, no FMA contraction (-ffp-contract=off)
//   cross(a,b)    = (a.y*b.z - a.z*b.y, a.z*b.x - a.x*b.z, a.x*b.y - a.y*b.x)
//   mul(M,v)      = row . vector, left to right, w term included
//   normalize(v)  = v / sqrt(dot(v,v)), IEEE sqrt and division per component
//   1 / x, a / b  = IEEE division
//   min / max     = IEEE minNum / maxNum (the non-NaN operand wins), as DXC's FMin / FMax
//   SampleLevel   = bilinear, clamp, texel centres at (i + 0.5) / size (the oracle's definition, usrt_oracle.cpp)
// Everything else -- control flow, operation order, integer widths, which buffer is read when -- is the reference's.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>

typedef uint32_t uint;

struct float2 {
    float x, y;
    float2() : x(0), y(0) {}
    float2(float x_, float y_) : x(x_), y(y_) {}
};
struct float4;
struct float3 {
    float x, y, z;
    float3() : x(0), y(0), z(0) {}
    float3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
    float3(const float4& v);                       // HLSL implicit truncation float4 -> float3 (Raytracing.compute:183)
    operator float() const { return x; }           // HLSL implicit truncation float3 -> float  (Raytracing.compute:181)
};
struct float4 {
    float x, y, z, w;
    float4() : x(0), y(0), z(0), w(0) {}
    float4(float x_, float y_, float z_, float w_) : x(x_), y(y_), z(z_), w(w_) {}
    float4(const float3& v, float w_) : x(v.x), y(v.y), z(v.z), w(w_) {}
    float3 xyz() const { return float3(x, y, z); }
};
inline float3::float3(const float4& v) : x(v.x), y(v.y), z(v.z) {}
struct int2 {
    int x, y;
    int2() : x(0), y(0) {}
    int2(int x_, int y_) : x(x_), y(y_) {}
};
struct uint2 { uint x, y; };
struct uint3 {
    uint x, y, z;
    uint3() : x(0), y(0), z(0) {}
    uint3(uint x_, uint y_, uint z_) : x(x_), y(y_), z(z_) {}
    uint2 xy() const { return uint2{x, y}; }
};
struct float4x4 { float m[4][4]; };

// ---- arithmetic (component-wise, fp32, one rounding per operation) --------------------------------------------------
inline float3 operator+(const float3& a, const float3& b) { return float3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline float3 operator-(const float3& a, const float3& b) { return float3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline float3 operator*(const float3& a, const float3& b) { return float3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline float3 operator*(float s, const float3& a) { return float3(s * a.x, s * a.y, s * a.z); }
inline float3 operator*(const float3& a, float s) { return float3(a.x * s, a.y * s, a.z * s); }
inline float3 operator/(const float3& a, float s) { return float3(a.x / s, a.y / s, a.z / s); }
inline float3 operator/(float s, const float3& a) { return float3(s / a.x, s / a.y, s / a.z); }
inline float3 operator/(int s, const float3& a) { return float3((float)s / a.x, (float)s / a.y, (float)s / a.z); }
inline float2 operator+(const float2& a, const float2& b) { return float2(a.x + b.x, a.y + b.y); }
inline float2 operator*(float s, const float2& a) { return float2(s * a.x, s * a.y); }
inline float4 operator*(const float4& a, float s) { return float4(a.x * s, a.y * s, a.z * s, a.w * s); }

inline float min(float a, float b) { return (a != a) ? b : ((b != b) ? a : (a < b ? a : b)); }
inline float max(float a, float b) { return (a != a) ? b : ((b != b) ? a : (a > b ? a : b)); }
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
inline float3 min(const float3& a, const float3& b) { return float3(min(a.x, b.x), min(a.y, b.y), min(a.z, b.z)); }
inline float3 max(const float3& a, const float3& b) { return float3(max(a.x, b.x), max(a.y, b.y), max(a.z, b.z)); }

inline float dot(const float3& a, const float3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float dot(float s, const float3& b) { return dot(float3(s, s, s), b); }      // a scalar operand is splatted
inline float3 cross(const float3& a, const float3& b) {
    return float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
inline float3 normalize(const float3& v) {
    const float len = sqrtf(dot(v, v));
    return float3(v.x / len, v.y / len, v.z / len);
}
inline float4 mul(const float4x4& M, const float4& v) {
    float r[4];
    for (int i = 0; i < 4; ++i) r[i] = M.m[i][0] * v.x + M.m[i][1] * v.y + M.m[i][2] * v.z + M.m[i][3] * v.w;
    return float4(r[0], r[1], r[2], r[3]);
}
inline int sign(int v) { return (v > 0) - (v < 0); }
// firstbithigh(0) is -1 (0xFFFFFFFF) in HLSL, so clz32(0) = 31 - 0xFFFFFFFF = 32 in uint arithmetic
inline uint firstbithigh(uint v) { return v ? 31u - (uint)__builtin_clz(v) : 0xFFFFFFFFu; }

// ---- resources ---------------------------------------------------------------------------------------------------------
template <typename T> struct StructuredBuffer {
    const T* data = nullptr;
    const T& operator[](size_t i) const { return data[i]; }
};
template <typename T> struct RWStructuredBuffer {
    T* data = nullptr;
    T& operator[](size_t i) const { return data[i]; }
};
struct SamplerState {};
template <typename T> struct Texture2D {
    const float* texels = nullptr;     // width x height float4 texels, row 0 at v = 0
    int width = 0, height = 0;
    float4 SampleLevel(const SamplerState&, const float2& uv, int) const {
        const float x = uv.x * (float)width - 0.5f, y = uv.y * (float)height - 0.5f;
        const float x0f = floorf(x), y0f = floorf(y);
        const float fx = x - x0f, fy = y - y0f;
        auto clampi = [](float f, int hi) { if (!(f >= 0.0f)) return 0; if (f > (float)hi) return hi; return (int)f; };
        const int x0 = clampi(x0f, width - 1), x1 = clampi(x0f + 1.0f, width - 1);
        const int y0 = clampi(y0f, height - 1), y1 = clampi(y0f + 1.0f, height - 1);
        const float* c00 = texels + ((size_t)y0 * width + x0) * 4; const float* c10 = texels + ((size_t)y0 * width + x1) * 4;
        const float* c01 = texels + ((size_t)y1 * width + x0) * 4; const float* c11 = texels + ((size_t)y1 * width + x1) * 4;
        float o[4];
        for (int k = 0; k < 4; ++k) {
            const float top = c00[k] + (c10[k] - c00[k]) * fx;
            const float bot = c01[k] + (c11[k] - c01[k]) * fx;
            o[k] = top + (bot - top) * fy;
        }
        return float4(o[0], o[1], o[2], o[3]);
    }
};
template <typename T> struct RWTexture2D {
    T* pixels = nullptr;
    int width = 0;
    T& operator[](const uint2& p) const { return pixels[(size_t)p.y * width + p.x]; }
};

inline void InterlockedCompareExchange(uint& dest, uint compare_value, uint value, uint& original_value) {
    original_value = dest;                          // threads of a dispatch run one after the other here
    if (dest == compare_value) dest = value;
}

#define uniform
