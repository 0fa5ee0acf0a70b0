// unity_shim.hpp -- TEST INFRASTRUCTURE ONLY (part of the oracle/_ref recipe, see oracle/build_ref.sh).
//
// The static functions of /root/reference/Assets/_Scripts/MeshBufferContainer.cs (ExpandBits :32-39, Morton3D
// :41-50, GetCentroidAndAABB :52-71, NormalizeCentroid :73-83) and the DistributeKeys loop (:154-169) are plain
// C-like C#. build_ref.sh cuts those line ranges out of the reference file, re-spells the C#-only syntax
// (`private static` -> `static`, `out T x` -> `T& x`, `new T(` -> `T(`, `Math.` -> `Math::`, object initialiser ->
// designated initialiser) and #includes the result below these declarations. No reference text lives in this repo.
#pragma once
#include <cstddef>
#include <cstdint>

typedef uint32_t uint;

// UnityEngine.Vector3: three fp32 fields; operator+ and operator*(Vector3, float) are component-wise fp32
struct Vector3 {
    float x, y, z;
    Vector3() : x(0), y(0), z(0) {}
    Vector3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
};
inline Vector3 operator+(const Vector3& a, const Vector3& b) { return Vector3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline Vector3 operator*(const Vector3& a, float d) { return Vector3(a.x * d, a.y * d, a.z * d); }

// SceneDataTypes.cs:4-16 ([StructLayout(Sequential, Pack = 16)], 32 bytes; unnamed fields default to 0)
struct AABB {
    Vector3 min;
    float _dummy0;
    Vector3 max;
    float _dummy1;
};
static_assert(sizeof(AABB) == 32, "MeshBufferContainer.cs:103");

// System.Math.Min/Max(float, float) on the finite values this path feeds them, and Math.Max(uint, uint):
// C# resolves Math.Max(uint, 1) to the unsigned overload (the literal converts), which this overload set reproduces.
struct Math {
    static float Min(float a, float b) { return a < b ? a : b; }
    static float Max(float a, float b) { return a > b ? a : b; }
    static uint Max(uint a, uint b) { return a > b ? a : b; }
};

// MeshBufferContainer.cs:9-15:  size = 125f;  Whole = { min = Vector3.one * -1 * size, max = Vector3.one * size }
static const float size = 125.0f;
static const AABB Whole = {Vector3(1.0f * -1 * size, 1.0f * -1 * size, 1.0f * -1 * size), 0.0f,
                           Vector3(1.0f * size, 1.0f * size, 1.0f * size), 0.0f};

// DataBuffer<uint> (DataBuffer.cs:5-76) as far as DistributeKeys touches it: a host array; GetData / Sync are the
// GPU <-> CPU copies, which have nothing to do here
struct KeysBufferShim {
    uint* LocalBuffer = nullptr;
    void GetData() {}
    void Sync() {}
};
static KeysBufferShim _keysBuffer;
static uint _trianglesLength = 0;
