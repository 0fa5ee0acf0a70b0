"""Byte-exact numpy mirrors of the reference's GPU/C# struct layouts (the ABI of this path).

Reference: Assets/_Shaders/Constants.cginc:9-54 (HLSL) and Assets/_Scripts/SceneDataTypes.cs:4-90
(C#, StructLayout(Sequential, Pack=16)); hit record: Assets/_Shaders/Raytracing/Raytracing.compute:30-35.
The same layouts are declared for C in include/usrt.h.
"""
import numpy as np

AABB = np.dtype([("min", "<f4", 3), ("_dummy0", "<f4"), ("max", "<f4", 3), ("_dummy1", "<f4")])

InternalNode = np.dtype([("leftNode", "<u4"), ("leftNodeType", "<u4"), ("rightNode", "<u4"),
                         ("rightNodeType", "<u4"), ("parent", "<u4"), ("index", "<u4")])

LeafNode = np.dtype([("parent", "<u4"), ("index", "<u4")])

Triangle = np.dtype([("a", "<f4", 3), ("_dummy0", "<f4"), ("b", "<f4", 3), ("_dummy1", "<f4"),
                     ("c", "<f4", 3), ("_dummy2", "<f4"),
                     ("a_uv", "<f4", 2), ("b_uv", "<f4", 2), ("c_uv", "<f4", 2), ("_dummy3", "<f4", 2),
                     ("a_normal", "<f4", 3), ("_dummy4", "<f4"), ("b_normal", "<f4", 3), ("_dummy5", "<f4"),
                     ("c_normal", "<f4", 3), ("_dummy6", "<f4")])

RaycastResult = np.dtype([("distance", "<f4"), ("triangleIndex", "<u4"), ("uv", "<f4", 2)])

# MeshBufferContainer.cs:98-106 checks these two; the rest follow from Constants.cginc.
assert Triangle.itemsize == 128 and AABB.itemsize == 32
assert InternalNode.itemsize == 24 and LeafNode.itemsize == 8 and RaycastResult.itemsize == 16

INTERNAL_NODE = 0            # Constants.cginc:17
LEAF_NODE = 1                # Constants.cginc:18
NULL = 0xFFFFFFFF            # SceneDataTypes.cs:63-89 NullLeaf, MeshBufferContainer.cs:108-109 padding
# Constants.cginc:7: MAX_FLOAT is the integer literal 0x7F7FFFFF converted to float (bits 0x4EFF0000).
MAX_FLOAT = np.float32(0x7F7FFFFF)
WHOLE_MIN, WHOLE_MAX = -125.0, 125.0   # MeshBufferContainer.cs:9-15
