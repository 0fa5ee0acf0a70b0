"""B200-native LBVH build + ray cast behind UnitySimpleRaytracing's entry points."""
