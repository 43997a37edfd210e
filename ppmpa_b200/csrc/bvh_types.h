// bvh_types.h -- the bounding-volume hierarchy over the BOUNDED primitives of a scene (spheres, polygons,
// parallelograms), shared by the host builder (host_bvh.cpp) and the device traversal (bvh.cuh).
//
// The reference tests every object for every ray (calc_intersection, tracer.rs:306-350).  Beyond PPM_MAX_PRIMS
// primitives (or with the "bvh" option) the engine walks this hierarchy instead; the candidates it does not cull are
// tested with exactly the same arithmetic and tie rule, so the nearest hit is the one the brute-force scan returns.
//
// Layout (HBM, read through L1/L2 with 16-byte loads):
//   BvhNode, 128 B: the f64 boxes of BOTH children (one fetch decides both) + two child references.
//   BvhPrim,  80 B: what the hit test reads (point + two edge vectors, or centre + radius), in leaf order, with the
//                   primitive's object index (the reference's tie rule needs it) and its shape.
// Infinite planes are not in the hierarchy; they stay in the constant-memory primitive list.
#ifndef PPM_BVH_TYPES_H_
#define PPM_BVH_TYPES_H_

#include <stdint.h>

#define PPM_BVH_LEAF 0x80000000u          // child reference: leaf flag | (count - 1) << 28 | first BvhPrim slot
#define PPM_BVH_NONE 0xFFFFFFFFu          // child reference: no child (a root with one leaf)
#ifndef PPM_BVH_LEAF_MAX
#define PPM_BVH_LEAF_MAX 2u               // primitives per leaf (1 / 2 / 3 / 4 / 8 swept: profiles/r2_bvh_leaf_sweep.txt; the child reference has 3 bits for the count)
#endif
#define PPM_BVH_STACK 64                  // traversal stack; the builder bounds the depth (PPM_BVH_SAH_DEPTH + log2 N)
#define PPM_BVH_SAH_DEPTH 30              // below this depth splits are by SAH, beyond it by object median
#define PPM_BVH_MAX_PRIMS (1u << 26)

struct BvhNode {
  double box[2][6];        // child k: lo.xyz, hi.xyz (padded, see host_bvh.cpp)
  uint32_t child[2];
  uint32_t _pad[6];
};
struct BvhPrim {
  double p0[3];            // Polygon / Parallelogram position | Sphere centre
  double d1[3];            // edge 1 | (radius, 0, 0)
  double d2[3];            // edge 2
  int32_t obj;             // object index in the scene (order of ppm_scene_set)
  int32_t type;            // PPM_SHAPE_* in bits 0..7; bit 8 + l: the primitive lies in the plane of area light l
};                         // (the emitter's own geometry: class (b) of the shadow-ray culling, kernels_eye.cuh)
#ifdef __cplusplus
static_assert(sizeof(BvhNode) == 128, "BvhNode layout");
static_assert(sizeof(BvhPrim) == 80, "BvhPrim layout");
#endif

#endif
