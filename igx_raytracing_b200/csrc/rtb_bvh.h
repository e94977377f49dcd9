// rtb_bvh.h — host-side BVH builder (new functionality: the reference has no acceleration structure).
#pragma once
#include <stdint.h>
#include <vector>
#include "rtb_types.h"

namespace rtb {

struct BvhStats {
    uint32_t nodeCount = 0, leafCount = 0, maxDepth = 0;
    float sahCost = 0.0f, buildMs = 0.0f;
    float lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};   // bounds of the padded triangle boxes
    float leafNodeExtent = 0.0f;                  // 8-wide tree: mean edge length of the nodes that hold only triangles
    std::vector<uint32_t> levelFirst;             // 8-wide tree: first node of every level, then nodeCount (levels are index ranges)
};

// Binned-SAH binary BVH over `count` reference-layout triangles.
//  nodes    : 64-byte records; the first min(nodeCount, topNodes) are the breadth-first top of the tree
//             (the traversal kernel stages them in shared memory), the rest depth-first.
//  travTris : 48-byte (p0, e1, e2, id) records in leaf order.
// Every box is padded by 2^-18 * (largest absolute vertex coordinate) so that the fused slab test can
// never cull a triangle the reference's Möller–Trumbore test would accept (DESIGN.md "conservative boxes").
void buildBvh(const TriangleRec* tris, uint32_t count, uint32_t topNodes, int threads,
              std::vector<BvhNode>& nodes, std::vector<TravTri>& travTris, BvhStats& stats);

// 8-wide compressed BVH (80-byte Node8 records, breadth-first so the children of a node are consecutive) and its
// traversal triangles (children's triangles consecutive per node).  Same padded boxes as buildBvh, rounded outwards
// onto each node's 8-bit grid.
void buildCwbvh(const TriangleRec* tris, uint32_t count, int threads, std::vector<Node8>& nodes, std::vector<TravTri>& travTris, BvhStats& stats);

constexpr uint32_t CWBVH_MAX_LEAF = 3;    // triangles per leaf slot (unary count in 3 bits)
constexpr uint32_t BVH_MAX_LEAF = 4;      // triangles per leaf (the link encodes count-1 in 3 bits)
constexpr uint32_t BVH_MAX_DEPTH = 60;    // traversal stack holds 64 entries

}  // namespace rtb
