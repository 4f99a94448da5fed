// Stand-in for the third-party header Simplify.h (Fast-Quadric-Mesh-Simplification; not in the reference tree) so that
// BaseLib/motion/Mesh.cpp compiles.  OBJloader::Mesh::simplify is never reached by the pinned paths; simplify_mesh aborts.
#pragma once
#include <vector>
#include <cstdio>
#include <cstdlib>
namespace Simplify {
enum Attributes { NONE = 0, NORMAL = 2, TEXCOORD = 4, COLOR = 8 };
struct Vec3 { double x, y, z; };
struct Triangle { int v[3]; int attr; int material; bool deleted; vectorn uvs[3]; };
struct Vertex { Vec3 p; };
static std::vector<Triangle> triangles;
static std::vector<Vertex> vertices;
inline void simplify_mesh(int, double, bool) { std::fprintf(stderr, "oracle/_ref: Simplify.h is a stub\n"); std::abort(); }
}  // namespace Simplify
