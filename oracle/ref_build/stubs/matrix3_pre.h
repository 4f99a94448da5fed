// Force-included in front of BaseLib/math/matrix3.cpp ONLY.  matrix3.cpp pulls PhysicsLib/TRL/eigenSupport.h (Eigen, absent
// here) for exactly two members, matrix3::inverse() and matrix3::inverse(matrix3 const&); every other member compiles from the
// reference source untouched.  Defining the include guard skips eigenSupport.h, and eigenView(matrix3) below gives those two
// members a cofactor inverse (what Eigen's fixed-size 3x3 inverse computes).  Neither member is on a pinned path.
#pragma once
#define EIGEN_SUPPORT_H_
#include "stdafx.h"
#include "math/mathclass.h"
#include "math/conversion.h"
#include "motion/Liegroup.h"
struct RefStubM3 {
  double* p;
  RefStubM3& noalias() { return *this; }
  struct Inv { double v[9]; };
  Inv inverse() const {
    const double *a = p; Inv o;
    double c00 = a[4] * a[8] - a[5] * a[7], c01 = a[5] * a[6] - a[3] * a[8], c02 = a[3] * a[7] - a[4] * a[6];
    double det = a[0] * c00 + a[1] * c01 + a[2] * c02, id = 1.0 / det;
    o.v[0] = c00 * id; o.v[1] = (a[2] * a[7] - a[1] * a[8]) * id; o.v[2] = (a[1] * a[5] - a[2] * a[4]) * id;
    o.v[3] = c01 * id; o.v[4] = (a[0] * a[8] - a[2] * a[6]) * id; o.v[5] = (a[2] * a[3] - a[0] * a[5]) * id;
    o.v[6] = c02 * id; o.v[7] = (a[1] * a[6] - a[0] * a[7]) * id; o.v[8] = (a[0] * a[4] - a[1] * a[3]) * id;
    return o;
  }
  void operator=(Inv const& o) { for (int i = 0; i < 9; i++) p[i] = o.v[i]; }
};
inline RefStubM3 eigenView(matrix3 const& m) { return RefStubM3{(double*)&m}; }
