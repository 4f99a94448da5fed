// Pre-include for BaseLib/math/Filter.cpp: its one Eigen user (Filter::deconvolution, two expressions) is not on any pinned path, and
// Eigen is absent from this image.  Skip PhysicsLib/TRL/eigenSupport.h by its include guard and give `eigenTView` a do-nothing type
// so that the file compiles; GetGaussFilter / LTIFilter / gaussFilter (what the harness calls) are untouched reference code.
#pragma once
#define EIGEN_SUPPORT_H_
class matrixn;
struct NoEigenExpr {
  NoEigenExpr& noalias() { return *this; }
  NoEigenExpr array() const { return *this; }
  NoEigenExpr matrix() const { return *this; }
  NoEigenExpr operator/(NoEigenExpr const&) const { return *this; }
  NoEigenExpr operator*(NoEigenExpr const&) const { return *this; }
};
inline NoEigenExpr eigenTView(matrixn const&) { return NoEigenExpr(); }
