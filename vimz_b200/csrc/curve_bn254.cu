// curve_bn254.cu -- instantiates every kernel for CurveBn254 (one translation unit per curve).
#include "curve_impl.cuh"
namespace vimz {
const CurveVTable* vtable_bn254() {
  static const CurveVTable t = make_vtable<CurveBn254>("bn254");
  return &t;
}
}  // namespace vimz
