// curve_pallas.cu -- instantiates every kernel for CurvePallas (one translation unit per curve).
#include "curve_impl.cuh"
namespace vimz {
const CurveVTable* vtable_pallas() {
  static const CurveVTable t = make_vtable<CurvePallas>("pallas");
  return &t;
}
}  // namespace vimz
