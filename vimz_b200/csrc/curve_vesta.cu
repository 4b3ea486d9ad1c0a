// curve_vesta.cu -- instantiates every kernel for CurveVesta (one translation unit per curve).
#include "curve_impl.cuh"
namespace vimz {
const CurveVTable* vtable_vesta() {
  static const CurveVTable t = make_vtable<CurveVesta>("vesta");
  return &t;
}
}  // namespace vimz
