// curve_grumpkin.cu -- instantiates every kernel for CurveGrumpkin (one translation unit per curve).
#include "curve_impl.cuh"
namespace vimz {
const CurveVTable* vtable_grumpkin() {
  static const CurveVTable t = make_vtable<CurveGrumpkin>("grumpkin");
  return &t;
}
}  // namespace vimz
