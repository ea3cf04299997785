// gx_render_kernel instantiations for the texture-only shade modes (see gvdbx_pick.cuh)
#include "gvdbx_pick.cuh"
GX_DEFINE_PICK(tricubic, GX_MODE_TRICUBIC, false, false)
GX_DEFINE_PICK(emptyskip, GX_MODE_EMPTYSKIP, false, false)
GX_DEFINE_PICK(section2d, GX_MODE_SECTION2D, false, false)
GX_DEFINE_PICK(section3d, GX_MODE_SECTION3D, false, false)
