// gx_render_kernel instantiations for one shade mode (see gvdbx_pick.cuh)
#include "gvdbx_pick.cuh"
GX_DEFINE_PICK(deepshadow, GX_MODE_DEEPSHADOW, true, false)
