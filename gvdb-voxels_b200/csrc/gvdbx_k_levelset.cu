// gx_render_kernel instantiations for one shade mode (see gvdbx_pick.cuh)
#include "gvdbx_pick.cuh"
GX_DEFINE_PICK(levelset, GX_MODE_LEVELSET, true, true)
