// gx_render_kernel instantiations for one shade mode (see gvdbx_pick.cuh)
#include "gvdbx_pick.cuh"
GX_DEFINE_PICK(voxel, GX_MODE_VOXEL, true, true)
