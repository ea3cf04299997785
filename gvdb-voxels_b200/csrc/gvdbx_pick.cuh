// gvdbx_pick.cuh — instantiation table of gx_render_kernel<MODE, SAMPLER, FLAGS, UNI>, one translation unit per shade
// mode (gvdbx_k_*.cu) so that the variants compile in parallel.  gvdbx_api.cu only sees the gx_pick_* functions.
#pragma once
#include "gvdbx_extra.cuh"

typedef void (*gx_kernel_t)(const GxParams);

// AB = also build the two A/B traversal variants (reference-shaped loops, packet traversal): core modes only
// the brick-queue variants (GX_FLAG_QUEUE) exist for the deep modes and for trilinear / level set
template <int MODE, int SAMPLER, bool UNI, bool AB>
static gx_kernel_t gx_pick_flags(int flags)
{
    if constexpr (MODE == GX_MODE_DEEP || MODE == GX_MODE_DEEPSHADOW || MODE == GX_MODE_TRILINEAR || MODE == GX_MODE_LEVELSET) {
        switch (flags) {
        case GX_FLAG_QUEUE: return gx_render_kernel<MODE, SAMPLER, GX_FLAG_QUEUE, UNI>;
        case GX_FLAG_QUEUE | GX_FLAG_TILES: return gx_render_kernel<MODE, SAMPLER, GX_FLAG_QUEUE | GX_FLAG_TILES, UNI>;
        case GX_FLAG_QUEUE | GX_FLAG_SPP: return gx_render_kernel<MODE, SAMPLER, GX_FLAG_QUEUE | GX_FLAG_SPP, UNI>;
        case GX_FLAG_QUEUE | GX_FLAG_TILES | GX_FLAG_SPP: return gx_render_kernel<MODE, SAMPLER, GX_FLAG_QUEUE | GX_FLAG_TILES | GX_FLAG_SPP, UNI>;
        }
    }
    switch (flags) {
    case 0: return gx_render_kernel<MODE, SAMPLER, 0, UNI>;
    case GX_FLAG_DEBUG | GX_FLAG_COUNT: return gx_render_kernel<MODE, SAMPLER, GX_FLAG_DEBUG | GX_FLAG_COUNT, UNI>;
    case GX_FLAG_TILES: return gx_render_kernel<MODE, SAMPLER, GX_FLAG_TILES, UNI>;
    case GX_FLAG_SPP: return gx_render_kernel<MODE, SAMPLER, GX_FLAG_SPP, UNI>;
    case GX_FLAG_TILES | GX_FLAG_SPP: return gx_render_kernel<MODE, SAMPLER, GX_FLAG_TILES | GX_FLAG_SPP, UNI>;
    }
    if constexpr (AB) {
        if (flags == GX_FLAG_LITERAL) return gx_render_kernel<MODE, SAMPLER, GX_FLAG_LITERAL, false>;   // A/B variants: generic tree only
        if (flags == GX_FLAG_PACKET) return gx_render_kernel<MODE, SAMPLER, GX_FLAG_PACKET, false>;
    }
    return nullptr;
}

// LINEAR = the brick-major linear-load sampler is available for this mode
#define GX_DEFINE_PICK(NAME, MODE, LINEAR, AB)                                                                          \
    gx_kernel_t gx_pick_##NAME(int sampler, int flags, bool uni)                                                        \
    {                                                                                                                   \
        if (sampler == GX_SAMPLER_TEX)                                                                                  \
            return uni ? gx_pick_flags<MODE, GX_SAMPLER_TEX, true, AB>(flags) : gx_pick_flags<MODE, GX_SAMPLER_TEX, false, AB>(flags); \
        if constexpr (LINEAR)                                                                                           \
            return uni ? gx_pick_flags<MODE, GX_SAMPLER_LINEAR, true, AB>(flags) : gx_pick_flags<MODE, GX_SAMPLER_LINEAR, false, AB>(flags); \
        return nullptr;                                                                                                 \
    }
