// ref_hostdump.cpp — TEST INFRASTRUCTURE (oracle/_ref).  CPU-only: compiles the reference's own Camera3D and Matrix4F
// (src/gvdb_camera.cpp, src/gvdb_vec.cpp, unmodified, in place) and dumps the host state the render path consumes,
// as uint32 bit patterns, for the parameter sets given on stdin.  Used by tests/make_golden_hoststate.py to pin the
// product's host mirror (gvdb-voxels_b200/csrc/gvdbx_host.cpp).
//
//   cam  fov w h  ax ay az  tx ty tz  dist      -> from_pos tl tr bl        (12 floats; Camera3D::setFov/setAspect/setOrbit)
//   xfm  px py pz  sx sy sz  ax ay az  tx ty tz -> xform invxform invxrot   (48 floats; the Matrix4F call sequence of
//                                                                           VolumeGVDB::SetTransform, gvdb_volume_gvdb.cpp:5770-5794)
#include "gvdb_camera.h"
#include "gvdb_vec.h"
#include <cstdio>
#include <cstring>
#include <cstdint>
using namespace nvdb;

static void put(const float* f, int n) { for (int i = 0; i < n; i++) { uint32_t u; memcpy(&u, f + i, 4); printf("%s%08x", i ? " " : "", u); } printf("\n"); }

int main()
{
    char tag[16];
    while (scanf("%15s", tag) == 1) {
        if (!strcmp(tag, "cam")) {
            float fov, w, h, ax, ay, az, tx, ty, tz, dist;
            if (scanf("%f %f %f %f %f %f %f %f %f %f", &fov, &w, &h, &ax, &ay, &az, &tx, &ty, &tz, &dist) != 10) return 1;
            Camera3D cam;
            cam.setFov(fov);
            cam.setOrbit(Vector3DF(ax, ay, az), Vector3DF(tx, ty, tz), dist, 1.0f);
            cam.setAspect(w / h);
            float o[12] = { cam.origRayWorld.x, cam.origRayWorld.y, cam.origRayWorld.z, cam.tlRayWorld.x, cam.tlRayWorld.y, cam.tlRayWorld.z,
                            cam.trRayWorld.x, cam.trRayWorld.y, cam.trRayWorld.z, cam.blRayWorld.x, cam.blRayWorld.y, cam.blRayWorld.z };
            put(o, 12);
        } else if (!strcmp(tag, "xfm")) {
            float v[12];
            for (int i = 0; i < 12; i++) if (scanf("%f", &v[i]) != 1) return 1;
            Vector3DF pre(v[0], v[1], v[2]), scal(v[3], v[4], v[5]), angs(v[6], v[7], v[8]), trans(v[9], v[10], v[11]);
            Matrix4F xrot, xform, invxform, invxrot;
            xrot.RotateZYX(angs);
            invxrot.Identity();
            invxrot.InvScaleInPlace(scal);
            invxrot.InvLeftMultiplyInPlace(xrot);
            xform.Identity();
            xform.RotateTZYXS(angs, trans, scal);
            xform.PreTranslate(pre);
            invxform = xform;
            invxform.InvertTRS();
            float o[48];
            memcpy(o, xform.GetDataF(), 64); memcpy(o + 16, invxform.GetDataF(), 64); memcpy(o + 32, invxrot.GetDataF(), 64);
            put(o, 48);
        }
    }
    return 0;
}
