// ref_harness.cpp — TEST INFRASTRUCTURE (oracle/_ref), not product.
//
// Drives the UNMODIFIED reference (oracle/_ref/libgvdb.so + its own cubin) through its own public API,
// the way source/gRenderToFile/main_rendertofile.cpp:17-83 does, on the synthetic scenes of oracle/scenes.h:
//
//   Initialize -> Configure(3,3,3,3,3) -> AddChannel(0,T_FLOAT,1) -> ActivateSpace per brick -> FinishTopology ->
//   UpdateAtlas -> atlas upload (mPool->AtlasCommitFromCPU) -> UpdateApron -> scene params -> AddRenderBuf ->
//   Render(mode,0,0) -> ReadRenderBuf
//
// and dumps, into <outdir>/ :
//   vdbinfo.bin (1232 B)  scninfo_<mode>.bin (416 B)  pool<g>_L<l>.bin  atlas.bin (+ meta.txt)  transfer.bin
//   out_<mode>.rgba       hit_<mode>.f32 (32 B/pixel via the RenderKernel wrappers of oracle_kernels.cu)
//   timing.json           (CPU topology-build seconds, render ms/frame by CUDA-event-free cuCtxSynchronize bracketing)
//
// These dumps are (a) the byte-identical inputs both renderers consume in the parity tests and (b) the golden outputs.
//
// usage: ref_harness <preset> <outdir> [--modes voxel,trilinear,levelset,deep] [--size WxH] [--frames N] [--warmup W]
//                    [--nodump] [--shadow 0|1] [--hits 0|1]
//        ref_harness <preset> <outdir> --bench --orbit F --steps K --warmup W [--mode levelset] [--size WxH]
//          bench mode: a step = F frames on an orbit (camera yaw + 360*j/F); prints one JSON line with the device-synchronised
//          time of K steps for Render() alone and for Render()+ReadRenderBuf() (the reference's end-to-end path).
#include "gvdb.h"
#ifdef GVDBX_SHIM
// ref_harness_x: the same harness with the product's Level-A shim (include/gvdbx_shim.h, the subclass INTEGRATION.md gives to
// a GVDB maintainer) compiled in and libgvdbx.so linked next to the UNMODIFIED libgvdb.so — one process, one CUcontext (the
// reference's own).  --gvdbx renders every requested mode twice, Render() and RenderX(), into the reference's mRenderBuf
// and compares the bytes.  The plain ref_harness never links the product.
#include "gvdbx_shim.h"
#endif
#include <cuda.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>
#include <unordered_map>
#include <sys/stat.h>

#include "scenes.h"

using namespace nvdb;

static double now_s()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
// order-sensitive checksum over 32-bit words (tests/common.py::cksum32 computes the same with numpy): proves that the volume the
// reference built itself (pools, atlas after its own UpdateApron) is byte-identical to the one the product imports, without
// writing gigabytes to disk
static void cksum32(const void* p, size_t bytes, unsigned long long out[2])
{
    const uint32_t* w = (const uint32_t*)p;
    const size_t n = bytes / 4;
    unsigned long long s1 = 0, s2 = 0;
    #pragma omp parallel for reduction(+ : s1, s2)
    for (long long i = 0; i < (long long)n; i++) { s1 += w[i]; s2 += (unsigned long long)w[i] * (unsigned long long)(i % 65521 + 1); }
    out[0] = s1; out[1] = s2;
}
static void dump(const std::string& path, const void* p, size_t n)
{
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) { fprintf(stderr, "cannot write %s\n", path.c_str()); exit(2); }
    fwrite(p, 1, n, f); fclose(f);
}
// kind 3: user kernel of the reference's own sample (source/gRenderKernel/render_custom.cu) through RenderKernel;
// kind 0: native Render(shade) [+ hit wrapper]; kind 1: composed RGBA8 kernel through RenderKernel (BASELINE config 4);
// kind 2: per-sample float colour kernel, `spp` samples averaged on the host (BASELINE config 5)
struct Mode { const char* name; int shade; const char* hitkernel; int kind; };
static const Mode kModes[] = {
    {"voxel", SHADE_VOXEL, "oracleHitVoxel", 0}, {"trilinear", SHADE_TRILINEAR, "oracleHitTrilinear", 0},
    {"levelset", SHADE_LEVELSET, "oracleHitLevelSet", 0}, {"deep", SHADE_VOLUME, "oracleDeepRaw", 0},
    {"tricubic", SHADE_TRICUBIC, "oracleHitTricubic", 0}, {"emptyskip", SHADE_EMPTYSKIP, "oracleHitEmptySkip", 0},
    {"section2d", SHADE_SECTION2D, "", 0}, {"section3d", SHADE_SECTION3D, "", 0},
    {"deepshadow", SHADE_VOLUME, "oracleDeepShadow", 1}, {"deepspp", SHADE_VOLUME, "oracleDeepSample", 2},
    {"voxelid", SHADE_VOXEL, "oracleVoxelId", 0},            // voxel int3 + depth t + leaf id through the instrumented brick function (never a default mode)
    {"custom", SHADE_TRILINEAR, "raycast_kernel", 3},     // the reference's gRenderKernel sample (render_custom.cubin) through RenderKernel
};

int main(int argc, char** argv)
{
    if (argc < 3) { fprintf(stderr, "usage: ref_harness <preset> <outdir> [opts]\n"); return 1; }
    std::string preset = argv[1], outdir = argv[2], modes = "";
    int W = 0, H = 0, frames = 1, warmup = 0, nodump = 0, shadow = -1, hits = 1, bench = 0, orbit = 8, steps = 5, spp = 4;
    std::string bmode = "";
    float xf[12] = {0, 0, 0, 1, 1, 1, 0, 0, 0, 0, 0, 0};
    int have_xf = 0, use_dbuf = 0, nrays = 0;
    std::string savevbx = "";
    int cfg[5] = {3, 3, 3, 3, 3};                  // Configure(q4..q0): log2 dims from the top level down to the brick
    int lightdump = 0, use_color = 0, gvdbx = 0;               // --color: second channel T_UCHAR4 + SetColorChannel (gFluidSurface / gSprayDeposit usage)
    std::string usermodule = "";                   // --module: render the native modes with the kernels of this cubin (RenderKernel)
    for (int i = 3; i < argc; i++) {
        std::string a = argv[i];
        if (a == "--modes" && i + 1 < argc) modes = argv[++i];
        else if (a == "--size" && i + 1 < argc) sscanf(argv[++i], "%dx%d", &W, &H);
        else if (a == "--frames" && i + 1 < argc) frames = atoi(argv[++i]);
        else if (a == "--warmup" && i + 1 < argc) warmup = atoi(argv[++i]);
        else if (a == "--nodump") nodump = 1;
        else if (a == "--lightdump") lightdump = 1;
        else if (a == "--color") use_color = 1;
        else if (a == "--gvdbx") gvdbx = 1;
        else if (a == "--module" && i + 1 < argc) { usermodule = argv[++i]; hits = 0; }        // images + ScnInfo + VDBInfo only (no pools / atlas: large volumes)
        else if (a == "--shadow" && i + 1 < argc) shadow = atoi(argv[++i]);
        else if (a == "--hits" && i + 1 < argc) hits = atoi(argv[++i]);
        else if (a == "--bench") { bench = 1; nodump = 1; hits = 0; }
        else if (a == "--orbit" && i + 1 < argc) orbit = atoi(argv[++i]);
        else if (a == "--steps" && i + 1 < argc) steps = atoi(argv[++i]);
        else if (a == "--mode" && i + 1 < argc) bmode = argv[++i];
        else if (a == "--xform" && i + 1 < argc) { have_xf = 1; sscanf(argv[++i], "%f,%f,%f,%f,%f,%f,%f,%f,%f,%f,%f,%f", xf, xf + 1, xf + 2, xf + 3, xf + 4, xf + 5, xf + 6, xf + 7, xf + 8, xf + 9, xf + 10, xf + 11); }
        else if (a == "--dbuf") use_dbuf = 1;
        else if (a == "--savevbx" && i + 1 < argc) savevbx = argv[++i];
        else if (a == "--config" && i + 1 < argc) sscanf(argv[++i], "%d,%d,%d,%d,%d", cfg, cfg + 1, cfg + 2, cfg + 3, cfg + 4);
        else if (a == "--spp" && i + 1 < argc) spp = atoi(argv[++i]);
        else if (a == "--raytrace" && i + 1 < argc) nrays = atoi(argv[++i]);
    }
    scene_preset P;
    if (scene_get_preset(preset.c_str(), &P)) { fprintf(stderr, "unknown preset %s\n", preset.c_str()); return 1; }
    if (W > 0) { P.width = W; P.height = H; }
    if (shadow == 0) { P.shadow[0] = 0; P.shadow[1] = 0; P.shadow[2] = 0; }
    mkdir(outdir.c_str(), 0777);

    // ---- synthetic volume (CPU, shared generator)
    scene_data S;
    double t0 = now_s();
    scene_generate(&P, &S);
    double t_gen = now_s() - t0;
    fprintf(stderr, "[ref] preset %s: %d bricks (gen %.2fs)\n", P.name, S.nbricks, t_gen);

    // ---- reference library
#ifdef GVDBX_SHIM
    VolumeGVDBX gvdb;
#else
    VolumeGVDB gvdb;
    if (gvdbx) { fprintf(stderr, "--gvdbx needs ref_harness_x (built with the shim)\n"); return 1; }
#endif
    gvdb.SetDebug(false);
    gvdb.SetVerbose(false);
    gvdb.SetCudaDevice(0);          // explicit device: never GVDB_DEV_FIRST (cuGLGetDevices) -> cuCtxCreate'd context
    gvdb.Initialize();
#ifdef GVDBX_SHIM
    if (gvdbx && !gvdb.InitX(0)) { fprintf(stderr, "InitX failed\n"); return 6; }
#endif

    // ---- CPU topology build (timed: the reference's host-side baseline)
    t0 = now_s();
    gvdb.Configure(cfg[0], cfg[1], cfg[2], cfg[3], cfg[4]);
    {   // 16 x 16 x N brick slots like the reference samples; 128 x 128 x N for the large volume (3-D array limit of 16384 along z)
        const int cxy = S.nbricks > 400000 ? 128 : 16;
        gvdb.SetChannelDefault(cxy, cxy, 1);
    }
    gvdb.AddChannel(0, T_FLOAT, 1);
    if (use_color) gvdb.AddChannel(1, T_UCHAR4, 1, F_POINT);    // as gPointFusion does (main_point_fusion.cpp:430); AddChannel's default
                                                                // F_LINEAR is rejected by CUDA for an integer-read texture
    double t_cfg = now_s() - t0;
    t0 = now_s();
    // one ActivateSpace per 8^3 scene brick; with Configure(.., q0 > 3) several scene bricks share one leaf (16^3 / 32^3 voxels)
    for (int n = 0; n < S.nbricks; n++)
        gvdb.ActivateSpace(Vector3DF((float)S.brick_pos[3 * n], (float)S.brick_pos[3 * n + 1], (float)S.brick_pos[3 * n + 2]));
    double t_act = now_s() - t0;
    t0 = now_s();
    gvdb.FinishTopology();
    gvdb.UpdateAtlas();
    cuCtxSynchronize();
    double t_fin = now_s() - t0;
    gvdb.SetEpsilon(P.epsilon, 256);

    // ---- atlas upload: every scene brick goes to its place inside its leaf's atlas brick (voxels no scene brick covers
    // keep the background value: 0 for the densities, +band for the signed distance field)
    Vector3DI ares = gvdb.mPool->getAtlasRes(0);
    size_t atexels = (size_t)ares.x * ares.y * ares.z;
    const int res0 = 1 << cfg[4];
    std::vector<float> atlas(atexels, 0.0f);
    int nleaf = (int)gvdb.mPool->getPoolTotalCnt(0, 0);
    if (res0 == 8 && nleaf != S.nbricks) { fprintf(stderr, "leaf count %d != bricks %d\n", nleaf, S.nbricks); return 3; }
    if (res0 != 8 && P.kind == SCN_KIND_SDF) {
        for (int l = 0; l < nleaf; l++) {
            Node* nd = gvdb.getNode(0, 0, l);
            for (int k = 0; k < res0; k++) for (int j = 0; j < res0; j++)
                std::fill_n(&atlas[((size_t)(nd->mValue.z + k) * ares.y + (nd->mValue.y + j)) * ares.x + nd->mValue.x], res0, 12.0f);
        }
    }
    std::unordered_map<unsigned long long, int> leaf_at;          // leaf min corner -> leaf index
    auto key = [](int x, int y, int z) { return ((unsigned long long)(unsigned)x << 42) ^ ((unsigned long long)(unsigned)y << 21) ^ (unsigned long long)(unsigned)z; };
    for (int l = 0; l < nleaf; l++) { Node* nd = gvdb.getNode(0, 0, l); leaf_at[key(nd->mPos.x, nd->mPos.y, nd->mPos.z)] = l; }
    for (int n = 0; n < S.nbricks; n++) {
        const int bx = S.brick_pos[3 * n], by = S.brick_pos[3 * n + 1], bz = S.brick_pos[3 * n + 2];
        auto it = leaf_at.find(key(bx / res0 * res0, by / res0 * res0, bz / res0 * res0));
        if (it == leaf_at.end()) { fprintf(stderr, "no leaf for brick %d\n", n); return 3; }
        const int lf = it->second;
        Node* nd = gvdb.getNode(0, 0, lf);
        const int ox = S.brick_pos[3 * n] - nd->mPos.x, oy = S.brick_pos[3 * n + 1] - nd->mPos.y, oz = S.brick_pos[3 * n + 2] - nd->mPos.z;
        if (ox < 0 || oy < 0 || oz < 0 || ox + 8 > res0 || oy + 8 > res0 || oz + 8 > res0) { fprintf(stderr, "brick %d outside its leaf\n", n); return 3; }
        if (res0 == 8 && lf != n) { fprintf(stderr, "leaf %d is not brick %d\n", lf, n); return 3; }
        const float* v = S.values + 512 * (size_t)n;
        for (int k = 0; k < 8; k++) for (int j = 0; j < 8; j++) {
            float* dst = &atlas[((size_t)(nd->mValue.z + oz + k) * ares.y + (nd->mValue.y + oy + j)) * ares.x + nd->mValue.x + ox];
            memcpy(dst, v + (k * 8 + j) * 8, 8 * sizeof(float));
        }
    }
    gvdb.mPool->AtlasCommitFromCPU(0, (uchar*)atlas.data());
    gvdb.UpdateApron(0, 0.0f);
    cuCtxSynchronize();
    // --color: a colour per interior voxel from its index-space position, same slot layout as channel 0
    std::vector<unsigned char> color;
    if (use_color) {
        color.assign(atexels * 4, 0);
        for (int n = 0; n < nleaf; n++) {
            Node* nd = gvdb.getNode(0, 0, n);
            for (int k = 0; k < res0; k++) for (int j = 0; j < res0; j++) for (int i = 0; i < res0; i++) {
                const int wx = nd->mPos.x + i, wy = nd->mPos.y + j, wz = nd->mPos.z + k;
                unsigned char* c = &color[4 * (((size_t)(nd->mValue.z + k) * ares.y + (nd->mValue.y + j)) * ares.x + nd->mValue.x + i)];
                c[0] = (unsigned char)(40 + (wx * 37 + wy * 11) % 216); c[1] = (unsigned char)(40 + (wy * 29 + wz * 7) % 216);
                c[2] = (unsigned char)(40 + (wz * 41 + wx * 3) % 216);  c[3] = (unsigned char)(128 + (wx + wy + wz) % 128);
            }
        }
        gvdb.mPool->AtlasCommitFromCPU(1, color.data());
        gvdb.SetColorChannel(1);
        cuCtxSynchronize();
    }

    // ---- scene
    Scene* scn = gvdb.getScene();
    scn->SetSteps(P.steps[0], P.steps[1], P.steps[2]);
    scn->SetExtinct(P.extinct[0], P.extinct[1], P.extinct[2]);
    scn->SetVolumeRange(P.thresh[0], P.thresh[1], P.thresh[2]);
    scn->SetCutoff(P.cutoff[0], P.cutoff[1], P.cutoff[2]);
    scn->SetBackgroundClr(P.backclr[0], P.backclr[1], P.backclr[2], P.backclr[3]);
    scn->SetShadowParams(P.shadow[0], P.shadow[1], P.shadow[2]);
    if (P.transfer == 1) {
        scn->LinearTransferFunc(0.00f, 0.25f, Vector4DF(0, 0, 0, 0), Vector4DF(1, 1, 0, 0.1f));
        scn->LinearTransferFunc(0.25f, 0.50f, Vector4DF(1, 1, 0, 0.4f), Vector4DF(1, 0, 0, 0.3f));
        scn->LinearTransferFunc(0.50f, 0.75f, Vector4DF(1, 0, 0, 0.3f), Vector4DF(.2f, .2f, 0.2f, 0.1f));
        scn->LinearTransferFunc(0.75f, 1.00f, Vector4DF(.2f, .2f, 0.2f, 0.1f), Vector4DF(0, 0, 0, 0.0));
        gvdb.CommitTransferFunc();
    }
    Camera3D* cam = new Camera3D;
    cam->setFov(P.fov);
    cam->setOrbit(Vector3DF(P.cam_angs[0], P.cam_angs[1], P.cam_angs[2]),
                  Vector3DF(P.cam_target[0], P.cam_target[1], P.cam_target[2]), P.cam_dist, 1.0f);
    scn->SetCamera(cam);
    scn->SetRes(P.width, P.height);
    Light* lgt = new Light;
    lgt->setOrbit(Vector3DF(P.light_angs[0], P.light_angs[1], P.light_angs[2]),
                  Vector3DF(P.light_target[0], P.light_target[1], P.light_target[2]), P.light_dist, 1.0f);
    scn->SetLight(0, lgt);

    const int w = P.width, h = P.height;
    gvdb.AddRenderBuf(0, w, h, 4);
    gvdb.AddRenderBuf(1, w, h, 32);
    gvdb.AddRenderBuf(3, w, h, 16);
    // --xform: VolumeGVDB::SetTransform(pretrans, scale, angles, trans)   (gvdb_volume_gvdb.cpp:5770)
    if (have_xf) gvdb.SetTransform(Vector3DF(xf[0], xf[1], xf[2]), Vector3DF(xf[3], xf[4], xf[5]), Vector3DF(xf[6], xf[7], xf[8]), Vector3DF(xf[9], xf[10], xf[11]));
    // --dbuf: a synthetic depth buffer in a plain render buffer (AddDepthBuf itself needs GL): NDC depth of a slanted
    // plane between 0.8 and 1.2 camera distances, bound with Scene::SetDepthBuf (consumed at gvdb_volume_gvdb.cpp:4300)
    if (use_dbuf) {
        gvdb.AddRenderBuf(2, w, h, 4);
        std::vector<float> z((size_t)w * h);
        const double n = cam->getNear(), f = cam->getFar();
        for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) {
            double lin = P.cam_dist * (0.8 + 0.4 * (double)x / w);
            z[(size_t)y * w + x] = (float)(f / (f - n) - (n * f / (f - n)) / lin);
        }
        cuMemcpyHtoD(gvdb.mRenderBuf[2].gpu, z.data(), z.size() * sizeof(float));
        scn->SetDepthBuf(2);
        if (!nodump) dump(outdir + "/dbuf.bin", z.data(), z.size() * sizeof(float));
    }
    // --raytrace N: VolumeGVDB::Raytrace on an explicit ray bundle (gvdb_volume_gvdb.cpp:4384-4408, gSprayDeposit usage)
    if (nrays > 0) {
        DataPtr rays;
        gvdb.AllocData(rays, nrays, sizeof(ScnRay));
        scene_make_rays(&P, nrays, rays.cpu);
        if (!nodump) dump(outdir + "/rays_in.bin", rays.cpu, (size_t)nrays * 64);
        gvdb.CommitData(rays);
        const float bias = -0.0001f;
        gvdb.Raytrace(rays, 0, SHADE_TRILINEAR, 0, bias);
        cuCtxSynchronize();
        gvdb.RetrieveData(rays);
        if (!nodump) {
            dump(outdir + "/rays_out.bin", rays.cpu, (size_t)nrays * 64);
            dump(outdir + "/scninfo_raytrace.bin", gvdb.getScnInfo(), 416);
        }
    }


    // --savevbx: the reference's own VBX writer (gvdb_volume_gvdb.cpp:1626-1767) on the finished volume (transform included)
    if (!savevbx.empty()) gvdb.SaveVBX(savevbx);

    // ---- bench mode (bench.py --impl reference): the reference's own CUDA render through its public API
    if (bench) {
        int shade = P.shade, kind = 0;
        const char* kname = "";
        for (const Mode& m : kModes) if (bmode == m.name) { shade = m.shade; kind = m.kind; kname = m.hitkernel; }
        CUmodule bmod = 0; CUfunction bfn = 0;
        if (kind != 0) {
            if (cuModuleLoad(&bmod, "oracle_kernels.cubin") != CUDA_SUCCESS || cuModuleGetFunction(&bfn, bmod, kname) != CUDA_SUCCESS) {
                fprintf(stderr, "cannot load %s from oracle_kernels.cubin\n", kname); return 4;
            }
            gvdb.SetModule(bmod);
            scn->SetShading(shade);
        }
#ifdef GVDBX_SHIM
        if (gvdbx) {
            if (kind != 0) gvdb.SetModule();
            gvdbx_set_option(gvdb.x(), GVDBX_OPT_DEEP_SHADOW, kind == 1);
            gvdbx_set_option(gvdb.x(), GVDBX_OPT_SPP, kind == 2 ? spp : 1);
        }
#endif
        std::vector<unsigned char> frame((size_t)w * h * 4);
        auto set_cam = [&](int j) {
            cam->setOrbit(Vector3DF(P.cam_angs[0] + 360.0f * (float)j / (float)orbit, P.cam_angs[1], P.cam_angs[2]),
                          Vector3DF(P.cam_target[0], P.cam_target[1], P.cam_target[2]), P.cam_dist, 1.0f);
        };
        // kind 1: the composed kernel writes RGBA8 into buffer 0; kind 2: `spp` per-sample launches into the float buffer
        // (the host-side average is not part of the timed region: it favours the reference)
        auto render = [&]() {
#ifdef GVDBX_SHIM
            if (gvdbx) { gvdb.RenderX(shade, 0, 0); return; }       // kind 1 / 2: GVDBX_OPT_DEEP_SHADOW / GVDBX_OPT_SPP set below
#endif
            if (kind == 0) gvdb.Render(shade, 0, 0);
            else if (kind == 1) gvdb.RenderKernel(bfn, 0, 0);
            else for (int sidx = 0; sidx < spp; sidx++) { scn->SetSample(sidx); scn->SetFrame(spp); gvdb.RenderKernel(bfn, 0, 3); }
        };
        for (int s = 0; s < warmup; s++) for (int j = 0; j < orbit; j++) { set_cam(j); render(); }
        cuCtxSynchronize();
        double tk0 = now_s();
        for (int s = 0; s < steps; s++) { for (int j = 0; j < orbit; j++) { set_cam(j); render(); } cuCtxSynchronize(); }
        double t_kernel = now_s() - tk0;
        double te0 = now_s();
        for (int s = 0; s < steps; s++) for (int j = 0; j < orbit; j++) { set_cam(j); render(); gvdb.ReadRenderBuf(0, frame.data()); }
        double t_e2e = now_s() - te0;
        unsigned long long sum = 0;
        for (size_t i = 0; i < frame.size(); i += 97) sum += frame[i];
        printf("{\"preset\":\"%s\",\"bricks\":%d,\"width\":%d,\"height\":%d,\"shade\":%d,\"orbit\":%d,\"steps\":%d,\"warmup\":%d,"
               "\"render_s\":%.6f,\"e2e_s\":%.6f,\"topology_build_s\":%.6f,\"scene_gen_s\":%.3f,\"checksum\":%llu,\"gvdbx\":%d}\n",
               preset.c_str(), S.nbricks, w, h, shade, orbit, steps, warmup, t_kernel, t_e2e, t_cfg + t_act + t_fin, t_gen, sum, gvdbx);
        scene_free(&S);
        return 0;
    }

    // ---- static dumps
    if (!nodump && lightdump) {
        gvdb.PrepareVDB();
        dump(outdir + "/vdbinfo.bin", gvdb.getVDBInfo(), 1232);
        FILE* mf = fopen((outdir + "/meta.txt").c_str(), "w");
        fprintf(mf, "preset %s\nbricks %d\nlevels %d\natlas_res %d %d %d\nwidth %d\nheight %d\n", P.name, S.nbricks, gvdb.mPool->getNumLevels(), ares.x, ares.y, ares.z, w, h);
        // checksums instead of the bytes: pools as the reference holds them, atlas after the reference's own UpdateApron
        gvdb.FetchPoolCPU();
        unsigned long long ck[2];
        for (int g = 0; g < 2; g++) for (int l = 0; l < gvdb.mPool->getNumLevels(); l++) {
            uint64 cnt = gvdb.mPool->getPoolTotalCnt(g, l), wid = gvdb.mPool->getPoolWidth(g, l);
            cksum32(cnt * wid ? gvdb.mPool->getPoolCPU(g, l) : "", cnt * wid, ck);
            fprintf(mf, "cksum_pool %d %d %llu %llu %llu\n", g, l, (unsigned long long)(cnt * wid), ck[0], ck[1]);
        }
        {
            std::vector<float> back(atexels);
            for (int z = 0; z < ares.z; z++)
                gvdb.mPool->AtlasRetrieveSlice(0, z, 0, 0, (uchar*)(back.data() + (size_t)z * ares.x * ares.y));
            cksum32(back.data(), atexels * sizeof(float), ck);
            fprintf(mf, "cksum_atlas %llu %llu %llu\n", (unsigned long long)(atexels * sizeof(float)), ck[0], ck[1]);
        }
        fclose(mf);
        dump(outdir + "/transfer.bin", scn->getTransferFunc(), 16384 * 16);
    }
    if (!nodump && !lightdump) {
        gvdb.PrepareVDB();
        dump(outdir + "/vdbinfo.bin", gvdb.getVDBInfo(), 1232);
        gvdb.FetchPoolCPU();
        int levs = gvdb.mPool->getNumLevels();
        FILE* mf = fopen((outdir + "/meta.txt").c_str(), "w");
        fprintf(mf, "preset %s\nbricks %d\nlevels %d\natlas_res %d %d %d\nwidth %d\nheight %d\n", P.name, S.nbricks, levs, ares.x, ares.y, ares.z, w, h);
        for (int g = 0; g < 2; g++) for (int l = 0; l < levs; l++) {
            uint64 cnt = gvdb.mPool->getPoolTotalCnt(g, l), wid = gvdb.mPool->getPoolWidth(g, l);
            fprintf(mf, "pool %d %d %llu %llu\n", g, l, (unsigned long long)cnt, (unsigned long long)wid);
            char nm[64]; snprintf(nm, sizeof nm, "/pool%d_L%d.bin", g, l);
            dump(outdir + nm, cnt * wid ? gvdb.mPool->getPoolCPU(g, l) : "", cnt * wid);
        }
        fclose(mf);
        // atlas after the reference's own UpdateApron
        std::vector<float> back(atexels);
        for (int z = 0; z < ares.z; z++)
            gvdb.mPool->AtlasRetrieveSlice(0, z, 0, 0, (uchar*)(back.data() + (size_t)z * ares.x * ares.y));
        dump(outdir + "/atlas.bin", back.data(), atexels * sizeof(float));
        dump(outdir + "/transfer.bin", scn->getTransferFunc(), 16384 * 16);
        if (use_color) dump(outdir + "/color.bin", color.data(), color.size());
        dump(outdir + "/brick_pos.bin", S.brick_pos, sizeof(int32_t) * 3 * (size_t)S.nbricks);
    }

    // ---- oracle wrapper kernels via the RenderKernel plugin API
    CUmodule omod = 0;
    if (cuModuleLoad(&omod, "oracle_kernels.cubin") != CUDA_SUCCESS) { fprintf(stderr, "cannot load oracle_kernels.cubin\n"); hits = 0; omod = 0; }

    std::string timing = "{\"preset\":\"" + preset + "\",\"bricks\":" + std::to_string(S.nbricks) +
        ",\"width\":" + std::to_string(w) + ",\"height\":" + std::to_string(h) +
        ",\"topology_build_s\":{\"configure\":" + std::to_string(t_cfg) + ",\"activate\":" + std::to_string(t_act) +
        ",\"finish_update_atlas\":" + std::to_string(t_fin) + ",\"threads\":1},\"scene_gen_s\":" + std::to_string(t_gen) + ",\"render\":{";
    bool first = true;
    std::vector<unsigned char> img((size_t)w * h * 4);
    std::vector<float> hitbuf((size_t)w * h * 8);
    const float cN = 0.5f * (float)P.N;
    std::vector<float> sample((size_t)w * h * 4), accum((size_t)w * h * 4);
    for (const Mode& m : kModes) {
        bool want = modes.empty() ? (m.shade == P.shade && m.kind == 0 && strcmp(m.name, "voxelid") != 0)
                                  : (("," + modes + ",").find(std::string(",") + m.name + ",") != std::string::npos);
        if (!want) continue;
        // section plane: 2D = centre + (u,0,v) * half extent (slice_norm is a per-axis scale there); 3D = tilted plane
        if (m.shade == SHADE_SECTION2D) scn->SetCrossSection(Vector3DF(cN, cN * 0.9f, cN), Vector3DF(cN * 0.8f, 1.0f, cN * 0.8f));
        if (m.shade == SHADE_SECTION3D) scn->SetCrossSection(Vector3DF(cN, cN, cN * 0.85f), Vector3DF(0.3f, 0.2f, 1.0f));
        CUfunction kfn = 0;
        static CUmodule umod = 0;
        if (m.kind == 0 && !usermodule.empty()) {
            // module-level drop-in: same kernel NAME as the native one, taken from the user's cubin, launched through the
            // reference's own plugin seam (SetModule + RenderKernel, 8 x 8 CTAs)
            static const char* names[][2] = {{"voxel", "gvdbRaySurfaceVoxel"}, {"trilinear", "gvdbRaySurfaceTrilinear"}, {"tricubic", "gvdbRaySurfaceTricubic"},
                                             {"levelset", "gvdbRayLevelSet"}, {"deep", "gvdbRayDeep"}, {"emptyskip", "gvdbRayEmptySkip"},
                                             {"section2d", "gvdbSection2D"}, {"section3d", "gvdbSection3D"}};
            const char* kname = "";
            for (auto& n : names) if (std::string(n[0]) == m.name) kname = n[1];
            if (!umod && cuModuleLoad(&umod, usermodule.c_str()) != CUDA_SUCCESS) { fprintf(stderr, "cannot load %s\n", usermodule.c_str()); return 5; }
            if (cuModuleGetFunction(&kfn, umod, kname) != CUDA_SUCCESS) { fprintf(stderr, "no kernel %s in %s\n", kname, usermodule.c_str()); return 5; }
            gvdb.SetModule(umod);
            scn->SetShading(m.shade);
        }
        if (m.kind != 0) {
            CUmodule use = omod;
            if (m.kind == 3) {
                static CUmodule cmod = 0;
                if (!cmod && cuModuleLoad(&cmod, "render_custom.cubin") != CUDA_SUCCESS) { fprintf(stderr, "cannot load render_custom.cubin\n"); continue; }
                use = cmod;
            }
            if (!use || cuModuleGetFunction(&kfn, use, m.hitkernel) != CUDA_SUCCESS) { fprintf(stderr, "no kernel %s\n", m.hitkernel); continue; }
            gvdb.SetModule(use);
            scn->SetShading(m.shade);
        }
        auto render = [&]() {
            if (m.kind == 0 && kfn) gvdb.RenderKernel(kfn, 0, 0);
            else if (m.kind == 0) gvdb.Render(m.shade, 0, 0);
            else if (m.kind == 1 || m.kind == 3) gvdb.RenderKernel(kfn, 0, 0);
            else for (int sidx = 0; sidx < spp; sidx++) { scn->SetSample(sidx); scn->SetFrame(spp); gvdb.RenderKernel(kfn, 0, 3); }
        };
        for (int i = 0; i < warmup; i++) render();
        cuCtxSynchronize();
        std::vector<double> ms;
        for (int i = 0; i < frames; i++) {
            double a = now_s();
            render();
            cuCtxSynchronize();
            ms.push_back((now_s() - a) * 1e3);
        }
        double a = now_s();
        if (m.kind == 2) {
            // sample colours summed in sample order, scaled by 1/spp, packed like make_uchar4(clr*255)
            std::fill(accum.begin(), accum.end(), 0.0f);
            for (int sidx = 0; sidx < spp; sidx++) {
                scn->SetSample(sidx); scn->SetFrame(spp);
                gvdb.RenderKernel(kfn, 0, 3);
                cuCtxSynchronize();
                gvdb.ReadRenderBuf(3, (unsigned char*)sample.data());
                for (size_t i = 0; i < accum.size(); i++) accum[i] = accum[i] + sample[i];
            }
            const float inv = 1.0f / (float)spp;
            for (size_t i = 0; i < accum.size(); i++) { volatile float c = accum[i] * inv; volatile float q = c * 255.0f; img[i] = (unsigned char)(int)q; }
            scn->SetSample(0); scn->SetFrame(0);
        } else {
            gvdb.ReadRenderBuf(0, img.data());
        }
        double read_ms = (now_s() - a) * 1e3;
        std::sort(ms.begin(), ms.end());
        double med = ms[ms.size() / 2], mn = ms[0];
        if (!nodump) {
            dump(outdir + "/out_" + m.name + ".rgba", img.data(), img.size());
            dump(outdir + "/scninfo_" + m.name + ".bin", gvdb.getScnInfo(), 416);
        }
        if (m.kind != 0 || kfn) gvdb.SetModule();
        if (hits && m.kind == 0 && m.hitkernel[0]) {
            CUfunction fn;
            if (cuModuleGetFunction(&fn, omod, m.hitkernel) == CUDA_SUCCESS) {
                gvdb.SetModule(omod);               // scn symbol of the wrapper module
                scn->SetShading(m.shade);
                gvdb.RenderKernel(fn, 0, 1);
                cuCtxSynchronize();
                gvdb.ReadRenderBuf(1, (unsigned char*)hitbuf.data());
                gvdb.SetModule();                   // back to the native module
                if (!nodump) dump(outdir + "/hit_" + m.name + ".f32", hitbuf.data(), hitbuf.size() * sizeof(float));
            }
        }
        long xdiff = -1, xnonbg = 0;
        double xms = 0;
#ifdef GVDBX_SHIM
        if (gvdbx && (m.kind == 0 || m.kind == 1 || m.kind == 2) && usermodule.empty()) {
            // the same frame through the shim: RenderX() into the reference's own render buffer 0, read back by the
            // reference's own ReadRenderBuf
            gvdbx_set_option(gvdb.x(), GVDBX_OPT_DEEP_SHADOW, m.kind == 1);
            gvdbx_set_option(gvdb.x(), GVDBX_OPT_SPP, m.kind == 2 ? spp : 1);
            std::vector<unsigned char> imgx((size_t)w * h * 4, 0x5A);
            cuMemsetD8(gvdb.mRenderBuf[0].gpu, 0x5A, (size_t)w * h * 4);
            gvdb.RenderX(m.shade, 0, 0);
            cuCtxSynchronize();
            std::vector<double> xs;
            for (int i = 0; i < std::max(frames, 1); i++) { double a0 = now_s(); gvdb.RenderX(m.shade, 0, 0); cuCtxSynchronize(); xs.push_back((now_s() - a0) * 1e3); }
            std::sort(xs.begin(), xs.end());
            xms = xs[xs.size() / 2];
            gvdb.ReadRenderBuf(0, imgx.data());
            xdiff = 0;
            for (size_t i = 0; i < imgx.size(); i += 4) {
                if (memcmp(&imgx[i], &img[i], 4)) xdiff++;
                if (memcmp(&img[i], &img[0], 4)) xnonbg++;
            }
            // and as bands with the overlapped read-back (RenderX(..., bands) + ReadRenderBufX), then once more through the
            // reference's own ReadRenderBuf behind a banded RenderX: any difference counts as a mismatch of the shim
            for (int pass = 0; pass < 2; pass++) {
                std::vector<unsigned char> imgb((size_t)w * h * 4, 0xA5);
                cuMemsetD8(gvdb.mRenderBuf[0].gpu, 0xA5, (size_t)w * h * 4);
                gvdb.RenderX(m.shade, 0, 0, pass == 0 ? 3 : 5);
                if (pass == 0) gvdb.ReadRenderBufX(0, imgb.data()); else gvdb.ReadRenderBuf(0, imgb.data());
                for (size_t i = 0; i < imgb.size(); i += 4) if (memcmp(&imgb[i], &img[i], 4)) xdiff++;
            }
            if (!nodump) dump(outdir + "/outx_" + m.name + ".rgba", imgx.data(), imgx.size());
            fprintf(stderr, "[ref] %s: RenderX %.3f ms/frame, %ld of %d pixels differ from Render() (plain + banded + banded read by the reference)\n", m.name, xms, xdiff, w * h);
        }
#endif
        char buf[384];
        snprintf(buf, sizeof buf, "%s\"%s\":{\"ms_median\":%.4f,\"ms_min\":%.4f,\"read_ms\":%.4f,\"frames\":%d,\"x_mismatch\":%ld,\"x_nonbg\":%ld,\"x_ms_median\":%.4f}",
                 first ? "" : ",", m.name, med, mn, read_ms, frames, xdiff, xnonbg, xms);
        timing += buf; first = false;
        fprintf(stderr, "[ref] %s: %.3f ms/frame (min %.3f), %.2f Mrays/s\n", m.name, med, mn, w * (double)h / med * 1e-3);
    }
    timing += "}}";
    dump(outdir + "/timing.json", timing.data(), timing.size());
    printf("%s\n", timing.c_str());
    scene_free(&S);
    return 0;
}
