// gvdbx_host.h — host-side mirror of the reference's render-facing interface, on top of the C ABI (include/gvdbx.h).
//
// Same names, argument meaning and state flow as the reference classes a Render() caller touches:
//   Camera3D / Light   src/gvdb_camera.{h,cpp}   (setFov, setAspect, setNearFar, setOrbit -> corner rays)
//   Scene              src/gvdb_scene.{h,cpp}    (SetSteps, SetExtinct, SetVolumeRange, SetCutoff, SetBackgroundClr,
//                                                 SetShadowParams, LinearTransferFunc, SetCamera, SetLight, SetRes)
//   VolumeGVDB         src/gvdb_volume_gvdb.cpp  (SetTransform :5770, CommitTransferFunc :4892, AddRenderBuf :4164,
//                                                 ResizeRenderBuf :4211, PrepareRender :4254, Render :4336,
//                                                 ReadRenderBuf :4241, SetEpsilon gvdb_volume_gvdb.h:334)
// The arithmetic that produces the 416 bytes of ScnInfo is restated operation by operation (float vs double exactly as
// the reference writes it) so that the bytes are identical; tests/test_host_mirror.py pins this against dumps of the
// reference's own Camera3D / Matrix4F (oracle/_ref/ref_hostdump, tests/golden/hoststate_*.bin).
#pragma once
#include <stdint.h>
#include <string>
#include <vector>
#include "gvdbx_types.h"
#include "../../include/gvdbx.h"

namespace gvdbx {

struct Vec3 { float x = 0, y = 0, z = 0; Vec3() {} Vec3(float a, float b, float c) : x(a), y(b), z(c) {} };
struct Vec4 { float x = 0, y = 0, z = 0, w = 0; Vec4() {} Vec4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {} };

// column-major 4x4, element (row, col) at data[4*col + row]   (src/gvdb_vec.h:537-543)
struct Matrix4 {
    float data[16];
    Matrix4() { Identity(); }
    Matrix4& Identity();
    Matrix4& Zero();
    Matrix4& RotateZYX(const Vec3& angs_deg);                                // src/gvdb_vec.cpp:218-250
    Matrix4& RotateTZYXS(const Vec3& angs_deg, const Vec3& t, const Vec3& s);// :257-281
    Matrix4& PreTranslate(const Vec3& t);                                    // :487-498
    Matrix4& MulAssign(const Matrix4& op);                                   // operator*= :164-193
    Matrix4& LeftMultiplyInPlace(const Matrix4& m);                          // :569-583
    Matrix4& ScaleInPlace(const Vec3& s);                                    // :585-598
    Matrix4& InvertTRS();                                                    // :427-465 (general inverse in double)
    Matrix4& Basis(const Vec3& c1, const Vec3& c2, const Vec3& c3);          // :375-382
};

class Camera3D {
public:
    Camera3D();                                                              // src/gvdb_camera.cpp:37-56
    void setFov(float fov) { mFov = fov; updateMatricies(); }
    void setAspect(float asp) { mAspect = asp; updateMatricies(); }
    void setNearFar(float n, float f) { mNear = n; mFar = f; updateMatricies(); }
    void setOrbit(Vec3 angs, Vec3 tp, float dist, float dolly) { setOrbit(angs.x, angs.y, angs.z, tp, dist, dolly); }
    void setOrbit(float ax, float ay, float az, Vec3 tp, float dist, float dolly);   // :90-105
    void updateMatricies();                                                  // :175-228 (+ updateFrustum corner rays :343-346)
    Vec3 inverseRayProj(float x, float y, float z) const;                    // :382-393
    float getNear() const { return mNear; }
    float getFar() const { return mFar; }
    Vec3& getPos() { return from_pos; }

    Vec3 from_pos, to_pos, up_dir, ang_euler;
    Vec3 dir_vec, side_vec, up_vec;
    Vec3 origRayWorld, tlRayWorld, trRayWorld, blRayWorld, brRayWorld;
    Matrix4 rotate_matrix, view_matrix, proj_matrix, invviewproj_matrix;
private:
    float mFov, mAspect, mNear, mFar, mOrbitDist, mDolly;
};
typedef Camera3D Light;                                                      // src/gvdb_camera.h:176

class Scene {
public:
    Scene();                                                                 // src/gvdb_scene.cpp:20-54
    ~Scene();
    Camera3D* SetCamera(Camera3D* cam);                                      // takes ownership, like the reference
    Light*    SetLight(int n, Light* light);
    Camera3D* getCamera() { return mCamera; }
    Light*    getLight() { return mLight; }
    void SetRes(int x, int y);                                               // gvdb_scene.h:127 (also camera aspect)
    void SetSteps(float direct, float shadow, float fine) { mSteps = Vec3(direct, shadow, fine); }
    void SetExtinct(float a, float b, float c) { mExtinct = Vec3(a, b, c); }
    void SetVolumeRange(float viso, float vmin, float vmax) { mVThreshold = Vec3(viso, vmin, vmax); }
    void SetCutoff(float a, float b, float c) { mCutoff = Vec3(a, b, c); }
    void SetBackgroundClr(float r, float g, float b, float a) { mBackgroundClr = Vec4(r, g, b, a); }
    void SetShadowParams(float x, float y, float z) { mShadowParams = Vec3(x, y, z); }
    void SetCrossSection(Vec3 pos, Vec3 norm) { mSectionPnt = pos; mSectionNorm = norm; }
    void LinearTransferFunc(float t0, float t1, Vec4 a, Vec4 b);             // src/gvdb_scene.cpp:66-88
    const float* getTransferFunc() const { return mTransferFunc; }

    Vec3 mSteps, mExtinct, mVThreshold, mCutoff, mShadowParams, mSectionPnt, mSectionNorm;
    Vec4 mBackgroundClr;
    int  mXres = 0, mYres = 0, mFrame = 0, mSample = 0, mFilterMode = 0, mDepthBuf = 255;
private:
    Camera3D* mCamera = nullptr;
    Light*    mLight = nullptr;
    float*    mTransferFunc = nullptr;   // 16384 x float4
};

// Render-facing subset of VolumeGVDB.  The volume itself (node pools + brick atlas, reference layouts) is handed
// over with ImportTopology* / ImportAtlas*; everything downstream of that is the reference's call sequence.
class VolumeGVDB {
public:
    VolumeGVDB();
    ~VolumeGVDB();
    int  SetCudaDevice(int devid, void* stream = nullptr);                   // gvdb_volume_gvdb.cpp:240 (no kernels to load here)
    int  Initialize();                                                       // :2320-2361 default scene / LUT
    Scene* getScene() { return mScene; }
    void SetTransform(Vec3 pretrans, Vec3 scal, Vec3 angs, Vec3 trans);      // :5770-5794
    int  ImportTopologyHost(const void* vdbinfo, const void* const* pool0, const void* const* pool1, const uint64_t* pool1_bytes);
    int  ImportTopologyDevice(const void* vdbinfo);
    int  ImportAtlasHost(int chan, const float* texels, int rx, int ry, int rz);
    int  ImportAtlasArray(int chan, void* cuarray, int rx, int ry, int rz);
    // VBX files (GVDB_FILESPEC.txt; LoadVBX gvdb_volume_gvdb.cpp:507-683, SaveVBX :1626-1767).  LoadVBX parses the file on
    // the host (transform, per-level pools, channel-0 float atlas), derives the VDBInfo block the way FinishTopology /
    // ComputeBounds / PrepareVDB do (:1579-1593, :1792-1816, :3946-3989) and imports everything (device needed unless
    // parse_only).  SaveVBX writes the pools kept from the last host import / LoadVBX plus the atlas read back from the GPU.
    int  LoadVBX(const char* fname, bool parse_only = false);
    int  SaveVBX(const char* fname);
    void SetEpsilon(float eps, int maxiter) { mEpsilon = eps; mMaxIter = maxiter; }   // gvdb_volume_gvdb.h:334
    const char* getVDBInfoHost() const { return (const char*)&mVDBHost; }             // as derived by LoadVBX / given to ImportTopologyHost
    int  CommitTransferFunc();                                               // :4892-4898
    int  AddRenderBuf(int chan, int width, int height, int byteperpix);      // :4164-4178
    int  ResizeRenderBuf(int chan, int width, int height, int byteperpix);   // :4211-4238
    int  ReadRenderBuf(int chan, unsigned char* outptr);                     // :4241-4251
    // extension: render buffer j lives on frame lane j % n (gvdbx_lanes): Render(.., rbuf) and the reads of rbuf are
    // enqueued there, so frames in different render buffers overlap.  ReadRenderBufAsync returns at once; the host
    // buffer (pinned) is valid after SyncRenderBuf(chan).
    int  SetRenderLanes(int n);
    void SetReadbackBands(int n) { mReadbackBands = n; }    // 0 = automatic (~2 MB of pixels per band), 1 = one launch per frame
    int  ReadRenderBufAsync(int chan, unsigned char* outptr);
    int  SyncRenderBuf(int chan);
    void PrepareRender(int w, int h, char shading);                          // :4254-4306 (fills mScnInfo)
    int  Render(char shading, uint8_t chan, uint8_t rbuf);                   // :4336-4381
    const char* getScnInfo() const { return (const char*)&mScnInfo; }
    uint64_t getRenderBufGPU(int chan) const { return chan < (int)mRenderBuf.size() ? mRenderBuf[chan].gpu : 0; }
    gvdbx_t* handle() { return mCtx; }
    const char* lastError() const;

    Matrix4 mXform, mInvXform, mInvXrot;
private:
    struct RenderBuf { uint64_t gpu = 0; size_t max = 0, size = 0, stride = 0; };
    gvdbx_t*  mCtx = nullptr;
    Scene*    mScene = nullptr;
    GxScnInfo mScnInfo;
    std::vector<RenderBuf> mRenderBuf;
    bool      mTransferCommitted = false;
    // host copies for SaveVBX
    GxVDBInfo mVDBHost;
    std::vector<std::vector<unsigned char>> mPool0, mPool1;
    uint64_t  mRoot = 0;
    int       mLevels = 0, mAtlasRes[3] = {0, 0, 0}, mAtlasCnt[3] = {0, 0, 0};
    Vec3      mPretrans, mAngs, mScale = Vec3(1, 1, 1), mTrans;
    float     mEpsilon = 0.001f;                                             // gvdb_volume_gvdb.cpp:71-72
    int       mMaxIter = 256;
    std::string mErr;
    int       mLanes = 0;
    int       mReadbackBands = 0;       // Render without lanes: bands per frame for the overlapped read-back (0 = ~2 MB per band, 1 = off)
};

}  // namespace gvdbx

// flat C view of the same objects for ctypes / FFI hosts
extern "C" {
typedef struct gvdbxh_volume gvdbxh_volume;
gvdbxh_volume* gvdbxh_create(int cuda_device);      // cuda_device < 0: host-state only (no device context)
void  gvdbxh_destroy(gvdbxh_volume*);
void  gvdbxh_set_transform(gvdbxh_volume*, const float pretrans[3], const float scal[3], const float angs[3], const float trans[3]);
void  gvdbxh_camera(gvdbxh_volume*, float fov, const float angs[3], const float target[3], float dist, float dolly);
void  gvdbxh_camera_nearfar(gvdbxh_volume*, float n, float f);
void  gvdbxh_light(gvdbxh_volume*, const float angs[3], const float target[3], float dist, float dolly);
void  gvdbxh_scene_params(gvdbxh_volume*, const float steps[3], const float extinct[3], const float thresh[3],
                          const float cutoff[3], const float backclr[4], const float shadow[3]);
void  gvdbxh_cross_section(gvdbxh_volume*, const float pnt[3], const float norm[3]);   // Scene::SetCrossSection, src/gvdb_scene.h:147
void  gvdbxh_linear_transfer(gvdbxh_volume*, float t0, float t1, const float a[4], const float b[4]);
const float* gvdbxh_transfer_table(gvdbxh_volume*);
void  gvdbxh_set_res(gvdbxh_volume*, int w, int h);
void  gvdbxh_prepare_render(gvdbxh_volume*, int w, int h, int shading, void* scninfo_out416);
int   gvdbxh_import_topology_host(gvdbxh_volume*, const void* vdbinfo, const void* const* pool0, const void* const* pool1, const uint64_t* pool1_bytes);
int   gvdbxh_import_atlas_host(gvdbxh_volume*, int chan, const float* texels, int rx, int ry, int rz);
int   gvdbxh_load_vbx(gvdbxh_volume*, const char* fname, int parse_only);
int   gvdbxh_save_vbx(gvdbxh_volume*, const char* fname);
void  gvdbxh_vdbinfo(gvdbxh_volume*, void* out1232);
void  gvdbxh_set_epsilon(gvdbxh_volume*, float eps, int maxiter);
int   gvdbxh_commit_transfer(gvdbxh_volume*);
int   gvdbxh_add_render_buf(gvdbxh_volume*, int chan, int w, int h, int bpp);
int   gvdbxh_render(gvdbxh_volume*, int shading, int chan, int rbuf);
int   gvdbxh_read_render_buf(gvdbxh_volume*, int chan, void* out);
int   gvdbxh_set_render_lanes(gvdbxh_volume*, int n);
int   gvdbxh_set_readback_bands(gvdbxh_volume*, int n);
int   gvdbxh_read_render_buf_async(gvdbxh_volume*, int chan, void* out);
int   gvdbxh_sync_render_buf(gvdbxh_volume*, int chan);
int   gvdbxh_set_option(gvdbxh_volume*, int option, int value);
const char* gvdbxh_last_error(gvdbxh_volume*);
}
