// gvdbx_host.cpp — see gvdbx_host.h.  Host state producers for the render path: their outputs are the bytes of ScnInfo.
#include "gvdbx_host.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

namespace gvdbx {

static const float kDegToRad = 3.141592f / 180.0f;       // src/gvdb_types.h:59 (note: not M_PI)

// ------------------------------------------------------------------------------------------------ Matrix4
Matrix4& Matrix4::Identity() { memset(data, 0, sizeof data); data[0] = data[5] = data[10] = data[15] = 1.0f; return *this; }
Matrix4& Matrix4::Zero() { memset(data, 0, sizeof data); return *this; }

Matrix4& Matrix4::RotateZYX(const Vec3& a)
{
    const float cx = cosf(a.x * kDegToRad), sx = sinf(a.x * kDegToRad);
    const float cy = cosf(a.y * kDegToRad), sy = sinf(a.y * kDegToRad);
    const float cz = cosf(a.z * kDegToRad), sz = sinf(a.z * kDegToRad);
    data[0] = cy * cz;                  data[1] = cy * sz;                   data[2] = -sy;      data[3] = 0;
    data[4] = cz * sx * sy - cx * sz;   data[5] = cx * cz + sx * sy * sz;    data[6] = cy * sx;  data[7] = 0;
    data[8] = cx * cz * sy + sx * sz;   data[9] = -cz * sx + cx * sy * sz;   data[10] = cx * cy; data[11] = 0;
    data[12] = 0; data[13] = 0; data[14] = 0; data[15] = 1;
    return *this;
}
Matrix4& Matrix4::RotateTZYXS(const Vec3& a, const Vec3& t, const Vec3& s)
{
    RotateZYX(a);
    data[12] = t.x; data[13] = t.y; data[14] = t.z;
    for (int i = 0; i < 3; i++) { data[i] *= s.x; data[i + 4] *= s.y; data[i + 8] *= s.z; }
    return *this;
}
Matrix4& Matrix4::PreTranslate(const Vec3& t)
{
    data[12] += data[0] * t.x + data[4] * t.y + data[8] * t.z;
    data[13] += data[1] * t.x + data[5] * t.y + data[9] * t.z;
    data[14] += data[2] * t.x + data[6] * t.y + data[10] * t.z;
    return *this;
}
Matrix4& Matrix4::MulAssign(const Matrix4& m)            // this = this * m
{
    float o[16];
    memcpy(o, data, sizeof o);
    const float* op = m.data;
    for (int c = 0; c < 4; c++)
        for (int r = 0; r < 4; r++)
            data[4 * c + r] = op[4 * c] * o[r] + op[4 * c + 1] * o[4 + r] + op[4 * c + 2] * o[8 + r] + op[4 * c + 3] * o[12 + r];
    return *this;
}
Matrix4& Matrix4::LeftMultiplyInPlace(const Matrix4& m)  // this = m * this
{
    Matrix4 right = *this;
    for (int k = 0; k < 4; k++)
        for (int i = 0; i < 4; i++)
            data[i + 4 * k] = m.data[i] * right.data[4 * k] + m.data[4 + i] * right.data[4 * k + 1]
                            + m.data[8 + i] * right.data[4 * k + 2] + m.data[12 + i] * right.data[4 * k + 3];
    return *this;
}
Matrix4& Matrix4::ScaleInPlace(const Vec3& s)
{
    for (int i = 0; i < 4; i++) { data[4 * i] *= s.x; data[4 * i + 1] *= s.y; data[4 * i + 2] *= s.z; }
    return *this;
}
Matrix4& Matrix4::Basis(const Vec3& c1, const Vec3& c2, const Vec3& c3)
{
    data[0] = c1.x; data[1] = c2.x; data[2] = c3.x; data[3] = 0;
    data[4] = c1.y; data[5] = c2.y; data[6] = c3.y; data[7] = 0;
    data[8] = c1.z; data[9] = c2.z; data[10] = c3.z; data[11] = 0;
    data[12] = 0; data[13] = 0; data[14] = 0; data[15] = 1;
    return *this;
}
// General 4x4 inverse (despite the name — the reference's InvertTRS is a full inverse too, src/gvdb_vec.cpp:427-465).
// Adjugate by the Leibniz expansion, evaluated in double and rounded to float once per element.  The products are
// accumulated in the reference's term order (table below: sign, i, j, k -> sign * m[i]*m[j]*m[k]) because both the
// last double bit and the SIGN OF ZERO entries of the result depend on it, and ScnInfo.invxform must match byte for byte.
static const signed char kAdjTerms[16][6][4] = {
    {{-1,11,14, 5}, { 1,10,15, 5}, { 1,11,13, 6}, {-1,10,13, 7}, {-1,15, 6, 9}, { 1,14, 7, 9}},
    {{ 1, 1,11,14}, {-1, 1,10,15}, {-1,11,13, 2}, { 1,10,13, 3}, { 1,15, 2, 9}, {-1,14, 3, 9}},
    {{-1,15, 2, 5}, { 1,14, 3, 5}, { 1, 1,15, 6}, {-1,13, 3, 6}, {-1, 1,14, 7}, { 1,13, 2, 7}},
    {{ 1,11, 2, 5}, {-1,10, 3, 5}, {-1, 1,11, 6}, { 1, 1,10, 7}, { 1, 3, 6, 9}, {-1, 2, 7, 9}},
    {{ 1,11,14, 4}, {-1,10,15, 4}, {-1,11,12, 6}, { 1,10,12, 7}, { 1,15, 6, 8}, {-1,14, 7, 8}},
    {{-1, 0,11,14}, { 1, 0,10,15}, { 1,11,12, 2}, {-1,10,12, 3}, {-1,15, 2, 8}, { 1,14, 3, 8}},
    {{ 1,15, 2, 4}, {-1,14, 3, 4}, {-1, 0,15, 6}, { 1,12, 3, 6}, { 1, 0,14, 7}, {-1,12, 2, 7}},
    {{-1,11, 2, 4}, { 1,10, 3, 4}, { 1, 0,11, 6}, {-1, 0,10, 7}, {-1, 3, 6, 8}, { 1, 2, 7, 8}},
    {{-1,11,13, 4}, { 1,11,12, 5}, {-1,15, 5, 8}, { 1,13, 7, 8}, { 1,15, 4, 9}, {-1,12, 7, 9}},
    {{-1, 1,11,12}, { 1, 0,11,13}, { 1, 1,15, 8}, {-1,13, 3, 8}, {-1, 0,15, 9}, { 1,12, 3, 9}},
    {{-1, 1,15, 4}, { 1,13, 3, 4}, { 1, 0,15, 5}, {-1,12, 3, 5}, { 1, 1,12, 7}, {-1, 0,13, 7}},
    {{ 1, 1,11, 4}, {-1, 0,11, 5}, { 1, 3, 5, 8}, {-1, 1, 7, 8}, {-1, 3, 4, 9}, { 1, 0, 7, 9}},
    {{ 1,10,13, 4}, {-1,10,12, 5}, { 1,14, 5, 8}, {-1,13, 6, 8}, {-1,14, 4, 9}, { 1,12, 6, 9}},
    {{ 1, 1,10,12}, {-1, 0,10,13}, {-1, 1,14, 8}, { 1,13, 2, 8}, { 1, 0,14, 9}, {-1,12, 2, 9}},
    {{ 1, 1,14, 4}, {-1,13, 2, 4}, {-1, 0,14, 5}, { 1,12, 2, 5}, {-1, 1,12, 6}, { 1, 0,13, 6}},
    {{-1, 1,10, 4}, { 1, 0,10, 5}, {-1, 2, 5, 8}, { 1, 1, 6, 8}, { 1, 2, 4, 9}, {-1, 0, 6, 9}},
};
Matrix4& Matrix4::InvertTRS()
{
    double m[16], adj[16];
    for (int i = 0; i < 16; i++) m[i] = (double)data[i];
    for (int e = 0; e < 16; e++) {
        double acc = 0;
        for (int t = 0; t < 6; t++) {
            const signed char* q = kAdjTerms[e][t];
            const double term = m[q[1]] * m[q[2]] * m[q[3]];
            if (t == 0) acc = (q[0] < 0) ? -term : term;
            else        acc = (q[0] < 0) ? acc - term : acc + term;
        }
        adj[e] = acc;
    }
    double det = data[0] * adj[0] + data[1] * adj[4] + data[2] * adj[8] + data[3] * adj[12];
    if (det == 0) return *this;
    det = 1.0f / det;
    for (int i = 0; i < 16; i++) data[i] = (float)(adj[i] * det);
    return *this;
}

// ------------------------------------------------------------------------------------------------ vector helpers
// Vector3DF::Normalize / Cross evaluate in double and round per component (src/gvdb_vec.h:125-133, 256-265)
static void normalize(Vec3& v)
{
    double n = (double)v.x * (double)v.x + (double)v.y * (double)v.y + (double)v.z * (double)v.z;
    if (n != 0.0) {
        double r = 1.0 / sqrt(n);
        v.x = (float)(v.x * r); v.y = (float)(v.y * r); v.z = (float)(v.z * r);
    }
}
static void cross(Vec3& a, const Vec3& v)
{
    double ax = a.x, ay = a.y, az = a.z;
    a.x = (float)(ay * (double)v.z - az * (double)v.y);
    a.y = (float)(-ax * (double)v.z + az * (double)v.x);
    a.z = (float)(ax * (double)v.y - ay * (double)v.x);
}

// ------------------------------------------------------------------------------------------------ Camera3D
Camera3D::Camera3D()
{
    up_dir = Vec3(0.0f, 1.0f, 0.0f);
    mAspect = (float)800.0f / 600.0f;
    mDolly = 5.0f; mFov = 40.0f; mNear = 0.1f; mFar = 5000.0f; mOrbitDist = 0;
    setOrbit(0, 45, 0, Vec3(0, 0, 0), 120.0f, 1.0f);
    updateMatricies();
}
void Camera3D::setOrbit(float ax, float ay, float az, Vec3 tp, float dist, float dolly)
{
    ang_euler = Vec3(ax, ay, az);
    mOrbitDist = dist;
    mDolly = dolly;
    // the reference calls cos()/sin() on float arguments (C++ overloads -> single precision) and widens the product
    double dx = cosf(ang_euler.y * kDegToRad) * sinf(ang_euler.x * kDegToRad);
    double dy = sinf(ang_euler.y * kDegToRad);
    double dz = cosf(ang_euler.y * kDegToRad) * cosf(ang_euler.x * kDegToRad);
    from_pos.x = tp.x + (float)dx * mOrbitDist;
    from_pos.y = tp.y + (float)dy * mOrbitDist;
    from_pos.z = tp.z + (float)dz * mOrbitDist;
    to_pos = tp;
    updateMatricies();
}
void Camera3D::updateMatricies()
{
    // gluLookAt basis
    dir_vec = Vec3(to_pos.x - from_pos.x, to_pos.y - from_pos.y, to_pos.z - from_pos.z);
    normalize(dir_vec);
    side_vec = dir_vec; cross(side_vec, up_dir); normalize(side_vec);
    up_vec = side_vec;  cross(up_vec, dir_vec);  normalize(up_vec);
    dir_vec.x *= -1; dir_vec.y *= -1; dir_vec.z *= -1;
    rotate_matrix.Basis(side_vec, up_vec, dir_vec);
    view_matrix = rotate_matrix;
    view_matrix.PreTranslate(Vec3(-from_pos.x, -from_pos.y, -from_pos.z));
    // gluPerspective
    float sx = (float)tanf(mFov * kDegToRad / 2.0f) * mNear;
    float sy = sx / mAspect;
    proj_matrix.Zero();
    proj_matrix.data[0] = 2.0f * mNear / sx;
    proj_matrix.data[5] = 2.0f * mNear / sy;
    proj_matrix.data[10] = -(mFar + mNear) / (mFar - mNear);
    proj_matrix.data[14] = -(2.0f * mFar * mNear) / (mFar - mNear);
    proj_matrix.data[11] = -1.0f;
    // (P * V_rotation_only)^-1
    Matrix4 vnt = view_matrix;
    vnt.data[12] = 0.0f; vnt.data[13] = 0.0f; vnt.data[14] = 0.0f;
    invviewproj_matrix = proj_matrix;
    invviewproj_matrix.MulAssign(vnt);
    invviewproj_matrix.InvertTRS();
    origRayWorld = from_pos;
    tlRayWorld = inverseRayProj(-1.0f, 1.0f, mNear);
    trRayWorld = inverseRayProj(1.0f, 1.0f, mNear);
    blRayWorld = inverseRayProj(-1.0f, -1.0f, mNear);
    brRayWorld = inverseRayProj(1.0f, -1.0f, mNear);
}
Vec3 Camera3D::inverseRayProj(float x, float y, float z) const
{
    const float* d = invviewproj_matrix.data;
    float wx = d[0] * x + d[4] * y + d[8] * z + d[12];
    float wy = d[1] * x + d[5] * y + d[9] * z + d[13];
    float wz = d[2] * x + d[6] * y + d[10] * z + d[14];
    float ww = d[3] * x + d[7] * y + d[11] * z + d[15];
    return Vec3(wx / ww, wy / ww, wz / ww);
}

// ------------------------------------------------------------------------------------------------ Scene
Scene::Scene()
{
    mShadowParams = Vec3(0.8f, 1.0f, 0);
    mSteps = Vec3(1.0f, 16.0f, 0.1f);
    mExtinct = Vec3(-1.1f, 1.5f, 0.0f);
    mVThreshold = Vec3(0.1f, 0.0f, 1.0f);
    mCutoff = Vec3(0.005f, 0.01f, 0.0f);
}
Scene::~Scene() { delete mCamera; delete mLight; free(mTransferFunc); }
Camera3D* Scene::SetCamera(Camera3D* cam) { if (mCamera != cam) delete mCamera; mCamera = cam; return cam; }
Light*    Scene::SetLight(int, Light* l) { if (mLight != l) delete mLight; mLight = l; return l; }
void Scene::SetRes(int x, int y) { mXres = x; mYres = y; if (mCamera) mCamera->setAspect((float)x / (float)y); }
void Scene::LinearTransferFunc(float t0, float t1, Vec4 a, Vec4 b)
{
    const int sz = GVDBX_TRANSFER_ENTRIES;
    int n0 = (int)(t0 * (float)sz), n1 = (int)(t1 * (float)sz);
    if (!mTransferFunc) mTransferFunc = (float*)calloc(sz, 4 * sizeof(float));
    for (int n = n0; n < n1; n++) {
        float u = float(n - n0) / float(n1 - n0);
        float* c = mTransferFunc + 4 * n;
        c[0] = a.x + u * (b.x - a.x); c[1] = a.y + u * (b.y - a.y); c[2] = a.z + u * (b.z - a.z); c[3] = a.w + u * (b.w - a.w);
    }
}

// ------------------------------------------------------------------------------------------------ VolumeGVDB
VolumeGVDB::VolumeGVDB()
{
    memset(&mScnInfo, 0, sizeof mScnInfo);
    SetTransform(Vec3(0, 0, 0), Vec3(1, 1, 1), Vec3(0, 0, 0), Vec3(0, 0, 0));   // gvdb_volume_gvdb.cpp:77
}
VolumeGVDB::~VolumeGVDB()
{
    if (mCtx) gvdbx_destroy(mCtx);
    delete mScene;
}
const char* VolumeGVDB::lastError() const { return mCtx ? gvdbx_last_error(mCtx) : "no device context"; }
int VolumeGVDB::SetCudaDevice(int devid, void* stream)
{
    if (mCtx) { gvdbx_destroy(mCtx); mCtx = nullptr; }
    return gvdbx_create(&mCtx, devid, stream);
}
int VolumeGVDB::Initialize()
{
    delete mScene;
    mScene = new Scene;
    mScene->SetCamera(new Camera3D);
    mScene->SetLight(0, new Light);
    mScene->SetVolumeRange(0.1f, 0, 1);
    mScene->LinearTransferFunc(0, 1, Vec4(0, 0, 0, 0), Vec4(1, 1, 1, 0.1f));
    return mCtx ? CommitTransferFunc() : GVDBX_OK;
}
void VolumeGVDB::SetTransform(Vec3 pretrans, Vec3 scal, Vec3 angs, Vec3 trans)
{
    Matrix4 xrot;
    xrot.RotateZYX(angs);
    mInvXrot.Identity();
    mInvXrot.ScaleInPlace(Vec3(1.0f / scal.x, 1.0f / scal.y, 1.0f / scal.z));
    xrot.InvertTRS();
    mInvXrot.LeftMultiplyInPlace(xrot);
    mXform.Identity();
    mXform.RotateTZYXS(angs, trans, scal);
    mXform.PreTranslate(pretrans);
    mInvXform = mXform;
    mInvXform.InvertTRS();
}
int VolumeGVDB::ImportTopologyHost(const void* v, const void* const* p0, const void* const* p1, const uint64_t* n1)
{
    return mCtx ? gvdbx_import_topology_host(mCtx, v, p0, p1, n1) : GVDBX_E_STATE;
}
int VolumeGVDB::ImportTopologyDevice(const void* v) { return mCtx ? gvdbx_import_topology(mCtx, v) : GVDBX_E_STATE; }
int VolumeGVDB::ImportAtlasHost(int chan, const float* t, int rx, int ry, int rz)
{
    return mCtx ? gvdbx_import_atlas_host(mCtx, chan, t, rx, ry, rz) : GVDBX_E_STATE;
}
int VolumeGVDB::ImportAtlasArray(int chan, void* arr, int rx, int ry, int rz)
{
    return mCtx ? gvdbx_import_atlas_array(mCtx, chan, arr, rx, ry, rz) : GVDBX_E_STATE;
}
int VolumeGVDB::CommitTransferFunc()
{
    if (!mCtx || !mScene || !mScene->getTransferFunc()) return GVDBX_E_STATE;
    int rc = gvdbx_set_transfer(mCtx, mScene->getTransferFunc());
    mTransferCommitted = (rc == GVDBX_OK);
    return rc;
}
int VolumeGVDB::AddRenderBuf(int chan, int w, int h, int bpp)
{
    if (chan < 0) return GVDBX_E_ARG;
    if ((int)mRenderBuf.size() < chan + 1) mRenderBuf.resize(chan + 1);
    return ResizeRenderBuf(chan, w, h, bpp);
}
}  // namespace gvdbx

// device buffer helpers live in the CUDA translation unit
extern "C" int gvdbx_internal_alloc(uint64_t* ptr, size_t bytes);
extern "C" int gvdbx_internal_free(uint64_t ptr);

namespace gvdbx {
int VolumeGVDB::ResizeRenderBuf(int chan, int w, int h, int bpp)
{
    if (chan < 0 || chan >= (int)mRenderBuf.size() || w <= 0 || h <= 0 || bpp <= 0) return GVDBX_E_ARG;
    if (!mCtx) return GVDBX_E_STATE;
    RenderBuf& b = mRenderBuf[chan];
    b.max = (size_t)w * h; b.size = (size_t)w * h * bpp; b.stride = (size_t)w;
    if (chan == 0 && mScene) mScene->SetRes(w, h);
    if (b.gpu) gvdbx_internal_free(b.gpu);
    b.gpu = 0;
    return gvdbx_internal_alloc(&b.gpu, b.size);
}
int VolumeGVDB::ReadRenderBuf(int chan, unsigned char* out)
{
    if (chan < 0 || chan >= (int)mRenderBuf.size() || !mRenderBuf[chan].gpu || !out) return GVDBX_E_ARG;
    if (mLanes) gvdbx_lane_select(mCtx, chan);
    return gvdbx_read_buffer(mCtx, mRenderBuf[chan].gpu, out, mRenderBuf[chan].size);
}
int VolumeGVDB::SetRenderLanes(int n)
{
    if (!mCtx) return GVDBX_E_STATE;
    int rc = gvdbx_lanes(mCtx, n);
    if (rc == GVDBX_OK) mLanes = n;
    return rc;
}
int VolumeGVDB::ReadRenderBufAsync(int chan, unsigned char* out)
{
    if (chan < 0 || chan >= (int)mRenderBuf.size() || !mRenderBuf[chan].gpu || !out) return GVDBX_E_ARG;
    if (mLanes) gvdbx_lane_select(mCtx, chan);
    return gvdbx_read_buffer_async(mCtx, mRenderBuf[chan].gpu, out, mRenderBuf[chan].size);
}
int VolumeGVDB::SyncRenderBuf(int chan)
{
    if (!mCtx) return GVDBX_E_STATE;
    if (mLanes) gvdbx_lane_select(mCtx, chan);
    return gvdbx_sync(mCtx);
}
void VolumeGVDB::PrepareRender(int w, int h, char shading)
{
    Camera3D* cam = mScene->getCamera();
    GxScnInfo& s = mScnInfo;
    s.width = w; s.height = h;
    s.camnear = cam->getNear(); s.camfar = cam->getFar();
    s.campos = {cam->origRayWorld.x, cam->origRayWorld.y, cam->origRayWorld.z};
    s.cams = {cam->tlRayWorld.x, cam->tlRayWorld.y, cam->tlRayWorld.z};
    s.camu = {cam->trRayWorld.x - s.cams.x, cam->trRayWorld.y - s.cams.y, cam->trRayWorld.z - s.cams.z};
    s.camv = {cam->blRayWorld.x - s.cams.x, cam->blRayWorld.y - s.cams.y, cam->blRayWorld.z - s.cams.z};
    // light position: application space -> voxel space, (x,y,z,1) * mInvXform   (gvdb_volume_gvdb.cpp:4271-4274)
    const Vec3& lp = mScene->getLight()->getPos();
    const float* m = mInvXform.data;
    const float lw = 1.0f;
    s.light_pos = { lp.x * m[0] + lp.y * m[4] + lp.z * m[8] + lw * m[12],
                    lp.x * m[1] + lp.y * m[5] + lp.z * m[9] + lw * m[13],
                    lp.x * m[2] + lp.y * m[6] + lp.z * m[10] + lw * m[14] };
    s.slice_pnt = {mScene->mSectionPnt.x, mScene->mSectionPnt.y, mScene->mSectionPnt.z};
    s.slice_norm = {mScene->mSectionNorm.x, mScene->mSectionNorm.y, mScene->mSectionNorm.z};
    s.shading = shading;
    s.filtering = (char)mScene->mFilterMode;
    s.frame = mScene->mFrame; s.samples = mScene->mSample;
    s.shadow_params = {mScene->mShadowParams.x, mScene->mShadowParams.y, mScene->mShadowParams.z};
    s.backclr = {mScene->mBackgroundClr.x, mScene->mBackgroundClr.y, mScene->mBackgroundClr.z, mScene->mBackgroundClr.w};
    s.extinct = {mScene->mExtinct.x, mScene->mExtinct.y, mScene->mExtinct.z};
    s.steps = {mScene->mSteps.x, mScene->mSteps.y, mScene->mSteps.z};
    s.cutoff = {mScene->mCutoff.x, mScene->mCutoff.y, mScene->mCutoff.z};
    s.thresh = {mScene->mVThreshold.x, mScene->mVThreshold.y, mScene->mVThreshold.z};
    memcpy(s.xform, mXform.data, sizeof s.xform);
    memcpy(s.invxform, mInvXform.data, sizeof s.invxform);
    memcpy(s.invxrot, mInvXrot.data, sizeof s.invxrot);
    s.transfer = 0;                      // the library holds its own device copy (gvdbx_set_transfer)
    s.outbuf = (uint64_t)-1;             // "NOT USED" in the reference
    s.dbuf = 0;                          // no depth buffer (mDepthBuf == 255)
}
int VolumeGVDB::Render(char shading, uint8_t chan, uint8_t rbuf)
{
    if (!mCtx || !mScene) return GVDBX_E_STATE;
    if (rbuf >= mRenderBuf.size() || !mRenderBuf[rbuf].gpu) return GVDBX_E_ARG;
    const int width = (int)mRenderBuf[rbuf].stride;
    const int height = (int)(mRenderBuf[rbuf].max / mRenderBuf[rbuf].stride);
    PrepareRender(width, height, shading);
    if (mLanes) gvdbx_lane_select(mCtx, rbuf);
    return gvdbx_render(mCtx, &mScnInfo, shading, chan, mRenderBuf[rbuf].gpu, 0, 0, 0, 0);
}
}  // namespace gvdbx

// ------------------------------------------------------------------------------------------------ flat C view
struct gvdbxh_volume { gvdbx::VolumeGVDB v; };
using gvdbx::Vec3; using gvdbx::Vec4;

extern "C" {
gvdbxh_volume* gvdbxh_create(int dev)
{
    gvdbxh_volume* h = new gvdbxh_volume;
    if (dev >= 0 && h->v.SetCudaDevice(dev) != GVDBX_OK) { delete h; return nullptr; }
    h->v.Initialize();
    return h;
}
void gvdbxh_destroy(gvdbxh_volume* h) { delete h; }
void gvdbxh_set_transform(gvdbxh_volume* h, const float p[3], const float s[3], const float a[3], const float t[3])
{
    h->v.SetTransform(Vec3(p[0], p[1], p[2]), Vec3(s[0], s[1], s[2]), Vec3(a[0], a[1], a[2]), Vec3(t[0], t[1], t[2]));
}
void gvdbxh_camera(gvdbxh_volume* h, float fov, const float a[3], const float t[3], float dist, float dolly)
{
    gvdbx::Camera3D* cam = new gvdbx::Camera3D;
    cam->setFov(fov);
    cam->setOrbit(Vec3(a[0], a[1], a[2]), Vec3(t[0], t[1], t[2]), dist, dolly);
    h->v.getScene()->SetCamera(cam);
}
void gvdbxh_camera_nearfar(gvdbxh_volume* h, float n, float f) { h->v.getScene()->getCamera()->setNearFar(n, f); }
void gvdbxh_light(gvdbxh_volume* h, const float a[3], const float t[3], float dist, float dolly)
{
    gvdbx::Light* l = new gvdbx::Light;
    l->setOrbit(Vec3(a[0], a[1], a[2]), Vec3(t[0], t[1], t[2]), dist, dolly);
    h->v.getScene()->SetLight(0, l);
}
void gvdbxh_scene_params(gvdbxh_volume* h, const float st[3], const float ex[3], const float th[3], const float cu[3],
                         const float bg[4], const float sh[3])
{
    gvdbx::Scene* s = h->v.getScene();
    s->SetSteps(st[0], st[1], st[2]); s->SetExtinct(ex[0], ex[1], ex[2]); s->SetVolumeRange(th[0], th[1], th[2]);
    s->SetCutoff(cu[0], cu[1], cu[2]); s->SetBackgroundClr(bg[0], bg[1], bg[2], bg[3]); s->SetShadowParams(sh[0], sh[1], sh[2]);
}
void gvdbxh_linear_transfer(gvdbxh_volume* h, float t0, float t1, const float a[4], const float b[4])
{
    h->v.getScene()->LinearTransferFunc(t0, t1, Vec4(a[0], a[1], a[2], a[3]), Vec4(b[0], b[1], b[2], b[3]));
}
void gvdbxh_cross_section(gvdbxh_volume* h, const float pnt[3], const float norm[3])
{
    h->v.getScene()->SetCrossSection({pnt[0], pnt[1], pnt[2]}, {norm[0], norm[1], norm[2]});
}
const float* gvdbxh_transfer_table(gvdbxh_volume* h) { return h->v.getScene()->getTransferFunc(); }
void gvdbxh_set_res(gvdbxh_volume* h, int w, int hh) { h->v.getScene()->SetRes(w, hh); }
void gvdbxh_prepare_render(gvdbxh_volume* h, int w, int hh, int shading, void* out)
{
    h->v.PrepareRender(w, hh, (char)shading);
    memcpy(out, h->v.getScnInfo(), GVDBX_SCNINFO_BYTES);
}
int gvdbxh_import_topology_host(gvdbxh_volume* h, const void* v, const void* const* p0, const void* const* p1, const uint64_t* n1)
{
    return h->v.ImportTopologyHost(v, p0, p1, n1);
}
int gvdbxh_import_atlas_host(gvdbxh_volume* h, int chan, const float* t, int rx, int ry, int rz) { return h->v.ImportAtlasHost(chan, t, rx, ry, rz); }
int gvdbxh_commit_transfer(gvdbxh_volume* h) { return h->v.CommitTransferFunc(); }
int gvdbxh_add_render_buf(gvdbxh_volume* h, int chan, int w, int hh, int bpp) { return h->v.AddRenderBuf(chan, w, hh, bpp); }
int gvdbxh_render(gvdbxh_volume* h, int shading, int chan, int rbuf) { return h->v.Render((char)shading, (uint8_t)chan, (uint8_t)rbuf); }
int gvdbxh_read_render_buf(gvdbxh_volume* h, int chan, void* out) { return h->v.ReadRenderBuf(chan, (unsigned char*)out); }
int gvdbxh_set_render_lanes(gvdbxh_volume* h, int n) { return h->v.SetRenderLanes(n); }
int gvdbxh_read_render_buf_async(gvdbxh_volume* h, int chan, void* out) { return h->v.ReadRenderBufAsync(chan, (unsigned char*)out); }
int gvdbxh_sync_render_buf(gvdbxh_volume* h, int chan) { return h->v.SyncRenderBuf(chan); }
int gvdbxh_set_option(gvdbxh_volume* h, int option, int value) { return h->v.handle() ? gvdbx_set_option(h->v.handle(), option, value) : GVDBX_E_STATE; }
const char* gvdbxh_last_error(gvdbxh_volume* h) { return h->v.lastError(); }
}
