// gvdbx_host.cpp — see gvdbx_host.h.  Host state producers for the render path: their outputs are the bytes of ScnInfo.
#include "gvdbx_host.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

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
    memset(&mVDBHost, 0, sizeof mVDBHost);
    SetTransform(Vec3(0, 0, 0), Vec3(1, 1, 1), Vec3(0, 0, 0), Vec3(0, 0, 0));   // gvdb_volume_gvdb.cpp:77
}
VolumeGVDB::~VolumeGVDB()
{
    if (mCtx) gvdbx_destroy(mCtx);
    delete mScene;
}
const char* VolumeGVDB::lastError() const
{
    if (!mErr.empty()) return mErr.c_str();
    return mCtx ? gvdbx_last_error(mCtx) : "no device context";
}
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
    // host copies of what was imported, so that SaveVBX can write them back out
    memcpy(&mVDBHost, v, sizeof mVDBHost);
    mLevels = 0;                                        // levels the tree was configured with (upper ones may be empty)
    while (mLevels < GX_MAXLEV && mVDBHost.nodewid[mLevels] > 0) mLevels++;
    mPool0.assign(mLevels, {}); mPool1.assign(mLevels, {});
    for (int l = 0; l < mLevels; l++) {
        const size_t b0 = size_t(mVDBHost.nodecnt[l]) * mVDBHost.nodewid[l];
        if (p0[l] && b0) mPool0[l].assign((const unsigned char*)p0[l], (const unsigned char*)p0[l] + b0);
        if (p1[l] && n1[l]) mPool1[l].assign((const unsigned char*)p1[l], (const unsigned char*)p1[l] + n1[l]);
    }
    mRoot = uint64_t(mVDBHost.top_lev) << 8;            // Elem(0, top_lev, 0): group | level << 8 | index << 16
    mAtlasRes[0] = mVDBHost.atlas_res.x; mAtlasRes[1] = mVDBHost.atlas_res.y; mAtlasRes[2] = mVDBHost.atlas_res.z;
    mAtlasCnt[0] = mVDBHost.atlas_cnt.x; mAtlasCnt[1] = mVDBHost.atlas_cnt.y; mAtlasCnt[2] = mVDBHost.atlas_cnt.z;
    return mCtx ? gvdbx_import_topology_host(mCtx, v, p0, p1, n1) : GVDBX_E_STATE;
}

// ------------------------------------------------------------------------------------------------ VBX
namespace {
struct Reader {
    FILE* fp; bool ok = true;
    template <class T> void get(T* dst, size_t n = 1) { if (ok && fread(dst, sizeof(T), n, fp) != n) ok = false; }
};
}

int VolumeGVDB::LoadVBX(const char* fname, bool parse_only)
{
    mErr.clear();
    FILE* fp = fopen(fname, "rb");
    if (!fp) { mErr = std::string("LoadVBX: unable to open ") + fname; return GVDBX_E_ARG; }
    Reader rd{fp};
    auto fail = [&](int code, const std::string& m) { fclose(fp); mErr = "LoadVBX: " + m; return code; };
    // every count / width of the header is checked against the file size before anything is allocated
    fseeko(fp, 0, SEEK_END);
    const uint64_t file_bytes = (uint64_t)ftello(fp);
    fseeko(fp, 0, SEEK_SET);

    //--- file header (gvdb_volume_gvdb.cpp:549-590)
    unsigned char major = 0, minor = 0;
    rd.get(&major); rd.get(&minor);
    if ((major == 1 && minor >= 11) || major > 1) {     // 1.11+ stores the grid transform
        rd.get(&mPretrans.x, 3); rd.get(&mAngs.x, 3); rd.get(&mScale.x, 3); rd.get(&mTrans.x, 3);
        SetTransform(mPretrans, mScale, mAngs, mTrans);
    }
    int num_grids = 0;
    rd.get(&num_grids);
    unsigned char read_masks = 0;
    if (major >= 2) rd.get(&read_masks);
    else if (major == 1 && minor == 0) read_masks = 1;  // GVDB 1.0 always used bitmasks
    if (!rd.ok || num_grids < 1 || num_grids > 1024) return fail(GVDBX_E_ARG, "bad header");
    if (read_masks) return fail(GVDBX_E_UNSUPPORTED, "bitmask child lists (GVDB 1.0 files) are not supported");
    std::vector<uint64_t> grid_offs(num_grids);
    rd.get(grid_offs.data(), num_grids);

    //--- grid header of the first grid (the reference loads every grid over the previous one; files it writes hold one)
    char grid_name[256];
    unsigned char dtype = 0, components = 0, compress = 0, topotype = 0, layout = 0;
    float voxelsize[3];
    int leafcnt = 0, leafdim[3], apron = 0, num_chan = 0, reuse = 0, axiscnt[3], axisres[3], levels = 0;
    uint64_t atlas_sz = 0, root = 0;
    rd.get(grid_name, 256); rd.get(&dtype); rd.get(&components); rd.get(&compress); rd.get(voxelsize, 3);
    rd.get(&leafcnt); rd.get(leafdim, 3); rd.get(&apron); rd.get(&num_chan); rd.get(&atlas_sz);
    rd.get(&topotype); rd.get(&reuse); rd.get(&layout); rd.get(axiscnt, 3); rd.get(axisres, 3);
    //--- topology section
    rd.get(&levels); rd.get(&root);
    if (!rd.ok || levels < 1 || levels > GX_MAXLEV) return fail(GVDBX_E_ARG, "bad grid header / level count");
    if (compress != 0) return fail(GVDBX_E_UNSUPPORTED, "compressed grids");
    int ld[GX_MAXLEV], res[GX_MAXLEV], range[GX_MAXLEV][3], cnt0[GX_MAXLEV], width0[GX_MAXLEV], cnt1[GX_MAXLEV], width1[GX_MAXLEV];
    for (int n = 0; n < levels; n++) {
        rd.get(&ld[n]); rd.get(&res[n]); rd.get(range[n], 3); rd.get(&cnt0[n]); rd.get(&width0[n]); rd.get(&cnt1[n]); rd.get(&width1[n]);
    }
    if (!rd.ok) return fail(GVDBX_E_ARG, "truncated topology header");
    if (apron < 0 || apron > 4 || num_chan < 1 || num_chan > 32) return fail(GVDBX_E_ARG, "bad apron / channel count");
    for (int a = 0; a < 3; a++)
        if (axiscnt[a] <= 0 || axisres[a] <= 0 || axisres[a] > 65536 || axisres[a] % axiscnt[a] != 0) return fail(GVDBX_E_ARG, "bad atlas geometry");
    {
        uint64_t need = 0;
        for (int n = 0; n < levels; n++) {
            if (ld[n] < 1 || ld[n] > 8 || range[n][0] <= 0 || range[n][1] <= 0 || range[n][2] <= 0) return fail(GVDBX_E_ARG, "bad level geometry (log2dim 1..8, positive range)");
            if (cnt0[n] < 0 || cnt1[n] < 0 || width0[n] < 0 || width1[n] < 0 || width0[n] > (1 << 20) || width1[n] > (1 << 28)) return fail(GVDBX_E_ARG, "bad pool size");
            need += uint64_t(cnt0[n]) * uint64_t(width0[n]) + uint64_t(cnt1[n]) * uint64_t(width1[n]);
        }
        need += 4ull * uint64_t(axisres[0]) * uint64_t(axisres[1]) * uint64_t(axisres[2]);
        if (need > file_bytes) return fail(GVDBX_E_ARG, "header announces more data than the file holds");
    }
    if (width0[0] != (int)sizeof(GxNode)) return fail(GVDBX_E_UNSUPPORTED, "node records of another library version (width != 64)");
    std::vector<std::vector<unsigned char>> pool0(levels), pool1(levels);
    for (int n = 0; n < levels; n++) {
        if (cnt0[n] < 0 || width0[n] < 0) return fail(GVDBX_E_ARG, "bad pool size");
        pool0[n].resize(size_t(cnt0[n]) * width0[n]);
        if (!pool0[n].empty()) rd.get(pool0[n].data(), pool0[n].size());
    }
    for (int n = 0; n < levels; n++) {
        if (cnt1[n] < 0 || width1[n] < 0) return fail(GVDBX_E_ARG, "bad pool size");
        pool1[n].resize(size_t(cnt1[n]) * width1[n]);
        if (!pool1[n].empty()) rd.get(pool1[n].data(), pool1[n].size());
    }
    if (!rd.ok) return fail(GVDBX_E_ARG, "truncated pools");

    //--- VDBInfo as FinishTopology + PrepareVDB produce it (:1579-1593, :1792-1816, :3946-3989)
    GxVDBInfo v;
    memset(&v, 0, sizeof v);
    int tlev = 1;
    for (int n = levels - 1; n >= 0; n--) {
        v.dim[n] = ld[n];
        v.res[n] = 1 << ld[n];
        v.noderange[n] = {range[n][0], range[n][1], range[n][2]};
        v.vdel[n] = {float(range[n][0]) / float(v.res[n]), float(range[n][1]) / float(v.res[n]), float(range[n][2]) / float(v.res[n])};
        v.nodecnt[n] = cnt0[n]; v.nodewid[n] = width0[n]; v.childwid[n] = width1[n];
        if (cnt0[n] == 1) tlev = n;
    }
    v.atlas_apron = apron;
    v.atlas_cnt = {axiscnt[0], axiscnt[1], axiscnt[2]};
    v.atlas_res = {axisres[0], axisres[1], axisres[2]};
    v.brick_res = axiscnt[0] > 0 ? axisres[0] / axiscnt[0] : 0;
    for (int n = 0; n < apron && n < 4; n++) { v.apron_table[n] = n; v.apron_table[(apron * 2 - 1) - n] = (v.brick_res - 1) - n; }
    v.top_lev = tlev; v.epsilon = mEpsilon; v.max_iter = mMaxIter;
    v.clr_chan = GX_CHAN_UNDEF;
    // ComputeBounds: union of the active leaves' boxes
    if (cnt0[0] > 0) {
        const GxNode* nd = (const GxNode*)pool0[0].data();
        float mn[3] = {float(nd->mPos.x), float(nd->mPos.y), float(nd->mPos.z)}, mx[3] = {mn[0], mn[1], mn[2]};
        for (int i = 0; i < cnt0[0]; i++) {
            const GxNode* c = (const GxNode*)(pool0[0].data() + size_t(i) * width0[0]);
            if (!c->mFlags) continue;
            const int p[3] = {c->mPos.x, c->mPos.y, c->mPos.z};
            for (int a = 0; a < 3; a++) {
                if (p[a] < mn[a]) mn[a] = float(p[a]);
                if (p[a] + range[0][a] > mx[a]) mx[a] = float(p[a] + range[0][a]);
            }
        }
        v.bmin = {mn[0], mn[1], mn[2]}; v.bmax = {mx[0], mx[1], mx[2]};
    }
    mVDBHost = v;
    mLevels = levels; mRoot = root;
    for (int a = 0; a < 3; a++) { mAtlasRes[a] = axisres[a]; mAtlasCnt[a] = axiscnt[a]; }

    //--- atlas section: channel 0 must be T_FLOAT (3); further channels are skipped
    std::vector<float> atlas;
    for (int chan = 0; chan < num_chan; chan++) {
        int chan_type = 0, chan_stride = 0;
        rd.get(&chan_type); rd.get(&chan_stride);
        const size_t bytes = size_t(chan_stride) * axisres[0] * axisres[1] * axisres[2];
        if (!rd.ok || chan_stride <= 0) return fail(GVDBX_E_ARG, "truncated atlas header");
        if (chan == 0) {
            if (chan_type != 3 || chan_stride != 4) return fail(GVDBX_E_UNSUPPORTED, "channel 0 is not T_FLOAT");
            atlas.resize(bytes / 4);
            rd.get(atlas.data(), atlas.size());
        } else if (fseeko(fp, (off_t)bytes, SEEK_CUR) != 0) rd.ok = false;
    }
    if (!rd.ok) return fail(GVDBX_E_ARG, "truncated atlas");
    fclose(fp);
    if (parse_only) { mPool0 = pool0; mPool1 = pool1; return GVDBX_OK; }
    if (!mCtx) { mErr = "LoadVBX: no device context"; return GVDBX_E_STATE; }

    const void* p0[10] = {}; const void* p1[10] = {}; uint64_t n1[10] = {};
    for (int n = 0; n < levels; n++) { p0[n] = pool0[n].empty() ? nullptr : pool0[n].data(); p1[n] = pool1[n].empty() ? nullptr : pool1[n].data(); n1[n] = pool1[n].size(); }
    int rc = ImportTopologyHost(&v, p0, p1, n1);
    mRoot = root;
    if (rc == GVDBX_OK && !atlas.empty()) rc = gvdbx_import_atlas_host(mCtx, 0, atlas.data(), axisres[0], axisres[1], axisres[2]);
    return rc;
}

int VolumeGVDB::SaveVBX(const char* fname)
{
    mErr.clear();
    if (!mCtx || mLevels < 1 || mPool0.empty()) { mErr = "SaveVBX: nothing imported from host pools / LoadVBX"; return GVDBX_E_STATE; }
    const size_t texels = size_t(mAtlasRes[0]) * mAtlasRes[1] * mAtlasRes[2];
    std::vector<float> atlas(texels);
    int rc = gvdbx_export_atlas_host(mCtx, 0, atlas.data(), mAtlasRes[0], mAtlasRes[1], mAtlasRes[2]);
    if (rc) return rc;
    FILE* fp = fopen(fname, "wb");
    if (!fp) { mErr = std::string("SaveVBX: unable to open ") + fname; return GVDBX_E_ARG; }
    auto put = [&](const void* p, size_t bytes) { fwrite(p, 1, bytes, fp); };
    const unsigned char major = 1, minor = 11;             // MAJOR_VERSION / MINOR_VERSION, gvdb_volume_gvdb.cpp:31-32
    put(&major, 1); put(&minor, 1);
    put(&mPretrans.x, 12); put(&mAngs.x, 12); put(&mScale.x, 12); put(&mTrans.x, 12);
    const int num_grids = 1;
    put(&num_grids, 4);
    uint64_t grid_off = 0;
    const long grid_table = ftell(fp);
    put(&grid_off, 8);
    grid_off = (uint64_t)ftell(fp);
    char grid_name[256];
    memset(grid_name, 0, sizeof grid_name);                // (the reference writes this field uninitialised)
    const unsigned char dtype = 'f', components = 1, compress = 0, topotype = 2, layout = 0;
    const float voxelsize[3] = {1, 1, 1};
    const int leafcnt = mVDBHost.nodecnt[0], leafdim[3] = {mVDBHost.res[0], mVDBHost.res[0], mVDBHost.res[0]};
    const int apron = mVDBHost.atlas_apron, num_chan = 1, reuse = 0;
    const uint64_t atlas_sz = uint64_t(texels) * 4;
    put(grid_name, 256); put(&dtype, 1); put(&components, 1); put(&compress, 1); put(voxelsize, 12);
    put(&leafcnt, 4); put(leafdim, 12); put(&apron, 4); put(&num_chan, 4); put(&atlas_sz, 8);
    put(&topotype, 1); put(&reuse, 4); put(&layout, 1); put(mAtlasCnt, 12); put(mAtlasRes, 12);
    put(&mLevels, 4); put(&mRoot, 8);
    for (int n = 0; n < mLevels; n++) {
        const int res = mVDBHost.res[n], range[3] = {mVDBHost.noderange[n].x, mVDBHost.noderange[n].y, mVDBHost.noderange[n].z};
        const int cnt0 = mVDBHost.nodecnt[n], width0 = mVDBHost.nodewid[n], width1 = mVDBHost.childwid[n];
        const int cnt1 = width1 > 0 ? int(mPool1[n].size() / size_t(width1)) : 0;
        put(&mVDBHost.dim[n], 4); put(&res, 4); put(range, 12); put(&cnt0, 4); put(&width0, 4); put(&cnt1, 4); put(&width1, 4);
    }
    for (int n = 0; n < mLevels; n++) put(mPool0[n].data(), mPool0[n].size());
    for (int n = 0; n < mLevels; n++) put(mPool1[n].data(), mPool1[n].size());
    const int chan_type = 3, chan_stride = 4;              // T_FLOAT
    put(&chan_type, 4); put(&chan_stride, 4);
    put(atlas.data(), atlas.size() * 4);
    fseek(fp, grid_table, SEEK_SET);
    put(&grid_off, 8);
    const bool bad = ferror(fp) != 0;
    fclose(fp);
    if (bad) { mErr = "SaveVBX: write error"; return GVDBX_E_ARG; }
    return GVDBX_OK;
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
extern "C" int gvdbx_internal_alloc(gvdbx_t* h, uint64_t* ptr, size_t bytes);
extern "C" int gvdbx_internal_free(gvdbx_t* h, uint64_t ptr);

namespace gvdbx {
int VolumeGVDB::ResizeRenderBuf(int chan, int w, int h, int bpp)
{
    if (chan < 0 || chan >= (int)mRenderBuf.size() || w <= 0 || h <= 0 || bpp <= 0) return GVDBX_E_ARG;
    if (!mCtx) return GVDBX_E_STATE;
    RenderBuf& b = mRenderBuf[chan];
    b.max = (size_t)w * h; b.size = (size_t)w * h * bpp; b.stride = (size_t)w;
    if (chan == 0 && mScene) mScene->SetRes(w, h);
    if (b.gpu) gvdbx_internal_free(mCtx, b.gpu);
    b.gpu = 0;
    return gvdbx_internal_alloc(mCtx, &b.gpu, b.size);
}
int VolumeGVDB::ReadRenderBuf(int chan, unsigned char* out)
{
    if (chan < 0 || chan >= (int)mRenderBuf.size() || !mRenderBuf[chan].gpu || !out) return GVDBX_E_ARG;
    if (mLanes) { gvdbx_lane_select(mCtx, chan); return gvdbx_read_buffer(mCtx, mRenderBuf[chan].gpu, out, mRenderBuf[chan].size); }
    return gvdbx_read_banded(mCtx, mRenderBuf[chan].gpu, out, mRenderBuf[chan].size);
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
    if (mLanes) { gvdbx_lane_select(mCtx, rbuf); return gvdbx_render(mCtx, &mScnInfo, shading, chan, mRenderBuf[rbuf].gpu, 0, 0, 0, 0); }
    // one frame at a time (the reference's calling sequence): render in bands so that the synchronous ReadRenderBuf that follows
    // overlaps its copy with the bands still rendering.  Automatic: ~2 MB of pixels per band, at most 8 bands, one launch for
    // frames below 4 MB (measured, ms per frame of Render + ReadRenderBuf, 1 band -> automatic: 1080p level set 1.25 -> 1.06,
    // 4K voxel 5.88 -> 4.45, 4K deep + shadow 22.65 -> 21.24; 1024x768 trilinear 0.44 -> 0.46, hence the threshold).
    int nbands = mReadbackBands;
    if (nbands <= 0) {
        const size_t bytes = size_t(width) * height * 4;
        nbands = bytes < (4u << 20) ? 1 : int((bytes + (1u << 20)) >> 21);
        nbands = nbands > 8 ? 8 : nbands;
    }
    return gvdbx_render_banded(mCtx, &mScnInfo, shading, chan, mRenderBuf[rbuf].gpu, nbands);
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
// no C++ exception may cross the C boundary (allocation failures on absurd sizes)
int gvdbxh_load_vbx(gvdbxh_volume* h, const char* fname, int parse_only)
{
    try { return h->v.LoadVBX(fname, parse_only != 0); } catch (...) { return GVDBX_E_ARG; }
}
int gvdbxh_save_vbx(gvdbxh_volume* h, const char* fname)
{
    try { return h->v.SaveVBX(fname); } catch (...) { return GVDBX_E_ARG; }
}
void gvdbxh_vdbinfo(gvdbxh_volume* h, void* out) { memcpy(out, h->v.getVDBInfoHost(), 1232); }
void gvdbxh_set_epsilon(gvdbxh_volume* h, float eps, int maxiter) { h->v.SetEpsilon(eps, maxiter); }
int gvdbxh_commit_transfer(gvdbxh_volume* h) { return h->v.CommitTransferFunc(); }
int gvdbxh_add_render_buf(gvdbxh_volume* h, int chan, int w, int hh, int bpp) { return h->v.AddRenderBuf(chan, w, hh, bpp); }
int gvdbxh_render(gvdbxh_volume* h, int shading, int chan, int rbuf) { return h->v.Render((char)shading, (uint8_t)chan, (uint8_t)rbuf); }
int gvdbxh_read_render_buf(gvdbxh_volume* h, int chan, void* out) { return h->v.ReadRenderBuf(chan, (unsigned char*)out); }
int gvdbxh_set_render_lanes(gvdbxh_volume* h, int n) { return h->v.SetRenderLanes(n); }
int gvdbxh_set_readback_bands(gvdbxh_volume* h, int n) { h->v.SetReadbackBands(n); return GVDBX_OK; }
int gvdbxh_read_render_buf_async(gvdbxh_volume* h, int chan, void* out) { return h->v.ReadRenderBufAsync(chan, (unsigned char*)out); }
int gvdbxh_sync_render_buf(gvdbxh_volume* h, int chan) { return h->v.SyncRenderBuf(chan); }
int gvdbxh_set_option(gvdbxh_volume* h, int option, int value) { return h->v.handle() ? gvdbx_set_option(h->v.handle(), option, value) : GVDBX_E_STATE; }
const char* gvdbxh_last_error(gvdbxh_volume* h) { return h->v.lastError(); }
}
