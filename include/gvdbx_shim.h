/* gvdbx_shim.h — zero-patch drop-in for VolumeGVDB::Render on top of the STOCK libgvdb (Level A of SURVEY.md §8b).
 *
 * `Render` is a non-virtual member of an exported class (src/gvdb_volume_gvdb.h:341), but everything it needs is public
 * or protected (PrepareRender / PrepareVDB :483-486, getVDBInfo / getScnInfo :519-525, mPool and mRenderBuf in
 * src/gvdb_volume_base.h:74-75), so an application swaps `nvdb::VolumeGVDB` for this subclass and `Render` for `RenderX`.
 * Everything else — Configure, ActivateSpace, LoadVBX, UpdateApron, Compute, Scene setters, AddRenderBuf, ReadRenderBuf —
 * stays the reference's own code; pools, atlas CUarrays and render buffers stay owned by libgvdb in ITS CUcontext
 * (cuCtxCreate in StartCuda, src/gvdb_allocator.cpp:1105-1109).  libgvdbx adopts that context in InitX().
 *
 * Proven in-process by oracle/ref_harness.cpp --gvdbx (tests/test_parity_gpu.py::test_level_a_shim_inside_reference):
 * RenderX() writes, into the reference's own mRenderBuf, the same bytes as Render() in all eight shade modes.
 */
#ifndef GVDBX_SHIM_H
#define GVDBX_SHIM_H

#include "gvdb.h"
#include "gvdbx.h"

class VolumeGVDBX : public nvdb::VolumeGVDB {
public:
    ~VolumeGVDBX() { if (mX) gvdbx_destroy(mX); }

    /* once, after SetCudaDevice() + Initialize(): `cuda_device` as given to SetCudaDevice.  Must be called while the GVDB
     * context is current on this thread (it is, from SetCudaDevice on: StartCuda ends with cuCtxSetCurrent). */
    bool InitX(int cuda_device) { return gvdbx_create(&mX, cuda_device, /*stream*/ nullptr) == 0; }

    /* after the ATLAS CONTENT changed other than through UpdateApronX (LoadVBX, Compute, AtlasCommit ...): the reference has
     * no dirty notification for atlas writes.  Topology changes need no call: RenderX watches mVDBInfo.update. */
    bool SyncX(uchar chan = 0)
    {
        mVDBInfo.update = true;
        PrepareVDB();                                       /* fills mVDBInfo; pools are on the device */
        if (gvdbx_import_topology(mX, getVDBInfo()) != 0) return fail();
        nvdb::DataPtr a = mPool->getAtlas(chan);
        nvdb::Vector3DI r = mPool->getAtlasRes(chan);
        if (gvdbx_import_atlas_array(mX, chan, (void*)a.garray, r.x, r.y, r.z) != 0) return fail();
        if (mVDBInfo.clr_chan != CHAN_UNDEF) {              /* SetColorChannel: uchar4 atlas, filter mode as given to AddChannel */
            nvdb::DataPtr c = mPool->getAtlas(mVDBInfo.clr_chan);
            if (gvdbx_import_color_array(mX, (void*)c.garray, c.filter == F_LINEAR ? 1 : 0) != 0) return fail();
        } else {
            gvdbx_clear_color(mX);
        }
        mSynced = true;
        return true;
    }

    /* drop-in for Render(): same arguments, same render buffer, same bytes out of ReadRenderBuf() */
    void RenderX(char shading, uchar chan = 0, uchar rbuf = 0)
    {
        int w = (int)mRenderBuf[rbuf].stride, h = (int)(mRenderBuf[rbuf].max / mRenderBuf[rbuf].stride);
        const bool topo_dirty = mVDBInfo.update || !mSynced;   /* FinishTopology / UpdateAtlas / SetEpsilon / SetColorChannel */
        PrepareRender(w, h, shading);                       /* fills mScnInfo exactly as Render() does */
        PrepareVDB();
        if (topo_dirty && !SyncX(chan)) return;
        /* the transfer function travels as the device pointer CommitTransferFunc put into ScnInfo.transfer */
        if (gvdbx_render(mX, getScnInfo(), shading, chan, (uint64_t)mRenderBuf[rbuf].gpu, 0, 0, 0, 0) != 0) fail();
    }

    /* The same frame rendered in `bands` horizontal bands + the read-back that overlaps with them (gvdbx_render_banded /
     * gvdbx_read_banded): RenderX(..., bands) followed by ReadRenderBufX() is Render() followed by ReadRenderBuf() with the
     * device-to-host copy of band b hidden behind the rendering of the bands below it.  The reference's own ReadRenderBuf
     * stays valid after a banded RenderX (work on the context's stream is ordered behind every band). */
    void RenderX(char shading, uchar chan, uchar rbuf, int bands)
    {
        int w = (int)mRenderBuf[rbuf].stride, h = (int)(mRenderBuf[rbuf].max / mRenderBuf[rbuf].stride);
        const bool topo_dirty = mVDBInfo.update || !mSynced;
        PrepareRender(w, h, shading);
        PrepareVDB();
        if (topo_dirty && !SyncX(chan)) return;
        if (gvdbx_render_banded(mX, getScnInfo(), shading, chan, (uint64_t)mRenderBuf[rbuf].gpu, bands) != 0) fail();
    }
    void ReadRenderBufX(uchar rbuf, unsigned char* outptr)
    {
        if (gvdbx_read_banded(mX, (uint64_t)mRenderBuf[rbuf].gpu, outptr, (size_t)mRenderBuf[rbuf].size) != 0) fail();
    }

    /* UpdateApron(chan, boundval) on the shared CUarray, keeping libgvdbx's derived tables coherent (no re-import) */
    void UpdateApronX(uchar chan = 0, float boundval = 0.0f)
    {
        if ((mVDBInfo.update || !mSynced) && !SyncX(chan)) return;
        if (gvdbx_update_apron(mX, chan, boundval) != 0) fail();
    }

    gvdbx_t* x() { return mX; }

private:
    bool fail() { gprintf("gvdbx: %s\n", gvdbx_last_error(mX)); gerror(); return false; }
    gvdbx_t* mX = nullptr;
    bool     mSynced = false;
};

#endif /* GVDBX_SHIM_H */
