// gvdbx_multi.cu — multi-GPU entry points of the C ABI (include/gvdbx.h): image-space partition of ONE frame across GPUs
// with the volume replicated.  The reference has no multi-GPU path (one VolumeGVDB per device, gvdb_volume_gvdb.h:325);
// north_star asks for tiles partitioned over 1/2/4/8 B200s and the frame assembled on rank 0.
//
//   gvdbx_render_multi    one process, several contexts: every context renders its tiles straight into the frame of
//                         context 0 (peer access over NVLink), joined by events — SURVEY.md 8b's gvdbx_render_multi
//   gvdbx_ring_*          one process PER GPU: a ring of frames in rank 0's memory, mapped by the other ranks with CUDA IPC;
//                         render kernels store their tiles there directly, two 4-byte flags per frame order producers and
//                         consumer (stream-ordered device operations only, no NCCL, no host synchronisation)
//   gvdbx_hostring_*      one process per GPU, frames wanted on the HOST: a ring of row-major frames in a POSIX shared-memory
//                         segment that every process page-locks; each rank renders full-width bands and copies ITS bands over
//                         ITS OWN PCIe link (N links instead of funnelling every frame through rank 0's)
#include "gvdbx_internal.h"

#include <atomic>
#include <chrono>
#include <fcntl.h>
#include <sched.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <thread>
#include <unistd.h>

// ------------------------------------------------------------------------------------------------ one process, n contexts
extern "C" int gvdbx_render_multi(gvdbx_t* const* ranks, int nranks, const void* scninfo, int shade_mode, int chan,
                                  uint64_t outbuf_rank0_d, int tile_size)
{
    if (!ranks || nranks < 1 || nranks > 64 || !scninfo || !outbuf_rank0_d) return GVDBX_E_ARG;
    for (int r = 0; r < nranks; r++) if (!ranks[r]) return GVDBX_E_ARG;
    gvdbx_t* h0 = ranks[0];
    // peer access from every other device to the device that owns the frame (idempotent)
    for (int r = 1; r < nranks; r++) {
        if (ranks[r]->device == h0->device) continue;
        GxCtx ctx_(ranks[r]);
        int can = 0;
        GX_CUDA(ranks[r], cudaDeviceCanAccessPeer(&can, ranks[r]->device, h0->device));
        if (!can) return gx_fail(ranks[r], GVDBX_E_UNSUPPORTED, "no peer access to the device that owns the frame");
        cudaError_t e = cudaDeviceEnablePeerAccess(h0->device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) GX_CUDA(ranks[r], e);
        cudaGetLastError();
    }
    // fork: every context's stream waits for what context 0's stream holds (the frame buffer may still be in use there);
    // render; join: context 0's stream waits for every renderer.  Events only.
    std::vector<cudaEvent_t> done(nranks, nullptr);
    cudaEvent_t start = nullptr;
    int rc = GVDBX_OK;
    {
        GxCtx ctx_(h0);
        GX_CUDA(h0, cudaEventCreateWithFlags(&start, cudaEventDisableTiming));
        GX_CUDA(h0, cudaEventRecord(start, h0->stream));
    }
    for (int r = 0; r < nranks && rc == GVDBX_OK; r++) {
        gvdbx_t* h = ranks[r];
        GxCtx ctx_(h);
        if (r > 0 && cudaStreamWaitEvent(h->stream, start, 0) != cudaSuccess) { rc = gx_fail(h, GVDBX_E_CUDA, "cudaStreamWaitEvent"); break; }
        rc = gvdbx_render_tiles_direct(h, scninfo, shade_mode, chan, outbuf_rank0_d, tile_size, r, nranks);
        if (rc == GVDBX_OK && r > 0) {
            if (cudaEventCreateWithFlags(&done[r], cudaEventDisableTiming) != cudaSuccess || cudaEventRecord(done[r], h->stream) != cudaSuccess)
                rc = gx_fail(h, GVDBX_E_CUDA, "cudaEventRecord");
        }
    }
    {
        GxCtx ctx_(h0);
        for (int r = 1; r < nranks; r++) if (done[r]) { if (rc == GVDBX_OK) cudaStreamWaitEvent(h0->stream, done[r], 0); }
    }
    for (int r = 1; r < nranks; r++) if (done[r]) { GxCtx ctx_(ranks[r]); cudaEventDestroy(done[r]); }
    { GxCtx ctx_(h0); cudaEventDestroy(start); }
    return rc;
}

// ------------------------------------------------------------------------------------------------ peer frame ring
// Protocol: rank 0 owns `nslots` row-major frames followed by one `done` counter per slot; every rank owns one
// `released` flag.  Frame q (1, 2, ...) uses slot (q-1) % nslots for the ((q-1) / nslots + 1)-th time.
//   producer, every rank : [wait released >= q - nslots]  ->  tiles of this rank into the slot  ->  done[slot] += 1
//   consumer, rank 0     : wait done[slot] >= nranks * uses  ->  ... use the frame ...  ->  released = q on every rank
struct GxRingExport {                     // GVDBX_RING_EXPORT_BYTES
    int32_t  rank, pid;
    uint64_t released_ptr, ring_ptr;      // raw pointers: used instead of the handles when two ranks share a process (tests)
    uint8_t  released_handle[GVDBX_IPC_HANDLE_BYTES], ring_handle[GVDBX_IPC_HANDLE_BYTES];
    uint8_t  pad[GVDBX_RING_EXPORT_BYTES - 24 - 2 * GVDBX_IPC_HANDLE_BYTES];
};
static_assert(sizeof(GxRingExport) == GVDBX_RING_EXPORT_BYTES, "ring export blob");

struct gvdbx_ring {
    gvdbx_t* h = nullptr;
    int w = 0, hgt = 0, ts = 0, rank = 0, nranks = 1, nslots = 0;
    size_t frame_bytes = 0;
    uint32_t seq = 0;
    uint64_t released_local = 0, ring_base = 0;
    std::vector<uint64_t> released_all;   // rank 0: every rank's flag
    std::vector<uint64_t> own, opened;
    bool connected = false;
    uint64_t frame_ptr(int slot) const { return ring_base + size_t(slot) * frame_bytes; }
    uint64_t done_ptr(int slot) const { return ring_base + size_t(nslots) * frame_bytes + size_t(slot) * 256; }
};

extern "C" int gvdbx_ring_create(gvdbx_t* h, int width, int height, int tile_size, int rank, int nranks, int nslots,
                                 gvdbx_ring_t** out, void* export_blob)
{
    if (!h || !out || !export_blob || width <= 0 || height <= 0 || tile_size <= 0 || nranks < 1 || nranks > 16 || rank < 0 || rank >= nranks || nslots < 1)
        return GVDBX_E_ARG;
    gvdbx_ring* g = new gvdbx_ring;
    g->h = h; g->w = width; g->hgt = height; g->ts = tile_size; g->rank = rank; g->nranks = nranks; g->nslots = nslots;
    g->frame_bytes = (size_t(width) * height * 4 + 255) / 256 * 256;
    GxRingExport e;
    memset(&e, 0, sizeof e);
    e.rank = rank; e.pid = (int32_t)getpid();
    int rc = gvdbx_peer_alloc(h, 256, &g->released_local, e.released_handle);
    if (rc == GVDBX_OK) {
        g->own.push_back(g->released_local);
        e.released_ptr = g->released_local;
        if (rank == 0) {
            rc = gvdbx_peer_alloc(h, g->frame_bytes * nslots + size_t(256) * nslots, &g->ring_base, e.ring_handle);
            if (rc == GVDBX_OK) { g->own.push_back(g->ring_base); e.ring_ptr = g->ring_base; }
        }
    }
    if (rc != GVDBX_OK) { for (uint64_t p : g->own) gvdbx_peer_free(h, p); delete g; return rc; }
    memcpy(export_blob, &e, sizeof e);
    *out = g;
    return GVDBX_OK;
}

extern "C" int gvdbx_ring_connect(gvdbx_ring_t* g, const void* all_exports)
{
    if (!g || !all_exports) return GVDBX_E_ARG;
    const GxRingExport* ex = (const GxRingExport*)all_exports;
    const int me = (int)getpid();
    auto open = [&](const GxRingExport& e, bool ring, uint64_t* p) -> int {
        if (e.pid == me) { *p = ring ? e.ring_ptr : e.released_ptr; return GVDBX_OK; }      // an IPC handle cannot be opened by its own process
        int rc = gvdbx_peer_open(g->h, ring ? e.ring_handle : e.released_handle, p);
        if (rc == GVDBX_OK) g->opened.push_back(*p);
        return rc;
    };
    for (int r = 0; r < g->nranks; r++) if (ex[r].rank != r) return gx_fail(g->h, GVDBX_E_ARG, "ring exports must be passed in rank order");
    int rc = GVDBX_OK;
    if (g->rank == 0) {
        g->released_all.assign(g->nranks, 0);
        g->released_all[0] = g->released_local;
        for (int r = 1; r < g->nranks && rc == GVDBX_OK; r++) rc = open(ex[r], false, &g->released_all[r]);
    } else {
        rc = open(ex[0], true, &g->ring_base);
    }
    g->connected = (rc == GVDBX_OK);
    return rc;
}

extern "C" int gvdbx_ring_submit(gvdbx_ring_t* g, const void* scninfo, int shade_mode, int chan, uint32_t* seq_out)
{
    if (!g || !scninfo) return GVDBX_E_ARG;
    if (!g->connected) return gx_fail(g->h, GVDBX_E_STATE, "gvdbx_ring_connect first");
    const uint32_t q = ++g->seq;
    const int slot = int((q - 1) % uint32_t(g->nslots));
    if (!g->h->lanes.empty()) gvdbx_lane_select(g->h, int((q - 1) % g->h->lanes.size()));     // consecutive frames on alternating streams
    // the slot's previous frame (q - nslots) must have been consumed before its pixels are overwritten
    const bool wait = q > uint32_t(g->nslots);
    int rc = gvdbx_render_tiles_ring(g->h, scninfo, shade_mode, chan, g->frame_ptr(slot), g->ts, g->rank, g->nranks,
                                     wait ? g->released_local : 0, wait ? q - uint32_t(g->nslots) : 0, g->done_ptr(slot));
    if (seq_out) *seq_out = q;
    return rc;
}

extern "C" int gvdbx_ring_acquire(gvdbx_ring_t* g, uint32_t seq, void* consumer_stream, uint64_t* frame_d)
{
    if (!g || seq == 0) return GVDBX_E_ARG;
    if (g->rank != 0) return gx_fail(g->h, GVDBX_E_STATE, "only rank 0 consumes");
    const int slot = int((seq - 1) % uint32_t(g->nslots));
    const uint32_t uses = (seq - 1) / uint32_t(g->nslots) + 1;
    if (frame_d) *frame_d = g->frame_ptr(slot);
    return gvdbx_stream_wait(g->h, consumer_stream, g->done_ptr(slot), uint32_t(g->nranks) * uses);
}

extern "C" int gvdbx_ring_release(gvdbx_ring_t* g, uint32_t seq, void* consumer_stream)
{
    if (!g || seq == 0) return GVDBX_E_ARG;
    if (g->rank != 0) return gx_fail(g->h, GVDBX_E_STATE, "only rank 0 consumes");
    return gvdbx_stream_signal_many(g->h, consumer_stream, g->released_all.data(), g->nranks, seq);
}

extern "C" int gvdbx_ring_frame(gvdbx_ring_t* g, uint32_t seq, uint64_t* frame_d)
{
    if (!g || !frame_d || seq == 0) return GVDBX_E_ARG;
    *frame_d = g->frame_ptr(int((seq - 1) % uint32_t(g->nslots)));
    return GVDBX_OK;
}

extern "C" int gvdbx_ring_destroy(gvdbx_ring_t* g)
{
    if (!g) return GVDBX_E_ARG;
    const int rc = gvdbx_sync(g->h);            // reports a wait that ran into its timeout (sticky)
    for (uint64_t p : g->opened) gvdbx_peer_close(g->h, p);
    for (uint64_t p : g->own) gvdbx_peer_free(g->h, p);
    delete g;
    return rc;
}

// ------------------------------------------------------------------------------------------------ host frame ring
struct GxHostRingHeader {
    uint32_t magic, width, height, nslots, nranks, band_rows;
    uint64_t frame_bytes, frames_offset;
    alignas(64) std::atomic<uint32_t> consumed;                 // highest frame sequence number the consumer has released
    alignas(64) std::atomic<uint32_t> attached;                 // processes that have mapped the segment
    // followed by nslots * nranks delivery flags, 64 bytes apart: done(slot, rank) = sequence number delivered
};
#define GX_HOSTRING_MAGIC 0x47585248u

struct gvdbx_hostring {
    gvdbx_t* h = nullptr;
    std::string name;
    int rank = 0, nranks = 1, nslots = 0, band_rows = 32, w = 0, hgt = 0, pitch = 0, bands_mine = 0;
    size_t seg_bytes = 0, packed_bytes = 0;
    uint8_t* seg = nullptr;
    GxHostRingHeader* hdr = nullptr;
    uint32_t seq = 0;
    std::vector<void*> packed;          // one device buffer of this rank's bands per slot
    uint32_t* seqvals = nullptr;        // page-locked: the value each slot's delivery flag takes
    bool registered = false;
    uint64_t done_d = 0;                // device address of done(0, 0) in the mapped segment (0: not mapped, flags travel by host-to-host copies)
    volatile uint32_t* done(int slot, int r) const { return (volatile uint32_t*)(seg + sizeof(GxHostRingHeader) + (size_t(slot) * nranks + r) * 64); }
    uint8_t* frame(int slot) const { return seg + hdr->frames_offset + size_t(slot) * hdr->frame_bytes; }
};

extern "C" int gvdbx_hostring_create(gvdbx_t* h, const char* shm_name, int width, int height, int band_rows, int rank, int nranks,
                                     int nslots, gvdbx_hostring_t** out)
{
    if (!h || !shm_name || !out || width <= 0 || height <= 0 || band_rows <= 0 || nranks < 1 || rank < 0 || rank >= nranks || nslots < 1) return GVDBX_E_ARG;
    if (band_rows % h->block_h) return gx_fail(h, GVDBX_E_ARG, "band_rows must be a multiple of the CTA tile height");
    GxCtx ctx_(h);
    gvdbx_hostring* g = new gvdbx_hostring;
    g->h = h; g->name = shm_name; g->rank = rank; g->nranks = nranks; g->nslots = nslots; g->band_rows = band_rows; g->w = width; g->hgt = height;
    g->pitch = (width + h->block_w - 1) / h->block_w * h->block_w;
    const int nbands = (height + band_rows - 1) / band_rows;
    g->bands_mine = (nbands + nranks - 1) / nranks;         // band slots per rank (the last one may be unused)
    const size_t frame_bytes = (size_t(width) * height * 4 + 4095) / 4096 * 4096;
    const size_t frames_offset = (sizeof(GxHostRingHeader) + size_t(nslots) * nranks * 64 + 4095) / 4096 * 4096;
    g->seg_bytes = frames_offset + frame_bytes * nslots;
    auto fail = [&](const std::string& m) { if (g->seg) munmap(g->seg, g->seg_bytes); delete g; return gx_fail(h, GVDBX_E_CUDA, m); };
    int fd = -1;
    if (rank == 0) {
        shm_unlink(shm_name);
        fd = shm_open(shm_name, O_CREAT | O_EXCL | O_RDWR, 0600);
        if (fd < 0 || ftruncate(fd, (off_t)g->seg_bytes) != 0) { if (fd >= 0) close(fd); return fail(std::string("shm_open / ftruncate ") + shm_name); }
    } else {
        for (int tries = 0; tries < 6000; tries++) {        // rank 0 creates the segment; wait for it (up to ~60 s)
            fd = shm_open(shm_name, O_RDWR, 0600);
            struct stat st;
            if (fd >= 0 && fstat(fd, &st) == 0 && (size_t)st.st_size >= g->seg_bytes) break;
            if (fd >= 0) { close(fd); fd = -1; }
            std::this_thread::sleep_for(std::chrono::milliseconds(10));
        }
        if (fd < 0) return fail(std::string("shared segment not found: ") + shm_name);
    }
    g->seg = (uint8_t*)mmap(nullptr, g->seg_bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (g->seg == MAP_FAILED) { g->seg = nullptr; return fail("mmap"); }
    g->hdr = (GxHostRingHeader*)g->seg;
    if (rank == 0) {
        memset(g->seg, 0, frames_offset);
        g->hdr->width = width; g->hdr->height = height; g->hdr->nslots = nslots; g->hdr->nranks = nranks; g->hdr->band_rows = band_rows;
        g->hdr->frame_bytes = frame_bytes; g->hdr->frames_offset = frames_offset;
        std::atomic_thread_fence(std::memory_order_release);
        g->hdr->magic = GX_HOSTRING_MAGIC;
    } else {
        for (int tries = 0; tries < 6000 && ((volatile GxHostRingHeader*)g->hdr)->magic != GX_HOSTRING_MAGIC; tries++)
            std::this_thread::sleep_for(std::chrono::milliseconds(10));
        std::atomic_thread_fence(std::memory_order_acquire);
        if (g->hdr->magic != GX_HOSTRING_MAGIC || (int)g->hdr->width != width || (int)g->hdr->height != height || (int)g->hdr->nslots != nslots ||
            (int)g->hdr->nranks != nranks || (int)g->hdr->band_rows != band_rows)
            return fail("shared segment was created with other parameters");
    }
    g->hdr->attached.fetch_add(1);
    // page-lock the segment in THIS process: copies into it are then true asynchronous DMA over this GPU's own PCIe link
    if (cudaHostRegister(g->seg, g->seg_bytes, cudaHostRegisterPortable | cudaHostRegisterMapped) != cudaSuccess) { cudaGetLastError(); g->hdr->attached.fetch_sub(1); return fail("cudaHostRegister of the shared segment"); }
    g->registered = true;
    {
        void* dp = nullptr;
        if (cudaHostGetDevicePointer(&dp, (void*)g->done(0, 0), 0) == cudaSuccess && dp) g->done_d = (uint64_t)dp;
        else cudaGetLastError();
    }
    g->packed_bytes = size_t(g->bands_mine) * band_rows * g->pitch * 4;
    for (int s = 0; s < nslots; s++) {
        void* p = nullptr;
        if (cudaMalloc(&p, g->packed_bytes) != cudaSuccess) { gvdbx_hostring_destroy(g); return gx_fail(h, GVDBX_E_CUDA, "cudaMalloc of the band buffers"); }
        g->packed.push_back(p);
    }
    if (cudaHostAlloc((void**)&g->seqvals, sizeof(uint32_t) * nslots, cudaHostAllocPortable) != cudaSuccess) { gvdbx_hostring_destroy(g); return gx_fail(h, GVDBX_E_CUDA, "cudaHostAlloc"); }
    *out = g;
    return GVDBX_OK;
}

// every rank: this rank's bands of the next frame -> device band buffer -> (own PCIe link) -> their rows in the shared host frame
// -> delivery flag.  Blocks on the HOST only while the slot's previous frame has not been released by the consumer.
extern "C" int gvdbx_hostring_submit(gvdbx_hostring_t* g, const void* scninfo, int shade_mode, int chan, uint32_t* seq_out)
{
    if (!g || !scninfo) return GVDBX_E_ARG;
    gvdbx_t* h = g->h;
    GxCtx ctx_(h);
    const uint32_t q = ++g->seq;
    const int slot = int((q - 1) % uint32_t(g->nslots));
    if (seq_out) *seq_out = q;
    if (q > uint32_t(g->nslots)) {
        const uint32_t need = q - uint32_t(g->nslots);
        const auto t0 = std::chrono::steady_clock::now();
        while (int32_t(g->hdr->consumed.load(std::memory_order_acquire) - need) < 0) {
            sched_yield();
            if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(30)) return gx_fail(h, GVDBX_E_STATE, "host frame ring: the consumer did not release a slot within 30 s");
        }
    }
    if (!h->lanes.empty()) gvdbx_lane_select(h, int((q - 1) % h->lanes.size()));
    int rc = gvdbx_render_bands(h, scninfo, shade_mode, chan, (uint64_t)g->packed[slot], g->band_rows, g->rank, g->nranks);
    if (rc) return rc;
    const int nbands = (g->hgt + g->band_rows - 1) / g->band_rows;
    uint8_t* frame = g->frame(slot);
    // A full-width band is one contiguous run of bytes in the packed buffer AND in the row-major frame when the row pitch equals
    // the width: all complete bands of this rank then travel as ONE pitched copy (a "row" = one band, destination pitch = nranks
    // bands); only a clipped last band of the frame goes separately.  Otherwise one pitched copy per band.
    const size_t band_bytes = size_t(g->band_rows) * g->w * 4;
    int k0 = 0;
    if (g->pitch == g->w) {
        int nfull = 0;
        while (nfull < g->bands_mine && (nfull * g->nranks + g->rank + 1) * g->band_rows <= g->hgt) nfull++;
        if (nfull > 0)
            GX_CUDA(h, cudaMemcpy2DAsync(frame + size_t(g->rank) * band_bytes, size_t(g->nranks) * band_bytes, g->packed[slot], band_bytes, band_bytes, nfull,
                                         cudaMemcpyDeviceToHost, h->stream));
        k0 = nfull;
    }
    for (int k = k0; k < g->bands_mine; k++) {
        const int b = k * g->nranks + g->rank;
        if (b >= nbands) break;
        const int y0 = b * g->band_rows, rows = std::min(g->band_rows, g->hgt - y0);
        GX_CUDA(h, cudaMemcpy2DAsync(frame + size_t(y0) * g->w * 4, size_t(g->w) * 4, (const uint8_t*)g->packed[slot] + size_t(k) * g->band_rows * g->pitch * 4,
                                     size_t(g->pitch) * 4, size_t(g->w) * 4, rows, cudaMemcpyDeviceToHost, h->stream));
    }
    // delivery flag: a one-thread kernel behind the band copies on the same stream release-stores the sequence number straight
    // into the mapped shared segment (no host thread of the CUDA driver involved, unlike a host-to-host cudaMemcpyAsync: eight
    // ranks polling on sixteen host cores otherwise delay each other's flags)
    if (g->done_d) return gvdbx_stream_signal(h, nullptr, g->done_d + (size_t(slot) * g->nranks + g->rank) * 64, q);
    g->seqvals[slot] = q;
    GX_CUDA(h, cudaMemcpyAsync((void*)g->done(slot, g->rank), &g->seqvals[slot], sizeof(uint32_t), cudaMemcpyHostToHost, h->stream));
    return GVDBX_OK;
}

// consumer (any ONE process, normally rank 0): blocks until every rank has delivered frame `seq`; the frame is row-major
// RGBA8 in the shared segment — the bytes VolumeGVDB::ReadRenderBuf returns
extern "C" int gvdbx_hostring_wait(gvdbx_hostring_t* g, uint32_t seq, const void** frame_host, int timeout_ms)
{
    if (!g || seq == 0) return GVDBX_E_ARG;
    const int slot = int((seq - 1) % uint32_t(g->nslots));
    const auto t0 = std::chrono::steady_clock::now();
    for (int r = 0; r < g->nranks; r++) {
        while (int32_t(*g->done(slot, r) - seq) < 0) {
            sched_yield();
            if (timeout_ms > 0 && std::chrono::steady_clock::now() - t0 > std::chrono::milliseconds(timeout_ms))
                return gx_fail(g->h, GVDBX_E_STATE, "host frame ring: a rank did not deliver its bands in time");
        }
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    if (frame_host) *frame_host = g->frame(slot);
    return GVDBX_OK;
}

extern "C" int gvdbx_hostring_release(gvdbx_hostring_t* g, uint32_t seq)
{
    if (!g || seq == 0) return GVDBX_E_ARG;
    g->hdr->consumed.store(seq, std::memory_order_release);
    return GVDBX_OK;
}

extern "C" int gvdbx_hostring_destroy(gvdbx_hostring_t* g)
{
    if (!g) return GVDBX_E_ARG;
    {
        GxCtx ctx_(g->h);
        for (cudaStream_t s : g->h->lanes) cudaStreamSynchronize(s);
        cudaStreamSynchronize(g->h->base_stream);
        for (void* p : g->packed) cudaFree(p);
        if (g->seqvals) cudaFreeHost(g->seqvals);
        if (g->registered) cudaHostUnregister(g->seg);
    }
    if (g->seg) {
        const bool last = g->hdr->attached.fetch_sub(1) == 1;
        munmap(g->seg, g->seg_bytes);
        if (last || g->rank == 0) shm_unlink(g->name.c_str());
    }
    delete g;
    return GVDBX_OK;
}
