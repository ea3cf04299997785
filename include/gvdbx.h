/* gvdbx.h — C ABI of libgvdbx.so: B200-native (sm_100a) ray-cast render path, drop-in for the device side of
 * VolumeGVDB::Render(shade mode, channel, render buffer) of NVIDIA/gvdb-voxels.
 *
 * Boundary replaced (reference, relative to /root/reference/source/gvdb_library):
 *   src/gvdb_volume_gvdb.cpp:4336-4381  VolumeGVDB::Render      -> gvdbx_render / gvdbx_render_tiles
 *   src/gvdb_volume_gvdb.cpp:4254-4306  PrepareRender (ScnInfo) -> the 416-byte block passed to gvdbx_render
 *   src/gvdb_volume_gvdb.cpp:3946-3989  PrepareVDB (VDBInfo)    -> gvdbx_import_topology (1232-byte block)
 *   src/gvdb_volume_gvdb.cpp:720-803    SetupAtlasAccess        -> gvdbx_import_atlas_array / _host
 *   src/gvdb_volume_gvdb.cpp:4892-4898  CommitTransferFunc      -> gvdbx_set_transfer
 *   src/gvdb_volume_gvdb.cpp:4241-4251  ReadRenderBuf           -> gvdbx_read_buffer
 *   kernels/cuda_gvdb_module.cu:60-181  gvdbRayDeep / gvdbRaySurfaceVoxel / gvdbRaySurfaceTrilinear / gvdbRayLevelSet
 *
 * Conventions: plain pointers and sizes only; every function returns 0 or a negative GVDBX_E_* code and never
 * exits the process (the reference prints and exit(-1)s: src/gvdb_types.cpp:86-90); gvdbx_last_error() gives text.
 * All work is enqueued on the stream given at creation (NULL = legacy default stream, what the reference launches
 * on: gvdb_volume_gvdb.cpp:4375) with no hidden synchronisation unless stated.  There is NO CPU fallback: without
 * a CUDA device every entry point that touches the device fails with GVDBX_E_CUDA.
 */
#ifndef GVDBX_H
#define GVDBX_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GVDBX_VDBINFO_BYTES 1232   /* sizeof(VDBInfo), kernels/cuda_gvdb_nodes.cuh:42-67 */
#define GVDBX_SCNINFO_BYTES 416    /* sizeof(ScnInfo), kernels/cuda_gvdb_scene.cuh:35-64 */
#define GVDBX_NODE_BYTES    64     /* sizeof(VDBNode), kernels/cuda_gvdb_nodes.cuh:24-35 */
#define GVDBX_TRANSFER_ENTRIES 16384 /* src/gvdb_scene.cpp:70 */

/* shade modes: numeric values of the reference's SHADE_* (src/gvdb_types.h:73-81) */
#define GVDBX_SHADE_VOXEL      0
#define GVDBX_SHADE_SECTION2D  1   /* kernels/cuda_gvdb_module.cu:272-298 (texture sampler only) */
#define GVDBX_SHADE_SECTION3D  2   /* kernels/cuda_gvdb_module.cu:225-269 (texture sampler only) */
#define GVDBX_SHADE_EMPTYSKIP  3   /* kernels/cuda_gvdb_module.cu:184-207 */
#define GVDBX_SHADE_TRILINEAR  4
#define GVDBX_SHADE_TRICUBIC   5   /* kernels/cuda_gvdb_module.cu:122-139 (texture sampler only) */
#define GVDBX_SHADE_LEVELSET   6
#define GVDBX_SHADE_VOLUME     7
#define GVDBX_SHADE_OFF        100

/* error codes */
#define GVDBX_OK            0
#define GVDBX_E_ARG        -1   /* bad argument / unsupported value */
#define GVDBX_E_CUDA       -2   /* CUDA runtime error (see gvdbx_last_error) */
#define GVDBX_E_STATE      -3   /* call order (e.g. render before import) */
#define GVDBX_E_UNSUPPORTED -4  /* feature of the reference outside this path (e.g. colour channel) */

/* options for gvdbx_set_option */
#define GVDBX_OPT_SAMPLER   1   /* 0 = hardware texture fetch (bit-exact vs reference), 1 = linear brick-major loads */
#define GVDBX_OPT_BLOCK_W   2   /* CTA pixel tile width  (default 8)  */
#define GVDBX_OPT_BLOCK_H   3   /* CTA pixel tile height (default 8); width x height: a multiple of 32, at most 128 threads */
#define GVDBX_OPT_COUNTERS  4   /* accumulate work counters during render (slower; for roofline accounting): 1 = the work of the ALGORITHM as the
                                   reference does it (no brick culling: SURVEY.md 8d units), 2 = the work the production kernel does (culling on) */
#define GVDBX_OPT_CULL      6   /* 1 (default) = skip bricks whose value range cannot satisfy the mode's acceptance test (exact) */
#define GVDBX_OPT_SPP       7   /* rays per pixel (default 1 = the reference's pixel-centre ray).  n > 1: samples on a g x g sub-pixel grid,
                                   g = ceil(sqrt(n)), sample s at ((s % g) + .5) / g, ((s / g) + .5) / g; float colours summed in sample
                                   order, scaled by 1/n, packed once (BASELINE.json config 5: 4 spp) */
#define GVDBX_OPT_DEEP_SHADOW 8 /* 1 = SHADE_VOLUME casts one shadow march (rayShadowBrick, kernels/cuda_gvdb_raycast.cuh:445-463) from
                                   the first sample towards the light and darkens the accumulated colour (BASELINE.json config 4) */
#define GVDBX_OPT_STREAM_MEMOPS 9 /* 1 = gvdbx_stream_wait uses cuStreamWaitValue32 (front-end wait, unbounded) instead of the bounded polling kernel */
#define GVDBX_OPT_VOXEL_MASK 10 /* 1 (default) = SHADE_VOXEL tests per-brick occupancy bits (value > THRESH, rebuilt when THRESH or the atlas
                                   changes) instead of one point fetch per voxel step; 0 = fetch (A/B) */
#define GVDBX_OPT_TRAVERSAL 5   /* 0 = default (four-samples-per-round brick marchers; deep modes and SHADE_TRILINEAR: brick-queue traversal),
                                   1 = reference-shaped loops, one sample at a time (A/B), 2 = vote-converged two-phase packet traversal (A/B),
                                   3 = brick queue also for SHADE_LEVELSET, 4 = four-sample rounds WITHOUT the brick queue (A/B) */

typedef struct gvdbx_ctx gvdbx_t;

/* work counters of the last counted render (SURVEY.md §8d units) */
typedef struct gvdbx_counters {
    uint64_t s_tri;    /* trilinear samples (brick march + gradient taps)            32 B each */
    uint64_t s_pt;     /* point samples (SHADE_VOXEL brick DDA steps)                 4 B each */
    uint64_t n_dda;    /* DDA steps at levels >= 1 (child-table reads)                8 B each */
    uint64_t n_desc;   /* node-record reads on descent / brick entry                 64 B each */
    uint64_t s_lut;    /* transfer-function reads (deep)                             16 B each (not in HBM figure) */
    uint64_t rays;     /* primary + shadow rays cast */
} gvdbx_counters;

int  gvdbx_create(gvdbx_t** h, int cuda_device, void* cuda_stream);
int  gvdbx_destroy(gvdbx_t* h);
const char* gvdbx_last_error(const gvdbx_t* h);
int  gvdbx_set_option(gvdbx_t* h, int option, int value);

/* Topology.  `vdbinfo` is a HOST copy of the reference's 1232-byte VDBInfo as PrepareVDB fills it
 * (VolumeGVDB::getVDBInfo(), gvdb_volume_gvdb.h:519): nodelist[] / childlist[] hold DEVICE pointers to the caller's
 * pool-0 node records and pool-1 child lists.  The library builds its own compact traversal tables from them (device
 * side) and keeps no reference to the caller's pools afterwards.  Call again whenever the reference would set
 * mVDBInfo.update (FinishTopology / UpdateAtlas / SetEpsilon). */
int  gvdbx_import_topology(gvdbx_t* h, const void* vdbinfo);
/* Same, but pool contents come from HOST memory (e.g. Allocator::getPoolCPU or a VBX file): pool0[l] / pool1[l] point
 * at nodecnt[l]*nodewid[l] and <lists>*childwid[l] bytes; pool1_bytes[l] gives the size of each child-list pool. */
int  gvdbx_import_topology_host(gvdbx_t* h, const void* vdbinfo, const void* const* pool0, const void* const* pool1,
                                const uint64_t* pool1_bytes);

/* Brick atlas of channel `chan` (T_FLOAT only).  `cuarray` is the reference's 3-D CUarray (DataPtr::garray,
 * gvdb_allocator.cpp:392-408); it is sampled in place by the texture path and re-laid out brick-major for the linear
 * path.  Call again after the atlas content changed (there is no dirty notification in the reference). */
int  gvdbx_import_atlas_array(gvdbx_t* h, int chan, void* cuarray, int res_x, int res_y, int res_z);
/* Same from a HOST image of the atlas, x fastest (the layout of Allocator::AtlasCommitFromCPU, gvdb_allocator.cpp:797). */
int  gvdbx_import_atlas_host(gvdbx_t* h, int chan, const float* texels, int res_x, int res_y, int res_z);

/* Same from a DEVICE image (x fastest), e.g. one received by a broadcast from the rank that built the volume. */
int  gvdbx_import_atlas_device(gvdbx_t* h, int chan, uint64_t texels_d, int res_x, int res_y, int res_z);

/* VolumeGVDB::UpdateApron(chan, boundval) (src/gvdb_volume_gvdb.cpp:4418-4453, kernels/cuda_gvdb_operators.cuh:72-126)
 * on the imported atlas: every apron texel takes the value of the voxel at its index-space position in whichever brick
 * contains it, else `boundval`.  Writes the 3-D array (the caller's, when imported with gvdbx_import_atlas_array —
 * exactly what the reference's kernel does) and keeps the library's brick-major copy and value ranges coherent. */
int  gvdbx_update_apron(gvdbx_t* h, int chan, float boundval);
/* VolumeGVDB::UpdateApronFaces(chan) (src/gvdb_volume_gvdb.cpp:4461-4496, kernels/cuda_gvdb_operators.cuh:27-61): the cheap
 * variant — face-adjacent bricks swap their boundary voxel layers into each other's face aprons (edge / corner texels and faces
 * without a neighbour stay as they are).  Neighbours are found by point query on the imported tree (the reference:
 * UpdateNeighbors table).  Same coherence guarantees as gvdbx_update_apron. */
int  gvdbx_update_apron_faces(gvdbx_t* h, int chan);
/* Read the atlas array back into a host image, x fastest (Allocator::AtlasRetrieveSlice for every slice). */
int  gvdbx_export_atlas_host(gvdbx_t* h, int chan, float* texels, int res_x, int res_y, int res_z);

/* Colour channel (VolumeGVDB::SetColorChannel, src/gvdb_volume_gvdb.cpp:686-690; getColorF, kernels/cuda_gvdb_raycast.cuh:
 * 200-209): a uchar4 atlas with the brick-slot layout of channel 0, fetched at the hit voxel (surface modes) / at every
 * sample (deep).  It is used exactly when the imported VDBInfo has clr_chan set; rendering such a volume without a colour
 * atlas is GVDBX_E_STATE.  `cuarray` = the reference's 3-D CUarray of that channel (sampled in place); `filter` = the
 * channel's filter mode as given to AddChannel: 0 = F_POINT (gPointFusion).  1 = F_LINEAR, AddChannel's default, is
 * passed on to CUDA unchanged and rejected there for an integer-read texture (cudaErrorInvalidFilterSetting ->
 * GVDBX_E_CUDA) — the reference's own cuTexObjectCreate fails the same way for such a channel. */
int  gvdbx_import_color_array(gvdbx_t* h, void* cuarray, int filter);
int  gvdbx_import_color_host(gvdbx_t* h, const void* rgba8_texels, int res_x, int res_y, int res_z, int filter);
int  gvdbx_clear_color(gvdbx_t* h);

/* Transfer function: 16384 float4 (Scene::getTransferFunc()).  Alternatively leave unset and pass a device pointer in
 * ScnInfo.transfer exactly like the reference does. */
int  gvdbx_set_transfer(gvdbx_t* h, const float* rgba_host);

/* Render one frame (or a sub-rectangle) into a caller-owned DEVICE buffer of width*height RGBA8 pixels, row-major,
 * y = 0 the top-left corner ray — the bytes VolumeGVDB::ReadRenderBuf returns.  `scninfo` is a HOST copy of the
 * 416-byte ScnInfo that PrepareRender fills (VolumeGVDB::getScnInfo()).  tile_w/tile_h <= 0 means the full frame.
 * Asynchronous on the context's stream. */
int  gvdbx_render(gvdbx_t* h, const void* scninfo, int shade_mode, int chan, uint64_t outbuf_d,
                  int tile_x0, int tile_y0, int tile_w, int tile_h);

/* Image-space tiling for multi-GPU: the frame is cut into tile_size x tile_size tiles, numbered row-major; this call
 * renders the tiles with (tile_id % nranks) == rank into `packed_d`, tile after tile (each tile_size*tile_size RGBA8,
 * edge tiles padded), i.e. ceil(ntiles/nranks) tile slots per rank.  gvdbx_assemble_tiles scatters the gathered
 * [nranks][slots][tile] buffer back into a row-major frame. */
int  gvdbx_render_tiles(gvdbx_t* h, const void* scninfo, int shade_mode, int chan, uint64_t packed_d,
                        int tile_size, int rank, int nranks);
/* Direct variant: the same tile list, but each pixel is stored at its place in the row-major frame `frame_d`
 * (width*height RGBA8).  `frame_d` may be a buffer of ANOTHER GPU opened with gvdbx_peer_open: the stores then go over
 * NVLink from inside the render kernel — render and "gather" are one kernel, nothing is packed or assembled. */
int  gvdbx_render_tiles_direct(gvdbx_t* h, const void* scninfo, int shade_mode, int chan, uint64_t frame_d,
                               int tile_size, int rank, int nranks);
/* The per-frame call of the peer frame ring (multigpu.py::PeerFrameRing): optional back-pressure wait, this rank's
 * tiles into frame_d, then *done_flag_d += 1 — three stream-ordered operations on the current stream / lane. */
int  gvdbx_render_tiles_ring(gvdbx_t* h, const void* scninfo, int shade_mode, int chan, uint64_t frame_d,
                             int tile_size, int rank, int nranks, uint64_t wait_flag_d, uint32_t wait_value,
                             uint64_t done_flag_d);
int  gvdbx_tiles_per_rank(int width, int height, int tile_size, int nranks);
int  gvdbx_assemble_tiles(gvdbx_t* h, uint64_t gathered_d, uint64_t frame_d, int width, int height, int tile_size, int nranks);

/* VolumeGVDB::RenderKernel plugin point (src/gvdb_volume_gvdb.cpp:4309-4333): user kernels built against
 * gvdb-voxels_b200/csrc/gvdbx_plugin.cuh take the frame's parameter block (GxParams, by value, __grid_constant__) instead
 * of the reference's (VDBInfo*, chan, outBuf).  This call fills it for `scninfo` (shade_mode selects the per-frame tables
 * that are prepared: occupancy bits for SHADE_VOXEL, derived transfer table for SHADE_VOLUME); params_bytes must equal
 * sizeof(GxParams) of the headers the kernel was built with.  Launch with GVDBX_KERNEL_SMEM(threads) = 36 bytes of dynamic shared memory per thread (the traversal stack). */
int  gvdbx_kernel_params(gvdbx_t* h, const void* scninfo, int shade_mode, int chan, uint64_t outbuf_d, void* params_out,
                         size_t params_bytes);

/* Debug / parity outputs: same traversal, additionally writes 48 B per pixel into dbg_d:
 *   float4 {hit.x, hit.y, hit.z, t_hit}   float4 {norm.x, norm.y, norm.z, as_float(leaf id)}   int4 {voxel.x, voxel.y, voxel.z, iterations}
 * (deep mode: float4 raw colour before compositing, float4 {hit.x, hit.y, hit.z, 0}, int4 0). */
int  gvdbx_render_debug(gvdbx_t* h, const void* scninfo, int shade_mode, int chan, uint64_t outbuf_d, uint64_t dbg_d);

/* VolumeGVDB::Raytrace (src/gvdb_volume_gvdb.cpp:4384-4408, kernel gvdbRaytrace kernels/cuda_gvdb_module.cu:211-222):
 * traces `num_rays` explicit rays given as 64-byte ScnRay records (hit@0, normal@12, orig@24, dir@36, clr@48, pnode@52,
 * pndx@56; src/gvdb_volume_gvdb.h:300-308) in DEVICE memory with the trilinear surface brick function; writes hit
 * (pulled back by `bias` along the ray) and normal in place, hit = (NOHIT,NOHIT,NOHIT) on a miss. */
int  gvdbx_raytrace(gvdbx_t* h, const void* scninfo, int chan, uint64_t rays_d, int num_rays, float bias);

/* Full-width bands of `band_rows` rows (band b of the frame belongs to rank b % nranks), packed band after band with a row
 * pitch of `width` rounded up to the CTA tile width: each band is one contiguous block of rows (gvdbx_hostring_*). */
int  gvdbx_render_bands(gvdbx_t* h, const void* scninfo, int shade_mode, int chan, uint64_t packed_d, int band_rows, int rank, int nranks);

/* ---- multi-GPU (the reference has none: one VolumeGVDB per device, src/gvdb_volume_gvdb.h:325) ------------------------------
 * One frame is partitioned in image space, the volume is replicated on every GPU; output bytes are a pure function of the pixel,
 * hence identical for 1/2/4/8 GPUs.
 *
 * (a) one process, several contexts — SURVEY.md 8b's gvdbx_render_multi: ranks[r] (each created on its own device, each with
 * the volume imported) renders the tile_size^2 tiles r, r + nranks, ... straight into `outbuf_rank0_d`, a width*height RGBA8
 * buffer on ranks[0]'s device (peer access over NVLink is enabled on first use).  Stream-ordered: ranks[0]'s stream continues
 * when every context has finished; the other contexts' streams start when ranks[0]'s stream reaches this call. */
int  gvdbx_render_multi(gvdbx_t* const* ranks, int nranks, const void* scninfo, int shade_mode, int chan, uint64_t outbuf_rank0_d,
                        int tile_size);

/* (b) one process PER GPU, frames wanted on rank 0's DEVICE: the peer frame ring.  Rank 0 owns `nslots` frames; the other ranks
 * map them with CUDA IPC and their render kernels store tiles there directly.  Bootstrap: every rank calls _create, the
 * GVDBX_RING_EXPORT_BYTES blobs are exchanged with whatever the host application has (MPI, sockets, torch.distributed) and passed,
 * in rank order, to _connect.  Per frame every rank calls _submit once (frames are numbered 1, 2, ...; consecutive frames
 * alternate between the context's frame lanes); rank 0 additionally _acquire (makes `consumer_stream` wait until every rank has
 * delivered the frame), consumes on that stream, and _release (hands the slot back).  All waits / signals are stream-ordered
 * device operations; a wait that runs into its ~20 s timeout raises a sticky error that gvdbx_sync / _destroy report. */
#define GVDBX_RING_EXPORT_BYTES 160
typedef struct gvdbx_ring gvdbx_ring_t;
int  gvdbx_ring_create(gvdbx_t* h, int width, int height, int tile_size, int rank, int nranks, int nslots, gvdbx_ring_t** ring, void* export_blob);
int  gvdbx_ring_connect(gvdbx_ring_t* ring, const void* all_exports);
int  gvdbx_ring_submit(gvdbx_ring_t* ring, const void* scninfo, int shade_mode, int chan, uint32_t* frame_seq);
int  gvdbx_ring_acquire(gvdbx_ring_t* ring, uint32_t frame_seq, void* consumer_stream, uint64_t* frame_d);
int  gvdbx_ring_release(gvdbx_ring_t* ring, uint32_t frame_seq, void* consumer_stream);
int  gvdbx_ring_frame(gvdbx_ring_t* ring, uint32_t frame_seq, uint64_t* frame_d);
int  gvdbx_ring_destroy(gvdbx_ring_t* ring);

/* (c) one process per GPU, frames wanted on the HOST (ReadRenderBuf of a multi-GPU render): a ring of `nslots` row-major frames
 * in the POSIX shared-memory segment `shm_name` (rank 0 creates it, every process maps and page-locks it).  Every rank renders
 * full-width bands of `band_rows` rows (band b belongs to rank b % nranks) and copies ITS bands over ITS OWN PCIe link to their
 * rows of the frame — N links instead of funnelling every frame through rank 0's.  _submit: every rank, once per frame (blocks on
 * the host only while the slot's previous frame is unreleased); _wait: the consumer (one process) blocks until all ranks have
 * delivered frame `frame_seq` and gets the host pointer; _release returns the slot. */
typedef struct gvdbx_hostring gvdbx_hostring_t;
int  gvdbx_hostring_create(gvdbx_t* h, const char* shm_name, int width, int height, int band_rows, int rank, int nranks, int nslots,
                           gvdbx_hostring_t** ring);
int  gvdbx_hostring_submit(gvdbx_hostring_t* ring, const void* scninfo, int shade_mode, int chan, uint32_t* frame_seq);
int  gvdbx_hostring_wait(gvdbx_hostring_t* ring, uint32_t frame_seq, const void** frame_host, int timeout_ms);
int  gvdbx_hostring_release(gvdbx_hostring_t* ring, uint32_t frame_seq);
int  gvdbx_hostring_destroy(gvdbx_hostring_t* ring);

/* Peer memory for one-process-per-GPU rendering into a frame owned by one rank (CUDA IPC; both GPUs in one NVLink /
 * NVSwitch domain).  gvdbx_peer_alloc: zero-filled device buffer + 64-byte handle to ship to the other processes;
 * gvdbx_peer_open: map another process's buffer (peer access is enabled on first use). */
#define GVDBX_IPC_HANDLE_BYTES 64
int  gvdbx_peer_alloc(gvdbx_t* h, size_t bytes, uint64_t* dptr, void* handle64);
int  gvdbx_peer_free(gvdbx_t* h, uint64_t dptr);
int  gvdbx_peer_open(gvdbx_t* h, const void* handle64, uint64_t* dptr);
int  gvdbx_peer_close(gvdbx_t* h, uint64_t dptr);
/* Stream-ordered 32-bit sequence flags (cuda_stream NULL = the context's stream).  signal: after everything enqueued
 * before it has completed, store `value` to *flag_d (local or peer memory) with system-scope release.  wait: hold back
 * everything enqueued after it until *flag_d >= value (flag in LOCAL memory; 8 bytes: [0] sequence, [1] timeout mark:
 * the default polling kernel gives up after ~20 s and stores 0xDEAD there).  When waiter and signaller are streams of
 * the same process, enqueue the signal before the wait (streams may share a hardware queue). */
int  gvdbx_stream_signal(gvdbx_t* h, void* cuda_stream, uint64_t flag_d, uint32_t value);
int  gvdbx_stream_signal_add(gvdbx_t* h, void* cuda_stream, uint64_t flag_d, uint32_t inc);      /* *flag_d += inc (atomic, release) */
int  gvdbx_stream_signal_many(gvdbx_t* h, void* cuda_stream, const uint64_t* flags_d, int n, uint32_t value); /* n <= 16 flags, one launch */
int  gvdbx_stream_wait(gvdbx_t* h, void* cuda_stream, uint64_t flag_d, uint32_t value);
/* Switch the stream all later calls of this context enqueue on (e.g. to alternate frames between two streams). */
int  gvdbx_set_stream(gvdbx_t* h, void* cuda_stream);

/* Frame lanes.  A frame's kernel ends with a tail of a few long rays during which most SMs idle; consecutive frames
 * rendered on alternating streams overlap that tail with the next frame's start (1080p: +18 %; a rank of an 8-GPU run
 * renders 1/8 of the rays and gains 2x).  gvdbx_lanes creates n internal streams (0 = destroy); gvdbx_lane_select makes
 * later calls enqueue on lane (lane % n), -1 = back on the creation stream; _fork: every lane waits for what the creation
 * stream holds so far; _join: the creation stream waits for all lanes.  All stream-ordered, no host synchronisation.
 * The caller provides one output buffer per lane. */
int   gvdbx_lanes(gvdbx_t* h, int n);
int   gvdbx_lane_select(gvdbx_t* h, int lane);
void* gvdbx_lane_stream(gvdbx_t* h, int lane);
int   gvdbx_lanes_fork(gvdbx_t* h);
int   gvdbx_lanes_join(gvdbx_t* h);

/* ReadRenderBuf: device -> host copy of `bytes`, synchronises the stream. */
int  gvdbx_read_buffer(gvdbx_t* h, uint64_t buf_d, void* host, size_t bytes);
int  gvdbx_read_buffer_async(gvdbx_t* h, uint64_t buf_d, void* host, size_t bytes);   /* no synchronisation; pair with gvdbx_sync */
int  gvdbx_sync(gvdbx_t* h);
int  gvdbx_get_counters(gvdbx_t* h, gvdbx_counters* out);

/* Calibration aid: samples channel `chan` at n atlas-space points (xyz triples, device pointer) with the hardware
 * texture path and with the linear-load emulation; writes n floats each. */
int  gvdbx_sample_points(gvdbx_t* h, int chan, uint64_t xyz_d, int n, uint64_t out_tex_d, uint64_t out_lin_d);

/* Measurement aid: fp32 trilinear samples per second (in 1e9) the texture units of this GPU deliver on L1-resident bricks of the
 * imported atlas with nothing else in the way (fetches only, 8 in flight per thread) when the 8x4 lanes of a warp sample
 * `lane_spacing` voxels apart — the unit's rate depends on how many distinct texel quads one warp request touches: ~0.2 = the
 * best case, the voxels-per-pixel of a camera = what a ray packet of that camera can get.  Roofline denominator of the
 * TEX-bound deep mode.  Synchronises. */
/* Strict drop-in sequence, overlapped inside the library: VolumeGVDB::Render is asynchronous and ReadRenderBuf a synchronous copy
 * into pageable memory (src/gvdb_volume_gvdb.cpp:4241-4251), kernel and copy in series.  gvdbx_render_banded renders the frame as
 * `nbands` (1..16) horizontal bands on two alternating internal streams, an event behind each band; gvdbx_read_banded then copies
 * band after band as they finish — all but the last band's copy hides behind the rendering below it.  Same bytes, same buffer;
 * later work on the context's stream is ordered behind every band.  gvdbx_read_banded of a buffer that was not rendered in bands
 * is gvdbx_read_buffer. */
int  gvdbx_render_banded(gvdbx_t* h, const void* scninfo, int shade_mode, int chan, uint64_t outbuf_d, int nbands);
int  gvdbx_read_banded(gvdbx_t* h, uint64_t buf_d, void* host, size_t bytes);
int  gvdbx_measure_tex_peak(gvdbx_t* h, float lane_spacing, double* gsamples_per_s);
/* Fetch + filter microbenchmarks of the four ways of reading a brick on the imported atlas (csrc/gvdbx_microbench.cuh), Gsamples/s:
 * [0] texture unit on the caller's array, [1] brick-major copy + scalar read-only loads, [2] x-pair layout + 8-byte loads,
 * [3] brick-major blocks staged into shared memory by TMA (cp.async.bulk + mbarrier, two stages per warp).  [1]-[3] run the
 * software model of the unit's filter.  The numbers behind the texture-versus-linear-load decision. */
int  gvdbx_measure_sampler_ab(gvdbx_t* h, float lane_spacing, double* gsamples_per_s4);
/* Gsamples/s of the deep marcher's inner loop alone — four fetches, four transfer indices, four 16-byte table gathers, four colour
 * updates per round, no traversal, every lane busy, L1-resident bricks: the ceiling of the sample loop; needs gvdbx_set_transfer. */
int  gvdbx_measure_deep_loop_peak(gvdbx_t* h, float lane_spacing, int table_through_texture /* A/B: 1 = float4 texture fetches instead of 16-byte loads */,
                                  double* gsamples_per_s);

#ifdef __cplusplus
}
#endif
#endif /* GVDBX_H */
