// ds_runtime.cu — host runtime + C ABI of libdronestitch_cuda (include/dronestitch.h).
//
// Product build: nvcc -gencode arch=compute_100a,code=sm_100a. There is no CPU fallback: without a
// usable CUDA device every computing entry point returns DS_ERR_NO_DEVICE.
// The same file compiles with -DDS_EMU under g++ into the tests-only emulator (tests/emu), where
// "device" memory is host memory and a launch is a loop over blocks.
#include "../../include/dronestitch.h"

#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>

#include <unistd.h>

#include <algorithm>
#include <map>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "ds_geometry.h"
#include "ds_kernels.h"
#if DS_CUDA
#include <cuda.h>   // CUtensorMap + cuTensorMapEncodeTiled prototype (resolved at run time, no libcuda link)
#endif

#ifndef DS_VERSION_STRING
#define DS_VERSION_STRING "0.1.0"
#endif

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

// ------------------------------------------------------------------ backend
#if DS_CUDA
typedef cudaStream_t stream_t;
#define DS_CK(call)                                                                          \
    do {                                                                                     \
        cudaError_t e_ = (call);                                                             \
        if (e_ != cudaSuccess)                                                               \
            return fail(e_ == cudaErrorMemoryAllocation ? DS_ERR_OOM : DS_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); \
    } while (0)

int dev_alloc(void** p, size_t bytes) {
    *p = nullptr;
    if (bytes == 0) return DS_OK;
    DS_CK(cudaMalloc(p, bytes));
    return DS_OK;
}
void dev_free(void* p) { if (p) cudaFree(p); }
int h2d(void* d, const void* h, size_t n, stream_t s) { DS_CK(cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, s)); return DS_OK; }
int h2d_2d(void* d, size_t dp, const void* h, size_t hp, size_t wbytes, size_t rows, stream_t s) {
    DS_CK(cudaMemcpy2DAsync(d, dp, h, hp, wbytes, rows, cudaMemcpyHostToDevice, s));
    return DS_OK;
}
int d2h_2d(void* h, size_t hp, const void* d, size_t dp, size_t wbytes, size_t rows, stream_t s) {
    DS_CK(cudaMemcpy2DAsync(h, hp, d, dp, wbytes, rows, cudaMemcpyDeviceToHost, s));
    return DS_OK;
}
int d2h(void* h, const void* d, size_t n, stream_t s) { DS_CK(cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, s)); return DS_OK; }
int stream_sync(stream_t s) { DS_CK(cudaStreamSynchronize(s)); return DS_OK; }
// events order the upload, compute and download streams against each other
typedef cudaEvent_t event_t;
int ev_make(event_t* e) { if (!*e) DS_CK(cudaEventCreateWithFlags(e, cudaEventDisableTiming)); return DS_OK; }
void ev_drop(event_t e) { if (e) cudaEventDestroy(e); }
int ev_record(event_t e, stream_t s) { DS_CK(cudaEventRecord(e, s)); return DS_OK; }
int ev_wait(stream_t s, event_t e) { DS_CK(cudaStreamWaitEvent(s, e, 0)); return DS_OK; }
int ev_sync(event_t e) { DS_CK(cudaEventSynchronize(e)); return DS_OK; }
// The stream waits until the 32-bit word at `addr` (device memory of this GPU) is >= value (cyclic compare): a
// stream memory operation, no kernel spins. Resolved from the driver at run time like the tensor-map encoder.
typedef int (*WaitValue32Fn)(cudaStream_t, unsigned long long, unsigned int, unsigned int);
int stream_wait_value(stream_t s, int* addr, int value) {
    static WaitValue32Fn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &f, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (WaitValue32Fn)f;
        else
            cudaGetLastError();
    }
    if (!fn) return fail(DS_ERR_P2P_UNAVAILABLE, "cuStreamWaitValue32 is not available");
    const int r = fn(s, (unsigned long long)(uintptr_t)addr, (unsigned int)value, 0u /* CU_STREAM_WAIT_VALUE_GEQ */);
    if (r != 0) return fail(DS_ERR_CUDA, "cuStreamWaitValue32 failed (%d)", r);
    return DS_OK;
}
int dev_zero(void* p, size_t n) { DS_CK(cudaMemset(p, 0, n)); return DS_OK; }
// pinned host memory the device can read directly (zero-copy source of the launch metadata)
int pinned_alloc(void** host, void** dev_view, size_t bytes) {
    DS_CK(cudaHostAlloc(host, bytes, cudaHostAllocMapped));
    DS_CK(cudaHostGetDevicePointer(dev_view, *host, 0));
    return DS_OK;
}
void pinned_free(void* host) { if (host) cudaFreeHost(host); }

template <class Body, int NT, class P>
int launch(const P& p, long long blocks, stream_t s, int smem_bytes) {
    if (blocks <= 0) return DS_OK;
    if (blocks > 2147483647LL) return fail(DS_ERR_BAD_ARG, "grid too large (%lld blocks)", blocks);
    static int configured_dev = -1;
    int dev = 0;
    cudaGetDevice(&dev);
    if (smem_bytes > 48 * 1024 && configured_dev != dev) {
        DS_CK(cudaFuncSetAttribute(KernelOf<Body, NT>::fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
        configured_dev = dev;
    }
    // programmatic dependent launch (see ds_grid_dependency_sync): DS_PDL=0 falls back to plain stream order
    static int pdl = -1;
    if (pdl < 0) { const char* e = getenv("DS_PDL"); pdl = (e && atoi(e) == 0) ? 0 : 1; }
    if (pdl) {
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3((unsigned)blocks); cfg.blockDim = dim3(NT); cfg.dynamicSmemBytes = (size_t)smem_bytes; cfg.stream = s;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        DS_CK(cudaLaunchKernelEx(&cfg, KernelOf<Body, NT>::fn, p));
    } else {
        KernelOf<Body, NT>::fn<<<(unsigned)blocks, NT, smem_bytes, s>>>(p);
    }
    DS_CK(cudaGetLastError());
    return DS_OK;
}
#else
typedef int stream_t;
int dev_alloc(void** p, size_t bytes) {
    *p = bytes ? malloc(bytes) : nullptr;
    if (bytes && !*p) return fail(DS_ERR_OOM, "malloc(%zu) failed", bytes);
    return DS_OK;
}
void dev_free(void* p) { free(p); }
int h2d(void* d, const void* h, size_t n, stream_t) { memcpy(d, h, n); return DS_OK; }
int h2d_2d(void* d, size_t dp, const void* h, size_t hp, size_t wbytes, size_t rows, stream_t) {
    for (size_t r = 0; r < rows; r++) memcpy((char*)d + r * dp, (const char*)h + r * hp, wbytes);
    return DS_OK;
}
int d2h_2d(void* h, size_t hp, const void* d, size_t dp, size_t wbytes, size_t rows, stream_t) {
    for (size_t r = 0; r < rows; r++) memcpy((char*)h + r * hp, (const char*)d + r * dp, wbytes);
    return DS_OK;
}
int d2h(void* h, const void* d, size_t n, stream_t) { memcpy(h, d, n); return DS_OK; }
int stream_sync(stream_t) { return DS_OK; }
typedef int event_t;
int ev_make(event_t* e) { *e = 1; return DS_OK; }
void ev_drop(event_t) {}
int ev_record(event_t, stream_t) { return DS_OK; }
int ev_wait(stream_t, event_t) { return DS_OK; }
int ev_sync(event_t) { return DS_OK; }
// emulation: everything is synchronous, so the word must already hold the value - anything else is a sequencing
// error of the caller (stage 0 of every connected handle has to run before any stage 1)
int stream_wait_value(stream_t, int* addr, int value) {
    if (*addr - value < 0) return fail(DS_ERR_STATE, "neighbour handle is behind (counter %d, expected %d): run stage 0 on every handle first", *addr, value);
    return DS_OK;
}
int dev_zero(void* p, size_t n) { memset(p, 0, n); return DS_OK; }
int pinned_alloc(void** host, void** dev_view, size_t bytes) {
    *host = malloc(bytes ? bytes : 1);
    if (!*host) return fail(DS_ERR_OOM, "malloc(%zu) failed", bytes);
    *dev_view = *host;
    return DS_OK;
}
void pinned_free(void* host) { free(host); }

template <class Body, int NT, class P>
int launch(const P& p, long long blocks, stream_t, int smem_bytes) {
    // emulation: one "thread" per block (NT = 1), blocks spread over host threads
#pragma omp parallel
    {
        std::vector<unsigned char> sm((size_t)smem_bytes + 64);
#pragma omp for schedule(dynamic, 4)
        for (long long b = 0; b < blocks; b++) Body::template run<1>(p, (int)b, 0, sm.data());
    }
    return DS_OK;
}
#endif

#if DS_CUDA
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn tma_encoder() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)f;
        else
            cudaGetLastError();
    }
    return fn;
}
// 2-D tile map over a pitched 4-byte-element array: dims (w, h), row pitch in elements, box (bw, bh).
bool encode_tile_map(CUtensorMap* m, bool is_float, void* base, int w, int h, int pitch_elems, int bw, int bh) {
    EncodeTiledFn fn = tma_encoder();
    if (!fn) return false;
    const cuuint64_t gdim[2] = {(cuuint64_t)w, (cuuint64_t)h};
    const cuuint64_t gstride[1] = {(cuuint64_t)pitch_elems * 4};
    const cuuint32_t box[2] = {(cuuint32_t)bw, (cuuint32_t)bh};
    const cuuint32_t estr[2] = {1, 1};
    return fn(m, is_float ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, base, gdim, gstride, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
#endif

const int DS_MAX_FRAMES = 1 << 16;   // frame slots per handle (the frame table is dense: an index allocates every slot below it)

template <class T>
int dev_alloc_t(T** p, size_t count) { return dev_alloc((void**)p, count * sizeof(T)); }

struct Range { int lo, hi; };  // [lo, hi)

struct Frame {
    bool used = false;
    int w = 0, h = 0;
    ds_transform xf;
    uint32_t* d_src = nullptr; int pitch = 0; size_t src_cap = 0;
    int corner_x = 0, corner_y = 0, bw = 0, bh = 0;  // absolute placement (cv corners[i], sizes[i])
    int rx = 0, ry = 0, rw = 0, rh = 0;              // feed ROI rel. to padded canvas origin
    void* d_pyr = nullptr; size_t pyr_cap = 0;       // G/W levels 1..L
    uint32_t* d_mbits = nullptr; size_t mbits_cap = 0;
    uint8_t* d_dist = nullptr; size_t dist_cap = 0;   // FEATHER: L1 distance plane of the warped mask over the bbox
    uint8_t* d_seam = nullptr; size_t seam_cap = 0;        // the frame's mask plane over its bbox (seam / content / soft mask)
    uint8_t* d_content = nullptr; size_t content_cap = 0;  // DS_MASK_CONTENT: the content mask, kept for ds_download_frame_mask
    float* d_gainmap = nullptr; size_t gainmap_cap = 0;
    FrameDev dev{};
    bool mask_done = false; // FEATHER: mask bit plane built for the current composite
    bool partial = false;   // only the source rows a row-band handle reads are resident
    // Source pixels still to be brought in (DS_UPLOAD_ASYNC): the copy is cut into chunks of source rows that
    // ds_composite_async issues in the order its row slices need them.
    struct Pending {
        const uint8_t* src = nullptr; size_t stride = 0; bool on_device = false;
        int nchunks = 0, left = 0;
        std::vector<unsigned char> issued;
    } pend;
};

struct LevelPlan {
    int TW = 0, TH = 0, tiles_x = 0, tiles_y = 0;   // tile of the level's gather kernel
    Range own{0, 0};   // level 0: rows whose tiles run; levels >= 1: rows of the per-frame G_l / W_l planes this handle produces (or pulls)
    Range acc{0, 0};   // rows whose dst is written / consumed by the collapse
    int* d_off = nullptr; int* d_fr = nullptr; int* d_ids = nullptr;   // inside the canvas meta arena
    int* d_rec = nullptr;   // per launched tile id: {tile, first and one-past-last entry of its frame list, first frame}
    int n_ids = 0;
};

// One internally pipelined slice of the handle's rows (DESIGN.md "Upload / compute / download pipeline").
struct SubBand {
    Range rows{0, 0};          // level-0 output rows
    Range own[DS_MAXL];        // level 0: feed rows; level l >= 1: rows of G_l / W_l produced - [previous slice's end, this slice's end), nothing is recomputed
    Range coll[DS_MAXL];       // rows of level l accumulated (l >= 1) and finalised by the collapse into level l in this slice
    int ids_first[DS_MAXL], ids_count[DS_MAXL];   // the slice's run of LevelPlan::d_ids per level (heaviest tiles first)
    event_t fed0 = 0;          // recorded after the slice's level-0 feed
    event_t fed = 0;           // (unused)
    event_t done = 0;          // recorded after the slice's last launch
};

}  // namespace

struct ds_canvas {
    ds_canvas_desc desc;
    int L = 0;
    int pw = 0, ph = 0;      // padded canvas
    Range band{0, 0};        // level-0 rows this handle owns
    stream_t stream = 0;     // compute
    stream_t bulk = 0, tail = 0;   // sliced schedule: level-0 feeds (low priority) / the rest of each slice (high)
    event_t ev_start = 0;
    stream_t up = 0, dl = 0; // host-to-device copies (and the optional per-frame inputs), device-to-host copies
    stream_t xp = 0;         // BGR -> BGRX expansion of the chunks the copies deliver
    static const int NSLOT = 4;          // staging ring: one chunk of dense BGR rows per slot
    uint8_t* d_slot[NSLOT] = {nullptr, nullptr, nullptr, nullptr}; size_t slot_cap = 0;
    event_t slot_copied[NSLOT] = {0, 0, 0, 0}, slot_free[NSLOT] = {0, 0, 0, 0};
    bool slot_used[NSLOT] = {false, false, false, false};
    int slot_next = 0;
    int chunk_rows = 512;    // source rows per upload chunk (DS_UPLOAD_CHUNK_ROWS)
    int n_pending = 0;       // frames with chunks still to issue
    // NVLink P2P halo exchange with the handles of the bands above (0) and below (1)
    struct PeerFrame { int idx, ry, rh, rw, gp1; const char* pyr; size_t g_off, w_off; };
    struct Peer {
        bool connected = false;
        bool same_process = false; uint64_t uid = 0; int gen = 0;   // liveness check of raw pointers
        int* flags = nullptr;            // the neighbour's counter words (peer-mapped)
        Range band{0, 0};
        std::vector<PeerFrame> frames;
        std::vector<void*> mapped;       // cudaIpcOpenMemHandle results to close
    } peer[2];
    int* d_flags = nullptr;              // [0] / [1]: level-1 rows ready, written by the neighbour above / below; [2] / [3]: pulled
    int p2p_seq = 0;                     // composites run in exchange mode
    bool exchange = false;               // the launch metadata (and plan[0].own) are those of the exchange mode
    bool exchange_now = false;           // the composite being queued runs in exchange mode
    bool stage_open = false;             // ds_composite_stage(c, 0) ran in exchange mode, its second half is outstanding
    bool stage1_due = false;             // ds_composite_stage(c, 0) ran (any mode): stage 1 is expected next
    uint64_t uid = 0;                    // identity in the registry of live handles (same-process peers check each other's generation)
    int p2p_gen = 0;                     // bumped whenever a per-frame pyramid is reallocated: exported pointers are then stale
    Range own0_recompute{0, 0};          // level-0 feed rows without / with the exchange
    PullSeg* d_segs = nullptr; int n_segs = 0;   // meta arena
    event_t ev_chunks = 0;   // scratch: recorded on xp after the chunks a slice needs
    event_t ev_opts = 0;     // scratch: recorded on up at the start of a composite
    int meta_slice_rows = -2;   // slice height the launch metadata was built for
    bool own_stream = false;
    event_t ev_done = 0;     // end of the last composite (compute stream)
    bool ev_done_valid = false;
    event_t ev_meta = 0;     // metadata copy kernel finished reading h_meta
    bool ev_meta_valid = false;
    bool async_pending = false;   // an upload with DS_UPLOAD_ASYNC since the last composite
    std::vector<SubBand> subs;    // slices of the last composite
    void* h_meta = nullptr; void* h_meta_dev = nullptr; size_t h_meta_cap = 0;
    void* d_meta = nullptr; size_t d_meta_cap = 0;
    int lw[DS_MAXL], lh[DS_MAXL];
    px16* d_lvl[DS_MAXL];    // allocated rows [lvl_rows[l].lo, lvl_rows[l].hi); pointer is the virtual row-0 base
    px16* d_lvl_alloc[DS_MAXL];
    Range lvl_rows[DS_MAXL];
    uint8_t* d_out = nullptr; size_t out_pitch = 0;    // rows [band.lo, out_hi)
    uint8_t* d_mask = nullptr; size_t mask_pitch = 0;
    int out_hi = 0;
    std::vector<Frame> frames;
    FrameDev* d_frames = nullptr;                    // inside the meta arena
    void* d_tmaps = nullptr;                         // CUtensorMap[frame][2] over the frame sources (meta arena)
    bool tmaps_ok = false;
    void* d_lmaps = nullptr;                         // CUtensorMap[frame][L + 1][AccumBody::LM_N] over the per-frame G_l / W_l planes (meta arena)
    LevelPlan plan[DS_MAXL];   // multiband: one per level; feather: plan[0]
    bool dirty = true;
    bool l0_has_affine = false;   // some frame is AFFINE_F64 or has a seam mask / gain map: level 0 runs the general variant of the fast kernel
    bool l0_fast_ok = false;   // level-0 fast kernel applicable (all frames PLANE_F32, <= 64 frames per tile)
    int feather_R = 0;
    void* scratch[3] = {nullptr, nullptr, nullptr}; size_t scratch_cap[3] = {0, 0, 0};   // mask chain temporaries (apply_opts), kept between calls
    int64_t device_bytes = 0;
    int64_t h2d_bytes = 0;   // frame bytes copied host -> device so far
    int64_t launches = 0;
    float last_ms = 0.f;
    bool composited = false;
    bool profiling = false;
    struct Prof { const char* name; int level; int64_t ab; float ms; };
    std::vector<Prof> prof;
#if DS_CUDA
    struct Trace { char label[40]; cudaEvent_t ev; };
    std::vector<Trace> trace;           // DS_TRACE=1: stream timeline of one step, printed by ds_synchronize
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::vector<cudaEvent_t> prof_ev;   // 2 per launch
#endif
};

namespace {

// Live handles of this process and the generation of their exported pointers: a same-process peer holds raw device
// pointers into its neighbour's pyramids, so before using them it checks that the neighbour still exists and has not
// reallocated them since the export.
std::mutex g_reg_mu;
std::map<uint64_t, int> g_reg;
uint64_t g_next_uid = 1;
void reg_set(ds_canvas* c) { std::lock_guard<std::mutex> lk(g_reg_mu); if (!c->uid) c->uid = g_next_uid++; g_reg[c->uid] = c->p2p_gen; }
void reg_drop(ds_canvas* c) { std::lock_guard<std::mutex> lk(g_reg_mu); if (c->uid) g_reg.erase(c->uid); }
bool reg_alive(uint64_t uid, int gen) { std::lock_guard<std::mutex> lk(g_reg_mu); auto it = g_reg.find(uid); return it != g_reg.end() && it->second == gen; }

// DS_TRACE=1 (measurement aid): timestamps on the upload / compute / download streams.
#if DS_CUDA
bool trace_on() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("DS_TRACE"); v = (e && atoi(e) > 0) ? 1 : 0; }
    return v == 1;
}
#endif
void trace_mark(ds_canvas* c, stream_t s, const char* what, int idx) {
#if DS_CUDA
    if (!trace_on() || c->trace.size() > 4096) return;
    ds_canvas::Trace t;
    snprintf(t.label, sizeof(t.label), "%s %d", what, idx);
    if (cudaEventCreate(&t.ev) != cudaSuccess) return;
    cudaEventRecord(t.ev, s);
    c->trace.push_back(t);
#else
    (void)c; (void)s; (void)what; (void)idx;
#endif
}
void trace_dump(ds_canvas* c) {
#if DS_CUDA
    if (c->trace.empty()) return;
    for (const ds_canvas::Trace& t : c->trace) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, c->trace[0].ev, t.ev);
        fprintf(stderr, "[ds_trace] %8.3f ms  %s\n", ms, t.label);
        if (&t != &c->trace[0]) cudaEventDestroy(t.ev);
    }
    cudaEventDestroy(c->trace[0].ev);
    c->trace.clear();
    cudaGetLastError();
#else
    (void)c;
#endif
}

int grow(ds_canvas* c, void** p, size_t* cap, size_t need) {
    if (need <= *cap && *p) return DS_OK;
    if (*p) { dev_free(*p); c->device_bytes -= (int64_t)*cap; *p = nullptr; *cap = 0; }
    int rc = dev_alloc(p, need);
    if (rc) return rc;
    *cap = need;
    c->device_bytes += (int64_t)need;
    return DS_OK;
}

int set_device(const ds_canvas* c) {
#if DS_CUDA
    DS_CK(cudaSetDevice(c->desc.device));
#else
    (void)c;
#endif
    return DS_OK;
}

Range clip(Range r, int n) { return Range{std::max(r.lo, 0), std::min(r.hi, n)}; }
Range meet(Range a, Range b) { return Range{std::max(a.lo, b.lo), std::max(std::max(a.lo, b.lo), std::min(a.hi, b.hi))}; }

// Row plan of a band (see DESIGN.md "Row bands"): which rows of every level this handle must
// accumulate (acc = what the collapse reads), which rows of the per-frame planes G_l / W_l (l >= 1) it must hold
// (own[l]) and which level-0 tile rows must run (own[0]) so that all of them are locally produced — no inter-band
// exchange needed. ds_mb_accum at level l reads G_l over acc[l] and G_{l+1} over acc[l+1] (the pyrUp taps, by the
// definition of acc); ds_mb_pyrdown reads G_l rows [2a - 2, 2b] to produce G_{l+1} rows [a, b).
void plan_rows(int L, const int* lh, Range band, Range* acc, Range* own) {
    acc[0] = band;
    for (int l = 1; l <= L; l++) {
        Range r{(acc[l - 1].lo >> 1) - 1, ((acc[l - 1].hi - 1) >> 1) + 2};
        acc[l] = clip(r, lh[l]);
    }
    own[L] = acc[L];
    for (int l = L - 1; l >= 1; l--) {
        Range need{2 * own[l + 1].lo - 2, 2 * own[l + 1].hi + 1};
        own[l] = clip(Range{std::min(acc[l].lo, need.lo), std::max(acc[l].hi, need.hi)}, lh[l]);
    }
    if (L >= 1) {
        // level-0 tile rows [a, b) (even) produce G_1 rows [a / 2, b / 2)
        Range r{std::min(acc[0].lo, 2 * own[1].lo), std::max(acc[0].hi, 2 * own[1].hi)};
        r.lo &= ~1; r.hi = (r.hi + 1) & ~1;
        own[0] = clip(r, lh[0]);
    } else {
        own[0] = acc[0];
    }
}

// Tiles of the gather kernels: 64 x 64 at level 0 (ds_mb_feed_l0), 32 x 16 above (ds_mb_accum, one 2 x 2 quad per thread).
int level_tile_w(int l) { return l == 0 ? 64 : AccumBody::TW; }
int level_tile_h(int l) { return l == 0 ? 64 : AccumBody::TH; }

int fill_frame_dev(ds_canvas* c, Frame& f) {
    FrameDev& d = f.dev;
    memset(&d, 0, sizeof(d));
    d.src = f.d_src; d.src_w = f.w; d.src_h = f.h; d.src_pitch = f.pitch;
    d.kind = f.xf.kind; d.border = f.xf.border;
    if (f.xf.kind == DS_XF_PLANE_F32) {
        const dsgeo::Projector P = dsgeo::make_projector(f.xf.K, f.xf.R, f.xf.affine_warper != 0);
        memcpy(d.k, P.k_rinv, sizeof(d.k));
        const float one = 1 - P.t[2];
        d.t0 = P.t[0]; d.t1 = P.t[1]; d.scale = f.xf.scale;
        d.k2one = P.k_rinv[2] * one; d.k5one = P.k_rinv[5] * one; d.k8one = P.k_rinv[8] * one;
        d.tlx = f.corner_x; d.tly = f.corner_y;
    } else if (f.xf.kind == DS_XF_AFFINE_F64) {
        dsgeo::invert_affine_f64(f.xf.M, d.inv);
    } else {
        if (!dsgeo::invert_3x3_f64(f.xf.M, d.inv)) return fail(DS_ERR_BAD_ARG, "singular homography");
    }
    d.cx = f.corner_x - c->desc.x; d.cy = f.corner_y - c->desc.y;
    d.w = f.bw; d.h = f.bh;
    d.rx = f.rx; d.ry = f.ry; d.rw = f.rw; d.rh = f.rh;
    if (c->desc.blend_mode == DS_BLEND_MULTIBAND && c->L > 0) {
        char* base = (char*)f.d_pyr;
        size_t off = 0;
        for (int l = 1; l <= c->L; l++) {
            d.gp[l] = ((f.rw >> l) + 3) & ~3;
            const size_t n = (size_t)d.gp[l] * (f.rh >> l);
            d.G[l] = (px8*)(base + off); off += ((n * sizeof(px8) + 127) & ~(size_t)127);
            d.W[l] = (float*)(base + off); off += ((n * sizeof(float) + 127) & ~(size_t)127);
        }
    }
    d.mbits = f.d_mbits; d.mbits_pitch = (f.bw + 31) / 32;
    d.dist = f.d_dist; d.dist_pitch = (f.bw + 15) & ~15;
    d.seam = f.d_seam; d.seam_pitch = f.bw;
    return DS_OK;
}

size_t pyr_bytes(const ds_canvas* c, const Frame& f) {
    size_t off = 0;
    for (int l = 1; l <= c->L; l++) {
        const size_t n = (size_t)(((f.rw >> l) + 3) & ~3) * (f.rh >> l);
        off += ((n * sizeof(px8) + 127) & ~(size_t)127);
        off += ((n * sizeof(float) + 127) & ~(size_t)127);
    }
    return off;
}

int placement(const ds_transform* xf, int w, int h, int* out) {
    if (xf->kind == DS_XF_PLANE_F32) {
        const dsgeo::Projector P = dsgeo::make_projector(xf->K, xf->R, xf->affine_warper != 0);
        int tlx, tly, brx, bry;
        dsgeo::plane_result_roi(P, xf->scale, w, h, tlx, tly, brx, bry);
        out[0] = tlx; out[1] = tly; out[2] = brx - tlx + 1; out[3] = bry - tly + 1;
    } else if (xf->kind == DS_XF_AFFINE_F64 || xf->kind == DS_XF_HOMOGRAPHY_F64) {
        out[0] = xf->corner_x; out[1] = xf->corner_y; out[2] = xf->width; out[3] = xf->height;
    } else {
        return fail(DS_ERR_BAD_ARG, "unknown transform kind %d", xf->kind);
    }
    if (out[2] <= 0 || out[3] <= 0) return fail(DS_ERR_BAD_ARG, "empty warped bbox %dx%d", out[2], out[3]);
    return DS_OK;
}

// Build per-level tile -> frame lists (feed order) on the host and upload them.
// Source rows a frame's bbox rows [v0, v1] can read: the backward map (in double) at the four corners of the row
// strip, +- the bilinear tap and a rounding margin, with the rows border reflection folds back in. Returns false
// when no bound can be given (degenerate projective maps): the caller then takes the whole frame.
bool source_row_range(const FrameDev& F, int v0, int v1, int* lo_out, int* hi_out) {
    const int h = F.src_h;
    double mn = 1e300, mx = -1e300;
    int zsign = 0;
    for (int ci = 0; ci < 4; ci++) {
        const double u = (ci & 1) ? (double)(F.w - 1) : 0.0, v = (ci & 2) ? (double)v1 : (double)v0;
        double y, z = 1.0;
        if (F.kind == XF_PLANE) {
            const double U = ((double)F.tlx + u) / (double)F.scale - (double)F.t0, V = ((double)F.tly + v) / (double)F.scale - (double)F.t1;
            y = (double)F.k[3] * U + (double)F.k[4] * V + (double)F.k5one;
            z = (double)F.k[6] * U + (double)F.k[7] * V + (double)F.k8one;
        } else if (F.kind == XF_AFFINE) {
            y = F.inv[3] * u + F.inv[4] * v + F.inv[5];
        } else {
            y = F.inv[3] * u + F.inv[4] * v + F.inv[5];
            z = F.inv[6] * u + F.inv[7] * v + F.inv[8];
        }
        if (!(fabs(z) > 1e-9)) return false;
        const int sg = z > 0 ? 1 : -1;
        if (zsign && sg != zsign) return false;   // the horizon crosses the strip
        zsign = sg;
        const double sy = y / z;
        if (!(fabs(sy) < 1e9)) return false;
        mn = std::min(mn, sy); mx = std::max(mx, sy);
    }
    mn -= 2.0; mx += 3.0;
    if (mn < -(double)h || mx > 2.0 * (double)h) return false;
    const int L = (int)floor(mn), H = (int)ceil(mx);
    int lo = std::max(L, 0), hi = std::min(H, h - 1);
    if (L < 0) { lo = 0; hi = std::max(hi, std::min(-L, h - 1)); }
    if (H > h - 1) { hi = h - 1; lo = std::min(lo, std::max(2 * (h - 1) - H - 1, 0)); }
    if (lo > hi) { lo = 0; hi = h - 1; }
    *lo_out = lo; *hi_out = hi;
    return true;
}

// ---- chunked source upload: host rows -> staging slot (copy stream) -> BGRX rows of the frame (expansion stream)

int issue_chunk(ds_canvas* c, int fi, int k) {
    Frame& f = c->frames[(size_t)fi];
    Frame::Pending& pd = f.pend;
    if (k < 0 || k >= pd.nchunks || pd.issued[(size_t)k]) return DS_OK;
    int rc;
    const int r0 = k * c->chunk_rows, nr = std::min(c->chunk_rows, f.h - r0);
    ExpandParams ep;
    ep.dst = f.d_src + (size_t)r0 * f.pitch; ep.dst_pitch = f.pitch; ep.w = f.w; ep.h = nr;
    if (pd.on_device) {
        ep.src = pd.src + (size_t)r0 * pd.stride; ep.src_stride = pd.stride;
        if ((rc = launch<ExpandBody, 256>(ep, ExpandBody::blocks(ep), c->xp, 0))) return rc;
    } else {
        const size_t dense = (size_t)f.w * 3;
        const int sl = c->slot_next;
        c->slot_next = (c->slot_next + 1) % ds_canvas::NSLOT;
        if (c->slot_used[sl] && (rc = ev_wait(c->up, c->slot_free[sl]))) return rc;   // its previous chunk is expanded
        trace_mark(c, c->up, "h2d begin, frame", fi);
        if ((rc = h2d_2d(c->d_slot[sl], dense, pd.src + (size_t)r0 * pd.stride, pd.stride, dense, (size_t)nr, c->up))) return rc;
        c->h2d_bytes += (int64_t)(dense * (size_t)nr);
        if ((rc = ev_make(&c->slot_copied[sl])) || (rc = ev_record(c->slot_copied[sl], c->up)) || (rc = ev_wait(c->xp, c->slot_copied[sl]))) return rc;
        ep.src = c->d_slot[sl]; ep.src_stride = dense;
        if ((rc = launch<ExpandBody, 256>(ep, ExpandBody::blocks(ep), c->xp, 0))) return rc;
        if ((rc = ev_make(&c->slot_free[sl])) || (rc = ev_record(c->slot_free[sl], c->xp))) return rc;
        c->slot_used[sl] = true;
    }
    pd.issued[(size_t)k] = 1;
    if (--pd.left == 0) { pd.src = nullptr; c->n_pending--; trace_mark(c, c->xp, "frame complete", fi); }
    return DS_OK;
}

// chunks of frame fi holding source rows [lo, hi]
int issue_rows(ds_canvas* c, int fi, int lo, int hi) {
    int rc;
    for (int k = lo / c->chunk_rows; k <= hi / c->chunk_rows; k++)
        if ((rc = issue_chunk(c, fi, k))) return rc;
    return DS_OK;
}

// everything still pending, in frame order (plain uploads, taps, ds_synchronize, the end of a composite)
int issue_all(ds_canvas* c) {
    int rc;
    for (size_t fi = 0; fi < c->frames.size() && c->n_pending > 0; fi++) {
        Frame& f = c->frames[fi];
        if (!f.used || f.pend.left == 0) continue;
        if ((rc = issue_rows(c, (int)fi, 0, f.h - 1))) return rc;
    }
    return DS_OK;
}

// The chunks slice `sb` reads, then an event on the expansion stream the slice's stream waits for.
int issue_for_slice(ds_canvas* c, const SubBand& sb, stream_t waiter) {
    if (c->n_pending == 0) return DS_OK;
    const bool mb = c->desc.blend_mode == DS_BLEND_MULTIBAND;
    int rc;
    bool any = false;
    for (size_t fi = 0; fi < c->frames.size(); fi++) {
        Frame& f = c->frames[fi];
        if (!f.used || f.pend.left == 0) continue;
        // bbox rows whose warped pixels the slice's kernels evaluate
        int v0, v1;
        if (mb) {
            if (sb.own[0].lo >= sb.own[0].hi) continue;
            v0 = sb.own[0].lo - 5 - f.dev.cy; v1 = sb.own[0].hi + 4 - f.dev.cy;
        } else {
            const int TH = FeatherBody::TH;
            v0 = sb.rows.lo / TH * TH - 1 - f.dev.cy; v1 = (sb.rows.hi + TH - 1) / TH * TH + 1 - f.dev.cy;
        }
        if (mb) {
            // MultiBandBlender::feed extends the warped image over the gap of its ROI with BORDER_REFLECT
            // (copyMakeBorder): ROI rows above / below the bbox read the bbox rows mirrored at its edge
            const int r0 = f.ry - f.dev.cy, r1 = f.ry + f.rh - 1 - f.dev.cy;   // ROI rows, bbox-relative
            const int a = std::max(v0, r0), b = std::min(v1, r1);
            if (a > b) continue;
            v0 = std::max(a, 0); v1 = std::min(b, f.bh - 1);
            if (a < 0) { v0 = 0; v1 = std::max(v1, std::min(-a - 1, f.bh - 1)); }
            if (b > f.bh - 1) { v1 = f.bh - 1; v0 = std::min(v0, std::max(2 * f.bh - b - 1, 0)); }
        } else {
            v0 = std::max(v0, 0); v1 = std::min(v1, f.bh - 1);
        }
        if (v0 > v1) continue;
        int lo = 0, hi = f.h - 1;
        if (!source_row_range(f.dev, v0, v1, &lo, &hi)) { lo = 0; hi = f.h - 1; }
        const int before = f.pend.left;
        static const bool debug_plan = getenv("DS_DEBUG_PLAN") != nullptr;
        if (debug_plan) fprintf(stderr, "slice rows [%d,%d) own0 [%d,%d): frame %zu bbox rows [%d,%d] -> src rows [%d,%d]\n", sb.rows.lo, sb.rows.hi, sb.own[0].lo, sb.own[0].hi, fi, v0, v1, lo, hi);
        if ((rc = issue_rows(c, (int)fi, lo, hi))) return rc;
        any = any || f.pend.left != before;
    }
    if (!any) return DS_OK;   // everything it reads was ordered before an earlier slice's event (same waiter stream)
    if ((rc = ev_make(&c->ev_chunks)) || (rc = ev_record(c->ev_chunks, c->xp)) || (rc = ev_wait(waiter, c->ev_chunks))) return rc;
    return DS_OK;
}

// ---- NVLink P2P halo exchange (include/dronestitch.h, DESIGN.md "Row bands")

// Level-1 canvas rows this handle's level >= 1 feeds read beyond the band, per side: they come from the neighbour.
Range pulled_rows(const ds_canvas* c, int side) {
    const Range own1 = c->plan[1].own;
    const int b_lo = c->band.lo >> 1, b_hi = std::min(c->band.hi >> 1, c->lh[1]);
    if (side == 0) return Range{std::max(own1.lo, 0), b_lo};
    return Range{b_hi, std::min(own1.hi, c->lh[1])};
}

// Can the next composite exchange halos instead of recomputing them?
bool exchange_ready(const ds_canvas* c) {
    if (c->desc.blend_mode != DS_BLEND_MULTIBAND || c->L < 1) return false;
    const bool need_up = c->band.lo > 0, need_down = c->band.hi < c->ph;
    if (!need_up && !need_down) return false;
    return (!need_up || c->peer[0].connected) && (!need_down || c->peer[1].connected);
}

// Segments of the pull kernel: whole rows of G_1 / W_1 of every frame that reaches into the pulled rows.
int build_pull_segments(ds_canvas* c, std::vector<PullSeg>& segs) {
    segs.clear();
    for (int side = 0; side < 2; side++) {
        const ds_canvas::Peer& pr = c->peer[side];
        if (!pr.connected) continue;
        const Range rows = pulled_rows(c, side);
        if (rows.lo >= rows.hi) continue;
        for (size_t fi = 0; fi < c->frames.size(); fi++) {
            const Frame& f = c->frames[fi];
            if (!f.used) continue;
            const int ry1 = f.ry >> 1, rh1 = f.rh >> 1;
            const int a = std::max(rows.lo, ry1) - ry1, b = std::min(rows.hi, ry1 + rh1) - ry1;   // frame-local rows
            if (a >= b) continue;
            const ds_canvas::PeerFrame* pf = nullptr;
            for (const ds_canvas::PeerFrame& q : pr.frames) if (q.idx == (int)fi) { pf = &q; break; }
            if (!pf) return fail(DS_ERR_STATE, "the neighbour %s does not hold frame %zu, which reaches into its band", side ? "below" : "above", fi);
            if (pf->ry != f.ry || pf->rh != f.rh || pf->rw != f.rw || pf->gp1 != f.dev.gp[1])
                return fail(DS_ERR_STATE, "frame %zu has a different geometry on the neighbouring handle", fi);
            const size_t pitch = (size_t)f.dev.gp[1] * 4;   // bytes per row of either array
            const size_t g_local = (size_t)((const char*)f.dev.G[1] - (const char*)f.d_pyr), w_local = (size_t)((const char*)f.dev.W[1] - (const char*)f.d_pyr);
            PullSeg g{(const uint4*)(pf->pyr + pf->g_off + (size_t)a * pitch), (uint4*)((char*)f.d_pyr + g_local + (size_t)a * pitch), (long long)((size_t)(b - a) * pitch / 16)};
            PullSeg w{(const uint4*)(pf->pyr + pf->w_off + (size_t)a * pitch), (uint4*)((char*)f.d_pyr + w_local + (size_t)a * pitch), (long long)((size_t)(b - a) * pitch / 16)};
            segs.push_back(g); segs.push_back(w);
        }
    }
    return DS_OK;
}

// Uploads declared with DS_UPLOAD_ASYNC that no composite has consumed yet: copy them now and drain the streams.
int flush_uploads(ds_canvas* c) {
    int rc;
    if (c->n_pending > 0) {
        if (c->ev_done_valid && (rc = ev_wait(c->xp, c->ev_done))) return rc;
        if ((rc = issue_all(c))) return rc;
    }
    if ((rc = stream_sync(c->up))) return rc;
    return stream_sync(c->xp);
}

int default_pipeline_rows() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("DS_PIPELINE_ROWS");
        v = e ? atoi(e) : 0;
        if (v <= 0) v = 768;
    }
    return v;
}

// Cuts the handle's rows into slices that run one after the other on the compute stream. Slice b feeds, at every
// level, the rows between the end of slice b-1 and its own end (band rows + the pyramid halo below them, as
// plan_rows gives for a band ending there): the per-frame pyramids and the Laplacian levels a slice leaves behind
// are final, so the next slice continues from them and no row is computed twice. Its collapse finalises the rows
// up to its band edge (+ the 1-2 rows the next level down taps), again continuing where the previous one ended.
void plan_subbands(ds_canvas* c, int rows_per_slice) {
    const bool mb = c->desc.blend_mode == DS_BLEND_MULTIBAND;
    const int align = mb ? (1 << c->L) : FeatherBody::TH;
    const Range band{c->band.lo, mb ? c->band.hi : c->out_hi};
    std::vector<int> edges{band.lo, band.hi};
    if (rows_per_slice > 0 && band.hi - band.lo >= 2 * rows_per_slice) {
        const int step = std::max((rows_per_slice + align - 1) / align * align, align);
        // (1) edges where the set of frames a slice waits for changes: the last row a slice may end at without
        // reading frame f is f's first row minus the pyramid halo below a band edge
        int halo = 0;
        if (mb) {
            Range acc[DS_MAXL], own[DS_MAXL];
            const int mid = band.lo + (band.hi - band.lo) / 2 / align * align;
            plan_rows(c->L, c->lh, Range{band.lo, mid}, acc, own);
            halo = std::max(own[0].hi - mid, 0);
        }
        std::vector<int> dep;
        for (const Frame& f : c->frames) {
            if (!f.used) continue;
            const int top = mb ? f.ry : f.dev.cy;
            int y = (top - halo) / align * align;
            if (top - halo < 0) y = band.lo;
            if (y > band.lo + step / 3 && y < band.hi - step / 3) dep.push_back(y);
        }
        std::sort(dep.begin(), dep.end());
        int prev = band.lo;   // keep an edge only if it is at least a third of a slice below the previous one
        for (int y : dep)
            if (y - prev >= step / 3) { edges.insert(edges.end() - 1, y); prev = y; }
        // (2) cut what is left into pieces of about `step` rows
        std::vector<int> fine;
        for (size_t i = 0; i + 1 < edges.size(); i++) {
            const int lo = edges[i], hi = edges[i + 1];
            const int n = std::max(1, (hi - lo + step / 2) / step);
            const int piece = ((hi - lo + n - 1) / n + align - 1) / align * align;
            for (int k = 0; k < n && lo + k * piece < hi; k++) fine.push_back(lo + k * piece);
        }
        fine.push_back(band.hi);
        edges.swap(fine);
    }
    std::vector<event_t> keep_done, keep_fed, keep_fed0;
    for (SubBand& sb : c->subs) { keep_done.push_back(sb.done); keep_fed.push_back(sb.fed); keep_fed0.push_back(sb.fed0); }
    c->subs.clear();
    Range prev_own[DS_MAXL], prev_coll[DS_MAXL];
    for (int l = 0; l <= c->L; l++) { prev_own[l] = Range{0, c->plan[l].own.lo}; prev_coll[l] = Range{0, c->plan[l].acc.lo}; }
    for (size_t i = 0; i + 1 < edges.size(); i++) {
        SubBand sb;
        sb.rows = Range{edges[i], edges[i + 1]};
        const bool last = i + 2 == edges.size();
        for (int l = 0; l < DS_MAXL; l++) { sb.own[l] = sb.coll[l] = Range{0, 0}; sb.ids_first[l] = sb.ids_count[l] = 0; }
        if (mb) {
            Range acc[DS_MAXL], own[DS_MAXL];
            plan_rows(c->L, c->lh, sb.rows, acc, own);
            for (int l = 0; l <= c->L; l++) {
                const Range O = c->plan[l].own, A = c->plan[l].acc;
                const int h = std::max(last ? O.hi : std::min(own[l].hi, O.hi), prev_own[l].hi);
                sb.own[l] = Range{prev_own[l].hi, h};
                const int k = std::max(last ? A.hi : std::min(acc[l].hi, A.hi), prev_coll[l].hi);
                sb.coll[l] = Range{prev_coll[l].hi, k};
                prev_own[l] = sb.own[l]; prev_coll[l] = sb.coll[l];
            }
        }
        if (c->subs.size() < keep_done.size()) {
            sb.done = keep_done[c->subs.size()]; sb.fed = keep_fed[c->subs.size()]; sb.fed0 = keep_fed0[c->subs.size()];
        }
        c->subs.push_back(sb);
    }
    for (size_t i = c->subs.size(); i < keep_done.size(); i++) { ev_drop(keep_done[i]); ev_drop(keep_fed[i]); ev_drop(keep_fed0[i]); }
}

// Launch metadata of a composite: per level the CSR tile -> frames lists and the ids of the tiles that run,
// the frame descriptors and the TMA tensor maps. Everything is assembled in one host buffer and moved to one
// device arena by a copy kernel on the compute stream (ds_meta_copy), so rebuilding it neither blocks the host
// nor waits behind frame uploads in flight.
struct MetaBuilder {
    std::vector<unsigned char> buf;
    size_t add(const void* src, size_t bytes) {
        const size_t off = (buf.size() + 255) & ~(size_t)255;
        buf.resize(off + std::max<size_t>(bytes, 16), 0);
        if (bytes) memcpy(buf.data() + off, src, bytes);
        return off;
    }
};

int build_lists(ds_canvas* c) {
    const bool mb = c->desc.blend_mode == DS_BLEND_MULTIBAND;
    const int nl = mb ? c->L + 1 : 1;
    MetaBuilder mbd;
    size_t off_off[DS_MAXL], fr_off[DS_MAXL], ids_off[DS_MAXL], rec_off[DS_MAXL];
    std::vector<int> counts, off, fr, ids, recs;
    for (int l = 0; l < nl; l++) {
        LevelPlan& pl = c->plan[l];
        const int TW = mb ? pl.TW : FeatherBody::TW, TH = mb ? pl.TH : FeatherBody::TH;
        const int ntiles = pl.tiles_x * pl.tiles_y;
        counts.assign((size_t)ntiles + 1, 0);
        const Range run = (mb && l >= 1) ? pl.acc : pl.own;   // rows whose tiles run
        const int ty_lo = run.lo / TH, ty_hi = (run.hi + TH - 1) / TH;
        auto frame_rect = [&](const Frame& f, int& x0, int& y0, int& x1, int& y1) {
            if (mb) { x0 = f.rx >> l; y0 = f.ry >> l; x1 = x0 + (f.rw >> l); y1 = y0 + (f.rh >> l); }
            else { x0 = f.dev.cx; y0 = f.dev.cy; x1 = x0 + f.bw; y1 = y0 + f.bh; }
        };
        for (int pass = 0; pass < 2; pass++) {
            for (size_t fi = 0; fi < c->frames.size(); fi++) {
                const Frame& f = c->frames[fi];
                if (!f.used) continue;
                int x0, y0, x1, y1;
                frame_rect(f, x0, y0, x1, y1);
                x0 = std::max(x0, 0); y0 = std::max(y0, 0);
                x1 = std::min(x1, pl.tiles_x * TW); y1 = std::min(y1, pl.tiles_y * TH);
                if (x0 >= x1 || y0 >= y1) continue;
                const int tx0 = x0 / TW, tx1 = (x1 - 1) / TW;
                const int ty0 = std::max(y0 / TH, ty_lo), ty1 = std::min((y1 - 1) / TH, ty_hi - 1);
                for (int ty = ty0; ty <= ty1; ty++)
                    for (int tx = tx0; tx <= tx1; tx++) {
                        const int t = ty * pl.tiles_x + tx;
                        if (pass == 0) counts[(size_t)t + 1]++;
                        else fr[(size_t)off[t]++] = (int)fi;
                    }
            }
            if (pass == 0) {
                off.assign((size_t)ntiles + 1, 0);
                for (int t = 0; t < ntiles; t++) off[(size_t)t + 1] = off[t] + counts[(size_t)t + 1];
                fr.assign((size_t)std::max(off[ntiles], 1), 0);
                counts = off;  // keep the prefix sums; `off` is consumed as a cursor in pass 1
            }
        }
        if (mb) {
            // the fast kernels pack two 16-bit accumulator lanes per int: exact while a tile sees <= 64 frames
            int longest = 0;
            for (int t = 0; t < ntiles; t++) longest = std::max(longest, counts[(size_t)t + 1] - counts[t]);
            if (l == 0) {
                // the fast level-0 kernel builds plane (float) and warpAffine (integer) coordinates in its pipelined loops;
                // warpPerspective (double) coordinates go through the per-pixel loop of its general variant
                c->l0_fast_ok = longest <= MBFastBody<64, true>::MAXF;
                c->l0_has_affine = false;
                for (const Frame& f : c->frames) if (f.used && (f.xf.kind != DS_XF_PLANE_F32 || f.d_seam || f.d_gainmap || f.dev.any_gain)) c->l0_has_affine = true;
            }
        }
        // tiles to run, per slice: every tile in the slice's tile rows (empty ones still write zeros), the ones with
        // the most frames first - blocks are dispatched in id order, so the launch drains on its cheapest tiles
        ids.clear();
        for (SubBand& sb : c->subs) {
            const Range rows = mb ? (l >= 1 ? sb.coll[l] : sb.own[l]) : meet(sb.rows, pl.own);
            sb.ids_first[l] = (int)ids.size();
            if (rows.lo < rows.hi) {
                const int ty0 = rows.lo / TH, ty1 = (rows.hi + TH - 1) / TH;
                for (int ty = ty0; ty < ty1; ty++)
                    for (int tx = 0; tx < pl.tiles_x; tx++) ids.push_back(ty * pl.tiles_x + tx);
                static const bool sort_tiles = getenv("DS_SORT_TILES") && atoi(getenv("DS_SORT_TILES")) != 0;   // experiment: row-major order wins (L2 locality)
                if (sort_tiles) std::stable_sort(ids.begin() + sb.ids_first[l], ids.end(), [&](int a, int b) {
                    return counts[(size_t)a + 1] - counts[a] > counts[(size_t)b + 1] - counts[b];
                });
            }
            sb.ids_count[l] = (int)ids.size() - sb.ids_first[l];
        }
        pl.n_ids = (int)ids.size();
        off_off[l] = mbd.add(counts.data(), counts.size() * sizeof(int));
        fr_off[l] = mbd.add(fr.data(), fr.size() * sizeof(int));
        ids_off[l] = mbd.add(ids.data(), ids.size() * sizeof(int));
        // one 16-byte record per launched tile: what a CTA otherwise finds through three dependent loads
        recs.resize(ids.size() * 4);
        for (size_t k = 0; k < ids.size(); k++) {
            const int t = ids[k];
            recs[4 * k] = t; recs[4 * k + 1] = counts[t]; recs[4 * k + 2] = counts[(size_t)t + 1];
            recs[4 * k + 3] = counts[(size_t)t + 1] > counts[t] ? fr[(size_t)counts[t]] : -1;
        }
        rec_off[l] = mbd.add(recs.data(), recs.size() * sizeof(int));
    }
    // frame descriptors
    std::vector<FrameDev> fd(c->frames.size());
    for (size_t i = 0; i < c->frames.size(); i++) {
        if (c->frames[i].used) fd[i] = c->frames[i].dev; else memset(&fd[i], 0, sizeof(FrameDev));
    }
    const size_t frames_off = mbd.add(fd.data(), fd.size() * sizeof(FrameDev));
    std::vector<PullSeg> segs;
    if (c->exchange) { int rcs = build_pull_segments(c, segs); if (rcs) return rcs; }
    const size_t segs_off = mbd.add(segs.data(), segs.size() * sizeof(PullSeg));
    size_t tmaps_off = 0, lmaps_off = 0;
    bool lmaps_ok = false;
    c->tmaps_ok = false;
#if DS_CUDA
    if (c->desc.blend_mode == DS_BLEND_MULTIBAND && c->L >= 1 && !c->frames.empty()) {
        // TMA descriptors of the frame sources (level-0 kernel: shared-memory box loads and L2 prefetches)
        std::vector<CUtensorMap> tm(c->frames.size() * 2);
        memset(tm.data(), 0, tm.size() * sizeof(CUtensorMap));
        bool ok = true;
        for (size_t i = 0; i < c->frames.size() && ok; i++) {
            const Frame& f = c->frames[i];
            if (!f.used) continue;
            // slot [frame][0]: the BGRX source, box = footprint of a 71x71 tile region under a few degrees of
            // rotation (L2 prefetch only: correctness never depends on it)
            ok = encode_tile_map(&tm[i * 2], false, f.d_src, f.w, f.h, f.pitch, 96, 88) &&
                 // slot [frame][1]: the same source with the box the level-0 kernel loads into shared memory
                 encode_tile_map(&tm[i * 2 + 1], false, f.d_src, f.w, f.h, f.pitch, MBFastBody<64, true>::BOXW, MBFastBody<64, true>::BOXH);
        }
        if (ok) {
            tmaps_off = mbd.add(tm.data(), tm.size() * sizeof(CUtensorMap));
            c->tmaps_ok = true;
        }
        // ... and of the per-frame planes G_l / W_l, l = 1 .. L, with the boxes ds_mb_accum stages in its shared-memory ring
        if (ok) {
            const size_t ls = (size_t)c->L + 1;
            std::vector<CUtensorMap> lm(c->frames.size() * ls * AccumBody::LM_N);
            memset(lm.data(), 0, lm.size() * sizeof(CUtensorMap));
            for (size_t i = 0; i < c->frames.size() && ok; i++) {
                const Frame& f = c->frames[i];
                if (!f.used) continue;
                for (int l = 1; l <= c->L && ok; l++) {
                    const int w = f.rw >> l, h = f.rh >> l;
                    CUtensorMap* m = &lm[(i * ls + l) * AccumBody::LM_N];
                    ok = encode_tile_map(m + AccumBody::LM_G, false, f.dev.G[l], w, h, f.dev.gp[l], AccumBody::GBW, AccumBody::GBH) &&
                         encode_tile_map(m + AccumBody::LM_W, true, f.dev.W[l], w, h, f.dev.gp[l], AccumBody::GBW, AccumBody::GBH) &&
                         encode_tile_map(m + AccumBody::LM_C, false, f.dev.G[l], w, h, f.dev.gp[l], AccumBody::CBW, AccumBody::CBH) &&
                         encode_tile_map(m + AccumBody::LM_PG, false, f.dev.G[l], w, h, f.dev.gp[l], PyrDownBody::XW, PyrDownBody::XH) &&
                         encode_tile_map(m + AccumBody::LM_PW, true, f.dev.W[l], w, h, f.dev.gp[l], PyrDownBody::XW, PyrDownBody::XH);
                }
            }
            if (ok) { lmaps_off = mbd.add(lm.data(), lm.size() * sizeof(CUtensorMap)); lmaps_ok = true; }
        }
    }
#endif
    // arenas (grown geometrically so that steady-state rebuilds allocate nothing)
    int rc;
    const size_t need = (mbd.buf.size() + 255) & ~(size_t)255;
    if (c->ev_meta_valid && (rc = ev_sync(c->ev_meta))) return rc;   // the previous copy kernel has read h_meta
    if (need > c->h_meta_cap) {
        pinned_free(c->h_meta); c->h_meta = nullptr; c->h_meta_cap = 0;
        const size_t cap = need + need / 2;
        if ((rc = pinned_alloc(&c->h_meta, &c->h_meta_dev, cap))) return rc;
        c->h_meta_cap = cap;
    }
    if (need > c->d_meta_cap && (rc = grow(c, &c->d_meta, &c->d_meta_cap, need + need / 2))) return rc;
    memcpy(c->h_meta, mbd.buf.data(), mbd.buf.size());
    MetaCopyParams cp{(const uint4*)c->h_meta_dev, (uint4*)c->d_meta, (long long)(need / 16)};
    if ((rc = launch<MetaCopyBody, 256>(cp, (cp.n16 + MetaCopyBody::PER_BLOCK - 1) / MetaCopyBody::PER_BLOCK, c->stream, 0))) return rc;
    if ((rc = ev_make(&c->ev_meta)) || (rc = ev_record(c->ev_meta, c->stream))) return rc;
    c->ev_meta_valid = true;
    char* base = (char*)c->d_meta;
    for (int l = 0; l < nl; l++) {
        c->plan[l].d_off = (int*)(base + off_off[l]);
        c->plan[l].d_fr = (int*)(base + fr_off[l]);
        c->plan[l].d_ids = (int*)(base + ids_off[l]);
        c->plan[l].d_rec = (int*)(base + rec_off[l]);
    }
    c->d_frames = (FrameDev*)(base + frames_off);
    c->d_segs = (PullSeg*)(base + segs_off); c->n_segs = (int)segs.size();
    c->d_tmaps = c->tmaps_ok ? (void*)(base + tmaps_off) : nullptr;
    c->d_lmaps = lmaps_ok ? (void*)(base + lmaps_off) : nullptr;
    c->dirty = false;
    return DS_OK;
}

OutParams out_params(const ds_canvas* c) {
    OutParams o;
    const int bpp = c->desc.out_format == DS_OUT_BGRA8 ? 4 : 3;
    (void)bpp;
    // virtual row-0 bases: rows [band.lo, out_hi) are allocated
    o.out = c->d_out - (size_t)c->band.lo * c->out_pitch;
    o.out_pitch = c->out_pitch;
    o.mask = c->d_mask ? c->d_mask - (size_t)c->band.lo * c->mask_pitch : nullptr;
    o.mask_pitch = c->mask_pitch;
    o.fmt = c->desc.out_format == DS_OUT_BGRA8 ? 1 : 0;
    o.w = c->desc.width; o.h = c->out_hi;
    return o;
}

// Marks the start / end of one launch for the per-kernel profile (no-ops unless profiling is on).
int prof_mark(ds_canvas* c, stream_t st, bool begin, const char* name, int level, int64_t ab) {
    if (!c->profiling) return DS_OK;
    if (begin) c->prof.push_back(ds_canvas::Prof{name, level, ab, 0.f});
#if DS_CUDA
    const size_t need = c->prof.size() * 2;
    while (c->prof_ev.size() < need) {
        cudaEvent_t e;
        DS_CK(cudaEventCreate(&e));
        c->prof_ev.push_back(e);
    }
    DS_CK(cudaEventRecord(c->prof_ev[(c->prof.size() - 1) * 2 + (begin ? 0 : 1)], st));
#else
    (void)st;
#endif
    return DS_OK;
}

// Algorithmic bytes of the SURVEY.md 8(d) model attributed per launch (see DESIGN.md "Roofline"):
//   feed level 0   : 3*S (source) + A*(20 RMW dst0 + 10/4 write G1,W1)
//   feed level l>=1: A/4^l * (20 read G,W twice + 20 RMW dst_l) + A/4^(l+1) * 10 (write next level)
//   top level      : A/4^L * (20 + 20)
//   collapse       : C * 17.5 / 4^(l-1) per step into level l-1, minus 2*C for the final uchar4 store
struct ABModel { double S, A, C; };
// canvas pixels of the model: the rows this handle owns (a row-band handle composites its band only)
double ab_canvas_px(const ds_canvas* c) {
    const int rows = std::max(0, std::min(c->band.hi, c->desc.height) - c->band.lo);
    return (double)c->desc.width * rows;
}
// share of a frame that falls into the rows this handle owns (1 for a whole-canvas handle): a frame straddling two bands
// counts once across the handles, and the halo a band recomputes is overhead, not algorithmic work
double ab_frame_share(const ds_canvas* c, const Frame& f) {
    const int top = f.corner_y - c->desc.y;
    const int rows = std::min(top + f.bh, std::min(c->band.hi, c->desc.height)) - std::max(top, c->band.lo);
    return f.bh > 0 ? (double)std::max(rows, 0) / (double)f.bh : 0.0;
}
ABModel ab_inputs(const ds_canvas* c) {
    ABModel m{0, 0, ab_canvas_px(c)};
    for (const Frame& f : c->frames) if (f.used) { const double k = ab_frame_share(c, f); m.S += k * f.w * f.h; m.A += k * f.bw * f.bh; }
    return m;
}

// ---- launches of one composite, per row slice (SubBand). A handle that is not pipelined has one slice.

int launch_feed(ds_canvas* c, stream_t st, int l, const SubBand& sb, const ABModel& abm) {
    int rc;
    LevelPlan& pl = c->plan[l];
    // level 0: the tile rows that run (every row that runs also writes its dst, if this handle stores it);
    // levels >= 1: the rows of the level this slice accumulates and finalises
    const Range own = l == 0 ? sb.own[0] : sb.coll[l];
    if (own.lo >= own.hi) return DS_OK;
    const int first = sb.ids_first[l], count = sb.ids_count[l];
    const Range acc = meet(own, pl.acc);
    MBParams mp;
    mp.frames = c->d_frames; mp.tile_off = pl.d_off; mp.tile_frames = pl.d_fr; mp.tile_ids = pl.d_ids + first;
    mp.tile_rec = (const int4*)(pl.d_rec + 4 * (size_t)first);
    mp.tiles_x = pl.tiles_x; mp.level = l; mp.L = c->L;
    mp.dst = c->d_lvl[l]; mp.dst_w = c->lw[l]; mp.dst_h = c->lh[l];
    mp.acc_y0 = acc.lo; mp.acc_y1 = acc.hi;
    mp.own_y0 = own.lo; mp.own_y1 = own.hi;
    mp.tmaps = c->tmaps_ok ? c->d_tmaps : nullptr;
    static const bool gap_fast = !(getenv("DS_GAP_FAST") && atoi(getenv("DS_GAP_FAST")) == 0);
    // bit 1: source footprints staged in shared memory by TMA box loads (needs the tensor maps; DS_SRC_BOX=0 turns it off)
    static const bool src_box = !(getenv("DS_SRC_BOX") && atoi(getenv("DS_SRC_BOX")) == 0);
    mp.flags = (gap_fast ? 1 : 0) | ((src_box && mp.tmaps) ? 2 : 0);
    // bit 2: ds_mb_accum brings the planes in through its shared-memory ring of TMA boxes (DS_ACC_RING=0: direct loads)
    static const bool acc_ring = !(getenv("DS_ACC_RING") && atoi(getenv("DS_ACC_RING")) == 0);
    mp.lmaps = c->d_lmaps; mp.lstride = c->L + 1;
    mp.rnd_bias = DS_RND_BIAS;
    mp.m_tiles_x = pl.tiles_x > 1 ? (uint32_t)(0x100000000ull / (uint32_t)pl.tiles_x) : 0xffffffffu;
    if (acc_ring && mp.lmaps) mp.flags |= 4;
    static const bool gap_full = !(getenv("DS_GAP_FULL") && atoi(getenv("DS_GAP_FULL")) == 0);
    if (!gap_full) mp.flags |= 8;
#if !DS_CUDA
    if (src_box) mp.flags |= 2;   // the emulator stages the boxes with a plain copy loop
    if (acc_ring) mp.flags |= 4;
#endif
    const double q = 1.0 / (double)(1ull << (2 * l));
    double ab;
    // (levels >= 1: the model's "read G, W twice" is split between ds_mb_pyrdown, which reads them once, and this launch)
    if (l == 0) ab = 3.0 * abm.S + abm.A * (c->L > 0 ? 22.5 : 20.0);
    else if (l < c->L) ab = abm.A * q * 30.0;
    else ab = abm.A * q * 40.0;
    ab *= (double)count / (double)std::max(pl.n_ids, 1);
    if ((rc = prof_mark(c, st, true, l == 0 ? "mb_feed" : "mb_accum", l, (int64_t)ab))) return rc;
    if (l == 0 && c->L > 0 && c->l0_fast_ok && c->l0_has_affine) rc = launch<MBFastBody<64, true, true>, 512>(mp, count, st, MBFastBody<64, true, true>::smem_bytes());
    else if (l == 0 && c->L > 0 && c->l0_fast_ok) rc = launch<MBFastBody<64, true>, 512>(mp, count, st, MBFastBody<64, true>::smem_bytes());
    else if (l == 0) rc = launch<MBBody<64, true>, 512>(mp, count, st, MBBody<64, true>::smem_bytes());
    else rc = launch<AccumBody, 128>(mp, count, st, AccumBody::smem_bytes());
    if (rc) return rc;
    if ((rc = prof_mark(c, st, false, nullptr, 0, 0))) return rc;
    c->launches++;
    return DS_OK;
}

// G_{l+1} / W_{l+1} = pyrDown(G_l / W_l) of every frame, over the rows of level l + 1 this slice produces
int launch_pyrdown(ds_canvas* c, stream_t st, int l, const SubBand& sb, const ABModel& abm) {
    int rc;
    const Range rows = sb.own[l + 1];
    if (rows.lo >= rows.hi || c->frames.empty()) return DS_OK;
    // CTAs per frame: the column blocks of the widest plane x the row blocks a frame can have inside the rows
    int wmax = 0, hmax = 0;
    for (const Frame& f : c->frames) if (f.used) { wmax = std::max(wmax, f.rw >> (l + 1)); hmax = std::max(hmax, f.rh >> (l + 1)); }
    if (wmax <= 0 || hmax <= 0) return DS_OK;
    PyrParams pp;
    pp.frames = c->d_frames; pp.nframes = (int)c->frames.size(); pp.level = l;
    pp.txmax = (wmax + PyrDownBody::BW - 1) / PyrDownBody::BW;
    pp.R = (std::min(rows.hi - rows.lo, hmax) + PyrDownBody::BH - 1) / PyrDownBody::BH + 1;   // + 1: the first needed row is anywhere in its block
    pp.m_per = (uint32_t)(0x100000000ull / (uint32_t)(pp.txmax * pp.R)); pp.m_tx = (uint32_t)(0x100000000ull / (uint32_t)pp.txmax);
    if (pp.txmax * pp.R == 1) pp.m_per = 0xffffffffu;
    if (pp.txmax == 1) pp.m_tx = 0xffffffffu;
    pp.own_y0 = rows.lo; pp.own_y1 = rows.hi;
    static const bool pyr_box = !(getenv("DS_PYR_BOX") && atoi(getenv("DS_PYR_BOX")) == 0);   // DS_PYR_BOX=0: direct loads everywhere
    pp.lmaps = pyr_box ? c->d_lmaps : nullptr; pp.lstride = c->L + 1;
#if DS_CUDA
    pp.boxes = (pyr_box && pp.lmaps) ? 1 : 0;
#else
    pp.boxes = pyr_box ? 1 : 0;
#endif
    const double q = 1.0 / (double)(1ull << (2 * l));
    const Range all = c->plan[l + 1].own;
    const double ab = abm.A * q * 12.5 * (double)(rows.hi - rows.lo) / (double)std::max(all.hi - all.lo, 1);
    if ((rc = prof_mark(c, st, true, "mb_pyrdown", l, (int64_t)ab))) return rc;
    if ((rc = launch<PyrDownBody, 128>(pp, (long long)pp.nframes * pp.txmax * pp.R, st, PyrDownBody::smem_bytes()))) return rc;
    if ((rc = prof_mark(c, st, false, nullptr, 0, 0))) return rc;
    c->launches++;
    return DS_OK;
}

// levels 1 .. L of one slice: every per-frame plane first (each reads the one below), then the accumulations
int launch_upper_levels(ds_canvas* c, stream_t st, const SubBand& sb, const ABModel& abm) {
    int rc;
    for (int l = 1; l < c->L; l++) if ((rc = launch_pyrdown(c, st, l, sb, abm))) return rc;
    for (int l = 1; l <= c->L; l++) if ((rc = launch_feed(c, st, l, sb, abm))) return rc;
    return DS_OK;
}

// collapse step l: lvl_{l-1} = sat(pyrUp(lvl_l) + lvl_{l-1}) over the rows of level l-1 this slice finalises
int launch_collapse(ds_canvas* c, stream_t st, int l, const SubBand& sb, const ABModel& abm) {
    int rc;
    const Range rows = sb.coll[l - 1];
    if (rows.lo >= rows.hi) return DS_OK;
    CollapseParams cp;
    cp.coarse = c->d_lvl[l]; cp.cw = c->lw[l]; cp.ch = c->lh[l];
    cp.fine = c->d_lvl[l - 1]; cp.fw = c->lw[l - 1]; cp.fh = c->lh[l - 1];
    cp.y0 = rows.lo; cp.y1 = rows.hi;
    cp.final = (l == 1);
    cp.o = out_params(c);
    const long long items = CollapseBody::items(cp);
    const double q = 1.0 / (double)(1ull << (2 * (l - 1)));
    const Range all = c->plan[l - 1].acc;
    const double ab = (abm.C * 17.5 * q - (l == 1 ? 2.0 * abm.C : 0.0)) * (double)(rows.hi - rows.lo) / (double)std::max(all.hi - all.lo, 1);
    if ((rc = prof_mark(c, st, true, "mb_collapse", l - 1, (int64_t)ab))) return rc;
    if ((rc = launch<CollapseBody, 256>(cp, (items + CollapseBody::PER_BLOCK - 1) / CollapseBody::PER_BLOCK, st, 0))) return rc;
    if ((rc = prof_mark(c, st, false, nullptr, 0, 0))) return rc;
    c->launches++;
    return DS_OK;
}

int launch_finalize_l0(ds_canvas* c, stream_t st, const SubBand& sb) {
    const Range rows = meet(sb.rows, Range{c->band.lo, c->out_hi});
    if (rows.lo >= rows.hi) return DS_OK;
    FinalizeL0Params fp{c->d_lvl[0], c->lw[0], c->lh[0], rows.lo, rows.hi, out_params(c)};
    const long long n = (long long)fp.fw * (fp.y1 - fp.y0);
    int rc = launch<FinalizeL0Body, 256>(fp, (n + FinalizeL0Body::PER_BLOCK - 1) / FinalizeL0Body::PER_BLOCK, st, 0);
    if (!rc) c->launches++;
    return rc;
}

int launch_feather(ds_canvas* c, stream_t st, const SubBand& sb, const ABModel& abm) {
    int rc;
    LevelPlan& pl = c->plan[0];
    const Range rows = meet(sb.rows, pl.own);
    if (rows.lo >= rows.hi) return DS_OK;
    const int TH = FeatherBody::TH;
    const int ty0 = rows.lo / TH, ty1 = (rows.hi + TH - 1) / TH;
    // mask bit planes of the frames this slice blends (once per composite)
    for (size_t i = 0; i < c->frames.size(); i++) {
        Frame& f = c->frames[i];
        if (!f.used || f.mask_done) continue;
        if (f.dev.cy >= ty1 * TH || f.dev.cy + f.bh <= ty0 * TH) continue;
        const long long nwords = (long long)f.dev.mbits_pitch * f.bh;
        if ((rc = prof_mark(c, st, true, "feather_mask", -1, 0))) return rc;
        if (MaskBitsRowBody::eligible(f.dev)) {
            // plane maps without a seam mask: one thread per 32-pixel word, two coordinate evaluations in the interior
            MaskBitsRowParams mp{c->d_frames, (int)i, f.d_mbits};
            rc = launch<MaskBitsRowBody, 256>(mp, (nwords + MaskBitsRowBody::PER_BLOCK - 1) / MaskBitsRowBody::PER_BLOCK, st, 0);
        } else {
            MaskBitsParams mp{c->d_frames, (int)i, f.d_mbits};
            rc = launch<MaskBitsBody, 256>(mp, (nwords + MaskBitsBody::WORDS_PER_BLOCK - 1) / MaskBitsBody::WORDS_PER_BLOCK, st, 0);
        }
        if (rc) return rc;
        if ((rc = prof_mark(c, st, false, nullptr, 0, 0))) return rc;
        c->launches++;
        // L1 distance to the nearest zero of that mask, once per frame (the blend kernel reads it per pixel)
        FeatherDistParams dp{c->d_frames, (int)i, f.d_dist, c->feather_R};
        if ((rc = prof_mark(c, st, true, "feather_dist", -1, 0))) return rc;
        if ((rc = launch<FeatherDistBody, 256>(dp, FeatherDistBody::blocks(f.dev), st, FeatherDistBody::smem_bytes()))) return rc;
        if ((rc = prof_mark(c, st, false, nullptr, 0, 0))) return rc;
        c->launches++;
        f.mask_done = true;
    }
    const int first = sb.ids_first[0], count = sb.ids_count[0];
    FeatherParams fp;
    fp.frames = c->d_frames; fp.tile_off = pl.d_off; fp.tile_frames = pl.d_fr; fp.tile_ids = pl.d_ids + first;
    fp.tiles_x = pl.tiles_x; fp.sharpness = c->desc.sharpness; fp.R = c->feather_R;
    fp.row0 = rows.lo; fp.row1 = rows.hi;
    fp.o = out_params(c);
    const double ab = (3.0 * abm.S + 4.0 * abm.C) * (double)count / (double)std::max(pl.n_ids, 1);
    if ((rc = prof_mark(c, st, true, "feather_blend", 0, (int64_t)ab))) return rc;
    if ((rc = launch<FeatherBody, 256>(fp, count, st, FeatherBody::smem_bytes()))) return rc;
    if ((rc = prof_mark(c, st, false, nullptr, 0, 0))) return rc;
    c->launches++;
    return DS_OK;
}

// Tail of every composite: uploads nobody asked for, timing events, the "done" event later uploads wait for.
int composite_epilogue(ds_canvas* c) {
    int rc;
    const bool feather = c->desc.blend_mode == DS_BLEND_FEATHER;
    if (c->n_pending > 0) {
        const bool whole = c->band.lo == 0 && c->band.hi >= (feather ? c->desc.height : c->ph);
        if (whole) {
            // source rows no slice read (none, normally) still belong to the resident frame
            if ((rc = issue_all(c))) return rc;
            if ((rc = ev_make(&c->ev_chunks)) || (rc = ev_record(c->ev_chunks, c->xp)) || (rc = ev_wait(c->stream, c->ev_chunks))) return rc;
        } else {
            // a row-band handle never reads the source rows that map outside its band + halo: they are not
            // transferred at all (the frame stays partially resident; only the whole-frame debug taps mind)
            for (Frame& f : c->frames)
                if (f.used && f.pend.left > 0) { f.pend.left = 0; f.pend.src = nullptr; f.partial = true; c->n_pending--; }
        }
    }
#if DS_CUDA
    DS_CK(cudaEventRecord(c->ev1, c->stream));
#endif
    if ((rc = ev_make(&c->ev_done)) || (rc = ev_record(c->ev_done, c->stream))) return rc;
    c->ev_done_valid = true;
    c->composited = true;
    c->stage_open = false;
    return DS_OK;
}

int launch_signal(ds_canvas* c, stream_t st, int* a, int* b, int value) {
    if (!a && !b) return DS_OK;
    SignalParams sp{a, b, value};
    int rc = launch<SignalBody, 32>(sp, 1, st, 0);
    if (!rc) c->launches++;
    return rc;
}

// Exchange mode, first half: level-0 feed over exactly the band's rows, then tell the neighbours that this
// handle's level-1 rows of composite `seq` are in place.
int composite_exchange_begin(ds_canvas* c, const ABModel& abm) {
    int rc;
    const int seq = ++c->p2p_seq;
    SubBand& sb = c->subs[0];
    // the neighbours have pulled what they needed of the previous composite's rows (which this feed overwrites)
    for (int side = 0; side < 2; side++)
        if (c->peer[side].connected && (rc = stream_wait_value(c->stream, c->d_flags + 2 + side, seq - 1))) return rc;
    if ((rc = launch_feed(c, c->stream, 0, sb, abm))) return rc;
    // seen from the neighbour above this handle is the one below: its word [1]; and vice versa
    return launch_signal(c, c->stream, c->peer[0].connected ? c->peer[0].flags + 1 : nullptr, c->peer[1].connected ? c->peer[1].flags + 0 : nullptr, seq);
}

// Second half: wait for the neighbours' level-1 rows, pull the halo rows over NVLink, release the neighbours, and
// run the levels above and the collapse from local memory.
int composite_exchange_finish(ds_canvas* c) {
    int rc;
    const ABModel abm = ab_inputs(c);
    const int seq = c->p2p_seq;
    SubBand& sb = c->subs[0];
    for (int side = 0; side < 2; side++)
        if (c->peer[side].connected && (rc = stream_wait_value(c->stream, c->d_flags + side, seq))) return rc;
    if (c->n_segs > 0) {
        PullParams pp{c->d_segs, c->n_segs};
        if ((rc = prof_mark(c, c->stream, true, "p2p_pull", 1, 0))) return rc;
        if ((rc = launch<PullBody, 256>(pp, (long long)c->n_segs * PullBody::BLOCKS_PER_SEG, c->stream, 0))) return rc;
        if ((rc = prof_mark(c, c->stream, false, nullptr, 0, 0))) return rc;
        c->launches++;
    }
    if ((rc = launch_signal(c, c->stream, c->peer[0].connected ? c->peer[0].flags + 3 : nullptr, c->peer[1].connected ? c->peer[1].flags + 2 : nullptr, seq))) return rc;
    if ((rc = launch_upper_levels(c, c->stream, sb, abm))) return rc;
    for (int l = c->L; l >= 1; l--) if ((rc = launch_collapse(c, c->stream, l, sb, abm))) return rc;
    if ((rc = ev_make(&sb.done)) || (rc = ev_record(sb.done, c->stream))) return rc;
    return composite_epilogue(c);
}

// stage: -1 = the whole composite; 0 / 1 = its two halves (ds_composite_stage)
int run_composite(ds_canvas* c, int stage) {
    int rc;
    if (stage == 1) {
        if (!c->stage1_due) return fail(DS_ERR_STATE, "ds_composite_stage(c, 1) without stage 0");
        if (c->stage_open) {
            rc = composite_exchange_finish(c);
            if (!rc) c->stage1_due = false;
            return rc;
        }
        c->stage1_due = false;
        return DS_OK;   // stage 0 ran the whole composite (no exchange for this handle): nothing was left to do
    }
    if (c->stage_open || c->stage1_due) return fail(DS_ERR_STATE, "the previous composite is half done: call ds_composite_stage(c, 1)");
    // same-process neighbours: their exported pointers must still be what was exported
    for (int side = 0; side < 2; side++) {
        const ds_canvas::Peer& pr = c->peer[side];
        if (pr.connected && pr.same_process && !reg_alive(pr.uid, pr.gen))
            return fail(DS_ERR_STATE, "the neighbour %s was destroyed or reallocated its frames: export and connect again", side ? "below" : "above");
    }
    // pipelined (sliced) when asked for, or when frames are still arriving on the upload stream
    int slice_rows = 0;
    if (c->desc.pipeline_rows > 0) slice_rows = c->desc.pipeline_rows;
    else if (c->desc.pipeline_rows == 0 && c->async_pending) slice_rows = default_pipeline_rows();
    c->async_pending = false;
    // halos from the neighbours instead of recomputing them: only the unsliced schedule can wait for a neighbour
    const bool xmode = slice_rows == 0 && exchange_ready(c);
    const bool replan = c->dirty || slice_rows != c->meta_slice_rows || xmode != c->exchange || c->subs.empty();
    if (replan) {
        c->exchange = xmode;
        if (c->desc.blend_mode == DS_BLEND_MULTIBAND) c->plan[0].own = xmode ? c->plan[0].acc : c->own0_recompute;
        plan_subbands(c, slice_rows);
    }
    c->exchange_now = xmode;
    const bool feather = c->desc.blend_mode == DS_BLEND_FEATHER;
    const bool two = c->subs.size() > 1 && !c->profiling && c->bulk && c->tail;
    const stream_t P = two ? c->bulk : c->stream, Q = two ? c->tail : c->stream;
    if (c->n_pending > 0) {
        // the expansions overwrite frame sources the previous composite may still read
        if (c->ev_done_valid && (rc = ev_wait(c->xp, c->ev_done))) return rc;
        // get the copy engine going before the host builds the launch metadata
        if ((rc = issue_for_slice(c, c->subs[0], P))) return rc;
    }
    if (replan) {
        if ((rc = build_lists(c))) return rc;
        c->meta_slice_rows = slice_rows;
    }
    // optional per-frame inputs (seam masks, gain maps) were copied on the upload stream when the frame was declared
    if ((rc = ev_make(&c->ev_opts)) || (rc = ev_record(c->ev_opts, c->up)) || (rc = ev_wait(c->stream, c->ev_opts))) return rc;
    c->launches = 0;
    const ABModel abm = ab_inputs(c);
    for (Frame& f : c->frames) f.mask_done = false;
#if DS_CUDA
    DS_CK(cudaEventRecord(c->ev0, c->stream));
#endif
    c->stage1_due = (stage == 0);
    if (xmode) {
        c->stage_open = true;
        if ((rc = composite_exchange_begin(c, abm))) return rc;
        return stage == -1 ? composite_exchange_finish(c) : DS_OK;
    }
    // Connected handles that run the recompute schedule this time (row slices, uploads in flight) keep the hand-over
    // protocol going, so a neighbour that does exchange never waits for nothing: this composite counts, the neighbours'
    // "pulled" marks of the previous one are waited for before the level-1 rows are overwritten, and after the last
    // level-0 feed the rows are announced - and, since this handle pulls nothing, released at once.
    const bool keepalive = c->peer[0].connected || c->peer[1].connected;
    int ka_seq = 0;
    if (keepalive) {
        ka_seq = ++c->p2p_seq;
        for (int side = 0; side < 2; side++)
            if (c->peer[side].connected && (rc = stream_wait_value(c->stream, c->d_flags + 2 + side, ka_seq - 1))) return rc;
    }
    // Sliced schedule: the level-0 feeds (the bulk of the work) of all slices run back to back on a low-priority
    // stream; everything else of a slice - its feeds of level >= 1 and its collapse, small launches that are
    // latency-bound - follows on a high-priority stream as soon as the slice's level-0 feed is done, and overlaps
    // the next slice's level-0 feed. Stream order on the second stream gives the cross-slice order for free: a
    // slice's level >= 1 feeds read the pyramid rows the previous slice left, its collapse continues from the rows
    // the previous collapse finalised. With one slice, or per-kernel profiling, everything is on the canvas stream.
    if (two) {
        // after the previous composite and the metadata copy (both on the canvas stream)
        if ((rc = ev_make(&c->ev_start)) || (rc = ev_record(c->ev_start, c->stream))) return rc;
        if ((rc = ev_wait(P, c->ev_start)) || (rc = ev_wait(Q, c->ev_start))) return rc;
    }
    for (size_t b = 0; b < c->subs.size(); b++) {
        SubBand& sb = c->subs[b];
        if (b > 0 && (rc = issue_for_slice(c, sb, P))) return rc;
        trace_mark(c, P, "slice begin, row", sb.rows.lo);
        if (feather) {
            if ((rc = launch_feather(c, P, sb, abm))) return rc;
            if ((rc = ev_make(&sb.done)) || (rc = ev_record(sb.done, P))) return rc;
            trace_mark(c, P, "slice end, row", sb.rows.hi);
            continue;
        }
        if ((rc = launch_feed(c, P, 0, sb, abm))) return rc;
        trace_mark(c, P, "  level-0 feed end, row", sb.rows.hi);
        if (keepalive && b + 1 == c->subs.size()) {
            if ((rc = launch_signal(c, P, c->peer[0].connected ? c->peer[0].flags + 1 : nullptr, c->peer[1].connected ? c->peer[1].flags + 0 : nullptr, ka_seq))) return rc;
            if ((rc = launch_signal(c, P, c->peer[0].connected ? c->peer[0].flags + 3 : nullptr, c->peer[1].connected ? c->peer[1].flags + 2 : nullptr, ka_seq))) return rc;
        }
        if (two && ((rc = ev_make(&sb.fed0)) || (rc = ev_record(sb.fed0, P)) || (rc = ev_wait(Q, sb.fed0)))) return rc;
        if ((rc = launch_upper_levels(c, Q, sb, abm))) return rc;
        for (int l = c->L; l >= 1; l--) if ((rc = launch_collapse(c, Q, l, sb, abm))) return rc;
        if (c->L == 0 && (rc = launch_finalize_l0(c, Q, sb))) return rc;
        if ((rc = ev_make(&sb.done)) || (rc = ev_record(sb.done, Q))) return rc;
        trace_mark(c, Q, "slice end, row", sb.rows.hi);
    }
    if (two) {
        // join: the canvas stream is the one callers synchronise on
        if ((rc = ev_wait(c->stream, c->subs.back().done))) return rc;
    }
    return composite_epilogue(c);
}

struct BlobHeader {
    uint32_t magic, version;
    int64_t pid;
    int32_t device, L, pw, ph, band_lo, band_hi, nframes, pad;
    uint64_t flags_ptr;
    unsigned char flags_handle[64];
    uint64_t uid; int32_t gen, pad2;
};
struct BlobFrame {
    int32_t idx, ry, rh, rw, gp1, pad;
    uint64_t g_off, w_off, pyr_ptr;
    unsigned char pyr_handle[64];
};
const uint32_t BLOB_MAGIC = 0x50325044u;   // "DP2P"

void peer_close(ds_canvas* c, int side) {
    ds_canvas::Peer& pr = c->peer[side];
#if DS_CUDA
    for (void* m : pr.mapped) cudaIpcCloseMemHandle(m);
    cudaGetLastError();
#endif
    pr = ds_canvas::Peer();
}

int check_device() {
#if DS_CUDA
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        cudaGetLastError();
        return fail(DS_ERR_NO_DEVICE, "no usable CUDA device (%s); libdronestitch_cuda has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    }
#endif
    return DS_OK;
}

// cv::getGaussianKernel(n, sigma, CV_32F) (bit-exact variant): values exp(-x^2 / (2 sigma^2)) for the lower half, their
// sum doubled + the centre's 1, every tap scaled by 1 / sum in double, rounded to float once.
int gaussian_kernel_f32(double sigma, float* k, int* radius) {
    const int n = (int)lrint(sigma * 8.0 + 1.0) | 1;   // float depth: cvRound(sigma * 4 * 2 + 1) | 1
    const int R = n / 2;
    if (sigma <= 0.0 || R > DS_SOFT_MAXR) return fail(DS_ERR_UNSUPPORTED, "soft_sigma %g needs a kernel of %d taps (supported: sigma in (0, 10])", sigma, n);
    const double scale2x = -0.125 / (sigma * sigma);
    double vals[DS_SOFT_MAXR], sum = 0.0;
    int x = 1 - n;
    for (int i = 0; i < R; i++, x += 2) { vals[i] = exp((double)(x * x) * scale2x); sum += vals[i]; }
    sum *= 2.0; sum += 1.0;
    const double mul = 1.0 / sum;
    for (int i = 0; i < R; i++) k[i] = k[n - 1 - i] = (float)(vals[i] * mul);
    k[R] = (float)mul;
    *radius = R;
    return DS_OK;
}

// cv::resize(INTER_NEAREST): source index of every destination column / row.
void nearest_table(int s_, int d_, int* idx) {
    const double inv_scale = (double)d_ / (double)s_;
    const double ifx = 1.0 / inv_scale;
    for (int v = 0; v < d_; v++) idx[v] = std::min((int)floor((double)v * ifx), s_ - 1);
}

// Optional per-frame inputs (ds_frame_opts): seam masks, the global stage's content / soft masks, gains.
// The frame's geometry and source buffer are in place; the pixels may still be on their way (they are waited for only
// when the content mask reads them).
int apply_opts(ds_canvas* c, Frame& f, int idx, const ds_frame_opts* opts) {
    int rc;
    const uint32_t fl = opts ? opts->flags : 0u;
    const bool want_content = (fl & DS_MASK_CONTENT) != 0, want_soft = (fl & DS_MASK_SOFT) != 0;
    const bool has_low = opts && opts->seam_lowres, has_full = opts && opts->seam_mask && !has_low;
    const bool nearest = has_low && (fl & DS_SEAM_NEAREST);
    const size_t plane = (size_t)f.bw * f.bh;
    SoftMaskParams sp;
    memset(&sp, 0, sizeof(sp));
    // arguments are checked before anything about the frame changes
    if (want_soft && (rc = gaussian_kernel_f32(opts->soft_sigma > 0.0 ? opts->soft_sigma : 10.0, sp.k, &sp.R))) return rc;
    if (has_low && (opts->seam_lowres_w <= 0 || opts->seam_lowres_h <= 0)) return fail(DS_ERR_BAD_ARG, "seam_lowres size %dx%d", opts->seam_lowres_w, opts->seam_lowres_h);
    if (opts && opts->gain_blocks && (opts->gain_blocks_w <= 0 || opts->gain_blocks_h <= 0)) return fail(DS_ERR_BAD_ARG, "gain_blocks size %dx%d", opts->gain_blocks_w, opts->gain_blocks_h);
    // a block map with a single row or column (a frame under 33 px) takes another code path inside cv::resize that is not pinned
    if (opts && opts->gain_blocks && (opts->gain_blocks_w == 1 || opts->gain_blocks_h == 1))
        return fail(DS_ERR_UNSUPPORTED, "gain_blocks %dx%d: single-row / single-column block maps are not supported, pass the resized map as gain_map", opts->gain_blocks_w, opts->gain_blocks_h);
    // temporaries live in the handle's scratch slots (no allocation per call); the caller's buffers are only borrowed, so
    // these (rare) paths drain the upload stream before returning
    uint8_t* d_low = nullptr; int* d_tab = nullptr; uint8_t* d_tmp = nullptr;
    bool drain = false;
    auto cleanup = [&](int code) {
        if (drain) { const int r2 = stream_sync(c->up); if (!code) code = r2; }
        return code;
    };
    auto scratch = [&](int slot, size_t bytes, void** out) {
        const int r = grow(c, &c->scratch[slot], &c->scratch_cap[slot], bytes);
        *out = c->scratch[slot];
        return r;
    };
    if (has_low) {
        const int sw = opts->seam_lowres_w, sh = opts->seam_lowres_h;
        const size_t sst = opts->seam_lowres_stride ? opts->seam_lowres_stride : (size_t)sw;
        if ((rc = grow(c, (void**)&f.d_seam, &f.seam_cap, plane))) return rc;
        std::vector<int> tab(2 * ((size_t)f.bw + f.bh));
        if (nearest) {
            // stitch_global.cpp:649-655: resize(seam_masks[i], warped size, INTER_NEAREST), threshold(> 1)
            nearest_table(sw, f.bw, tab.data());
            nearest_table(sh, f.bh, tab.data() + f.bw);
        } else {
            // composePanorama: dilate(masks_warped[i]) -> resize(mask_warped.size(), INTER_LINEAR_EXACT) -> AND
            // coefficient tables in double on the host (OpenCV computes them in softdouble)
            auto coefs = [](int s_, int d_, int* idx, int* c1) {
                const double scale = (double)s_ / (double)d_;
                for (int v = 0; v < d_; v++) {
                    const double fv = scale * ((double)v + 0.5) - 0.5;
                    int i = (int)floor(fv);
                    int k1 = (int)lrint((fv - (double)i) * 256.0);
                    if (i < 0 || s_ <= 1) { i = 0; k1 = 0; }
                    if (i >= s_ - 1) { i = s_ - 1; k1 = 0; }
                    idx[v] = i; c1[v] = k1;
                }
            };
            coefs(sw, f.bw, tab.data(), tab.data() + f.bw);
            coefs(sh, f.bh, tab.data() + 2 * f.bw, tab.data() + 2 * f.bw + f.bh);
        }
        if ((rc = scratch(0, (size_t)sw * sh, (void**)&d_low))) return rc;
        if ((rc = scratch(1, tab.size() * sizeof(int), (void**)&d_tab))) return rc;
        drain = true;
        if ((rc = h2d_2d(d_low, (size_t)sw, opts->seam_lowres, sst, (size_t)sw, (size_t)sh, c->up))) return cleanup(rc);
        if ((rc = h2d(d_tab, tab.data(), tab.size() * sizeof(int), c->up))) return cleanup(rc);
        if (!nearest) {
            SeamUpParams up{d_low, sw, sh, sw, d_tab, d_tab + f.bw, d_tab + 2 * f.bw, d_tab + 2 * f.bw + f.bh, f.d_seam, f.bw, f.bh};
            if ((rc = launch<SeamUpBody, 256>(up, ((long long)plane + SeamUpBody::PER_BLOCK - 1) / SeamUpBody::PER_BLOCK, c->up, 0))) return cleanup(rc);
        }
    } else if (has_full) {
        if ((rc = grow(c, (void**)&f.d_seam, &f.seam_cap, plane))) return rc;
        if ((rc = h2d_2d(f.d_seam, (size_t)f.bw, opts->seam_mask, opts->seam_mask_stride ? opts->seam_mask_stride : (size_t)f.bw,
                         (size_t)f.bw, (size_t)f.bh, c->up))) return rc;
    } else if (f.d_seam && !want_content && !want_soft) {
        dev_free(f.d_seam); c->device_bytes -= (int64_t)f.seam_cap; f.d_seam = nullptr; f.seam_cap = 0;
    }
    if (!want_content && f.d_content) {
        dev_free(f.d_content); c->device_bytes -= (int64_t)f.content_cap; f.d_content = nullptr; f.content_cap = 0;
    }
    if ((rc = fill_frame_dev(c, f))) return cleanup(rc);
    if (want_content || want_soft || nearest) {
        // the global stage's mask chain (ds_mask_kernels.h), on the upload stream behind the copies it reads
        MaskPrepParams mp;
        memset(&mp, 0, sizeof(mp));
        mp.F = f.dev;
        mp.want_content = want_content ? 1 : 0;
        if (nearest) { mp.low = d_low; mp.low_pitch = opts->seam_lowres_w; mp.ix = d_tab; mp.iy = d_tab + f.bw; }
        else if (has_low || has_full) { mp.seam = f.d_seam; mp.seam_pitch = f.bw; }
        mp.binarize = (nearest || want_soft) ? 1 : 0;
        if ((rc = grow(c, (void**)&f.d_seam, &f.seam_cap, plane))) return cleanup(rc);
        if (want_content) {
            if ((rc = grow(c, (void**)&f.d_content, &f.content_cap, plane))) return cleanup(rc);
            // the content mask reads the frame's pixels: bring them in now and queue behind their expansion
            if (f.partial) return cleanup(fail(DS_ERR_STATE, "frame %d is only partially resident", idx));
            if (f.pend.left > 0 && (rc = issue_rows(c, idx, 0, f.h - 1))) return cleanup(rc);
            if ((rc = ev_make(&c->ev_chunks)) || (rc = ev_record(c->ev_chunks, c->xp)) || (rc = ev_wait(c->up, c->ev_chunks))) return cleanup(rc);
        }
        mp.content_out = f.d_content;
        // with a soft mask to follow, the binary plane goes to scratch and the blurred plane lands in the frame's mask
        if (want_soft && (rc = scratch(2, plane, (void**)&d_tmp))) return cleanup(rc);
        mp.out = want_soft ? d_tmp : f.d_seam;
        drain = true;
        if ((rc = launch<MaskPrepBody, 256>(mp, ((long long)plane + MaskPrepBody::PER_BLOCK - 1) / MaskPrepBody::PER_BLOCK, c->up, 0))) return cleanup(rc);
        if (want_soft) {
            sp.bin = d_tmp; sp.bin_pitch = f.bw; sp.out = f.d_seam; sp.out_pitch = f.bw; sp.w = f.bw; sp.h = f.bh;
            const long long tiles = (long long)((f.bw + SoftMaskBody::T - 1) / SoftMaskBody::T) * ((f.bh + SoftMaskBody::T - 1) / SoftMaskBody::T);
            if ((rc = launch<SoftMaskBody, 256>(sp, tiles, c->up, SoftMaskBody::smem_bytes()))) return cleanup(rc);
        }
        f.dev.seam = f.d_seam; f.dev.seam_pitch = f.bw;
    }
    if ((rc = cleanup(DS_OK))) return rc;
    if (opts && opts->gain_blocks) {
        // BlocksGainCompensator::apply: resize(gain_maps_[i], image size, INTER_LINEAR) on the device (fractions in double here)
        const int gw = opts->gain_blocks_w, gh = opts->gain_blocks_h;
        const size_t gst = opts->gain_blocks_stride ? opts->gain_blocks_stride : (size_t)gw * sizeof(float);
        if ((rc = grow(c, (void**)&f.d_gainmap, &f.gainmap_cap, (size_t)f.bw * f.bh * sizeof(float)))) return rc;
        std::vector<int> tab(2 * ((size_t)f.bw + f.bh));   // ix | ax (float bits) | iy | ay
        auto coefs = [](int s_, int d_, int* idx, int* frac_bits) {
            const double scale = (double)s_ / (double)d_;
            for (int v = 0; v < d_; v++) {
                const double fv = ((double)v + 0.5) * scale - 0.5;
                int i = (int)floor(fv);
                float fr = (float)(fv - (double)i);
                if (i < 0) { i = 0; fr = 0.f; }
                if (i >= s_ - 1) { i = s_ - 1; fr = 0.f; }
                idx[v] = i; memcpy(&frac_bits[v], &fr, 4);
            }
        };
        coefs(gw, f.bw, tab.data(), tab.data() + f.bw);
        coefs(gh, f.bh, tab.data() + 2 * f.bw, tab.data() + 2 * f.bw + f.bh);
        float* d_blocks = nullptr; int* d_gtab = nullptr;
        if ((rc = grow(c, &c->scratch[0], &c->scratch_cap[0], (size_t)gw * gh * sizeof(float)))) return rc;
        if ((rc = grow(c, &c->scratch[1], &c->scratch_cap[1], tab.size() * sizeof(int)))) return rc;
        d_blocks = (float*)c->scratch[0]; d_gtab = (int*)c->scratch[1];
        if ((rc = h2d_2d(d_blocks, (size_t)gw * sizeof(float), opts->gain_blocks, gst, (size_t)gw * sizeof(float), (size_t)gh, c->up))) return rc;
        if ((rc = h2d(d_gtab, tab.data(), tab.size() * sizeof(int), c->up))) { stream_sync(c->up); return rc; }
        GainResizeParams gp{d_blocks, gw, gh, gw, d_gtab, (const float*)(d_gtab + f.bw), d_gtab + 2 * f.bw, (const float*)(d_gtab + 2 * f.bw + f.bh),
                            f.d_gainmap, f.bw, f.bh};
        const long long n = (long long)f.bw * f.bh;
        rc = launch<GainResizeBody, 256>(gp, (n + GainResizeBody::PER_BLOCK - 1) / GainResizeBody::PER_BLOCK, c->up, 0);
        const int rs = stream_sync(c->up);   // the caller's block map and the host tables are only borrowed
        if (rc || rs) return rc ? rc : rs;
        f.dev.gainmap = f.d_gainmap; f.dev.gainmap_pitch = f.bw;
    } else if (opts && opts->gain_map) {
        const size_t gst = opts->gain_map_stride ? opts->gain_map_stride : (size_t)f.bw * sizeof(float);
        if ((rc = grow(c, (void**)&f.d_gainmap, &f.gainmap_cap, (size_t)f.bw * f.bh * sizeof(float)))) return rc;
        if ((rc = h2d_2d(f.d_gainmap, (size_t)f.bw * sizeof(float), opts->gain_map, gst, (size_t)f.bw * sizeof(float), (size_t)f.bh, c->up))) return rc;
        f.dev.gainmap = f.d_gainmap; f.dev.gainmap_pitch = f.bw;
    } else if (f.d_gainmap) {
        dev_free(f.d_gainmap); c->device_bytes -= (int64_t)f.gainmap_cap; f.d_gainmap = nullptr; f.gainmap_cap = 0;
    }
    if (opts && opts->channel_gain) {
        f.dev.has_gain = 1;
        for (int k = 0; k < 3; k++) f.dev.gain[k] = opts->channel_gain[k];
    }
    if (opts && opts->compensator_gain) {
        f.dev.has_cgain = 1;
        for (int k = 0; k < 3; k++) f.dev.cgain[k] = opts->compensator_gain[k];
    }
    f.dev.any_gain = (f.dev.has_gain || f.dev.has_cgain || f.dev.gainmap) ? 1 : 0;
    return DS_OK;
}

int do_upload(ds_canvas* c, int idx, const void* bgr, bool on_device, int w, int h, size_t stride,
              const ds_transform* xf, const ds_frame_opts* opts) {
    if (!c || !bgr || !xf) return fail(DS_ERR_BAD_ARG, "null argument");
    if (idx < 0 || idx >= DS_MAX_FRAMES) return fail(DS_ERR_BAD_ARG, "frame_idx %d out of range [0, %d)", idx, DS_MAX_FRAMES);
    if (w <= 0 || h <= 0 || w > 32767 || h > 32767) return fail(DS_ERR_BAD_ARG, "frame size %dx%d (cv::remap needs < 32768)", w, h);
    if (stride < (size_t)w * 3) return fail(DS_ERR_BAD_ARG, "stride %zu < 3*w", stride);
    if (c->stage_open) return fail(DS_ERR_STATE, "a composite is half done: call ds_composite_stage(c, 1) before uploading");
    int rc;
    if ((rc = set_device(c))) return rc;
    int pl[4];
    if ((rc = placement(xf, w, h, pl))) return rc;
    // the frame must lie inside the canvas ROI (prepare(resultRoi(corners, sizes)) guarantees it in the reference);
    // checked before anything of the handle changes, so that a refused upload leaves the previous frame in place
    if (pl[0] < c->desc.x || pl[1] < c->desc.y || (long long)pl[0] + pl[2] > (long long)c->desc.x + c->desc.width ||
        (long long)pl[1] + pl[3] > (long long)c->desc.y + c->desc.height)
        return fail(DS_ERR_BAD_ARG, "frame %d bbox (%d,%d %dx%d) leaves the canvas ROI (%d,%d %dx%d)", idx, pl[0], pl[1],
                    pl[2], pl[3], c->desc.x, c->desc.y, c->desc.width, c->desc.height);
    if ((size_t)idx >= c->frames.size()) c->frames.resize((size_t)idx + 1);
    Frame& f = c->frames[(size_t)idx];
    // from here on the frame's state changes: whatever fails below, the launch metadata must be rebuilt
    struct DirtyOnFailure { ds_canvas* c; bool armed; ~DirtyOnFailure() { if (armed) { c->dirty = true; c->composited = false; } } } guard{c, true};
    const bool async = opts && (opts->flags & DS_UPLOAD_ASYNC);
    const bool was_used = f.used;
    const FrameDev before = f.dev;
    // the frame's buffers may still be read by the previous composite
    if (c->ev_done_valid && ((rc = ev_wait(c->up, c->ev_done)) || (rc = ev_wait(c->xp, c->ev_done)))) return rc;
    f.used = true; f.w = w; f.h = h; f.xf = *xf;
    f.corner_x = pl[0]; f.corner_y = pl[1]; f.bw = pl[2]; f.bh = pl[3];
    // source: dense BGR rows -> staging slot -> BGRX expansion, in chunks of source rows
    f.pitch = (w + 31) & ~31;
    if ((rc = grow(c, (void**)&f.d_src, &f.src_cap, (size_t)f.pitch * h * sizeof(uint32_t)))) return rc;
#if !DS_CUDA
    memset(f.d_src, 0xA5, (size_t)f.pitch * h * sizeof(uint32_t));   // emulator: a slice reading rows that were not delivered yet shows
#endif
    if (!on_device) {
        const size_t need = (size_t)std::min(c->chunk_rows, h) * w * 3;
        if (need > c->slot_cap) {
            for (int k = 0; k < ds_canvas::NSLOT; k++) {
                size_t cap = c->slot_cap;
                if ((rc = grow(c, (void**)&c->d_slot[k], &cap, need))) return rc;   // cudaFree waits for chunks in flight
                c->slot_used[k] = false;
            }
            c->slot_cap = need;
        }
    }
    if (f.pend.left > 0) c->n_pending--;   // replaced before it was copied
    f.pend.src = (const uint8_t*)bgr; f.pend.stride = stride; f.pend.on_device = on_device;
    f.pend.nchunks = (h + c->chunk_rows - 1) / c->chunk_rows; f.pend.left = f.pend.nchunks;
    f.pend.issued.assign((size_t)f.pend.nchunks, 0);
    f.partial = false;
    c->n_pending++;
    // Without DS_UPLOAD_ASYNC the caller's buffers are only borrowed for the call: copy now, drain before returning.
    if (!async && (rc = issue_rows(c, idx, 0, h - 1))) return rc;
    // blend-mode specific geometry and buffers
    if (c->desc.blend_mode == DS_BLEND_MULTIBAND) {
        dsgeo::feed_roi(c->desc.x, c->desc.y, c->pw, c->ph, c->L, f.corner_x, f.corner_y, f.bw, f.bh, f.rx, f.ry, f.rw, f.rh);
        {
            void* const before_pyr = f.d_pyr;
            if ((rc = grow(c, &f.d_pyr, &f.pyr_cap, pyr_bytes(c, f)))) return rc;
            if (before_pyr && f.d_pyr != before_pyr) { c->p2p_gen++; reg_set(c); }   // pointers exported earlier are stale now
        }
    } else {
        if ((rc = grow(c, (void**)&f.d_mbits, &f.mbits_cap, (size_t)((f.bw + 31) / 32) * f.bh * sizeof(uint32_t)))) return rc;
        if ((rc = grow(c, (void**)&f.d_dist, &f.dist_cap, (size_t)((f.bw + 15) & ~15) * f.bh))) return rc;
    }
    if ((rc = apply_opts(c, f, idx, opts))) return rc;
    if (async) c->async_pending = true;
    else if ((rc = stream_sync(c->up)) || (rc = stream_sync(c->xp))) return rc;
    // same geometry, buffers and gains as before (a new image for the same slot): the launch metadata stands
    guard.armed = false;
    if (!was_used || memcmp(&before, &f.dev, sizeof(FrameDev)) != 0) c->dirty = true;
    c->composited = false;
    return DS_OK;
}

int smallest_feather_radius(float sharpness) {
    for (int r = 1; r <= 4096; r++)
        if ((float)r * sharpness >= 1.f) return r;
    return 4097;
}

}  // namespace

// ====================================================================== C ABI

extern "C" {

DS_API const char* ds_last_error(void) { return g_err.c_str(); }
DS_API const char* ds_version(void) {
#if DS_CUDA
    return "libdronestitch_cuda " DS_VERSION_STRING " (sm_100a)";
#else
    return "libdronestitch_emu " DS_VERSION_STRING " (TEST-ONLY CPU emulation of the kernels)";
#endif
}

DS_API int ds_warp_roi(const ds_transform* xf, int src_w, int src_h, int32_t out_xywh[4]) {
    if (!xf || !out_xywh || src_w <= 0 || src_h <= 0) return fail(DS_ERR_BAD_ARG, "null / empty argument");
    int o[4];
    int rc = placement(xf, src_w, src_h, o);
    if (rc) return rc;
    for (int i = 0; i < 4; i++) out_xywh[i] = o[i];
    return DS_OK;
}

static int band_plan_from_desc(const ds_canvas_desc* d, int* L_out, int* pw_out, int* ph_out, Range* band_out, int* lw, int* lh,
                               Range* acc, Range* own) {
    if (d->width <= 0 || d->height <= 0) return fail(DS_ERR_BAD_ARG, "empty canvas %dx%d", d->width, d->height);
    int L = 0, pw = d->width, ph = d->height;
    if (d->blend_mode == DS_BLEND_MULTIBAND) {
        if (d->num_bands < 0 || d->num_bands >= DS_MAXL) return fail(DS_ERR_BAD_ARG, "num_bands %d not in [0, %d]", d->num_bands, DS_MAXL - 1);
        L = dsgeo::effective_bands(d->num_bands, d->width, d->height);
        pw = dsgeo::pad_to(d->width, 1 << L); ph = dsgeo::pad_to(d->height, 1 << L);
    } else if (d->blend_mode != DS_BLEND_FEATHER) {
        return fail(DS_ERR_BAD_ARG, "unknown blend mode %d", d->blend_mode);
    }
    Range band{0, ph};
    if (d->band_y1 > 0) {
        band.lo = d->band_y0; band.hi = std::min(d->band_y1, ph);
        if (band.lo < 0 || band.lo >= band.hi) return fail(DS_ERR_BAD_ARG, "bad band [%d, %d)", d->band_y0, d->band_y1);
        const int m = 1 << L;
        if (band.lo % m || (band.hi % m && band.hi != ph)) return fail(DS_ERR_BAD_ARG, "band edges must be multiples of %d", m);
    }
    lw[0] = pw; lh[0] = ph;
    for (int l = 1; l <= L; l++) { lw[l] = (lw[l - 1] + 1) / 2; lh[l] = (lh[l - 1] + 1) / 2; }
    plan_rows(L, lh, band, acc, own);
    *L_out = L; *pw_out = pw; *ph_out = ph; *band_out = band;
    return DS_OK;
}

DS_API int ds_frame_touches_band(const ds_canvas_desc* desc, const int32_t fr[4]) {
    if (!desc || !fr) return 0;
    int L, pw, ph, lw[DS_MAXL], lh[DS_MAXL];
    Range band, acc[DS_MAXL], own[DS_MAXL];
    if (band_plan_from_desc(desc, &L, &pw, &ph, &band, lw, lh, acc, own)) return 0;
    int y0, y1;
    if (desc->blend_mode == DS_BLEND_MULTIBAND) {
        int rx, ry, rw, rh;
        dsgeo::feed_roi(desc->x, desc->y, pw, ph, L, fr[0], fr[1], fr[2], fr[3], rx, ry, rw, rh);
        y0 = ry; y1 = ry + rh;
        return (y0 < own[0].hi && y1 > own[0].lo) ? 1 : 0;
    }
    y0 = fr[1] - desc->y; y1 = y0 + fr[3];
    return (y0 < band.hi && y1 > band.lo) ? 1 : 0;
}

DS_API int ds_global_blend_bands(int canvas_w, int canvas_h, int configured_bands) {
    if (canvas_w <= 0 || canvas_h <= 0) return -1;
    // std::ceil(std::log2(double(max))) - 1, capped at 12 (stitch_global.cpp:632-634)
    const int auto_bands = std::min(12, (int)std::ceil(std::log2((double)std::max(canvas_w, canvas_h))) - 1);
    return std::max(std::max(5, configured_bands), auto_bands);
}

DS_API int ds_plan_row_bands(const ds_canvas_desc* desc, const int32_t* fr, int n_frames, int n_bands, int32_t* edges) {
    if (!desc || !edges || n_bands < 1 || n_frames < 0 || (n_frames > 0 && !fr)) return fail(DS_ERR_BAD_ARG, "null / empty argument");
    ds_canvas_desc d = *desc;
    d.band_y0 = d.band_y1 = 0;
    int L, pw, ph, lw[DS_MAXL], lh[DS_MAXL];
    Range band, acc[DS_MAXL], own[DS_MAXL];
    int rc = band_plan_from_desc(&d, &L, &pw, &ph, &band, lw, lh, acc, own);
    if (rc) return rc;
    const int m = d.blend_mode == DS_BLEND_MULTIBAND ? (1 << L) : FeatherBody::TH;
    const int nunits = (ph + m - 1) / m;   // rows are handed out in units of m
    if (n_bands > nunits) return fail(DS_ERR_BAD_ARG, "%d bands for a canvas of %d row units", n_bands, nunits);
    // work per unit: canvas pixels (collapse, stores) + bbox pixels of every frame crossing it (warp, pyramids, accumulate)
    std::vector<double> work((size_t)nunits, 0.0);
    for (int u = 0; u < nunits; u++) work[u] = (double)d.width * std::min(m, ph - u * m);
    for (int i = 0; i < n_frames; i++) {
        const int y0 = fr[4 * i + 1] - d.y, y1 = y0 + fr[4 * i + 3], w = fr[4 * i + 2];
        for (int u = std::max(y0, 0) / m; u < nunits && u * m < y1; u++) {
            const int a = std::max(y0, u * m), b = std::min(y1, (u + 1) * m);
            if (b > a) work[u] += 4.0 * (double)w * (b - a);   // a frame pixel costs about four times a bare canvas pixel
        }
    }
    double total = 0;
    for (double v : work) total += v;
    edges[0] = 0;
    double run = 0;
    int u = 0;
    for (int k = 1; k < n_bands; k++) {
        const double target = total * k / n_bands;
        while (u < nunits && run + work[u] * 0.5 < target) run += work[u++];
        u = std::max(u, edges[k - 1] / m + 1);            // every band keeps at least one unit ...
        u = std::min(u, nunits - (n_bands - k));          // ... and leaves one for each band below
        run = 0;
        for (int q = 0; q < u; q++) run += work[q];
        edges[k] = u * m;
    }
    edges[n_bands] = ph;
    return DS_OK;
}

DS_API int ds_create_canvas(const ds_canvas_desc* desc, ds_canvas** out) {
    if (!desc || !out) return fail(DS_ERR_BAD_ARG, "null argument");
    *out = nullptr;
    int rc;
    if ((rc = check_device())) return rc;
    if (desc->out_format != DS_OUT_BGR8 && desc->out_format != DS_OUT_BGRA8) return fail(DS_ERR_BAD_ARG, "unknown out_format %d", desc->out_format);
    ds_canvas* c = new (std::nothrow) ds_canvas();
    if (!c) return fail(DS_ERR_OOM, "host allocation failed");
    c->desc = *desc;
    for (int l = 0; l < DS_MAXL; l++) { c->d_lvl[l] = nullptr; c->d_lvl_alloc[l] = nullptr; c->lw[l] = c->lh[l] = 0; c->lvl_rows[l] = Range{0, 0}; }
    Range acc[DS_MAXL], own[DS_MAXL];
    if ((rc = band_plan_from_desc(desc, &c->L, &c->pw, &c->ph, &c->band, c->lw, c->lh, acc, own))) { delete c; return rc; }
    if (desc->blend_mode == DS_BLEND_FEATHER) {
        if (!(desc->sharpness > 0.f)) { delete c; return fail(DS_ERR_BAD_ARG, "sharpness must be > 0"); }
        c->feather_R = smallest_feather_radius(desc->sharpness);
        if (c->feather_R > FeatherBody::RMAX) {
            const int need = c->feather_R;
            delete c;
            return fail(DS_ERR_UNSUPPORTED, "feather sharpness %g needs a %d px window (max %d)", desc->sharpness, need, FeatherBody::RMAX);
        }
    }
#if DS_CUDA
    if (cudaSetDevice(desc->device) != cudaSuccess) { delete c; cudaGetLastError(); return fail(DS_ERR_NO_DEVICE, "cudaSetDevice(%d) failed", desc->device); }
    if (desc->stream) { c->stream = (cudaStream_t)desc->stream; }
    else {
        if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return fail(DS_ERR_CUDA, "cudaStreamCreate failed"); }
        c->own_stream = true;
    }
    {
        // uploads get the higher priority: the expansion kernel of an arriving frame must not queue behind the
        // blocks of a running slice, or the copy engine idles
        int lo_prio = 0, hi_prio = 0;   // numerically lower = higher priority
        cudaDeviceGetStreamPriorityRange(&lo_prio, &hi_prio);
        const int mid_prio = hi_prio < lo_prio - 1 ? hi_prio + 1 : hi_prio;
        if (cudaStreamCreateWithPriority(&c->up, cudaStreamNonBlocking, hi_prio) != cudaSuccess ||
            cudaStreamCreateWithFlags(&c->dl, cudaStreamNonBlocking) != cudaSuccess ||
            cudaStreamCreateWithPriority(&c->xp, cudaStreamNonBlocking, hi_prio) != cudaSuccess ||
            cudaStreamCreateWithPriority(&c->bulk, cudaStreamNonBlocking, lo_prio) != cudaSuccess ||
            cudaStreamCreateWithPriority(&c->tail, cudaStreamNonBlocking, mid_prio) != cudaSuccess) {
            ds_destroy_canvas(c);
            return fail(DS_ERR_CUDA, "cudaStreamCreate failed");
        }
    }
    cudaEventCreate(&c->ev0); cudaEventCreate(&c->ev1);
#endif
    if (const char* e = getenv("DS_UPLOAD_CHUNK_ROWS")) { if (atoi(e) > 0) c->chunk_rows = atoi(e); }
    // output rows [band.lo, out_hi)
    c->out_hi = std::min(c->band.hi, desc->height);
    const int out_rows = std::max(c->out_hi - c->band.lo, 0);
    const int bpp = desc->out_format == DS_OUT_BGRA8 ? 4 : 3;
    c->out_pitch = ((size_t)desc->width * bpp + 255) & ~(size_t)255;
    size_t cap = 0;
    rc = grow(c, (void**)&c->d_out, &cap, c->out_pitch * (size_t)std::max(out_rows, 1));
    if (!rc && desc->out_format == DS_OUT_BGR8) {
        c->mask_pitch = ((size_t)desc->width + 255) & ~(size_t)255;
        cap = 0;
        rc = grow(c, (void**)&c->d_mask, &cap, c->mask_pitch * (size_t)std::max(out_rows, 1));
    }
    if (!rc && desc->blend_mode == DS_BLEND_MULTIBAND) {
        for (int l = 0; l <= c->L && !rc; l++) {
            LevelPlan& pl = c->plan[l];
            pl.TW = level_tile_w(l); pl.TH = level_tile_h(l);
            pl.tiles_x = (c->lw[l] + pl.TW - 1) / pl.TW; pl.tiles_y = (c->lh[l] + pl.TH - 1) / pl.TH;
            pl.own = own[l]; pl.acc = acc[l];
            if (l == 0) c->own0_recompute = own[0];
            // rows stored: what the collapse touches (acc) — the feed writes only those
            c->lvl_rows[l] = acc[l];
            cap = 0;
            const size_t rows = (size_t)std::max(acc[l].hi - acc[l].lo, 1);
            rc = grow(c, (void**)&c->d_lvl_alloc[l], &cap, rows * c->lw[l] * sizeof(px16));
            if (!rc) c->d_lvl[l] = c->d_lvl_alloc[l] - (size_t)acc[l].lo * c->lw[l];
        }
    } else if (!rc) {
        LevelPlan& pl = c->plan[0];
        pl.TW = pl.TH = 0;
        pl.tiles_x = (desc->width + FeatherBody::TW - 1) / FeatherBody::TW;
        pl.tiles_y = (desc->height + FeatherBody::TH - 1) / FeatherBody::TH;
        pl.own = Range{c->band.lo, c->out_hi}; pl.acc = pl.own;
    }
    if (rc) { ds_destroy_canvas(c); return rc; }
    *out = c;
    return DS_OK;
}

DS_API void ds_destroy_canvas(ds_canvas* c) {
    if (!c) return;
    reg_drop(c);
    set_device(c);
#if DS_CUDA
    if (c->up) cudaStreamSynchronize(c->up);
    if (c->xp) cudaStreamSynchronize(c->xp);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->bulk) cudaStreamSynchronize(c->bulk);
    if (c->tail) cudaStreamSynchronize(c->tail);
    if (c->dl) cudaStreamSynchronize(c->dl);
#endif
    for (Frame& f : c->frames) {
        dev_free(f.d_src); dev_free(f.d_pyr); dev_free(f.d_mbits); dev_free(f.d_dist); dev_free(f.d_seam); dev_free(f.d_content); dev_free(f.d_gainmap);
    }
    for (int l = 0; l < DS_MAXL; l++) dev_free(c->d_lvl_alloc[l]);
    for (SubBand& sb : c->subs) { ev_drop(sb.done); ev_drop(sb.fed); ev_drop(sb.fed0); }
    dev_free(c->d_out); dev_free(c->d_mask); dev_free(c->d_meta);
    for (int k = 0; k < 3; k++) dev_free(c->scratch[k]);
    peer_close(c, 0); peer_close(c, 1);
    dev_free(c->d_flags);
    for (int k = 0; k < ds_canvas::NSLOT; k++) { dev_free(c->d_slot[k]); ev_drop(c->slot_copied[k]); ev_drop(c->slot_free[k]); }
    ev_drop(c->ev_chunks); ev_drop(c->ev_opts);
    pinned_free(c->h_meta);
    ev_drop(c->ev_done); ev_drop(c->ev_meta); ev_drop(c->ev_start);
#if DS_CUDA
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    for (cudaEvent_t e : c->prof_ev) cudaEventDestroy(e);
    if (c->up) cudaStreamDestroy(c->up);
    if (c->xp) cudaStreamDestroy(c->xp);
    if (c->dl) cudaStreamDestroy(c->dl);
    if (c->bulk) cudaStreamDestroy(c->bulk);
    if (c->tail) cudaStreamDestroy(c->tail);
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
#endif
    delete c;
}

DS_API int ds_upload_frame(ds_canvas* c, int frame_idx, const uint8_t* bgr, int w, int h, size_t stride,
                           const ds_transform* xf, const ds_frame_opts* opts) {
    return do_upload(c, frame_idx, bgr, false, w, h, stride, xf, opts);
}

DS_API int ds_upload_frame_device(ds_canvas* c, int frame_idx, const void* dev_bgr, int w, int h, size_t stride,
                                  const ds_transform* xf, const ds_frame_opts* opts) {
    return do_upload(c, frame_idx, dev_bgr, true, w, h, stride, xf, opts);
}

DS_API int ds_update_frame_opts(ds_canvas* c, int frame_idx, const ds_frame_opts* opts) {
    if (!c) return fail(DS_ERR_BAD_ARG, "null canvas");
    if (frame_idx < 0 || (size_t)frame_idx >= c->frames.size() || !c->frames[(size_t)frame_idx].used)
        return fail(DS_ERR_BAD_ARG, "frame %d not uploaded", frame_idx);
    if (c->stage_open) return fail(DS_ERR_STATE, "a composite is half done: call ds_composite_stage(c, 1) first");
    int rc;
    if ((rc = set_device(c))) return rc;
    if ((rc = flush_uploads(c))) return rc;
    // the frame's planes may still be read by the previous composite
    if (c->ev_done_valid && (rc = ev_wait(c->up, c->ev_done))) return rc;
    Frame& f = c->frames[(size_t)frame_idx];
    const FrameDev before = f.dev;
    if ((rc = apply_opts(c, f, frame_idx, opts))) return rc;
    if ((rc = stream_sync(c->up))) return rc;
    if (memcmp(&before, &f.dev, sizeof(FrameDev)) != 0) c->dirty = true;
    c->composited = false;
    return DS_OK;
}

DS_API int ds_download_frame_mask(ds_canvas* c, int frame_idx, int which, uint8_t* out, size_t stride) {
    if (!c || !out) return fail(DS_ERR_BAD_ARG, "null argument");
    if (frame_idx < 0 || (size_t)frame_idx >= c->frames.size() || !c->frames[(size_t)frame_idx].used)
        return fail(DS_ERR_BAD_ARG, "frame %d not uploaded", frame_idx);
    if (which != 0 && which != 1) return fail(DS_ERR_BAD_ARG, "which = %d (0: blend mask, 1: content mask)", which);
    int rc;
    if ((rc = set_device(c))) return rc;
    if ((rc = flush_uploads(c))) return rc;
    Frame& f = c->frames[(size_t)frame_idx];
    const size_t st = stride ? stride : (size_t)f.bw;
    if (st < (size_t)f.bw) return fail(DS_ERR_BAD_ARG, "stride %zu < width %d", st, f.bw);
    if (which == 1) {
        if (!f.d_content) return fail(DS_ERR_STATE, "frame %d was not uploaded with DS_MASK_CONTENT", frame_idx);
        if ((rc = d2h_2d(out, st, f.d_content, (size_t)f.bw, (size_t)f.bw, (size_t)f.bh, c->up))) return rc;
        return stream_sync(c->up);
    }
    const size_t plane = (size_t)f.bw * f.bh;
    uint8_t* d_tmp = nullptr;
    if ((rc = dev_alloc_t(&d_tmp, plane))) return rc;
    MaskPrepParams mp;
    memset(&mp, 0, sizeof(mp));
    mp.F = f.dev; mp.seam = f.d_seam; mp.seam_pitch = f.bw; mp.and_nearest = 1; mp.out = d_tmp;
    rc = launch<MaskPrepBody, 256>(mp, ((long long)plane + MaskPrepBody::PER_BLOCK - 1) / MaskPrepBody::PER_BLOCK, c->up, 0);
    if (!rc) rc = d2h_2d(out, st, d_tmp, (size_t)f.bw, (size_t)f.bw, (size_t)f.bh, c->up);
    const int r2 = stream_sync(c->up);
    dev_free(d_tmp);
    return rc ? rc : r2;
}

DS_API int ds_composite_async(ds_canvas* c) {
    if (!c) return fail(DS_ERR_BAD_ARG, "null canvas");
    int rc;
    if ((rc = set_device(c))) return rc;
    return run_composite(c, -1);
}

DS_API int ds_synchronize(ds_canvas* c) {
    if (!c) return fail(DS_ERR_BAD_ARG, "null canvas");
    int rc;
    if ((rc = set_device(c))) return rc;
    if ((rc = flush_uploads(c))) return rc;
    if ((rc = stream_sync(c->stream))) return rc;
    if ((rc = stream_sync(c->dl))) return rc;
    trace_dump(c);
#if DS_CUDA
    if (c->composited) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, c->ev0, c->ev1) == cudaSuccess) c->last_ms = ms; else cudaGetLastError();
    }
#endif
    return DS_OK;
}

DS_API int ds_composite(ds_canvas* c) {
    int rc = ds_composite_async(c);
    if (rc) return rc;
    return ds_synchronize(c);
}

DS_API int ds_download_tile(ds_canvas* c, int x, int y, int w, int h, uint8_t* out, size_t stride, uint8_t* mask_out, size_t mask_stride) {
    if (!c || !out) return fail(DS_ERR_BAD_ARG, "null argument");
    if (!c->composited) return fail(DS_ERR_STATE, "ds_download_tile before ds_composite");
    if (w <= 0 || h <= 0 || x < 0 || w > c->desc.width - x || y < c->band.lo || y > c->out_hi || h > c->out_hi - y)
        return fail(DS_ERR_BAD_ARG, "tile (%d,%d %dx%d) outside this handle's rows [%d,%d) x [0,%d)", x, y, w, h, c->band.lo, c->out_hi, c->desc.width);
    const int bpp = c->desc.out_format == DS_OUT_BGRA8 ? 4 : 3;
    if (stride < (size_t)w * bpp) return fail(DS_ERR_BAD_ARG, "stride too small");
    int rc;
    if ((rc = set_device(c))) return rc;
    if (mask_out) {
        if (c->desc.out_format == DS_OUT_BGRA8) return fail(DS_ERR_UNSUPPORTED, "BGRA8 canvases carry the mask in alpha");
        if (mask_stride < (size_t)w) return fail(DS_ERR_BAD_ARG, "mask stride too small");
    }
    // slice by slice, each as soon as its rows are final: the copies of the first slices overlap the compute of
    // the later ones (and the uploads those still wait for)
    for (const SubBand& sb : c->subs) {
        const int y0 = std::max(y, sb.rows.lo), y1 = std::min(y + h, sb.rows.hi);
        if (y0 >= y1) continue;
        if ((rc = ev_wait(c->dl, sb.done))) return rc;
        trace_mark(c, c->dl, "d2h begin, row", y0);
        const uint8_t* src = c->d_out + (size_t)(y0 - c->band.lo) * c->out_pitch + (size_t)x * bpp;
        if ((rc = d2h_2d(out + (size_t)(y0 - y) * stride, stride, src, c->out_pitch, (size_t)w * bpp, (size_t)(y1 - y0), c->dl))) return rc;
        if (mask_out) {
            const uint8_t* ms = c->d_mask + (size_t)(y0 - c->band.lo) * c->mask_pitch + x;
            if ((rc = d2h_2d(mask_out + (size_t)(y0 - y) * mask_stride, mask_stride, ms, c->mask_pitch, (size_t)w, (size_t)(y1 - y0), c->dl))) return rc;
        }
    }
    trace_mark(c, c->dl, "d2h end, row", y + h);
    if ((rc = stream_sync(c->dl))) return rc;
#if DS_CUDA
    if (cudaEventQuery(c->ev1) == cudaSuccess) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, c->ev0, c->ev1) == cudaSuccess) c->last_ms = ms; else cudaGetLastError();
    } else {
        cudaGetLastError();
    }
#endif
    return DS_OK;
}

DS_API int ds_warp_frame(int device, const uint8_t* bgr, int w, int h, size_t stride, const ds_transform* xf,
                         int32_t out_xywh[4], uint8_t* out_bgr, uint8_t* out_mask) {
    if (!bgr || !xf || !out_xywh) return fail(DS_ERR_BAD_ARG, "null argument");
    int rc;
    if ((rc = ds_warp_roi(xf, w, h, out_xywh))) return rc;
    if (!out_bgr && !out_mask) return DS_OK;
    if (!out_bgr || !out_mask) return fail(DS_ERR_BAD_ARG, "out_bgr and out_mask go together");
    // a scratch handle whose canvas is the frame's own bbox: the warp is the one every composite runs
    ds_canvas_desc d;
    memset(&d, 0, sizeof(d));
    d.x = out_xywh[0]; d.y = out_xywh[1]; d.width = out_xywh[2]; d.height = out_xywh[3];
    d.blend_mode = DS_BLEND_FEATHER; d.sharpness = 0.02f; d.out_format = DS_OUT_BGR8; d.device = device;
    ds_canvas* c = nullptr;
    if ((rc = ds_create_canvas(&d, &c))) return rc;
    rc = ds_upload_frame(c, 0, bgr, w, h, stride, xf, nullptr);
    if (!rc) rc = ds_debug_get_warped(c, 0, out_bgr, out_mask);
    const std::string keep = g_err;
    ds_destroy_canvas(c);
    if (rc) g_err = keep;
    return rc;
}

DS_API int ds_auto_crop_rect(ds_canvas* c, int32_t out_xywh[4]) {
    if (!c || !out_xywh) return fail(DS_ERR_BAD_ARG, "null argument");
    if (!c->composited) return fail(DS_ERR_STATE, "ds_auto_crop_rect before ds_composite");
    const int W = c->desc.width, H = c->desc.height;
    if (c->band.lo != 0 || c->out_hi < H) return fail(DS_ERR_UNSUPPORTED, "ds_auto_crop_rect needs a whole-canvas handle (this one owns rows [%d,%d))", c->band.lo, c->out_hi);
    int rc;
    if ((rc = set_device(c))) return rc;
    if ((rc = stream_sync(c->stream))) return rc;
    for (const SubBand& sb : c->subs) if (sb.done && (rc = ev_sync(sb.done))) return rc;
    const int cap = 256;   // events per row: up to 128 foreground runs (dark specks inside the content split runs)
    int* d_count = nullptr; int* d_events = nullptr;
    if ((rc = dev_alloc_t(&d_count, (size_t)H))) return rc;
    if ((rc = dev_alloc_t(&d_events, (size_t)H * cap))) { dev_free(d_count); return rc; }
    std::vector<int> count((size_t)H), events((size_t)H * cap);
    rc = dev_zero(d_count, (size_t)H * sizeof(int));
    if (!rc) {
        RowRunsParams rp{c->d_out, c->out_pitch, c->desc.out_format == DS_OUT_BGRA8 ? 4 : 3, W, H, cap, d_count, d_events};
        rc = launch<RowRunsBody, 256>(rp, H, c->stream, 0);
    }
    if (!rc) rc = d2h(count.data(), d_count, (size_t)H * sizeof(int), c->stream);
    if (!rc) rc = d2h(events.data(), d_events, (size_t)H * cap * sizeof(int), c->stream);
    const int r2 = stream_sync(c->stream);
    dev_free(d_count); dev_free(d_events);
    if (rc || r2) return rc ? rc : r2;
    c->launches++;
    // runs per row, in x order
    struct Run { int y, s, e, parent; };
    std::vector<Run> runs;
    std::vector<int> row_first((size_t)H + 1, 0);
    for (int y = 0; y < H; y++) {
        row_first[(size_t)y] = (int)runs.size();
        const int n = count[(size_t)y];
        if (n > cap) return fail(DS_ERR_UNSUPPORTED, "row %d of the canvas has more than %d foreground runs: use cv::findContours on the downloaded panorama", y, cap / 2);
        int* ev = events.data() + (size_t)y * cap;
        std::sort(ev, ev + n);   // x << 1 | is_end: a one-pixel run sorts as start, end
        for (int k = 0; k + 1 < n; k += 2) runs.push_back(Run{y, ev[k] >> 1, ev[k + 1] >> 1, (int)runs.size()});
    }
    row_first[(size_t)H] = (int)runs.size();
    if (runs.empty()) { out_xywh[0] = 0; out_xywh[1] = 0; out_xywh[2] = W; out_xywh[3] = H; return DS_OK; }
    // 8-connected components of the runs (union-find)
    auto find = [&](int i) { while (runs[(size_t)i].parent != i) { runs[(size_t)i].parent = runs[(size_t)runs[(size_t)i].parent].parent; i = runs[(size_t)i].parent; } return i; };
    for (int y = 1; y < H; y++) {
        int a = row_first[(size_t)y - 1];
        const int a_end = row_first[(size_t)y];
        for (int b = row_first[(size_t)y]; b < row_first[(size_t)y + 1]; b++) {
            const Run& rb = runs[(size_t)b];
            while (a < a_end && runs[(size_t)a].e < rb.s - 1) a++;
            for (int k = a; k < a_end && runs[(size_t)k].s <= rb.e + 1; k++) {
                const int ra = find(k), rbb = find(b);
                if (ra != rbb) runs[(size_t)std::max(ra, rbb)].parent = std::min(ra, rbb);
            }
        }
    }
    // per component: bounding box and core pixels (foreground with all four 4-neighbours foreground; never on a border)
    struct Comp { int x0, y0, x1, y1; long long core; };
    std::vector<Comp> comp(runs.size(), Comp{INT32_MAX, INT32_MAX, -1, -1, 0});
    auto overlap_rows = [&](int y, int lo, int hi, std::vector<std::pair<int, int>>& out) {
        out.clear();
        if (y < 0 || y >= H) return;
        for (int k = row_first[(size_t)y]; k < row_first[(size_t)y + 1]; k++) {
            const int s_ = std::max(lo, runs[(size_t)k].s), e_ = std::min(hi, runs[(size_t)k].e);
            if (s_ <= e_) out.push_back({s_, e_});
        }
    };
    std::vector<std::pair<int, int>> up, dn;
    for (size_t i = 0; i < runs.size(); i++) {
        const Run& r = runs[i];
        Comp& cc = comp[(size_t)find((int)i)];
        cc.x0 = std::min(cc.x0, r.s); cc.x1 = std::max(cc.x1, r.e);
        cc.y0 = std::min(cc.y0, r.y); cc.y1 = std::max(cc.y1, r.y);
        if (r.e - r.s >= 2) {
            overlap_rows(r.y - 1, r.s + 1, r.e - 1, up);
            overlap_rows(r.y + 1, r.s + 1, r.e - 1, dn);
            size_t j = 0;
            for (const auto& u : up) {
                while (j < dn.size() && dn[j].second < u.first) j++;
                for (size_t k = j; k < dn.size() && dn[k].first <= u.second; k++)
                    cc.core += std::min(u.second, dn[k].second) - std::max(u.first, dn[k].first) + 1;
            }
        }
    }
    int best = -1;
    for (size_t i = 0; i < comp.size(); i++)
        if (comp[i].x1 >= 0 && (best < 0 || comp[i].core > comp[(size_t)best].core)) best = (int)i;
    // contourArea bounds: Pick's theorem gives area >= interior lattice points - 1 >= core - 1; any contour's area is at
    // most that of its bounding box of pixel centres
    const long long best_lo = comp[(size_t)best].core - 1;
    for (size_t i = 0; i < comp.size(); i++) {
        if ((int)i == best || comp[i].x1 < 0) continue;
        const long long hi = (long long)(comp[i].x1 - comp[i].x0) * (comp[i].y1 - comp[i].y0);
        if (hi >= best_lo)
            return fail(DS_ERR_UNSUPPORTED, "two foreground regions of comparable size (%lld vs %lld): use cv::findContours on the downloaded panorama", best_lo, hi);
    }
    const Comp& b = comp[(size_t)best];
    out_xywh[0] = b.x0; out_xywh[1] = b.y0; out_xywh[2] = b.x1 - b.x0 + 1; out_xywh[3] = b.y1 - b.y0 + 1;
    return DS_OK;
}

DS_API int ds_get_info(const ds_canvas* c, ds_canvas_info* info) {
    if (!c || !info) return fail(DS_ERR_BAD_ARG, "null argument");
    memset(info, 0, sizeof(*info));
    info->padded_width = c->pw; info->padded_height = c->ph; info->num_bands = c->L;
    int n = 0;
    double src_px = 0, bbox_px = 0;
    for (const Frame& f : c->frames) if (f.used) { n++; const double k = ab_frame_share(c, f); src_px += k * f.w * f.h; bbox_px += k * f.bw * f.bh; }
    info->num_frames = n;
    info->band_y0 = c->band.lo; info->band_y1 = c->band.hi;
    info->device_bytes = c->device_bytes;
    info->launches_last_composite = c->launches;
    info->ms_last_composite = c->last_ms;
    info->h2d_bytes_total = c->h2d_bytes;
    // SURVEY.md §8(d) algorithmic-bytes model
    const double C = ab_canvas_px(c);
    if (c->desc.blend_mode == DS_BLEND_FEATHER) info->algorithmic_bytes = (int64_t)(3.0 * src_px + 4.0 * C);
    else {
        double g = 0, q = 1;
        for (int l = 0; l <= c->L; l++) { g += q; q *= 0.25; }
        info->algorithmic_bytes = (int64_t)(3.0 * src_px + bbox_px * (20.0 * g + 30.0 * (g - 1.0)) + C * (17.5 * g - 2.0));
    }
    return DS_OK;
}

DS_API int ds_set_profiling(ds_canvas* c, int on) {
    if (!c) return fail(DS_ERR_BAD_ARG, "null canvas");
    c->profiling = on != 0;
    c->prof.clear();   // a profile accumulates over every composite until profiling is switched again
    return DS_OK;
}

DS_API int ds_get_kernel_times(ds_canvas* c, ds_kernel_time* out, int cap, int* n) {
    if (!c || !n) return fail(DS_ERR_BAD_ARG, "null argument");
    int rc;
    if ((rc = set_device(c))) return rc;
    if ((rc = stream_sync(c->stream))) return rc;
    *n = (int)c->prof.size();
    for (size_t i = 0; i < c->prof.size(); i++) {
#if DS_CUDA
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, c->prof_ev[2 * i], c->prof_ev[2 * i + 1]) != cudaSuccess) { cudaGetLastError(); ms = -1.f; }
        c->prof[i].ms = ms;
#endif
        if (out && (int)i < cap) {
            memset(&out[i], 0, sizeof(out[i]));
            snprintf(out[i].name, sizeof(out[i].name), "%s", c->prof[i].name);
            out[i].level = c->prof[i].level; out[i].ms = c->prof[i].ms; out[i].algorithmic_bytes = c->prof[i].ab;
        }
    }
    return DS_OK;
}

// ---------------------------------------------------------------- NVLink P2P halo exchange

DS_API int ds_p2p_export(ds_canvas* c, void* blob, size_t capacity, size_t* size) {
    if (!c || !size) return fail(DS_ERR_BAD_ARG, "null argument");
    if (c->desc.blend_mode != DS_BLEND_MULTIBAND || c->L < 1) return fail(DS_ERR_UNSUPPORTED, "the halo exchange exists for multi-band canvases with at least one band");
    int rc;
    if ((rc = set_device(c))) return rc;
    std::vector<const Frame*> fr;
    for (const Frame& f : c->frames) if (f.used && f.d_pyr) fr.push_back(&f);
    const size_t need = sizeof(BlobHeader) + fr.size() * sizeof(BlobFrame);
    *size = need;
    if (!blob) return DS_OK;   // size query
    if (capacity < need) return fail(DS_ERR_BAD_ARG, "blob needs %zu bytes", need);
    if (c->stage_open) return fail(DS_ERR_STATE, "a composite is half done");
    // An export starts a new connection epoch: old neighbours are dropped and the hand-over counters restart from zero,
    // so every handle of the chain exports again before any of them reconnects (ds_p2p_connect checks it).
    if ((rc = stream_sync(c->stream))) return rc;
    if (c->peer[0].connected || c->peer[1].connected) c->dirty = true;
    peer_close(c, 0); peer_close(c, 1);
    if (!c->d_flags && (rc = dev_alloc_t(&c->d_flags, 64))) return rc;   // own allocation: it gets its own IPC handle
    if ((rc = dev_zero(c->d_flags, 64 * sizeof(int)))) return rc;
    c->p2p_seq = 0;
    reg_set(c);
    BlobHeader h;
    memset(&h, 0, sizeof(h));
    h.magic = BLOB_MAGIC; h.version = 2; h.pid = (int64_t)getpid();
    h.uid = c->uid; h.gen = c->p2p_gen;
    h.device = c->desc.device; h.L = c->L; h.pw = c->pw; h.ph = c->ph; h.band_lo = c->band.lo; h.band_hi = c->band.hi;
    h.nframes = (int32_t)fr.size();
    h.flags_ptr = (uint64_t)(uintptr_t)c->d_flags;
#if DS_CUDA
    { cudaIpcMemHandle_t mh; DS_CK(cudaIpcGetMemHandle(&mh, c->d_flags)); static_assert(sizeof(mh) == 64, "ipc handle size"); memcpy(h.flags_handle, &mh, 64); }
#endif
    memcpy(blob, &h, sizeof(h));
    BlobFrame* out = (BlobFrame*)((char*)blob + sizeof(h));
    for (size_t i = 0; i < fr.size(); i++) {
        const Frame& f = *fr[i];
        BlobFrame bf;
        memset(&bf, 0, sizeof(bf));
        bf.idx = (int32_t)(&f - c->frames.data()); bf.ry = f.ry; bf.rh = f.rh; bf.rw = f.rw; bf.gp1 = f.dev.gp[1];
        bf.g_off = (uint64_t)((const char*)f.dev.G[1] - (const char*)f.d_pyr);
        bf.w_off = (uint64_t)((const char*)f.dev.W[1] - (const char*)f.d_pyr);
        bf.pyr_ptr = (uint64_t)(uintptr_t)f.d_pyr;
#if DS_CUDA
        { cudaIpcMemHandle_t mh; DS_CK(cudaIpcGetMemHandle(&mh, f.d_pyr)); memcpy(bf.pyr_handle, &mh, 64); }
#endif
        memcpy(&out[i], &bf, sizeof(bf));
    }
    return DS_OK;
}

DS_API int ds_p2p_connect(ds_canvas* c, int side, const void* blob, size_t size) {
    if (!c || !blob) return fail(DS_ERR_BAD_ARG, "null argument");
    if (side != 0 && side != 1) return fail(DS_ERR_BAD_ARG, "side must be 0 (the band above) or 1 (the band below)");
    if (c->desc.blend_mode != DS_BLEND_MULTIBAND || c->L < 1) return fail(DS_ERR_UNSUPPORTED, "the halo exchange exists for multi-band canvases with at least one band");
    if (c->stage_open) return fail(DS_ERR_STATE, "a composite is half done");
    if (size < sizeof(BlobHeader)) return fail(DS_ERR_BAD_ARG, "blob too small");
    BlobHeader h;
    memcpy(&h, blob, sizeof(h));
    if (h.magic != BLOB_MAGIC || h.version != 2 || size < sizeof(BlobHeader) + (size_t)h.nframes * sizeof(BlobFrame))
        return fail(DS_ERR_BAD_ARG, "not a ds_p2p_export blob");
    if (h.L != c->L || h.pw != c->pw || h.ph != c->ph) return fail(DS_ERR_BAD_ARG, "the neighbour describes a different canvas");
    if (side == 0 ? h.band_hi != c->band.lo : h.band_lo != c->band.hi)
        return fail(DS_ERR_BAD_ARG, "band [%d, %d) is not the neighbour %s of [%d, %d)", h.band_lo, h.band_hi, side ? "below" : "above", c->band.lo, c->band.hi);
    // the rows pulled from this neighbour must all be its own
    const Range rows = pulled_rows(c, side);
    if (rows.lo < (h.band_lo >> 1) || rows.hi > ((h.band_hi + 1) >> 1))
        return fail(DS_ERR_P2P_UNAVAILABLE, "the neighbour's band [%d, %d) is thinner than the pyramid halo (level-1 rows [%d, %d))", h.band_lo, h.band_hi, rows.lo, rows.hi);
    int rc;
    if ((rc = set_device(c))) return rc;
    if ((rc = stream_sync(c->stream))) return rc;
    // both ends count composites from the same epoch: this handle must not have composited in exchange mode since its own
    // last ds_p2p_export (the neighbour's blob is fresh by construction: exporting reset its counters)
    if (c->p2p_seq != 0) return fail(DS_ERR_STATE, "this handle has composited since its last ds_p2p_export: export again on every handle before reconnecting");
    if (!c->d_flags) {
        if ((rc = dev_alloc_t(&c->d_flags, 64))) return rc;
        if ((rc = dev_zero(c->d_flags, 64 * sizeof(int)))) return rc;
    }
    peer_close(c, side);
    ds_canvas::Peer& pr = c->peer[side];
    const bool same_process = h.pid == (int64_t)getpid();
    if (same_process && !reg_alive(h.uid, h.gen)) return fail(DS_ERR_STATE, "the exporting handle was destroyed or changed its frames since the export");
    pr.same_process = same_process; pr.uid = h.uid; pr.gen = h.gen;
    auto map = [&](uint64_t raw, const unsigned char* handle, void** out) -> int {
        if (same_process) {
#if DS_CUDA
            if (h.device != c->desc.device) {
                cudaError_t e = cudaDeviceEnablePeerAccess(h.device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return fail(DS_ERR_P2P_UNAVAILABLE, "no peer access from device %d to %d", c->desc.device, h.device); }
                cudaGetLastError();
            }
#endif
            *out = (void*)(uintptr_t)raw;
            return DS_OK;
        }
#if DS_CUDA
        cudaIpcMemHandle_t mh;
        memcpy(&mh, handle, 64);
        cudaError_t e = cudaIpcOpenMemHandle(out, mh, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) { cudaGetLastError(); return fail(DS_ERR_P2P_UNAVAILABLE, "cudaIpcOpenMemHandle: %s", cudaGetErrorString(e)); }
        pr.mapped.push_back(*out);
        return DS_OK;
#else
        (void)handle;
        return fail(DS_ERR_P2P_UNAVAILABLE, "no inter-process mapping in the emulator");
#endif
    };
    void* fl = nullptr;
    if ((rc = map(h.flags_ptr, h.flags_handle, &fl))) { peer_close(c, side); return rc; }
    pr.flags = (int*)fl;
    pr.band = Range{h.band_lo, h.band_hi};
    const BlobFrame* in = (const BlobFrame*)((const char*)blob + sizeof(h));
    for (int i = 0; i < h.nframes; i++) {
        BlobFrame bf;
        memcpy(&bf, &in[i], sizeof(bf));
        // only frames that reach into the rows pulled from this neighbour are mapped
        const int ry1 = bf.ry >> 1, rh1 = bf.rh >> 1;
        if (std::max(rows.lo, ry1) >= std::min(rows.hi, ry1 + rh1)) continue;
        void* base = nullptr;
        if ((rc = map(bf.pyr_ptr, bf.pyr_handle, &base))) { peer_close(c, side); return rc; }
        pr.frames.push_back(ds_canvas::PeerFrame{bf.idx, bf.ry, bf.rh, bf.rw, bf.gp1, (const char*)base, (size_t)bf.g_off, (size_t)bf.w_off});
    }
    pr.connected = true;
    c->dirty = true;
    return DS_OK;
}

DS_API int ds_p2p_disconnect(ds_canvas* c) {
    if (!c) return fail(DS_ERR_BAD_ARG, "null canvas");
    int rc;
    if ((rc = set_device(c))) return rc;
    if ((rc = stream_sync(c->stream))) return rc;
    peer_close(c, 0); peer_close(c, 1);
    c->dirty = true;
    return DS_OK;
}

DS_API int ds_composite_stage(ds_canvas* c, int stage) {
    if (!c) return fail(DS_ERR_BAD_ARG, "null canvas");
    if (stage != 0 && stage != 1) return fail(DS_ERR_BAD_ARG, "stage must be 0 or 1");
    int rc;
    if ((rc = set_device(c))) return rc;
    return run_composite(c, stage);
}

// ---------------------------------------------------------------- debug taps

static int tap_common(ds_canvas* c, int frame_idx, Frame** f) {
    if (!c) return fail(DS_ERR_BAD_ARG, "null canvas");
    if (frame_idx < 0 || (size_t)frame_idx >= c->frames.size() || !c->frames[(size_t)frame_idx].used)
        return fail(DS_ERR_BAD_ARG, "frame %d not uploaded", frame_idx);
    int rc;
    if ((rc = set_device(c))) return rc;
    if ((rc = flush_uploads(c))) return rc;
    if (c->dirty || c->subs.empty()) {
        plan_subbands(c, c->meta_slice_rows > 0 ? c->meta_slice_rows : 0);
        if ((rc = build_lists(c))) return rc;
    }
    *f = &c->frames[(size_t)frame_idx];
    return DS_OK;
}

DS_API int ds_debug_get_placement(ds_canvas* c, int frame_idx, int32_t out_xywh[4]) {
    if (!c || !out_xywh) return fail(DS_ERR_BAD_ARG, "null argument");
    if (frame_idx < 0 || (size_t)frame_idx >= c->frames.size() || !c->frames[(size_t)frame_idx].used)
        return fail(DS_ERR_BAD_ARG, "frame %d not uploaded", frame_idx);
    const Frame& f = c->frames[(size_t)frame_idx];
    out_xywh[0] = f.corner_x; out_xywh[1] = f.corner_y; out_xywh[2] = f.bw; out_xywh[3] = f.bh;
    return DS_OK;
}

DS_API int ds_debug_get_maps(ds_canvas* c, int frame_idx, int16_t* xy, uint16_t* a) {
    Frame* f;
    int rc = tap_common(c, frame_idx, &f);
    if (rc) return rc;
    if (!xy || !a) return fail(DS_ERR_BAD_ARG, "null output");
    const size_t n = (size_t)f->bw * f->bh;
    int16_t* d_xy = nullptr; uint16_t* d_a = nullptr;
    if ((rc = dev_alloc_t(&d_xy, n * 2))) return rc;
    if ((rc = dev_alloc_t(&d_a, n))) { dev_free(d_xy); return rc; }
    TapParams tp{c->d_frames, frame_idx, d_xy, d_a, nullptr, nullptr};
    rc = launch<TapBody, 256>(tp, (long long)((n + TapBody::PER_BLOCK - 1) / TapBody::PER_BLOCK), c->stream, 0);
    if (!rc) rc = d2h(xy, d_xy, n * 2 * sizeof(int16_t), c->stream);
    if (!rc) rc = d2h(a, d_a, n * sizeof(uint16_t), c->stream);
    if (!rc) rc = stream_sync(c->stream);
    dev_free(d_xy); dev_free(d_a);
    return rc;
}

DS_API int ds_debug_get_warped(ds_canvas* c, int frame_idx, uint8_t* bgr, uint8_t* mask) {
    Frame* f;
    int rc = tap_common(c, frame_idx, &f);
    if (rc) return rc;
    if (!bgr || !mask) return fail(DS_ERR_BAD_ARG, "null output");
    if (f->partial) return fail(DS_ERR_STATE, "frame %d is only partially resident (asynchronous upload into a row-band handle)", frame_idx);
    const size_t n = (size_t)f->bw * f->bh;
    uint8_t* d_b = nullptr; uint8_t* d_m = nullptr;
    if ((rc = dev_alloc_t(&d_b, n * 3))) return rc;
    if ((rc = dev_alloc_t(&d_m, n))) { dev_free(d_b); return rc; }
    TapParams tp{c->d_frames, frame_idx, nullptr, nullptr, d_b, d_m};
    rc = launch<TapBody, 256>(tp, (long long)((n + TapBody::PER_BLOCK - 1) / TapBody::PER_BLOCK), c->stream, 0);
    if (!rc) rc = d2h(bgr, d_b, n * 3, c->stream);
    if (!rc) rc = d2h(mask, d_m, n, c->stream);
    if (!rc) rc = stream_sync(c->stream);
    dev_free(d_b); dev_free(d_m);
    return rc;
}

DS_API int ds_debug_get_frame_level(ds_canvas* c, int frame_idx, int level, int16_t* g, float* w, int32_t dims_out[4]) {
    Frame* f;
    int rc = tap_common(c, frame_idx, &f);
    if (rc) return rc;
    if (c->desc.blend_mode != DS_BLEND_MULTIBAND) return fail(DS_ERR_STATE, "not a multiband canvas");
    if (level < 1 || level > c->L) return fail(DS_ERR_BAD_ARG, "level %d not in [1, %d]", level, c->L);
    const int lw = f->rw >> level, lh = f->rh >> level;
    if (dims_out) { dims_out[0] = f->rx >> level; dims_out[1] = f->ry >> level; dims_out[2] = lw; dims_out[3] = lh; }
    if (!g && !w) return DS_OK;
    if (!c->composited) return fail(DS_ERR_STATE, "frame pyramids exist only after ds_composite");
    const size_t pitch = (size_t)f->dev.gp[level];
    if (g) {
        std::vector<px8> tmp((size_t)lw * lh);
        if ((rc = d2h_2d(tmp.data(), (size_t)lw * sizeof(px8), f->dev.G[level], pitch * sizeof(px8), (size_t)lw * sizeof(px8), (size_t)lh, c->stream))) return rc;
        if ((rc = stream_sync(c->stream))) return rc;
        for (size_t i = 0; i < (size_t)lw * lh; i++) { g[3 * i] = tmp[i].b; g[3 * i + 1] = tmp[i].g; g[3 * i + 2] = tmp[i].r; }
    }
    if (w) {
        if ((rc = d2h_2d(w, (size_t)lw * sizeof(float), f->dev.W[level], pitch * sizeof(float), (size_t)lw * sizeof(float), (size_t)lh, c->stream))) return rc;
        if ((rc = stream_sync(c->stream))) return rc;
    }
    return DS_OK;
}

}  // extern "C"
