// C-ABI implementation (include/radfoam_b200.h) of the B200 foam tracer.
// Host side: argument checks, cached scene re-layout, kernel dispatch.
#include "../../include/radfoam_b200.h"

#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <atomic>
#include <cstring>
#include <new>
#include <string>

#include "foam_kernels.cuh"

// Kernel launch.  In the product this is the <<< >>> launch; the CPU kernel-logic emulator under tests/emu
// (test infrastructure, never loaded by the product) defines RFB_LAUNCH itself before this file is compiled.
#ifndef RFB_LAUNCH
#define RFB_UNPAREN(...) __VA_ARGS__
#define RFB_LAUNCH(kernel, grid, block, smem, stream, ...) \
    (RFB_UNPAREN kernel)<<<grid, block, smem, stream>>>(__VA_ARGS__)
#endif


namespace {

thread_local std::string g_last_error;
std::atomic<uint64_t> g_launches{0}; // process-wide: autograd runs backward on its own thread
std::atomic<uint64_t> g_device_allocs{0}, g_device_frees{0}; // cudaMalloc / cudaFree calls made by this library

int fail(const std::string &msg) {
    g_last_error = msg;
    return 1;
}

#define RFB_CUDA(call)                                                                    \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess)                                                            \
            return fail(std::string("CUDA call failed at " __FILE__ ":") +                \
                        std::to_string(__LINE__) + ": " + cudaGetErrorString(e_));        \
    } while (0)

#define RFB_LAUNCHED()                                                                    \
    do {                                                                                  \
        ++g_launches;                                                                     \
        RFB_CUDA(cudaGetLastError());                                                     \
    } while (0)

struct DeviceBuffer {
    void *ptr = nullptr;
    size_t bytes = 0;
    // grow-only; reallocation synchronises, steady state does not
    cudaError_t ensure(size_t need) {
        if (need <= bytes)
            return cudaSuccess;
        if (ptr) {
            ++g_device_frees;
            cudaError_t e = cudaFree(ptr);
            ptr = nullptr;
            bytes = 0;
            if (e != cudaSuccess)
                return e;
        }
        size_t want = need + need / 16 + 256;
        ++g_device_allocs;
        cudaError_t e = cudaMalloc(&ptr, want);
        if (e == cudaSuccess)
            bytes = want;
        return e;
    }
    void release() {
        if (ptr)
            cudaFree(ptr);
        ptr = nullptr;
        bytes = 0;
    }
};

struct SceneKey {
    const void *points = nullptr, *attrs = nullptr, *adj = nullptr, *off = nullptr;
    const void *diff = nullptr; // non-null: faces were copied from this caller-built array
    uint32_t n = 0, e = 0;
    uint64_t version = 0;
    bool faces = false;
    bool operator==(const SceneKey &o) const {
        return points == o.points && attrs == o.attrs && adj == o.adj && off == o.off && diff == o.diff &&
               n == o.n && e == o.e && version == o.version;
    }
};

int grid_for(uint64_t work_items, int block, int max_blocks = 148 * 16) {
    uint64_t b = (work_items + block - 1) / block;
    if (b > (uint64_t)max_blocks)
        b = max_blocks;
    if (b < 1)
        b = 1;
    return (int)b;
}

} // namespace

struct rfb_pipeline {
    int sh_degree;
    int attr_dtype;
    int device = -1;
    DeviceBuffer cells, sh_rows, faces, nbr, acc;
    SceneKey key;
    bool key_valid = false;
    // producer stream + completion event of the mirrors / of the tape: a consumer on another stream waits on it
    cudaStream_t scene_stream = nullptr, tape_stream = nullptr;
    cudaEvent_t scene_ready = nullptr, tape_ready = nullptr;
    // RFB_DEBUG=1: checksum of (points, attributes) taken when the mirrors were built, re-checked on every cache hit
    DeviceBuffer debug_sum;
    uint64_t scene_checksum = 0;
    uint32_t acc_points = 0;
    // caller-provided accumulator memory (rfb_set_grad_accumulator: a peer-mapped allocation for the fused
    // multi-GPU reduction); nullptr = the pipeline's own `acc`
    float *acc_external = nullptr;
    uint64_t acc_external_floats = 0;
    // parameter-form scene (rfb_bind_scene_params): attributes are derived from these inside the re-layout kernel
    rfb_scene_params params = {nullptr, nullptr, nullptr, 1.0f};
    bool params_bound = false;
    // walk tape (forward records, backward replays; see foam_kernels.cuh)
    DeviceBuffer scan_flag; // one word: a fast forward met a ray outside the ranked scan's domain (ScanMode)
    DeviceBuffer tape_pool, tape_table, tape_per_ray, tape_ctrl, tape_sched; // sched: [tile_steps | order]
    uint32_t tape_blocks = 0;
    bool tape_scheduled = false; // the last recording forward also built a tile order for the replay
    uint32_t tape_capacity = 0;      // chunks
    uint32_t tape_table_stride = 0;
    bool tape_valid = false;
    struct TapeKey {
        const void *rays = nullptr, *start = nullptr;
        uint32_t num_rays = 0, image_width = 0, max_steps = 0;
        float weight_threshold = 0.0f;
        uint64_t scene_version = 0;
    } tape_key;
    uint32_t *tape_readback = nullptr; // pinned host copy of ctrl, refreshed after every recording
    cudaEvent_t tape_readback_done = nullptr;
    bool tape_readback_pending = false;
    // optional live kernel timing (rfb_set_profiling): events around the ray kernels
    bool profiling = false;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr}; // fwd start/stop, bwd start/stop
    bool ev_valid[2] = {false, false};
};

namespace {

using namespace rfb;

int check_device(rfb_pipeline *p) {
    int dev = 0;
    RFB_CUDA(cudaGetDevice(&dev));
    if (p->device < 0)
        p->device = dev;
    else if (p->device != dev)
        return fail("pipeline was first used on CUDA device " + std::to_string(p->device) +
                    " but the current device is " + std::to_string(dev) +
                    " (create one pipeline per device)");
    return 0;
}

bool debug_checks() {
    static const bool on = [] {
        const char *e = getenv("RFB_DEBUG");
        return e && *e && *e != '0';
    }();
    return on;
}

// RFB_DEBUG only (synchronises): order-independent checksum of the caller's points and attributes
int scene_checksum(rfb_pipeline *p, uint32_t n, const float *points, const void *attrs, cudaStream_t stream,
                   uint64_t &out) {
    RFB_CUDA(p->debug_sum.ensure(sizeof(unsigned long long)));
    RFB_CUDA(cudaMemsetAsync(p->debug_sum.ptr, 0, sizeof(unsigned long long), stream));
    const int A = attr_dim(p->sh_degree);
    const uint64_t attr_words = (uint64_t)n * A * (p->attr_dtype == RFB_FLOAT16 ? 2 : 4) / 4;
    auto *sum = reinterpret_cast<unsigned long long *>(p->debug_sum.ptr);
    if (n) {
        RFB_LAUNCH((checksum_kernel), grid_for((uint64_t)n * 3, 256), 256, 0, stream,
                   reinterpret_cast<const uint32_t *>(points), (uint64_t)n * 3, sum);
        RFB_LAUNCH((checksum_kernel), grid_for(attr_words, 256), 256, 0, stream,
                   reinterpret_cast<const uint32_t *>(attrs), attr_words, sum);
    }
    unsigned long long host = 0;
    RFB_CUDA(cudaMemcpyAsync(&host, sum, sizeof(host), cudaMemcpyDeviceToHost, stream));
    RFB_CUDA(cudaStreamSynchronize(stream));
    out = host;
    return 0;
}

// Is `stream` being captured into a CUDA graph?  Then nothing host-visible may be attached to it: no events
// that the host (or another stream) later queries, no read-back.  A captured step replays with the buffers it
// was captured with; the library's bookkeeping (tape growth, cross-stream ordering) resumes outside the graph.
bool capturing(cudaStream_t stream) {
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(stream, &st) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return st != cudaStreamCaptureStatusNone;
}

// make `stream` wait for work recorded in `ev` on `producer` (no-op on the same stream)
int wait_for(cudaEvent_t ev, cudaStream_t producer, cudaStream_t stream) {
    if (ev && producer != stream && !capturing(stream))
        RFB_CUDA(cudaStreamWaitEvent(stream, ev, 0));
    return 0;
}

int mark(cudaEvent_t &ev, cudaStream_t &producer, cudaStream_t stream) {
    if (capturing(stream))
        return 0;
    if (!ev)
        RFB_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    RFB_CUDA(cudaEventRecord(ev, stream));
    producer = stream;
    return 0;
}

// (Re)build the internal mirrors of the scene unless the caller vouches they are current.
int ensure_scene(rfb_pipeline *p, uint32_t n, const float *points, const void *attrs, uint32_t e,
                 const uint32_t *adj, const uint32_t *off, bool need_faces,
                 const rfb_launch_opts *opts, cudaStream_t stream, const void *caller_diff = nullptr) {
    if (int rc = check_device(p))
        return rc;
    SceneKey k;
    k.points = points;
    k.attrs = attrs;
    k.adj = adj;
    k.off = off;
    k.diff = caller_diff;
    k.n = n;
    k.e = e;
    k.version = opts ? opts->scene_version : 0;
    if (p->key_valid && k.version != 0 && k == p->key && (p->key.faces || !need_faces)) {
        if (debug_checks()) {
            uint64_t now = 0;
            if (int rc = scene_checksum(p, n, points, attrs, stream, now))
                return rc;
            if (now != p->scene_checksum)
                return fail("RFB_DEBUG: points/attributes changed but scene_version did not (stale scene mirrors; "
                            "bump rfb_launch_opts.scene_version whenever the scene tensors are written)");
        }
        return wait_for(p->scene_ready, p->scene_stream, stream);
    }
    p->key_valid = false;
    if (debug_checks())
        if (int rc = scene_checksum(p, n, points, attrs, stream, p->scene_checksum))
            return rc;

    const int A = attr_dim(p->sh_degree), SR = sh_row(p->sh_degree);
    RFB_CUDA(p->cells.ensure((size_t)n * sizeof(float4)));
    RFB_CUDA(p->sh_rows.ensure((size_t)n * SR * sizeof(float)));
    if (n && p->params_bound) {
        const int grid = grid_for((uint64_t)n * (SR / 4), 256 * 4, 148 * 8);
        float4 *cells = reinterpret_cast<float4 *>(p->cells.ptr), *rows = reinterpret_cast<float4 *>(p->sh_rows.ptr);
#define RFB_BUILD_PARAMS(DEG)                                                                                      \
    do {                                                                                                           \
        if (p->attr_dtype == RFB_FLOAT16)                                                                          \
            RFB_LAUNCH((build_cells_params_kernel<__half, DEG>), grid, 256, 0, stream, points, p->params.att_dc,   \
                       p->params.att_sh, p->params.density, p->params.activation_scale, n, cells, rows);           \
        else                                                                                                       \
            RFB_LAUNCH((build_cells_params_kernel<float, DEG>), grid, 256, 0, stream, points, p->params.att_dc,    \
                       p->params.att_sh, p->params.density, p->params.activation_scale, n, cells, rows);           \
    } while (0)
        switch (p->sh_degree) {
        case 0: RFB_BUILD_PARAMS(0); break;
        case 1: RFB_BUILD_PARAMS(1); break;
        case 2: RFB_BUILD_PARAMS(2); break;
        default: RFB_BUILD_PARAMS(3); break;
        }
#undef RFB_BUILD_PARAMS
        RFB_LAUNCHED();
    } else if (n) {
        const int grid = grid_for((uint64_t)n * (SR / 4), 256 * 4, 148 * 8);
        float4 *cells = reinterpret_cast<float4 *>(p->cells.ptr), *rows = reinterpret_cast<float4 *>(p->sh_rows.ptr);
#define RFB_BUILD_CELLS(DEG)                                                                                       \
    do {                                                                                                           \
        if (p->attr_dtype == RFB_FLOAT16)                                                                          \
            RFB_LAUNCH((build_cells_kernel<__half, DEG>), grid, 256, 0, stream, points,                            \
                       reinterpret_cast<const __half *>(attrs), n, cells, rows);                                   \
        else                                                                                                       \
            RFB_LAUNCH((build_cells_kernel<float, DEG>), grid, 256, 0, stream, points,                             \
                       reinterpret_cast<const float *>(attrs), n, cells, rows);                                    \
    } while (0)
        switch (p->sh_degree) {
        case 0: RFB_BUILD_CELLS(0); break;
        case 1: RFB_BUILD_CELLS(1); break;
        case 2: RFB_BUILD_CELLS(2); break;
        default: RFB_BUILD_CELLS(3); break;
        }
#undef RFB_BUILD_CELLS
        RFB_LAUNCHED();
    }
    if (need_faces) {
        size_t slots = (size_t)padded_slots(n, e); // rows padded to 4 faces, <= 3 slack per row
        RFB_CUDA(p->faces.ensure(slots * sizeof(uint2)));
        RFB_CUDA(p->nbr.ensure(slots * sizeof(uint32_t)));
        if (n && caller_diff) {
            RFB_LAUNCH((build_faces_from_diff_kernel), grid_for((uint64_t)n * 16, 256), 256, 0, stream,
                       reinterpret_cast<const uint2 *>(caller_diff), n, adj, off,
                       reinterpret_cast<uint2 *>(p->faces.ptr), reinterpret_cast<uint32_t *>(p->nbr.ptr));
            RFB_LAUNCHED();
        } else if (n) {
            RFB_LAUNCH((build_faces_kernel), grid_for((uint64_t)n * 16, 256), 256, 0, stream, points, n, adj,
                       off, reinterpret_cast<uint2 *>(p->faces.ptr),
                       reinterpret_cast<uint32_t *>(p->nbr.ptr));
            RFB_LAUNCHED();
        }
    }
    k.faces = need_faces;
    p->key = k;
    p->key_valid = true;
    return mark(p->scene_ready, p->scene_stream, stream);
}

void ray_grid(uint32_t num_rays, uint32_t image_width, uint32_t &blocks, uint32_t &blocks_x,
              uint32_t &width_used) {
    width_used = (image_width && num_rays % image_width == 0) ? image_width : 0;
    if (width_used) {
        uint32_t h = num_rays / width_used;
        blocks_x = (width_used + kTileW - 1) / kTileW;
        blocks = blocks_x * ((h + kTileH - 1) / kTileH);
    } else {
        blocks_x = 0;
        blocks = (num_rays + kBlock - 1) / kBlock;
    }
}

int profile_mark(rfb_pipeline *p, int which, cudaStream_t stream) {
    if (!p->profiling || capturing(stream))
        return 0;
    if (!p->ev[which])
        RFB_CUDA(cudaEventCreate(&p->ev[which]));
    RFB_CUDA(cudaEventRecord(p->ev[which], stream));
    if (which & 1)
        p->ev_valid[which >> 1] = true;
    return 0;
}

template <typename Faces, int MODE>
int launch_forward_mode(int deg, const ForwardParams &fp, const Faces &fa, uint32_t blocks, cudaStream_t stream) {
    switch (deg) {
    case 0: RFB_LAUNCH((forward_kernel<0, Faces, MODE>), blocks, kBlock, 0, stream, fp, fa); break;
    case 1: RFB_LAUNCH((forward_kernel<1, Faces, MODE>), blocks, kBlock, 0, stream, fp, fa); break;
    case 2: RFB_LAUNCH((forward_kernel<2, Faces, MODE>), blocks, kBlock, 0, stream, fp, fa); break;
    default: RFB_LAUNCH((forward_kernel<3, Faces, MODE>), blocks, kBlock, 0, stream, fp, fa); break;
    }
    RFB_LAUNCHED();
    return 0;
}

// The fast kernel and, right behind it, its exact twin (a no-op unless the fast one raised the flag); with a
// contribution output (accumulated by atomics: a second pass would double it) one launch that chooses per ray.
template <typename Faces>
int launch_forward(int deg, const ForwardParams &fp, const Faces &fa, uint32_t blocks, cudaStream_t stream) {
    if (fp.contrib)
        return launch_forward_mode<Faces, kScanPerRay>(deg, fp, fa, blocks, stream);
    if (int rc = launch_forward_mode<Faces, kScanFast>(deg, fp, fa, blocks, stream))
        return rc;
    return launch_forward_mode<Faces, kScanExactTwin>(deg, fp, fa, blocks, stream);
}

template <typename Faces, int MODE>
int launch_forward_record_mode(int deg, const ForwardParams &fp, const Faces &fa, const Tape &tape, uint32_t blocks,
                               cudaStream_t stream) {
    switch (deg) {
    case 0: RFB_LAUNCH((forward_record_kernel<0, Faces, MODE>), blocks, kBlock, 0, stream, fp, fa, tape); break;
    case 1: RFB_LAUNCH((forward_record_kernel<1, Faces, MODE>), blocks, kBlock, 0, stream, fp, fa, tape); break;
    case 2: RFB_LAUNCH((forward_record_kernel<2, Faces, MODE>), blocks, kBlock, 0, stream, fp, fa, tape); break;
    default: RFB_LAUNCH((forward_record_kernel<3, Faces, MODE>), blocks, kBlock, 0, stream, fp, fa, tape); break;
    }
    RFB_LAUNCHED();
    return 0;
}

template <typename Faces>
int launch_forward_record(int deg, const ForwardParams &fp, const Faces &fa, const Tape &tape,
                          uint32_t blocks, cudaStream_t stream) {
    if (fp.contrib)
        return launch_forward_record_mode<Faces, kScanPerRay>(deg, fp, fa, tape, blocks, stream);
    if (int rc = launch_forward_record_mode<Faces, kScanFast>(deg, fp, fa, tape, blocks, stream))
        return rc;
    return launch_forward_record_mode<Faces, kScanExactTwin>(deg, fp, fa, tape, blocks, stream);
}

// Size / reset the tape for a recording forward of `blocks` CTAs; fills `tape`.  The tape is an optimisation:
// when its memory cannot be had (the pool is a raw cudaMalloc next to the caller's own allocator), `ok` comes
// back false, the forward runs without recording and the backward re-walks.
int prepare_tape(rfb_pipeline *p, uint32_t blocks, uint32_t num_rays, uint32_t max_steps, Tape &tape,
                 cudaStream_t stream, bool &ok) {
    ok = false;
    const uint32_t num_warps = blocks * (kBlock / 32);
    const uint32_t stride = max_steps / kTapeChunk + 2;
    // grow the pool if the previous recording needed more chunks than it had
    // (no event query while a graph is being captured: "unsafe" runtime calls invalidate a global-mode capture)
    if (p->tape_readback_pending && !capturing(stream) && cudaEventQuery(p->tape_readback_done) == cudaSuccess) {
        p->tape_readback_pending = false;
        uint32_t wanted = p->tape_readback[0] > p->tape_readback[2] ? p->tape_readback[0] : p->tape_readback[2];
        if (wanted > p->tape_capacity)
            p->tape_capacity = wanted + wanted / 4;
    }
    uint64_t want = (uint64_t)num_warps; // first guess: one chunk (32 steps) per warp; grows on demand
    if (p->tape_capacity < want)
        p->tape_capacity = (uint32_t)(want < 0xFFFFFFF0ull ? want : 0xFFFFFFF0ull);
    const size_t chunk_bytes = (size_t)kTapeChunk * 32 * sizeof(uint2);
    size_t pool_bytes = (size_t)p->tape_capacity * chunk_bytes;
    if (pool_bytes > p->tape_pool.bytes) {
        // growing: never take more than half of what is free now (plus what the pool already holds)
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
            const size_t cap = (free_b + p->tape_pool.bytes) / 2;
            if (pool_bytes + pool_bytes / 16 + 256 > cap) {
                const uint64_t fit = cap / (chunk_bytes + chunk_bytes / 16 + 1);
                p->tape_capacity = (uint32_t)(fit < p->tape_capacity ? fit : p->tape_capacity);
                pool_bytes = (size_t)p->tape_capacity * chunk_bytes;
            }
        }
        cudaGetLastError();
    }
    if (p->tape_capacity < want || p->tape_pool.ensure(pool_bytes) != cudaSuccess ||
        p->tape_table.ensure((size_t)num_warps * stride * sizeof(uint32_t)) != cudaSuccess ||
        p->tape_per_ray.ensure((size_t)num_rays * sizeof(uint2)) != cudaSuccess ||
        p->tape_sched.ensure((size_t)blocks * 2 * sizeof(uint32_t)) != cudaSuccess ||
        p->tape_ctrl.ensure(4 * sizeof(uint32_t)) != cudaSuccess) {
        cudaGetLastError(); // out of memory is not sticky; forget it
        p->tape_capacity = 0; // start from the first guess next time
        return 0;
    }
    RFB_CUDA(cudaMemsetAsync(p->tape_ctrl.ptr, 0, 4 * sizeof(uint32_t), stream));
    p->tape_blocks = blocks;
    tape.tile_steps = reinterpret_cast<uint32_t *>(p->tape_sched.ptr);
    tape.order = tape.tile_steps + blocks;
    if (!p->tape_readback) {
        RFB_CUDA(cudaMallocHost(reinterpret_cast<void **>(&p->tape_readback), 4 * sizeof(uint32_t)));
        RFB_CUDA(cudaEventCreateWithFlags(&p->tape_readback_done, cudaEventDisableTiming));
    }
    p->tape_table_stride = stride;
    tape.pool = reinterpret_cast<uint2 *>(p->tape_pool.ptr);
    tape.table = reinterpret_cast<uint32_t *>(p->tape_table.ptr);
    tape.per_ray = reinterpret_cast<uint2 *>(p->tape_per_ray.ptr);
    tape.ctrl = reinterpret_cast<uint32_t *>(p->tape_ctrl.ptr);
    tape.capacity = p->tape_capacity;
    tape.table_stride = stride;
    ok = true;
    return 0;
}

template <typename Faces>
int launch_backward(int deg, const BackwardParams &bp, const Faces &fa, uint32_t blocks,
                    cudaStream_t stream) {
    switch (deg) {
    case 0: RFB_LAUNCH((backward_kernel<0, Faces>), blocks, kBlock, 0, stream, bp, fa); break;
    case 1: RFB_LAUNCH((backward_kernel<1, Faces>), blocks, kBlock, 0, stream, bp, fa); break;
    case 2: RFB_LAUNCH((backward_kernel<2, Faces>), blocks, kBlock, 0, stream, bp, fa); break;
    default: RFB_LAUNCH((backward_kernel<3, Faces>), blocks, kBlock, 0, stream, bp, fa); break;
    }
    RFB_LAUNCHED();
    return 0;
}

template <int DEG, typename Faces, int SLOTS, int MIN_GROUP, int MIN_BLOCKS, bool REPLAY>
int launch_backward_cached_one(const BackwardParams &bp, const Faces &fa, const Tape &tape,
                               uint32_t blocks, cudaStream_t stream) {
    constexpr int GR = grad_row(DEG);
    constexpr size_t smem = (size_t)(kBlock / 32) * ((32 * GR + SLOTS * GR + SLOTS + 3) & ~3) * sizeof(float);
    auto kernel = backward_cached_kernel<DEG, Faces, SLOTS, MIN_GROUP, MIN_BLOCKS, REPLAY>;
    // the attribute is per device and idempotent: set it once per (instantiation, device)
    static std::atomic<uint64_t> configured{0};
    int dev = 0;
    RFB_CUDA(cudaGetDevice(&dev));
    const uint64_t bit = 1ull << (dev & 63);
    if (!(configured.load(std::memory_order_relaxed) & bit)) {
        RFB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured.fetch_or(bit, std::memory_order_relaxed);
    }
    RFB_LAUNCH((kernel), blocks, kBlock, smem, stream, bp, fa, tape);
    RFB_LAUNCHED();
    return 0;
}

// With a tape: the replay kernel and the re-walk kernel are both launched; the tape's overflow
// flag (device side) decides which one works.  Without: the re-walk kernel alone.
template <int DEG, typename Faces, int SLOTS, int MIN_GROUP, int MIN_BLOCKS>
int launch_backward_cached_cfg(const BackwardParams &bp, const Faces &fa, const Tape &tape,
                               uint32_t blocks, cudaStream_t stream) {
    if (tape.pool)
        if (int rc = launch_backward_cached_one<DEG, Faces, SLOTS, MIN_GROUP, MIN_BLOCKS, true>(
                bp, fa, tape, blocks, stream))
            return rc;
    return launch_backward_cached_one<DEG, Faces, SLOTS, MIN_GROUP, MIN_BLOCKS, false>(bp, fa, tape, blocks,
                                                                                     stream);
}

// Shipped configuration: 4 cache slots per warp, groups of >= 6 lanes go through the cache, 5 CTAs/SM
// (<= 102 registers).  Measured on the 1M-point / 1080p frame (B200, profiles/r01_backward_variants*.json,
// profiles/r02_variant_bench.json):
//   re-walk backward: direct 29.3 ms; (32 slots, >=2 lanes, 4 CTAs) 21.5; (16, >=4, 5) 17.7;
//                     (16, >=8, 5) 16.6; (8, >=8, 5) 15.7; 6 CTAs/SM spills and is slower;
//   replay backward:  (8, >=8, 5) 11.2; (4, >=8) 11.2; (2, >=8) 11.2; (4, >=4) 10.8;
//                     (4, >=5) 10.5; (2, >=6) 10.5; (4, >=6) 10.45  <- shipped
//   a pooled-row kernel (every group through a 8/16/32-row cache, compact records, quarter-warp rounds):
//                     12.4 - 14.5 ms -- 0.36x the reductions but +50 % instructions; built, measured, removed.
// With 4 slots the CTA's shared memory drops below 32 KB, so 5 CTAs fit the 164 KB carve-out and the SM keeps
// 92 KB of L1 instead of 60.
template <typename Faces>
int launch_backward_cached(int deg, const BackwardParams &bp, const Faces &fa, const Tape &tape,
                           uint32_t blocks, cudaStream_t stream) {
    switch (deg) {
    case 0: return launch_backward_cached_cfg<0, Faces, 4, 6, 5>(bp, fa, tape, blocks, stream);
    case 1: return launch_backward_cached_cfg<1, Faces, 4, 6, 5>(bp, fa, tape, blocks, stream);
    case 2: return launch_backward_cached_cfg<2, Faces, 4, 6, 5>(bp, fa, tape, blocks, stream);
    default: return launch_backward_cached_cfg<3, Faces, 4, 6, 5>(bp, fa, tape, blocks, stream);
    }
}

template <typename Faces, int MODE>
int launch_benchmark_mode(int deg, const BenchmarkParams &bp, const Faces &fa, uint32_t blocks, cudaStream_t stream) {
    switch (deg) {
    case 0: RFB_LAUNCH((benchmark_kernel<0, Faces, MODE>), blocks, kBlock, 0, stream, bp, fa); break;
    case 1: RFB_LAUNCH((benchmark_kernel<1, Faces, MODE>), blocks, kBlock, 0, stream, bp, fa); break;
    case 2: RFB_LAUNCH((benchmark_kernel<2, Faces, MODE>), blocks, kBlock, 0, stream, bp, fa); break;
    default: RFB_LAUNCH((benchmark_kernel<3, Faces, MODE>), blocks, kBlock, 0, stream, bp, fa); break;
    }
    RFB_LAUNCHED();
    return 0;
}

template <typename Faces>
int launch_benchmark(int deg, const BenchmarkParams &bp, const Faces &fa, uint32_t blocks, cudaStream_t stream) {
    if (int rc = launch_benchmark_mode<Faces, kScanFast>(deg, bp, fa, blocks, stream))
        return rc;
    return launch_benchmark_mode<Faces, kScanExactTwin>(deg, bp, fa, blocks, stream);
}

rfb_trace_settings settings_or_default(const rfb_trace_settings *s) {
    rfb_trace_settings d;
    d.weight_threshold = 0.001f; // default_trace_settings(), pipeline.h:15-20
    d.max_intersections = 1024;
    return s ? *s : d;
}

} // namespace

extern "C" {

const char *rfb_last_error(void) { return g_last_error.c_str(); }
int rfb_abi_version(void) { return RFB_ABI_VERSION; }
uint64_t rfb_launch_count(void) { return g_launches.load(); }
void rfb_reset_launch_count(void) { g_launches.store(0); }
void rfb_device_alloc_counts(uint64_t *allocs, uint64_t *frees) {
    if (allocs)
        *allocs = g_device_allocs.load();
    if (frees)
        *frees = g_device_frees.load();
}

int rfb_create_pipeline(int sh_degree, int attr_dtype, rfb_pipeline **out) {
    if (!out)
        return fail("rfb_create_pipeline: out is NULL");
    *out = nullptr;
    if (attr_dtype != RFB_FLOAT32 && attr_dtype != RFB_FLOAT16)
        return fail("Unsupported attribute type");
    if (sh_degree < 0 || sh_degree > 3)
        return fail("Unsupported SH degree");
    rfb_pipeline *p = new (std::nothrow) rfb_pipeline();
    if (!p)
        return fail("out of host memory");
    p->sh_degree = sh_degree;
    p->attr_dtype = attr_dtype;
    *out = p;
    return 0;
}

void rfb_destroy_pipeline(rfb_pipeline *p) {
    if (!p)
        return;
    p->cells.release();
    p->sh_rows.release();
    p->faces.release();
    p->nbr.release();
    p->acc.release();
    p->scan_flag.release();
    p->tape_pool.release();
    p->tape_sched.release();
    p->tape_table.release();
    p->tape_per_ray.release();
    p->tape_ctrl.release();
    if (p->tape_readback)
        cudaFreeHost(p->tape_readback);
    if (p->tape_readback_done)
        cudaEventDestroy(p->tape_readback_done);
    for (auto &e : p->ev)
        if (e)
            cudaEventDestroy(e);
    if (p->scene_ready)
        cudaEventDestroy(p->scene_ready);
    if (p->tape_ready)
        cudaEventDestroy(p->tape_ready);
    p->debug_sum.release();
    delete p;
}

uint32_t rfb_attribute_dim(const rfb_pipeline *p) { return p ? (uint32_t)rfb::attr_dim(p->sh_degree) : 0; }
int rfb_attribute_type(const rfb_pipeline *p) { return p ? p->attr_dtype : -1; }
uint32_t rfb_grad_row_floats(const rfb_pipeline *p) { return p ? (uint32_t)rfb::grad_row(p->sh_degree) : 0; }
void rfb_invalidate_cache(rfb_pipeline *p) {
    if (p)
        p->key_valid = false;
}

int rfb_tape_status(rfb_pipeline *p, uint32_t *capacity_chunks, uint32_t *used_chunks,
                    uint32_t *overflowed) {
    if (!p || !capacity_chunks || !used_chunks || !overflowed)
        return fail("rfb_tape_status: NULL argument");
    if (!p->tape_valid || !p->tape_readback)
        return fail("rfb_tape_status: no recorded tape");
    RFB_CUDA(cudaEventSynchronize(p->tape_readback_done));
    *capacity_chunks = p->tape_capacity;
    *used_chunks = p->tape_readback[0] > p->tape_readback[2] ? p->tape_readback[0] : p->tape_readback[2];
    *overflowed = p->tape_readback[1];
    return 0;
}

void rfb_set_profiling(rfb_pipeline *p, int enabled) {
    if (p) {
        p->profiling = enabled != 0;
        p->ev_valid[0] = p->ev_valid[1] = false;
    }
}

int rfb_last_kernel_ms(rfb_pipeline *p, int which, float *ms) {
    if (!p || !ms || which < 0 || which > 1)
        return fail("rfb_last_kernel_ms: bad argument");
    if (!p->ev_valid[which])
        return fail("rfb_last_kernel_ms: no timed launch recorded (enable rfb_set_profiling first)");
    RFB_CUDA(cudaEventSynchronize(p->ev[2 * which + 1]));
    RFB_CUDA(cudaEventElapsedTime(ms, p->ev[2 * which], p->ev[2 * which + 1]));
    return 0;
}

int rfb_prefetch_adjacent_diff(const float *points, uint32_t num_points, uint32_t point_adjacency_size,
                               const uint32_t *point_adjacency, const uint32_t *point_adjacency_offsets,
                               void *adjacent_diff, void *stream) {
    (void)point_adjacency_size;
    if (num_points == 0)
        return 0;
    if (!points || !point_adjacency || !point_adjacency_offsets || !adjacent_diff)
        return fail("rfb_prefetch_adjacent_diff: NULL argument");
    RFB_LAUNCH((rfb::adjacent_diff_kernel), grid_for((uint64_t)num_points * 16, 256), 256, 0,
               (cudaStream_t)stream, points, num_points, point_adjacency, point_adjacency_offsets,
               reinterpret_cast<uint2 *>(adjacent_diff));
    RFB_LAUNCHED();
    return 0;
}

namespace {
// Stream-ordered scratch from the device's default memory pool; freed blocks stay in the pool (the default
// release threshold of 0 hands them back to the driver at the next synchronisation: ~2 ms per call measured).
int stream_scratch(void **ptr, size_t bytes, cudaStream_t stream) {
    static std::once_flag pool_once;
    std::call_once(pool_once, [] {
        int dev = 0;
        cudaMemPool_t pool = nullptr;
        uint64_t keep = ~0ull;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess)
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        cudaGetLastError();
    });
    if (cudaMallocAsync(ptr, bytes, stream) != cudaSuccess) {
        cudaGetLastError();
        return fail("out of device memory for " + std::to_string(bytes) + " bytes of scratch");
    }
    return 0;
}
} // namespace

int rfb_nearest_point(const float *points, uint32_t num_points, const float *queries,
                      uint32_t num_queries, uint32_t *indices, void *stream_) {
    if (num_queries == 0)
        return 0;
    if (!points || !queries || !indices)
        return fail("rfb_nearest_point: NULL argument");
    if (num_points == 0)
        return fail("rfb_nearest_point: empty point set");
    cudaStream_t stream = (cudaStream_t)stream_;
    unsigned long long *best = nullptr;
    if (int rc = stream_scratch(reinterpret_cast<void **>(&best), sizeof(unsigned long long) * num_queries, stream))
        return rc;
    cudaMemsetAsync(best, 0xFF, sizeof(unsigned long long) * num_queries, stream);
    const int grid = grid_for((uint64_t)num_points, rfb::kNNTile, 148 * 4);
    RFB_LAUNCH((rfb::nearest_points_kernel), grid, 256, 0, stream, points, num_points, queries, num_queries,
               (const uint32_t *)nullptr, best);
    RFB_LAUNCH((rfb::nearest_points_finish_kernel), (num_queries + 255) / 256, 256, 0, stream, best, num_queries,
               indices);
    g_launches += 2;
    const cudaError_t launched = cudaGetLastError();
    cudaFreeAsync(best, stream); // stream-ordered: after the kernels above, also on the error path
    RFB_CUDA(launched);
    return 0;
}

int rfb_start_points(const float *points, uint32_t num_points, const float *rays, uint32_t num_rays,
                     uint32_t *start_point_index, void *stream_) {
    if (num_rays == 0)
        return 0;
    if (!points || !rays || !start_point_index)
        return fail("rfb_start_points: NULL argument");
    if (num_points == 0)
        return fail("rfb_start_points: empty point set");
    if (num_rays > 0x7FFFFFF0u)
        return fail("rfb_start_points: too many rays");
    cudaStream_t stream = (cudaStream_t)stream_;
    uint32_t table_size = 1024;
    while (table_size < 2u * num_rays)
        table_size <<= 1;
    // one scratch block: [best: R x u64][queries: R x 3 f32][table: T u32][query_of_slot: T u32]
    //                    [slot_of_ray: R u32][count: 4 u32]
    const size_t best_b = sizeof(unsigned long long) * (size_t)num_rays, query_b = 12 * (size_t)num_rays,
                 table_b = 4 * (size_t)table_size, slot_b = 4 * (size_t)num_rays;
    char *base = nullptr;
    if (int rc = stream_scratch(reinterpret_cast<void **>(&base), best_b + query_b + 2 * table_b + slot_b + 16, stream))
        return rc;
    auto *best = reinterpret_cast<unsigned long long *>(base);
    auto *queries = reinterpret_cast<float *>(base + best_b);
    auto *table = reinterpret_cast<uint32_t *>(base + best_b + query_b);
    auto *query_of_slot = table + table_size;
    auto *slot_of_ray = query_of_slot + table_size;
    auto *count = slot_of_ray + num_rays;
    cudaMemsetAsync(best, 0xFF, best_b, stream);
    cudaMemsetAsync(table, 0, table_b, stream);
    cudaMemsetAsync(count, 0, 16, stream);
    const uint32_t ray_grid = (num_rays + 255) / 256;
    RFB_LAUNCH((rfb::origin_slots_kernel), ray_grid, 256, 0, stream, rays, num_rays, table, table_size - 1,
               slot_of_ray, query_of_slot, queries, count);
    RFB_LAUNCH((rfb::nearest_points_kernel), grid_for((uint64_t)num_points, rfb::kNNTile, 148 * 4), 256, 0, stream,
               points, num_points, (const float *)queries, 0u, (const uint32_t *)count, best);
    RFB_LAUNCH((rfb::start_points_scatter_kernel), ray_grid, 256, 0, stream, (const uint32_t *)slot_of_ray, num_rays,
               (const uint32_t *)query_of_slot, (const unsigned long long *)best, start_point_index);
    g_launches += 3;
    const cudaError_t launched = cudaGetLastError();
    cudaFreeAsync(base, stream);
    RFB_CUDA(launched);
    return 0;
}

int rfb_farthest_neighbor(const float *points, uint32_t num_points, const uint32_t *point_adjacency,
                          const uint32_t *point_adjacency_offsets, uint32_t *indices, float *cell_radius,
                          void *stream_) {
    if (num_points == 0)
        return 0;
    if (!points || !point_adjacency_offsets || !indices || !cell_radius)
        return fail("rfb_farthest_neighbor: NULL argument");
    // point_adjacency may be NULL only when every row is empty; the kernels never read it then
    cudaStream_t stream = (cudaStream_t)stream_;
    const uint32_t grid = (uint32_t)(((uint64_t)num_points + 255) / 256);
    RFB_LAUNCH((rfb::farthest_neighbor_kernel), grid, 256, 0, stream, points, point_adjacency,
               point_adjacency_offsets, num_points, indices, cell_radius);
    RFB_LAUNCHED();
    return 0;
}

int rfb_trace_forward(rfb_pipeline *p, const rfb_trace_settings *settings, uint32_t num_points,
                      const float *points, const void *attributes, uint32_t point_adjacency_size,
                      const uint32_t *point_adjacency, const uint32_t *point_adjacency_offsets,
                      uint32_t num_rays, const float *rays, const uint32_t *start_point_index,
                      uint32_t num_depth_quantiles, const float *depth_quantiles, void *ray_rgba,
                      float *quantile_depths, uint32_t *quantile_point_indices,
                      uint32_t *num_intersections, void *point_contribution,
                      const rfb_launch_opts *opts, void *stream_) {
    if (!p)
        return fail("rfb_trace_forward: pipeline is NULL");
    if (num_rays == 0)
        return 0;
    if (p->params_bound)
        attributes = p->params.density; // identity for the mirror cache; the values come from the bound parameters
    if (!points || !attributes || !point_adjacency || !point_adjacency_offsets || !rays ||
        !start_point_index || !ray_rgba)
        return fail("rfb_trace_forward: NULL argument");
    if (depth_quantiles && num_depth_quantiles && (!quantile_depths || !quantile_point_indices))
        return fail("rfb_trace_forward: depth_quantiles given without output buffers");
    cudaStream_t stream = (cudaStream_t)stream_;
    if (int rc = ensure_scene(p, num_points, points, attributes, point_adjacency_size, point_adjacency,
                              point_adjacency_offsets, true, opts, stream))
        return rc;
    rfb_trace_settings s = settings_or_default(settings);
    ForwardParams fp;
    fp.cells = reinterpret_cast<const float4 *>(p->cells.ptr);
    fp.sh_rows = reinterpret_cast<const float *>(p->sh_rows.ptr);
    fp.rays = rays;
    fp.start = start_point_index;
    fp.quantiles = num_depth_quantiles ? depth_quantiles : nullptr;
    fp.rgba = ray_rgba;
    fp.qdepth = quantile_depths;
    fp.qidx = quantile_point_indices;
    fp.nint = num_intersections;
    fp.contrib = point_contribution;
    fp.num_rays = num_rays;
    fp.num_q = num_depth_quantiles;
    fp.weight_threshold = s.weight_threshold;
    fp.max_steps = s.max_intersections;
    fp.out_half = p->attr_dtype == RFB_FLOAT16;
    uint32_t blocks;
    ray_grid(num_rays, opts ? opts->image_width : 0, blocks, fp.blocks_x, fp.image_width);
    PaddedFaces fa;
    fa.faces = reinterpret_cast<const uint2 *>(p->faces.ptr);
    fa.nbr = reinterpret_cast<const uint32_t *>(p->nbr.ptr);
    fa.off = point_adjacency_offsets;
    RFB_CUDA(p->scan_flag.ensure(4 * sizeof(uint32_t)));
    RFB_CUDA(cudaMemsetAsync(p->scan_flag.ptr, 0, 4 * sizeof(uint32_t), stream));
    fp.exact_flag = reinterpret_cast<uint32_t *>(p->scan_flag.ptr);
    bool record = opts && (opts->flags & RFB_FLAG_RECORD_TAPE) && opts->scene_version != 0;
    p->tape_valid = false;
    Tape tape = {};
    if (record)
        if (int rc = prepare_tape(p, blocks, num_rays, s.max_intersections, tape, stream, record))
            return rc;
    if (int rc = profile_mark(p, 0, stream))
        return rc;
    if (int rc = record ? launch_forward_record(p->sh_degree, fp, fa, tape, blocks, stream)
                        : launch_forward(p->sh_degree, fp, fa, blocks, stream))
        return rc;
    if (int rc = profile_mark(p, 1, stream))
        return rc;
    if (record) {
        // the replay's longest-first schedule (small launches only: see tape_order_kernel)
        const char *sched = getenv("RFB_REPLAY_ORDER"); // "0" never, "1" always (for measuring)
        p->tape_scheduled = sched ? sched[0] == '1' : blocks <= 148u * 5u * 8u;
        if (p->tape_scheduled) {
            RFB_LAUNCH((tile_steps_kernel), (blocks + 7) / 8, 256, 0, stream, (const uint2 *)tape.per_ray, num_rays,
                       fp.image_width, fp.blocks_x, blocks, tape.tile_steps);
            RFB_LAUNCH((tape_order_kernel), 1, 1024, 0, stream, (const uint32_t *)tape.tile_steps, blocks, blocks,
                       tape.order);
            g_launches += 1;
            RFB_LAUNCHED();
        }
        if (!capturing(stream)) {
            RFB_CUDA(cudaMemcpyAsync(p->tape_readback, p->tape_ctrl.ptr, 4 * sizeof(uint32_t),
                                     cudaMemcpyDeviceToHost, stream));
            RFB_CUDA(cudaEventRecord(p->tape_readback_done, stream));
            p->tape_readback_pending = true;
        }
        p->tape_key.rays = rays;
        p->tape_key.start = start_point_index;
        p->tape_key.num_rays = num_rays;
        p->tape_key.image_width = fp.image_width;
        p->tape_key.max_steps = s.max_intersections;
        p->tape_key.weight_threshold = s.weight_threshold;
        p->tape_key.scene_version = opts->scene_version;
        p->tape_valid = true;
        if (int rc = mark(p->tape_ready, p->tape_stream, stream))
            return rc;
    }
    return 0;
}

int rfb_trace_backward_accumulate(rfb_pipeline *p, const rfb_trace_settings *settings,
                                  uint32_t num_points, const float *points, const void *attributes,
                                  uint32_t point_adjacency_size, const uint32_t *point_adjacency,
                                  const uint32_t *point_adjacency_offsets, uint32_t num_rays,
                                  const float *rays, const uint32_t *start_point_index,
                                  uint32_t num_depth_quantiles, const float *depth_quantiles,
                                  const uint32_t *quantile_point_indices, const void *ray_rgba,
                                  const void *ray_rgba_grad, const float *depth_grad,
                                  const void *ray_error, void *point_error,
                                  const rfb_launch_opts *opts, void *stream_) {
    if (!p)
        return fail("rfb_trace_backward: pipeline is NULL");
    cudaStream_t stream = (cudaStream_t)stream_;
    if (int rc = check_device(p))
        return rc;
    const int GR = grad_row(p->sh_degree);
    size_t acc_bytes = (size_t)num_points * GR * sizeof(float);
    float *acc_ptr = p->acc_external;
    if (acc_ptr) {
        if ((uint64_t)num_points * GR > p->acc_external_floats)
            return fail("rfb_trace_backward: the accumulator given to rfb_set_grad_accumulator is too small");
    } else {
        RFB_CUDA(p->acc.ensure(acc_bytes));
        acc_ptr = reinterpret_cast<float *>(p->acc.ptr);
    }
    RFB_CUDA(cudaMemsetAsync(acc_ptr, 0, acc_bytes, stream));
    p->acc_points = num_points;
    if (num_rays == 0)
        return 0;
    if (p->params_bound)
        attributes = p->params.density;
    if (!points || !attributes || !point_adjacency || !point_adjacency_offsets || !rays ||
        !start_point_index || !ray_rgba || !ray_rgba_grad)
        return fail("rfb_trace_backward: NULL argument");
    if (depth_quantiles && num_depth_quantiles && (!quantile_point_indices || !depth_grad))
        return fail("depth_grad must be provided if depth_quantiles is provided");
    if (int rc = ensure_scene(p, num_points, points, attributes, point_adjacency_size, point_adjacency,
                              point_adjacency_offsets, true, opts, stream))
        return rc;
    rfb_trace_settings s = settings_or_default(settings);
    BackwardParams bp;
    bp.cells = reinterpret_cast<const float4 *>(p->cells.ptr);
    bp.sh_rows = reinterpret_cast<const float *>(p->sh_rows.ptr);
    bp.rays = rays;
    bp.start = start_point_index;
    bp.quantiles = num_depth_quantiles ? depth_quantiles : nullptr;
    bp.qidx = quantile_point_indices;
    bp.rgba = ray_rgba;
    bp.rgba_grad = ray_rgba_grad;
    bp.depth_grad = depth_grad;
    bp.ray_error = ray_error;
    bp.point_error = ray_error ? point_error : nullptr;
    bp.acc = acc_ptr;
    bp.num_rays = num_rays;
    bp.num_q = num_depth_quantiles;
    bp.weight_threshold = s.weight_threshold;
    bp.max_steps = s.max_intersections;
    bp.io_half = p->attr_dtype == RFB_FLOAT16;
    uint32_t blocks;
    ray_grid(num_rays, opts ? opts->image_width : 0, blocks, bp.blocks_x, bp.image_width);
    PaddedFaces fa;
    fa.faces = reinterpret_cast<const uint2 *>(p->faces.ptr);
    fa.nbr = reinterpret_cast<const uint32_t *>(p->nbr.ptr);
    fa.off = point_adjacency_offsets;
    if (int rc = profile_mark(p, 2, stream))
        return rc;
    const char *mode = getenv("RFB_BWD_MODE"); // "direct" | "cached" (experiment switch)
    bool cached = mode ? strcmp(mode, "direct") != 0 : true;
    // replay the forward's walk tape when the caller vouches the inputs are those of the last
    // recording forward (and everything the walk depends on matches)
    Tape tape = {};
    const rfb_pipeline::TapeKey &tk = p->tape_key;
    if (cached && opts && (opts->flags & RFB_FLAG_USE_TAPE) && p->tape_valid && opts->scene_version != 0 &&
        tk.scene_version == opts->scene_version && tk.rays == rays && tk.start == start_point_index &&
        tk.num_rays == num_rays && tk.image_width == bp.image_width && tk.max_steps == s.max_intersections &&
        tk.weight_threshold == s.weight_threshold) {
        tape.pool = reinterpret_cast<uint2 *>(p->tape_pool.ptr);
        tape.table = reinterpret_cast<uint32_t *>(p->tape_table.ptr);
        tape.per_ray = reinterpret_cast<uint2 *>(p->tape_per_ray.ptr);
        tape.ctrl = reinterpret_cast<uint32_t *>(p->tape_ctrl.ptr);
        tape.capacity = p->tape_capacity;
        tape.table_stride = p->tape_table_stride;
        tape.tile_steps = reinterpret_cast<uint32_t *>(p->tape_sched.ptr);
        tape.order = (blocks == p->tape_blocks && p->tape_scheduled) ? tape.tile_steps + blocks : nullptr;
        if (int rc = wait_for(p->tape_ready, p->tape_stream, stream))
            return rc;
    }
    if (int rc = cached ? launch_backward_cached(p->sh_degree, bp, fa, tape, blocks, stream)
                        : launch_backward(p->sh_degree, bp, fa, blocks, stream))
        return rc;
    return profile_mark(p, 3, stream);
}

int rfb_grad_accumulator(rfb_pipeline *p, float **ptr, uint64_t *num_floats) {
    if (!p || !ptr || !num_floats)
        return fail("rfb_grad_accumulator: NULL argument");
    *ptr = p->acc_external ? p->acc_external : reinterpret_cast<float *>(p->acc.ptr);
    *num_floats = (uint64_t)p->acc_points * grad_row(p->sh_degree);
    return 0;
}

int rfb_set_grad_accumulator(rfb_pipeline *p, float *ptr, uint64_t num_floats) {
    if (!p)
        return fail("rfb_set_grad_accumulator: pipeline is NULL");
    if (ptr && (reinterpret_cast<uintptr_t>(ptr) & 15u))
        return fail("rfb_set_grad_accumulator: the accumulator must be 16-byte aligned");
    p->acc_external = ptr;
    p->acc_external_floats = ptr ? num_floats : 0;
    p->acc_points = 0;
    return 0;
}

int rfb_reduce_finalize_peers(rfb_pipeline *p, uint32_t world, uint32_t rank, uint32_t num_points,
                              const float *const *peer_acc, void *const *peer_attribute_grad,
                              float *const *peer_points_grad, const rfb_multicast *multicast, uint32_t flags,
                              void *stream_) {
    if (!p || !peer_acc || !peer_attribute_grad || !peer_points_grad)
        return fail("rfb_reduce_finalize_peers: NULL argument");
    if (multicast && (!multicast->acc || !multicast->attribute_grad || !multicast->points_grad))
        return fail("rfb_reduce_finalize_peers: incomplete multicast addresses");
    if (multicast && ((reinterpret_cast<uintptr_t>(multicast->acc) | reinterpret_cast<uintptr_t>(multicast->attribute_grad) |
                       reinterpret_cast<uintptr_t>(multicast->points_grad)) & 15u))
        return fail("rfb_reduce_finalize_peers: multicast addresses must be 16-byte aligned");
    if (world == 0 || world > (uint32_t)kMaxPeers || rank >= world)
        return fail("rfb_reduce_finalize_peers: world must be 1.." + std::to_string(kMaxPeers) + " and rank < world");
    if (num_points == 0)
        return 0;
    PeerReduceParams pr;
    for (uint32_t w = 0; w < world; ++w) {
        if (!peer_acc[w] || !peer_attribute_grad[w] || !peer_points_grad[w])
            return fail("rfb_reduce_finalize_peers: NULL peer pointer");
        if ((reinterpret_cast<uintptr_t>(peer_acc[w]) | reinterpret_cast<uintptr_t>(peer_attribute_grad[w]) |
             reinterpret_cast<uintptr_t>(peer_points_grad[w])) & 15u)
            return fail("rfb_reduce_finalize_peers: peer arrays must be 16-byte aligned");
        pr.acc[w] = peer_acc[w];
        pr.attr_grad[w] = peer_attribute_grad[w];
        pr.points_grad[w] = peer_points_grad[w];
    }
    pr.world = world;
    pr.num_points = num_points;
    // this rank's share: whole blocks of kPeerRows rows, dealt round-robin so that every rank streams from
    // the whole array (no rank's HBM serves one reader only)
    const uint32_t num_blocks = (num_points + kPeerRows - 1) / kPeerRows;
    pr.first_block = rank;
    pr.block_stride = world;
    pr.scrub = (flags & RFB_FLAG_SCRUB_NONFINITE) ? 1 : 0;
    pr.rank = rank;
    const char *dbg = getenv("RFB_PEER_DEBUG");
    pr.debug_mode = dbg ? atoi(dbg) : 0;
    pr.mc_acc = multicast ? multicast->acc : nullptr;
    pr.mc_attr_grad = multicast ? multicast->attribute_grad : nullptr;
    pr.mc_points_grad = multicast ? multicast->points_grad : nullptr;
    const uint32_t mine = (num_blocks + world - 1 - rank) / world;
    if (mine == 0)
        return 0;
    int grid = (int)(mine < 148u * 16u ? mine : 148u * 16u);
    cudaStream_t stream = (cudaStream_t)stream_;
    const bool half = p->attr_dtype == RFB_FLOAT16;
#define RFB_PEER_LAUNCH(DEG, W)                                                                          \
    do {                                                                                                 \
        if (half)                                                                                        \
            RFB_LAUNCH((reduce_finalize_peers_kernel<DEG, __half, W>), grid, 256, 0, stream, pr);        \
        else                                                                                             \
            RFB_LAUNCH((reduce_finalize_peers_kernel<DEG, float, W>), grid, 256, 0, stream, pr);         \
    } while (0)
#define RFB_PEER_WORLD(DEG)                                                                              \
    do {                                                                                                 \
        switch (world) {                                                                                 \
        case 2: RFB_PEER_LAUNCH(DEG, 2); break;                                                          \
        case 4: RFB_PEER_LAUNCH(DEG, 4); break;                                                          \
        case 8: RFB_PEER_LAUNCH(DEG, 8); break;                                                          \
        default: RFB_PEER_LAUNCH(DEG, 0); break;                                                         \
        }                                                                                                \
    } while (0)
    switch (p->sh_degree) {
    case 0: RFB_PEER_WORLD(0); break;
    case 1: RFB_PEER_WORLD(1); break;
    case 2: RFB_PEER_WORLD(2); break;
    default: RFB_PEER_WORLD(3); break;
    }
#undef RFB_PEER_WORLD
#undef RFB_PEER_LAUNCH
    RFB_LAUNCHED();
    return 0;
}

int rfb_trace_backward_finalize(rfb_pipeline *p, uint32_t num_points, float *points_grad,
                                void *attribute_grad, uint32_t flags, void *stream_) {
    if (!p)
        return fail("rfb_trace_backward_finalize: pipeline is NULL");
    if (num_points == 0)
        return 0;
    if (!points_grad || !attribute_grad)
        return fail("rfb_trace_backward_finalize: NULL argument");
    if (num_points != p->acc_points)
        return fail("rfb_trace_backward_finalize: no accumulated gradients for this point count");
    cudaStream_t stream = (cudaStream_t)stream_;
    const int A = attr_dim(p->sh_degree);
    const int grid = grid_for((uint64_t)num_points * A, 256 * 4, 148 * 8);
    int scrub = (flags & RFB_FLAG_SCRUB_NONFINITE) ? 1 : 0;
    const float *acc_ptr = p->acc_external ? p->acc_external : reinterpret_cast<const float *>(p->acc.ptr);
#define RFB_FINALIZE(DEG)                                                                                          \
    do {                                                                                                           \
        if (p->attr_dtype == RFB_FLOAT16)                                                                          \
            RFB_LAUNCH((finalize_grads_kernel<__half, DEG>), grid, 256, 0, stream, acc_ptr, num_points,            \
                       points_grad, reinterpret_cast<__half *>(attribute_grad), scrub);                            \
        else                                                                                                       \
            RFB_LAUNCH((finalize_grads_kernel<float, DEG>), grid, 256, 0, stream, acc_ptr, num_points,             \
                       points_grad, reinterpret_cast<float *>(attribute_grad), scrub);                             \
    } while (0)
    switch (p->sh_degree) {
    case 0: RFB_FINALIZE(0); break;
    case 1: RFB_FINALIZE(1); break;
    case 2: RFB_FINALIZE(2); break;
    default: RFB_FINALIZE(3); break;
    }
#undef RFB_FINALIZE
    RFB_LAUNCHED();
    return 0;
}

int rfb_bind_scene_params(rfb_pipeline *p, const rfb_scene_params *params) {
    if (!p)
        return fail("rfb_bind_scene_params: pipeline is NULL");
    p->key_valid = false;
    if (!params) {
        p->params_bound = false;
        return 0;
    }
    if (!params->att_dc || !params->density || (p->sh_degree > 0 && !params->att_sh))
        return fail("rfb_bind_scene_params: NULL parameter array");
    p->params = *params;
    p->params_bound = true;
    return 0;
}

int rfb_trace_backward_finalize_params(rfb_pipeline *p, uint32_t num_points, float *points_grad,
                                       float *att_dc_grad, float *att_sh_grad, float *density_grad,
                                       uint32_t flags, void *stream_) {
    if (!p)
        return fail("rfb_trace_backward_finalize_params: pipeline is NULL");
    if (!p->params_bound)
        return fail("rfb_trace_backward_finalize_params: no parameter-form scene is bound");
    if (num_points == 0)
        return 0;
    if (!points_grad || !att_dc_grad || !density_grad || (p->sh_degree > 0 && !att_sh_grad))
        return fail("rfb_trace_backward_finalize_params: NULL argument");
    if (num_points != p->acc_points)
        return fail("rfb_trace_backward_finalize_params: no accumulated gradients for this point count");
    cudaStream_t stream = (cudaStream_t)stream_;
    const int A = attr_dim(p->sh_degree);
    const int grid = grid_for((uint64_t)num_points * A, 256 * 4, 148 * 8);
    int scrub = (flags & RFB_FLAG_SCRUB_NONFINITE) ? 1 : 0;
    const float *acc_ptr = p->acc_external ? p->acc_external : reinterpret_cast<const float *>(p->acc.ptr);
#define RFB_FINALIZE_PARAMS(DEG)                                                                                   \
    do {                                                                                                           \
        if (p->attr_dtype == RFB_FLOAT16)                                                                          \
            RFB_LAUNCH((finalize_params_kernel<__half, DEG>), grid, 256, 0, stream, acc_ptr, p->params.density,    \
                       p->params.activation_scale, num_points, points_grad, att_dc_grad, att_sh_grad,              \
                       density_grad, scrub);                                                                       \
        else                                                                                                       \
            RFB_LAUNCH((finalize_params_kernel<float, DEG>), grid, 256, 0, stream, acc_ptr, p->params.density,     \
                       p->params.activation_scale, num_points, points_grad, att_dc_grad, att_sh_grad,              \
                       density_grad, scrub);                                                                       \
    } while (0)
    switch (p->sh_degree) {
    case 0: RFB_FINALIZE_PARAMS(0); break;
    case 1: RFB_FINALIZE_PARAMS(1); break;
    case 2: RFB_FINALIZE_PARAMS(2); break;
    default: RFB_FINALIZE_PARAMS(3); break;
    }
#undef RFB_FINALIZE_PARAMS
    RFB_LAUNCHED();
    return 0;
}

int rfb_trace_backward(rfb_pipeline *p, const rfb_trace_settings *settings, uint32_t num_points,
                       const float *points, const void *attributes, uint32_t point_adjacency_size,
                       const uint32_t *point_adjacency, const uint32_t *point_adjacency_offsets,
                       uint32_t num_rays, const float *rays, const uint32_t *start_point_index,
                       uint32_t num_depth_quantiles, const float *depth_quantiles,
                       const uint32_t *quantile_point_indices, const void *ray_rgba,
                       const void *ray_rgba_grad, const float *depth_grad, const void *ray_error,
                       float *ray_grad, float *points_grad, void *attribute_grad, void *point_error,
                       const rfb_launch_opts *opts, void *stream) {
    (void)ray_grad; // never written, like the reference kernel
    if (p && p->params_bound)
        return fail("rfb_trace_backward: a parameter-form scene is bound; use rfb_trace_backward_accumulate + "
                    "rfb_trace_backward_finalize_params");
    if (int rc = rfb_trace_backward_accumulate(
            p, settings, num_points, points, attributes, point_adjacency_size, point_adjacency,
            point_adjacency_offsets, num_rays, rays, start_point_index, num_depth_quantiles,
            depth_quantiles, quantile_point_indices, ray_rgba, ray_rgba_grad, depth_grad, ray_error,
            point_error, opts, stream))
        return rc;
    return rfb_trace_backward_finalize(p, num_points, points_grad, attribute_grad,
                                       opts ? opts->flags : 0, stream);
}

int rfb_trace_benchmark(rfb_pipeline *p, const rfb_trace_settings *settings, uint32_t num_points,
                        const float *points, const void *attributes, const uint32_t *point_adjacency,
                        const uint32_t *point_adjacency_offsets, const void *adjacent_diff,
                        const rfb_camera *camera, const uint32_t *start_point_index,
                        uint32_t *output_rgba, const rfb_launch_opts *opts, void *stream_) {
    if (!p)
        return fail("rfb_trace_benchmark: pipeline is NULL");
    if (!points || !attributes || !point_adjacency || !point_adjacency_offsets || !adjacent_diff ||
        !camera || !start_point_index || !output_rgba)
        return fail("rfb_trace_benchmark: NULL argument");
    if (camera->model != RFB_PINHOLE && camera->model != RFB_FISHEYE)
        return fail("Invalid camera model");
    if (camera->width == 0 || camera->height == 0)
        return 0;
    cudaStream_t stream = (cudaStream_t)stream_;
    // The reference interface does not pass the adjacency size (pipeline.h:117-126); it is
    // offsets[N].  Reading it costs a stream sync, paid only when the mirrors are (re)built.
    uint32_t num_edges = p->key_valid ? p->key.e : 0;
    {
        SceneKey probe;
        probe.points = points;
        probe.attrs = attributes;
        probe.adj = point_adjacency;
        probe.off = point_adjacency_offsets;
        probe.diff = adjacent_diff;
        probe.n = num_points;
        probe.e = num_edges;
        probe.version = opts ? opts->scene_version : 0;
        bool hit = p->key_valid && probe.version != 0 && probe == p->key && p->key.faces;
        if (!hit) {
            RFB_CUDA(cudaMemcpyAsync(&num_edges, point_adjacency_offsets + num_points, sizeof(uint32_t),
                                     cudaMemcpyDeviceToHost, stream));
            RFB_CUDA(cudaStreamSynchronize(stream));
        }
    }
    if (int rc = ensure_scene(p, num_points, points, attributes, num_edges, point_adjacency,
                              point_adjacency_offsets, true, opts, stream, adjacent_diff))
        return rc;
    rfb_trace_settings s = settings_or_default(settings);
    BenchmarkParams bp;
    bp.cells = reinterpret_cast<const float4 *>(p->cells.ptr);
    bp.sh_rows = reinterpret_cast<const float *>(p->sh_rows.ptr);
    bp.start = start_point_index;
    bp.out = output_rgba;
    memcpy(bp.cam.position, camera->position, sizeof(float) * 3);
    memcpy(bp.cam.forward, camera->forward, sizeof(float) * 3);
    memcpy(bp.cam.right, camera->right, sizeof(float) * 3);
    memcpy(bp.cam.up, camera->up, sizeof(float) * 3);
    bp.cam.fov = camera->fov;
    bp.cam.width = camera->width;
    bp.cam.height = camera->height;
    bp.cam.model = camera->model;
    bp.weight_threshold = s.weight_threshold;
    bp.max_steps = s.max_intersections;
    RFB_CUDA(p->scan_flag.ensure(4 * sizeof(uint32_t)));
    RFB_CUDA(cudaMemsetAsync(p->scan_flag.ptr, 0, 4 * sizeof(uint32_t), stream));
    bp.exact_flag = reinterpret_cast<uint32_t *>(p->scan_flag.ptr);
    bp.blocks_x = (camera->width + kTileW - 1) / kTileW;
    uint32_t blocks = bp.blocks_x * ((camera->height + kTileH - 1) / kTileH);
    // the caller's offsets, re-laid-out (copied, not recomputed) into the padded face rows
    PaddedFaces fa;
    fa.faces = reinterpret_cast<const uint2 *>(p->faces.ptr);
    fa.nbr = reinterpret_cast<const uint32_t *>(p->nbr.ptr);
    fa.off = point_adjacency_offsets;
    return launch_benchmark(p->sh_degree, bp, fa, blocks, stream);
}

} // extern "C"
