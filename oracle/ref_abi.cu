// C-ABI shim around the reference's own tracing pipeline — TEST INFRASTRUCTURE ONLY.
//
// This file is compiled together with /root/reference/src/tracing/pipeline.cu
// (taken from where it lies, unmodified, never copied into this repo) into
// oracle/_ref/libradfoam_ref.so.  It supplies the single symbol the reference TU
// leaves unresolved, radfoam::allocate_buffer (declared src/utils/cuda_array.h:35,
// defined upstream in torch_bindings/torch_bindings.cpp:13-29 on top of
// torch::empty) and flat extern "C" entry points so tests / bench.py can drive
// the reference's Pipeline::trace_forward / trace_backward / trace_benchmark
// (src/tracing/pipeline.h:58-133) through ctypes with raw device pointers.
//
// The upstream allocator is torch's caching allocator; a per-call cudaMalloc /
// cudaFree would add a device synchronisation the reference does not pay, so
// allocate_buffer here keeps freed blocks in a small size-keyed cache.
#include <cuda_runtime.h>

#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "delaunay/triangulation_ops.h"
#include "tracing/pipeline.h"
#include "utils/cuda_array.h"

namespace {

std::mutex g_pool_mutex;
std::multimap<size_t, void *> g_pool; // free blocks by size
thread_local std::string g_last_error;

struct PooledBuffer : public radfoam::OpaqueBuffer {
    void *ptr;
    size_t bytes;
    PooledBuffer(void *p, size_t b) : ptr(p), bytes(b) {}
    ~PooledBuffer() override {
        // stream-ordered reuse on the legacy default stream, like the caching
        // allocator the reference runs on: no synchronisation on release.
        std::lock_guard<std::mutex> lock(g_pool_mutex);
        g_pool.emplace(bytes, ptr);
    }
    void *data() override { return ptr; }
};

} // namespace

namespace radfoam {

std::unique_ptr<OpaqueBuffer> allocate_buffer(size_t bytes) {
    {
        std::lock_guard<std::mutex> lock(g_pool_mutex);
        auto it = g_pool.find(bytes);
        if (it != g_pool.end()) {
            void *p = it->second;
            g_pool.erase(it);
            return std::make_unique<PooledBuffer>(p, bytes);
        }
    }
    void *p = nullptr;
    cuda_check(cudaMalloc(&p, bytes));
    return std::make_unique<PooledBuffer>(p, bytes);
}

} // namespace radfoam

namespace {

std::shared_ptr<radfoam::Pipeline> get_pipeline(int sh_degree, int attr_is_half) {
    static std::mutex m;
    static std::map<int, std::shared_ptr<radfoam::Pipeline>> cache;
    std::lock_guard<std::mutex> lock(m);
    int key = sh_degree * 2 + (attr_is_half ? 1 : 0);
    auto it = cache.find(key);
    if (it != cache.end())
        return it->second;
    auto p = radfoam::create_pipeline(
        sh_degree, attr_is_half ? radfoam::Float16 : radfoam::Float32);
    cache[key] = p;
    return p;
}

template <typename F>
int guarded(F &&f) {
    try {
        f();
        return 0;
    } catch (const std::exception &e) {
        g_last_error = e.what();
        return 1;
    }
}

} // namespace

extern "C" {

const char *rfref_last_error() { return g_last_error.c_str(); }

void rfref_release_pool() {
    std::lock_guard<std::mutex> lock(g_pool_mutex);
    cudaDeviceSynchronize();
    for (auto &kv : g_pool)
        cudaFree(kv.second);
    g_pool.clear();
}

int rfref_attribute_dim(int sh_degree, int attr_is_half) {
    int out = -1;
    guarded([&] { out = (int)get_pipeline(sh_degree, attr_is_half)->attribute_dim(); });
    return out;
}

// radfoam::prefetch_adjacent_diff, src/tracing/pipeline.h:50-56
int rfref_prefetch_adjacent_diff(const float *points, uint32_t num_points,
                                 uint32_t point_adjacency_size,
                                 const uint32_t *point_adjacency,
                                 const uint32_t *point_adjacency_offsets,
                                 void *adjacent_diff) {
    return guarded([&] {
        radfoam::prefetch_adjacent_diff(
            reinterpret_cast<const radfoam::Vec3f *>(points), num_points,
            point_adjacency_size, point_adjacency, point_adjacency_offsets,
            reinterpret_cast<radfoam::Vec4h *>(adjacent_diff), nullptr);
    });
}

// radfoam::farthest_neighbor, src/delaunay/triangulation_ops.h:8-15 (its TU, triangulation_ops.cu, is the
// second reference source compiled into this library, also unmodified)
int rfref_farthest_neighbor(const float *points, uint32_t num_points, const uint32_t *point_adjacency,
                            const uint32_t *point_adjacency_offsets, uint32_t *indices,
                            float *cell_radius) {
    return guarded([&] {
        radfoam::farthest_neighbor(radfoam::Float32, points, point_adjacency,
                                   point_adjacency_offsets, num_points, indices, cell_radius, nullptr);
    });
}

// Pipeline::trace_forward, src/tracing/pipeline.h:62-78
int rfref_trace_forward(int sh_degree, int attr_is_half, float weight_threshold,
                        uint32_t max_intersections, uint32_t num_points,
                        const float *points, const void *attributes,
                        uint32_t point_adjacency_size,
                        const uint32_t *point_adjacency,
                        const uint32_t *point_adjacency_offsets,
                        uint32_t num_rays, const float *rays,
                        const uint32_t *start_point_index,
                        uint32_t num_depth_quantiles,
                        const float *depth_quantiles, void *ray_rgba,
                        float *quantile_depths,
                        uint32_t *quantile_point_indices,
                        uint32_t *num_intersections, void *point_contribution) {
    return guarded([&] {
        radfoam::TraceSettings s;
        s.weight_threshold = weight_threshold;
        s.max_intersections = max_intersections;
        get_pipeline(sh_degree, attr_is_half)
            ->trace_forward(s, num_points,
                            reinterpret_cast<const radfoam::Vec3f *>(points),
                            attributes, point_adjacency_size, point_adjacency,
                            point_adjacency_offsets, num_rays,
                            reinterpret_cast<const radfoam::Ray *>(rays),
                            start_point_index, num_depth_quantiles,
                            depth_quantiles, ray_rgba, quantile_depths,
                            quantile_point_indices, num_intersections,
                            point_contribution);
    });
}

// Pipeline::trace_backward, src/tracing/pipeline.h:80-100
int rfref_trace_backward(int sh_degree, int attr_is_half, float weight_threshold,
                         uint32_t max_intersections, uint32_t num_points,
                         const float *points, const void *attributes,
                         uint32_t point_adjacency_size,
                         const uint32_t *point_adjacency,
                         const uint32_t *point_adjacency_offsets,
                         uint32_t num_rays, const float *rays,
                         const uint32_t *start_point_index,
                         uint32_t num_depth_quantiles,
                         const float *depth_quantiles,
                         const uint32_t *quantile_point_indices,
                         const void *ray_rgba, const void *ray_rgba_grad,
                         const float *depth_grad, const void *ray_error,
                         float *ray_grad, float *points_grad,
                         void *attribute_grad, void *point_error) {
    return guarded([&] {
        radfoam::TraceSettings s;
        s.weight_threshold = weight_threshold;
        s.max_intersections = max_intersections;
        get_pipeline(sh_degree, attr_is_half)
            ->trace_backward(s, num_points,
                             reinterpret_cast<const radfoam::Vec3f *>(points),
                             attributes, point_adjacency_size, point_adjacency,
                             point_adjacency_offsets, num_rays,
                             reinterpret_cast<const radfoam::Ray *>(rays),
                             start_point_index, num_depth_quantiles,
                             depth_quantiles, quantile_point_indices, ray_rgba,
                             ray_rgba_grad, depth_grad, ray_error,
                             reinterpret_cast<radfoam::Ray *>(ray_grad),
                             reinterpret_cast<radfoam::Vec3f *>(points_grad),
                             attribute_grad, point_error);
    });
}

// Pipeline::trace_benchmark, src/tracing/pipeline.h:117-126.
// camera: position[3], forward[3], right[3], up[3]; model 0 = pinhole, 1 = fisheye
// (src/tracing/camera.h:12-24).
int rfref_trace_benchmark(int sh_degree, int attr_is_half, float weight_threshold,
                          uint32_t max_intersections, uint32_t num_points,
                          const float *points, const void *attributes,
                          const uint32_t *point_adjacency,
                          const uint32_t *point_adjacency_offsets,
                          const void *adjacent_diff, const float *cam_position,
                          const float *cam_forward, const float *cam_right,
                          const float *cam_up, float fov, uint32_t width,
                          uint32_t height, int model,
                          const uint32_t *start_point_index,
                          uint32_t *output_rgba) {
    return guarded([&] {
        radfoam::TraceSettings s;
        s.weight_threshold = weight_threshold;
        s.max_intersections = max_intersections;
        radfoam::Camera camera;
        camera.position = radfoam::Vec3f(cam_position);
        camera.forward = radfoam::Vec3f(cam_forward);
        camera.right = radfoam::Vec3f(cam_right);
        camera.up = radfoam::Vec3f(cam_up);
        camera.fov = fov;
        camera.width = width;
        camera.height = height;
        camera.model = model == 0 ? radfoam::Pinhole : radfoam::Fisheye;
        get_pipeline(sh_degree, attr_is_half)
            ->trace_benchmark(s, num_points,
                              reinterpret_cast<const radfoam::Vec3f *>(points),
                              attributes, point_adjacency,
                              point_adjacency_offsets,
                              reinterpret_cast<const radfoam::Vec4h *>(adjacent_diff),
                              camera, start_point_index, output_rgba);
    });
}

} // extern "C"
