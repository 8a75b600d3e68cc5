// Density control on the device: split / clone / prune as one plan + one apply pass,
// stream compaction and append without the host round trips of the tensor-op formulation.
//
// Reference semantics (the working subset; gaussian_model.py:130-197, optimizer.py:43-71, restated in
// mini-3d-gaussian-splatting_b200/scene.py + training.py and pinned by tests/test_training.py):
//   clone:  |grad| > th and mean(sigma) < small  -> keep the splat, append a copy at xyz + noise * 0.5*mean(sigma)
//   split:  |grad| > th and mean(sigma) > large  -> remove the splat, append two children at
//           xyz -+ R[:,0] * 0.5*mean(sigma) with sigma * 0.75 and opacity logit clamped to [-6, 6]
//   prune:  every row of the result whose sigmoid(opacity) <= min_opacity is dropped
// Output order = the order the sequential formulation produces:
//   [ surviving originals | clone copies | "minus" children | "plus" children ], each in index order.
#include "common.cuh"

#include <cub/cub.cuh>

namespace gs {

struct Tri {
    int a, b, c;       // kept originals, clone copies, split parents (that survive the opacity test)
    int d;             // clone CANDIDATES (before the opacity test): the rank in this channel indexes the jitter noise,
                       // which the sequential formulation draws as randn(k, 3) for all k candidates
};
struct TriSum {
    __host__ __device__ __forceinline__ Tri operator()(const Tri& x, const Tri& y) const {
        return Tri{x.a + y.a, x.b + y.b, x.c + y.c, x.d + y.d};
    }
};

__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }
// the child's opacity parameter: logit(sigmoid(o)) clamped to [-6, 6] (scene.py density_and_split)
__device__ __forceinline__ float child_opacity(float o) {
    const float p = sigmoid_f(o);
    return fminf(fmaxf(logf(p / (1.0f - p)), -6.0f), 6.0f);
}

__global__ void __launch_bounds__(256)
densify_classify_kernel(int64_t n, const float* __restrict__ scaling_log, const float* __restrict__ opacity,
                        const float* __restrict__ grad, float th, float small_sigma, float large_sigma, float min_opacity,
                        Tri* __restrict__ flags) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float gx = grad[i * 3 + 0], gy = grad[i * 3 + 1], gz = grad[i * 3 + 2];
    const float gn = sqrtf(gx * gx + gy * gy + gz * gz);
    const float sig = (expf(scaling_log[i * 3 + 0]) + expf(scaling_log[i * 3 + 1]) + expf(scaling_log[i * 3 + 2])) / 3.0f;
    const bool hot = gn > th;
    const bool clone = hot && sig < small_sigma;
    const bool split = hot && sig > large_sigma;
    const float o = opacity[i];
    const bool keep_self = sigmoid_f(o) > min_opacity;
    const bool keep_child = sigmoid_f(child_opacity(o)) > min_opacity;
    Tri f;
    f.a = (!split && keep_self) ? 1 : 0;
    f.b = (clone && keep_self) ? 1 : 0;
    f.c = (split && keep_child) ? 1 : 0;
    f.d = clone ? 1 : 0;
    flags[i] = f;
}

__global__ void densify_totals_kernel(int64_t n, const Tri* __restrict__ flags, const Tri* __restrict__ pos, int64_t* counts) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        Tri t = {0, 0, 0, 0};
        if (n > 0) t = Tri{pos[n - 1].a + flags[n - 1].a, pos[n - 1].b + flags[n - 1].b, pos[n - 1].c + flags[n - 1].c,
                           pos[n - 1].d + flags[n - 1].d};
        counts[0] = t.a;
        counts[1] = t.b;
        counts[2] = t.c;
        counts[3] = (int64_t)t.a + t.b + 2 * (int64_t)t.c;
        counts[4] = t.d;
    }
}

struct ModelPtrs {
    const float *xyz, *dc, *rest, *scaling, *rotation, *opacity;
    float *o_xyz, *o_dc, *o_rest, *o_scaling, *o_rotation, *o_opacity;
};

__device__ __forceinline__ void copy_row(const float* __restrict__ src, float* __restrict__ dst, int64_t s, int64_t d, int w) {
    for (int k = 0; k < w; ++k) dst[d * w + k] = src[s * w + k];
}

// One thread per original splat writes all the rows it produces (0..2 of them besides itself).
__global__ void __launch_bounds__(256)
densify_apply_kernel(int64_t n, const Tri* __restrict__ flags, const Tri* __restrict__ pos, int64_t kept, int64_t cloned,
                     int64_t split, ModelPtrs m, const float* __restrict__ noise, int32_t* __restrict__ src_row) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Tri f = flags[i], p = pos[i];
    if (!(f.a | f.b | f.c)) return;
    const float s0 = expf(m.scaling[i * 3 + 0]), s1 = expf(m.scaling[i * 3 + 1]), s2 = expf(m.scaling[i * 3 + 2]);
    const float half_mean = (s0 + s1 + s2) / 3.0f * 0.5f;
    if (f.a) {                                      // surviving original
        const int64_t d = p.a;
        copy_row(m.xyz, m.o_xyz, i, d, 3); copy_row(m.dc, m.o_dc, i, d, 3); copy_row(m.rest, m.o_rest, i, d, 45);
        copy_row(m.scaling, m.o_scaling, i, d, 3); copy_row(m.rotation, m.o_rotation, i, d, 4);
        m.o_opacity[d] = m.opacity[i];
        if (src_row) src_row[d] = (int32_t)i;       // the only rows with a history (optimiser moments follow them)
    }
    if (f.b) {                                      // clone copy: jittered position, everything else equal
        const int64_t d = kept + p.b;
#pragma unroll
        for (int k = 0; k < 3; ++k) m.o_xyz[d * 3 + k] = m.xyz[i * 3 + k] + noise[(int64_t)p.d * 3 + k] * half_mean;
        copy_row(m.dc, m.o_dc, i, d, 3); copy_row(m.rest, m.o_rest, i, d, 45);
        copy_row(m.scaling, m.o_scaling, i, d, 3); copy_row(m.rotation, m.o_rotation, i, d, 4);
        m.o_opacity[d] = m.opacity[i];
        if (src_row) src_row[d] = -1;
    }
    if (f.c) {                                      // two children along the first principal axis
        const float4 q = *reinterpret_cast<const float4*>(m.rotation + i * 4);
        const float qn = fmaxf(sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w), 1e-12f);
        const float w = q.x / qn, x = q.y / qn, y = q.z / qn, z = q.w / qn;
        // first column of the rotation matrix (math_utils.py:9-26)
        const float ax = 1.f - 2.f * (y * y + z * z), ay = 2.f * (x * y + w * z), az = 2.f * (x * z - w * y);
        const float ox = ax * half_mean, oy = ay * half_mean, oz = az * half_mean;
        const float co = child_opacity(m.opacity[i]);
        const float c0 = logf(s0 * 0.75f), c1 = logf(s1 * 0.75f), c2 = logf(s2 * 0.75f);
        const int64_t base = kept + cloned;
#pragma unroll
        for (int side = 0; side < 2; ++side) {
            const int64_t d = base + side * split + p.c;
            const float sgn = side ? 1.f : -1.f;
            m.o_xyz[d * 3 + 0] = m.xyz[i * 3 + 0] + sgn * ox;
            m.o_xyz[d * 3 + 1] = m.xyz[i * 3 + 1] + sgn * oy;
            m.o_xyz[d * 3 + 2] = m.xyz[i * 3 + 2] + sgn * oz;
            copy_row(m.dc, m.o_dc, i, d, 3); copy_row(m.rest, m.o_rest, i, d, 45);
            m.o_scaling[d * 3 + 0] = c0; m.o_scaling[d * 3 + 1] = c1; m.o_scaling[d * 3 + 2] = c2;
            // the unit quaternion the accessor returns (scene.py: rot = get_rotation[mask])
            *reinterpret_cast<float4*>(m.o_rotation + d * 4) = make_float4(w, x, y, z);
            m.o_opacity[d] = co;
            if (src_row) src_row[d] = -1;
        }
    }
}

static int64_t align256(int64_t x) { return (x + 255) / 256 * 256; }
struct DensifyLayout {
    int64_t flags, pos, cub_temp, cub_bytes, total;
};
static DensifyLayout densify_layout(int64_t n) {
    DensifyLayout L;
    size_t scan_bytes = 0;
    cub::DeviceScan::ExclusiveScan(nullptr, scan_bytes, (const Tri*)nullptr, (Tri*)nullptr, TriSum(), Tri{0, 0, 0, 0}, n);
    int64_t o = 0;
    L.flags = o; o += align256(n * (int64_t)sizeof(Tri));
    L.pos = o;   o += align256(n * (int64_t)sizeof(Tri));
    L.cub_temp = o;
    L.cub_bytes = align256((int64_t)scan_bytes);
    o += L.cub_bytes;
    L.total = o;
    return L;
}

}  // namespace gs

using namespace gs;

extern "C" int64_t gs_densify_workspace_bytes(int64_t n) {
    if (n < 0) return GS_ERR_INVALID_ARGUMENT;
    return densify_layout(n > 0 ? n : 1).total + 256;
}

extern "C" int gs_densify_plan(int64_t n, const float* scaling_log, const float* opacity, const float* grad,
                               float grad_threshold, float small_sigma, float large_sigma, float min_opacity,
                               void* workspace, int64_t workspace_bytes, int64_t* counts, void* stream) {
    GS_REQUIRE(n >= 0, "n < 0");
    GS_REQUIRE(counts != nullptr, "counts is NULL");
    cudaStream_t st = (cudaStream_t)stream;
    DeviceGuard guard(counts);
    if (n == 0) {
        GS_CUDA_TRY(cudaMemsetAsync(counts, 0, 5 * sizeof(int64_t), st));
        return GS_OK;
    }
    GS_REQUIRE(n < (1ll << 30), "n too large");
    GS_REQUIRE(scaling_log && opacity && grad && workspace, "NULL array argument");
    const DensifyLayout L = densify_layout(n);
    if (workspace_bytes < L.total) {
        set_error("gs_densify_plan: workspace %lld B < required %lld B", (long long)workspace_bytes, (long long)L.total);
        return GS_ERR_WORKSPACE_TOO_SMALL;
    }
    char* w = (char*)workspace;
    Tri* flags = (Tri*)(w + L.flags);
    Tri* pos = (Tri*)(w + L.pos);
    densify_classify_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, scaling_log, opacity, grad, grad_threshold, small_sigma,
                                                                        large_sigma, min_opacity, flags);
    GS_CUDA_TRY(cudaGetLastError());
    size_t cub_bytes = (size_t)L.cub_bytes;
    GS_CUDA_TRY(cub::DeviceScan::ExclusiveScan(w + L.cub_temp, cub_bytes, (const Tri*)flags, pos, TriSum(), Tri{0, 0, 0, 0}, n, st));
    densify_totals_kernel<<<1, 32, 0, st>>>(n, flags, pos, counts);
    GS_CUDA_TRY(cudaGetLastError());
    count_launches(2);
    return GS_OK;
}

extern "C" int gs_densify_apply(int64_t n, const void* workspace, int64_t kept, int64_t cloned, int64_t split,
                                const float* xyz, const float* features_dc, const float* features_rest, const float* scaling_log,
                                const float* rotation, const float* opacity, const float* noise,
                                float* o_xyz, float* o_features_dc, float* o_features_rest, float* o_scaling_log,
                                float* o_rotation, float* o_opacity, int32_t* src_row, void* stream) {
    GS_REQUIRE(n >= 0 && kept >= 0 && cloned >= 0 && split >= 0, "negative size");
    if (n == 0 || kept + cloned + split == 0) return GS_OK;
    GS_REQUIRE(workspace && xyz && features_dc && features_rest && scaling_log && rotation && opacity, "NULL input array");
    GS_REQUIRE(o_xyz && o_features_dc && o_features_rest && o_scaling_log && o_rotation && o_opacity, "NULL output array");
    GS_REQUIRE(cloned == 0 || noise != nullptr, "clone copies need the noise array");
    DeviceGuard guard(xyz);
    const DensifyLayout L = densify_layout(n);
    const char* w = (const char*)workspace;
    ModelPtrs m = {xyz, features_dc, features_rest, scaling_log, rotation, opacity,
                   o_xyz, o_features_dc, o_features_rest, o_scaling_log, o_rotation, o_opacity};
    densify_apply_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        n, (const Tri*)(w + L.flags), (const Tri*)(w + L.pos), kept, cloned, split, m, noise, src_row);
    GS_CUDA_TRY(cudaGetLastError());
    count_launches(1);
    return GS_OK;
}
