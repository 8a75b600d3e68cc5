// Shared device/host helpers for libgsplat_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/gsplat_b200.h"

namespace gs {

// ---- error plumbing: thread-local message, int status across the C ABI -------------------
void set_error(const char* fmt, ...);
// statistics only (relaxed atomic): hand-written kernels launched by this process
void count_launches(int n);
int cuda_fail(cudaError_t e, const char* what);

#define GS_CUDA_TRY(expr)                                   \
    do {                                                    \
        cudaError_t _e = (expr);                            \
        if (_e != cudaSuccess) return gs::cuda_fail(_e, #expr); \
    } while (0)

#define GS_REQUIRE(cond, msg)                               \
    do {                                                    \
        if (!(cond)) {                                      \
            gs::set_error("%s: %s", __func__, msg);         \
            return GS_ERR_INVALID_ARGUMENT;                 \
        }                                                   \
    } while (0)

// Pins the current device to the one that owns `ptr` for the lifetime of the guard, so a call
// made from the autograd engine's thread lands on the right GPU.
struct DeviceGuard {
    int prev = -1;
    bool changed = false;
    explicit DeviceGuard(const void* ptr) {
        cudaPointerAttributes attr;
        if (ptr && cudaPointerGetAttributes(&attr, ptr) == cudaSuccess &&
            attr.type == cudaMemoryTypeDevice) {
            cudaGetDevice(&prev);
            if (prev != attr.device) {
                cudaSetDevice(attr.device);
                changed = true;
            }
        } else {
            cudaGetLastError();  // clear
        }
    }
    ~DeviceGuard() {
        if (changed) cudaSetDevice(prev);
    }
};

// ---- programmatic dependent launch (PDL) ----------------------------------------------------------------------
// A frame is a chain of ~12 dependent kernels; between two of them the GPU drains, the next grid is set up and its
// first CTAs are dispatched (2-4 us per boundary in the steady-state timeline, profiles/r2_timeline_steady_state.txt).
// Launched with programmatic stream serialization, the next kernel's CTAs are dispatched while the previous kernel's
// last CTAs retire; every such kernel begins with grid_dependency_wait(), which returns once the previous kernel has
// completed and its memory is visible, so nothing is read early.  GSPLAT_B200_PDL=0 launches the plain way.
__device__ __forceinline__ void grid_dependency_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---- camera block passed by value to kernels ------------------------------------------------
struct Camera {
    float r[9];
    float t[3];
    float fx, fy, cx, cy;
};

inline Camera camera_from_host(const float* c) {
    Camera cam;
    for (int i = 0; i < 9; ++i) cam.r[i] = c[i];
    for (int i = 0; i < 3; ++i) cam.t[i] = c[9 + i];
    cam.fx = c[12];
    cam.fy = c[13];
    cam.cx = c[14];
    cam.cy = c[15];
    return cam;
}

constexpr int kTile = 16;            // the raster kernels are specialised for 16x16 tiles
constexpr int kRecFloats = GS_SPLAT_REC_FLOATS;

// Arithmetic that must round exactly like the reference's separate torch ops: the compiler is
// not allowed to contract these into FMAs.
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float div_rn(float a, float b) { return __fdiv_rn(a, b); }

}  // namespace gs
