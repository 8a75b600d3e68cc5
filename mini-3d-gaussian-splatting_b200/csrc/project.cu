// Stage P+M+C: fused activation + 3-D covariance + EWA projection + conic + radius + culling
// + tile rectangle, forward and backward.  One thread per Gaussian; the work is a pure stream
// over the parameter SoA (HBM-bound, no reuse), so the kernels are sized for coalescing and
// many loads in flight rather than for shared-memory staging.
//
// Reference semantics: src/core/renderer.py:117-220, src/core/gaussian_model.py:200-207,
// src/utils/math_utils.py:9-26.
#include "common.cuh"

namespace gs {

struct Splat3D {
    float S[9];   // 3-D covariance, row-major (general 3x3: the covariance-mode input may be any matrix)
    // parameter mode only:
    float R[9];   // rotation matrix
    float sig[3]; // exp(scaling_log)
    float qh[4];  // normalised quaternion (w,x,y,z)
    float qn;     // max(|rotation|, 1e-12)
};

struct Proj {
    float X, Y, Z, iz;
    float mx, my;
    float J00, J02, J11, J12;
    float M[9];     // Rv S Rv^T
    float a, b01, b10, c;   // cov2d (+1e-6 on the diagonal)
    float det;
    float q00, q01, q10, q11;
    float radius;
};

__device__ __forceinline__ void mat3_mul(const float* A, const float* B, float* C) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            C[i * 3 + j] = A[i * 3 + 0] * B[0 * 3 + j] + A[i * 3 + 1] * B[1 * 3 + j] + A[i * 3 + 2] * B[2 * 3 + j];
}
__device__ __forceinline__ void mat3_mul_bt(const float* A, const float* B, float* C) {  // A * B^T
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            C[i * 3 + j] = A[i * 3 + 0] * B[j * 3 + 0] + A[i * 3 + 1] * B[j * 3 + 1] + A[i * 3 + 2] * B[j * 3 + 2];
}
__device__ __forceinline__ void mat3_mul_at(const float* A, const float* B, float* C) {  // A^T * B
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            C[i * 3 + j] = A[0 * 3 + i] * B[0 * 3 + j] + A[1 * 3 + i] * B[1 * 3 + j] + A[2 * 3 + i] * B[2 * 3 + j];
}

// gaussian_model.py:113-118,200-207 + math_utils.py:9-26
__device__ __forceinline__ void covariance_from_params(const float* __restrict__ scaling_log,
                                                       const float* __restrict__ rotation,
                                                       int64_t i, Splat3D& g) {
    const float s0 = scaling_log[i * 3 + 0], s1 = scaling_log[i * 3 + 1], s2 = scaling_log[i * 3 + 2];
    const float4 q = *reinterpret_cast<const float4*>(rotation + i * 4);
    g.sig[0] = expf(s0);
    g.sig[1] = expf(s1);
    g.sig[2] = expf(s2);
    // F.normalize: q / max(|q|, 1e-12); the reference applies it twice (get_rotation, then
    // build_rotation_matrix); the second pass only matters at rounding level but is kept.
    float n = sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
    n = fmaxf(n, 1e-12f);
    g.qn = n;
    float w = q.x / n, x = q.y / n, y = q.z / n, z = q.w / n;
    float n2 = fmaxf(sqrtf(w * w + x * x + y * y + z * z), 1e-12f);
    w /= n2; x /= n2; y /= n2; z /= n2;
    g.qh[0] = w; g.qh[1] = x; g.qh[2] = y; g.qh[3] = z;
    const float xx = x * x, yy = y * y, zz = z * z;
    const float wx = w * x, wy = w * y, wz = w * z;
    const float xy = x * y, xz = x * z, yz = y * z;
    g.R[0] = 1.f - 2.f * (yy + zz); g.R[1] = 2.f * (xy - wz);       g.R[2] = 2.f * (xz + wy);
    g.R[3] = 2.f * (xy + wz);       g.R[4] = 1.f - 2.f * (xx + zz); g.R[5] = 2.f * (yz - wx);
    g.R[6] = 2.f * (xz - wy);       g.R[7] = 2.f * (yz + wx);       g.R[8] = 1.f - 2.f * (xx + yy);
    const float d0 = g.sig[0] * g.sig[0], d1 = g.sig[1] * g.sig[1], d2 = g.sig[2] * g.sig[2];
    // Sigma = R diag(d) R^T (symmetric by construction)
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = r; c < 3; ++c) {
            const float v = g.R[r * 3 + 0] * d0 * g.R[c * 3 + 0] + g.R[r * 3 + 1] * d1 * g.R[c * 3 + 1] +
                            g.R[r * 3 + 2] * d2 * g.R[c * 3 + 2];
            g.S[r * 3 + c] = v;
            g.S[c * 3 + r] = v;
        }
}

// renderer.py:154-192.  The sub-expressions that feed integer casts and the culling compares
// (camera-space position, pixel centre) reproduce the reference's rounding exactly:
// MKL evaluates `Xw @ Rv.T` as x0*r0 followed by two FMAs, `+ Tv` is a separate add, and
// `fx * X / Z + cx` is three separately rounded ops.
__device__ __forceinline__ void project_point(const Camera& cam, float x0, float x1, float x2,
                                              const float* S, float rmin, float rmax, Proj& p) {
    p.X = add_rn(__fmaf_rn(x2, cam.r[2], __fmaf_rn(x1, cam.r[1], mul_rn(x0, cam.r[0]))), cam.t[0]);
    p.Y = add_rn(__fmaf_rn(x2, cam.r[5], __fmaf_rn(x1, cam.r[4], mul_rn(x0, cam.r[3]))), cam.t[1]);
    p.Z = add_rn(__fmaf_rn(x2, cam.r[8], __fmaf_rn(x1, cam.r[7], mul_rn(x0, cam.r[6]))), cam.t[2]);
    p.mx = add_rn(div_rn(mul_rn(cam.fx, p.X), p.Z), cam.cx);
    p.my = add_rn(div_rn(mul_rn(-cam.fy, p.Y), p.Z), cam.cy);

    float tmp[9];
    mat3_mul(cam.r, S, tmp);          // Rv S
    mat3_mul_bt(tmp, cam.r, p.M);     // (Rv S) Rv^T
    p.iz = div_rn(1.0f, p.Z);
    p.J00 = cam.fx * p.iz;
    p.J02 = -cam.fx * p.X * p.iz * p.iz;
    p.J11 = -cam.fy * p.iz;
    p.J12 = cam.fy * p.Y * p.iz * p.iz;
    // cov2d = J M J^T + 1e-6 I with J = [[J00,0,J02],[0,J11,J12]]
    const float* M = p.M;
    const float u0 = p.J00 * M[0] + p.J02 * M[6];   // (J M) row 0
    const float u1 = p.J00 * M[1] + p.J02 * M[7];
    const float u2 = p.J00 * M[2] + p.J02 * M[8];
    const float v0 = p.J11 * M[3] + p.J12 * M[6];   // row 1
    const float v1 = p.J11 * M[4] + p.J12 * M[7];
    const float v2 = p.J11 * M[5] + p.J12 * M[8];
    p.a = (u0 * p.J00 + u2 * p.J02) + 1e-6f;
    p.b01 = u1 * p.J11 + u2 * p.J12;
    p.b10 = v0 * p.J00 + v2 * p.J02;
    p.c = (v1 * p.J11 + v2 * p.J12) + 1e-6f;
    p.det = p.a * p.c - p.b01 * p.b10;
    const float id = 1.0f / p.det;
    p.q00 = p.c * id;
    p.q01 = -p.b01 * id;
    p.q10 = -p.b10 * id;
    p.q11 = p.a * id;
    // lambda_max of the symmetric matrix eigvalsh sees (lower triangle): SURVEY 8c closed form
    const float mid = 0.5f * (p.a + p.c);
    const float hd = 0.5f * (p.a - p.c);
    const float lam = mid + sqrtf(hd * hd + p.b10 * p.b10);
    const float r = 3.0f * sqrtf(lam);
    // torch.clamp propagates NaN; fminf/fmaxf would swallow it
    p.radius = (r != r) ? r : fminf(fmaxf(r, rmin), rmax);
}

__device__ __forceinline__ float sigmoidf(float x) { return 1.0f / (1.0f + expf(-x)); }

// ---- optional view-dependent colour: real spherical harmonics up to degree 3 -----------------------
// The reference renders sigmoid(features[:,0,:]) only (renderer.py:88-92; its SH evaluator is a stub,
// math_utils.py:44-49).  With sh_degree > 0 the colour becomes
//     sigmoid( features[:,0,:] + sum_{k=1}^{(deg+1)^2-1} Y_k(dir) * features[:,k,:] ),   dir = normalize(xyz - camera centre),
// with the usual real-SH basis; it equals the reference bit for bit whenever the higher-order rows are
// zero (the reference's initialisation, gaussian_model.py:84).  Default is sh_degree = 0.
struct ShParams {
    const float* rest;        // features row 1 of splat 0; row k (>=1), channel c of splat i at rest[i*stride + (k-1)*3 + c]
    long long stride;
    float* g_rest;            // backward only
    long long g_stride;
    int degree;               // 0..3
    int staged;               // rest is [n,15,3] contiguous: stage through shared memory with float4 loads
    float cx, cy, cz;         // camera centre in world space
};
constexpr int kShRest = 15;
constexpr int kShRowFloats = kShRest * 3;

__device__ __forceinline__ int sh_terms(int degree) { return (degree + 1) * (degree + 1) - 1; }

// Y[0..14] = basis for k = 1..15 at unit direction (x,y,z)
__device__ __forceinline__ void sh_basis(int degree, float x, float y, float z, float* Y) {
    const float C1 = 0.4886025119029199f;
    Y[0] = -C1 * y; Y[1] = C1 * z; Y[2] = -C1 * x;
    if (degree < 2) return;
    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
    Y[3] = 1.0925484305920792f * xy;
    Y[4] = -1.0925484305920792f * yz;
    Y[5] = 0.31539156525252005f * (2.f * zz - xx - yy);
    Y[6] = -1.0925484305920792f * xz;
    Y[7] = 0.5462742152960396f * (xx - yy);
    if (degree < 3) return;
    Y[8] = -0.5900435899266435f * y * (3.f * xx - yy);
    Y[9] = 2.890611442640554f * xy * z;
    Y[10] = -0.4570457994644658f * y * (4.f * zz - xx - yy);
    Y[11] = 0.3731763325901154f * z * (2.f * zz - 3.f * xx - 3.f * yy);
    Y[12] = -0.4570457994644658f * x * (4.f * zz - xx - yy);
    Y[13] = 1.445305721320277f * z * (xx - yy);
    Y[14] = -0.5900435899266435f * x * (xx - 3.f * yy);
}

// g_dir = sum_k s[k] * grad Y_k
__device__ __forceinline__ void sh_basis_grad(int degree, float x, float y, float z, const float* s, float& gx, float& gy, float& gz) {
    const float C1 = 0.4886025119029199f;
    gx = -C1 * s[2]; gy = -C1 * s[0]; gz = C1 * s[1];
    if (degree < 2) return;
    const float a = 1.0925484305920792f, b = 0.31539156525252005f, c = 0.5462742152960396f;
    gx += a * y * s[3] - 2.f * b * x * s[5] - a * z * s[6] + 2.f * c * x * s[7];
    gy += a * x * s[3] - a * z * s[4] - 2.f * b * y * s[5] - 2.f * c * y * s[7];
    gz += -a * y * s[4] + 4.f * b * z * s[5] - a * x * s[6];
    if (degree < 3) return;
    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
    const float d0 = -0.5900435899266435f, d1 = 2.890611442640554f, d2 = -0.4570457994644658f,
                d3 = 0.3731763325901154f, d5 = 1.445305721320277f;
    gx += d0 * 6.f * xy * s[8] + d1 * yz * s[9] + d2 * (-2.f * xy) * s[10] + d3 * (-6.f * xz) * s[11] +
          d2 * (4.f * zz - 3.f * xx - yy) * s[12] + d5 * 2.f * xz * s[13] + d0 * (3.f * xx - 3.f * yy) * s[14];
    gy += d0 * (3.f * xx - 3.f * yy) * s[8] + d1 * xz * s[9] + d2 * (4.f * zz - xx - 3.f * yy) * s[10] + d3 * (-6.f * yz) * s[11] +
          d2 * (-2.f * xy) * s[12] + d5 * (-2.f * yz) * s[13] + d0 * (-6.f * xy) * s[14];
    gz += d1 * xy * s[9] + d2 * 8.f * yz * s[10] + d3 * (6.f * zz - 3.f * xx - 3.f * yy) * s[11] + d2 * 8.f * xz * s[12] +
          d5 * (xx - yy) * s[13];
}

// Block-cooperative copy of this block's [256 x 45] rows between global and shared memory with
// 16-byte accesses (the rows are 180 B: per-thread row accesses would be 4-byte and uncoalesced).
__device__ __forceinline__ void sh_stage_load(const ShParams& sh, int64_t n, float* s_rows) {
    const int64_t first = (int64_t)blockIdx.x * blockDim.x;
    const int64_t rows = min((int64_t)blockDim.x, n - first);
    const int64_t floats = rows * kShRowFloats;
    const float* src = sh.rest + first * kShRowFloats;            // 256*45*4 B = 46080 B: 16-byte aligned per block
    const int64_t vec = floats / 4;
    for (int64_t v = threadIdx.x; v < vec; v += blockDim.x)
        reinterpret_cast<float4*>(s_rows)[v] = __ldg(reinterpret_cast<const float4*>(src) + v);
    for (int64_t f = vec * 4 + threadIdx.x; f < floats; f += blockDim.x) s_rows[f] = src[f];
    __syncthreads();
}
__device__ __forceinline__ void sh_stage_store(const ShParams& sh, int64_t n, const float* s_rows, bool accumulate) {
    __syncthreads();
    const int64_t first = (int64_t)blockIdx.x * blockDim.x;
    const int64_t rows = min((int64_t)blockDim.x, n - first);
    const int64_t floats = rows * kShRowFloats;
    float* dst = sh.g_rest + first * kShRowFloats;
    const int64_t vec = floats / 4;
    for (int64_t v = threadIdx.x; v < vec; v += blockDim.x) {
        float4 val = reinterpret_cast<const float4*>(s_rows)[v];
        if (accumulate) {
            const float4 old = reinterpret_cast<const float4*>(dst)[v];
            val.x += old.x; val.y += old.y; val.z += old.z; val.w += old.w;
        }
        reinterpret_cast<float4*>(dst)[v] = val;
    }
    for (int64_t f = vec * 4 + threadIdx.x; f < floats; f += blockDim.x) dst[f] = accumulate ? dst[f] + s_rows[f] : s_rows[f];
}

// Densification statistics fused into the projection backward (the reference allocates these
// buffers but never fills them: gaussian_model.py:29-31).  For every visible splat:
//   grad_norm[i] += |dL/d means2D[i]|,  count[i] += 1,  max_radii[i] = max(max_radii[i], radii[i]).
struct DensifyStats {
    const float* radii;
    const uint8_t* vis;
    float* grad_norm;
    float* count;
    float* max_radii;
};

// results are written, or added to what the buffer holds (multi-view gradient accumulation)
__device__ __forceinline__ void emit(float* p, float v, bool accumulate) { *p = accumulate ? *p + v : v; }

template <bool kParamMode, bool kSh>
__global__ void __launch_bounds__(256)
project_fwd_kernel(int64_t n, const float* __restrict__ xyz, const float* __restrict__ scaling_log,
                   const float* __restrict__ rotation, const float* __restrict__ cov3d,
                   const float* __restrict__ opacity, int opacity_is_logit,
                   const float* __restrict__ feat0, int64_t feat_stride, ShParams sh, Camera cam,
                   int img_w, int img_h, int tile_size, float rmin, float rmax,
                   float2* __restrict__ means2d, float* __restrict__ depths, float4* __restrict__ conics,
                   float* __restrict__ radii, float* __restrict__ colors, float* __restrict__ opac_out,
                   uint8_t* __restrict__ vis_out, int32_t* __restrict__ tiles_touched,
                   ushort4* __restrict__ tile_rect, uint32_t* __restrict__ depth_keys,
                   float4* __restrict__ rec) {
    extern __shared__ __align__(16) float s_sh_rows[];
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (kSh && sh.staged) sh_stage_load(sh, n, s_sh_rows);
    if (i >= n) return;

    Splat3D g;
    if (kParamMode) {
        covariance_from_params(scaling_log, rotation, i, g);
    } else {
#pragma unroll
        for (int k = 0; k < 9; ++k) g.S[k] = cov3d[i * 9 + k];
    }
    const float x0 = xyz[i * 3 + 0], x1 = xyz[i * 3 + 1], x2 = xyz[i * 3 + 2];
    Proj p;
    project_point(cam, x0, x1, x2, g.S, rmin, rmax, p);

    const float op_in = opacity[i];
    const float op = opacity_is_logit ? sigmoidf(op_in) : op_in;
    float pre_r = feat0[i * feat_stride + 0], pre_g = feat0[i * feat_stride + 1], pre_b = feat0[i * feat_stride + 2];
    if (kSh) {
        const float* row = sh.staged ? s_sh_rows + threadIdx.x * kShRowFloats : sh.rest + i * sh.stride;
        float dx = x0 - sh.cx, dy = x1 - sh.cy, dz = x2 - sh.cz;
        const float inv = 1.0f / fmaxf(sqrtf(dx * dx + dy * dy + dz * dz), 1e-12f);
        dx *= inv; dy *= inv; dz *= inv;
        float Y[kShRest];
        sh_basis(sh.degree, dx, dy, dz, Y);
        const int terms = sh_terms(sh.degree);
        float ar = 0.f, ag = 0.f, ab = 0.f;
        for (int k = 0; k < terms; ++k) {
            ar = fmaf(Y[k], row[k * 3 + 0], ar);
            ag = fmaf(Y[k], row[k * 3 + 1], ag);
            ab = fmaf(Y[k], row[k * 3 + 2], ab);
        }
        pre_r += ar; pre_g += ag; pre_b += ab;          // + 0 exactly when the rows are zero
    }
    const float cr = sigmoidf(pre_r);
    const float cg = sigmoidf(pre_g);
    const float cb = sigmoidf(pre_b);

    // renderer.py:218 -- every compare is false for NaN, exactly as in torch
    const float r = p.radius;
    const float Wf = (float)img_w, Hf = (float)img_h;
    const bool vis = (p.Z > 0.f) && (p.mx >= -r) && (p.mx < add_rn(Wf, r)) && (p.my >= -r) &&
                     (p.my < add_rn(Hf, r)) && (r > 0.f);

    // renderer.py:278-293 (Python int() == C truncation toward zero)
    int cnt = 0;
    ushort4 rect = make_ushort4(0, 0, 0, 0);
    if (vis) {
        const int ir = (int)r, ix = (int)p.mx, iy = (int)p.my;
        const int px0 = max(ix - ir, 0), px1 = min(ix + 1 + ir, img_w);
        const int py0 = max(iy - ir, 0), py1 = min(iy + 1 + ir, img_h);
        if (px0 < px1 && py0 < py1) {
            const int tx0 = px0 / tile_size, tx1 = (px1 - 1) / tile_size;      // any tile size (renderer.py:24, :286-289)
            const int ty0 = py0 / tile_size, ty1 = (py1 - 1) / tile_size;
            cnt = (tx1 - tx0 + 1) * (ty1 - ty0 + 1);
            rect = make_ushort4((unsigned short)tx0, (unsigned short)ty0, (unsigned short)tx1, (unsigned short)ty1);
        }
    }

    means2d[i] = make_float2(p.mx, p.my);
    depths[i] = p.Z;
    conics[i] = make_float4(p.q00, p.q01, p.q10, p.q11);
    radii[i] = r;
    colors[i * 3 + 0] = cr;
    colors[i * 3 + 1] = cg;
    colors[i * 3 + 2] = cb;
    opac_out[i] = op;
    vis_out[i] = vis ? 1 : 0;
    tiles_touched[i] = cnt;
    tile_rect[i] = rect;
    // valid depths are < 0x7F800001, so the two sentinels sort behind every binned splat
    depth_keys[i] = cnt > 0 ? __float_as_uint(p.Z) : (vis ? 0xFFFFFFFEu : 0xFFFFFFFFu);
    // raster record: conic pre-scaled by c = -0.5*log2(e) so that exp(-0.5*s) = exp2(quadratic form)
    const float kC = -0.72134752044448170f;
    const float s00 = kC * p.q00, s01 = kC * add_rn(p.q01, p.q10), s11 = kC * p.q11;
    rec[i * 3 + 0] = make_float4(p.mx, p.my, s00, s01);
    // depth kept finite in the record: masked-out pixels multiply it by 0 (0*inf would be NaN)
    rec[i * 3 + 1] = make_float4(s11, op, fminf(p.Z, 3.0e38f), cr);
    // "regular" flag for the compositing kernels' fast path: opacity in [0,1] and a positive-definite
    // conic whose eigenvalue ratio is below 1e5 (det >= 1e-5 * trace^2), so that the quadratic form is
    // <= 0 at every pixel up to its own rounding and both clamps of renderer.py:335,339 are identities.
    // Every compare is false for NaN/inf inputs, which therefore take the general path.
    const float det_s = s00 * s11 - 0.25f * s01 * s01, tr_s = s00 + s11;
    const bool regular = (op >= 0.f) && (op <= 1.f) && (s00 < 0.f) && (s11 < 0.f) && (det_s >= 1e-5f * tr_s * tr_s) &&
                         (tr_s > -1e30f);
    // log2(opacity): the fast path folds the opacity into the exponent, a = 2^(c*s + log2 opacity); -inf for the opacities the
    // compositing kernels stage as zero (<= 1e-30, renderer.py:340 skips a <= 0)
    const float lop = (op > 1e-30f) ? log2f(op) : __int_as_float(0xff800000);
    rec[i * 3 + 2] = make_float4(cg, cb, regular ? 1.f : 0.f, lop);
}

// Backward (derived from the forward above).
template <bool kParamMode, bool kSh>
// (a register cap for more resident warps was measured and rejected: 57 -> 60 us at 48 registers, 68 us at 40)
__global__ void __launch_bounds__(256)
project_bwd_kernel(int64_t n, const float* __restrict__ xyz, const float* __restrict__ scaling_log,
                   const float* __restrict__ rotation, const float* __restrict__ cov3d,
                   const float* __restrict__ opacity, int opacity_is_logit,
                   const float* __restrict__ feat0, int64_t feat_stride, ShParams sh, Camera cam,
                   const float2* __restrict__ g_means2d, const float4* __restrict__ g_conics,
                   const float* __restrict__ g_depths, const float* __restrict__ g_colors,
                   const float* __restrict__ g_opac,
                   float* __restrict__ g_xyz, float* __restrict__ g_scaling, float4* __restrict__ g_rotation,
                   float* __restrict__ g_cov3d, float* __restrict__ g_opacity,
                   float* __restrict__ g_feat0, int64_t g_feat_stride, int accumulate_i, DensifyStats stats) {
    extern __shared__ __align__(16) float s_sh_rows[];
    grid_dependency_wait();                          // launched early (PDL) behind the compositing backward
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool staged = kSh && sh.staged;
    const bool acc = accumulate_i != 0;
    if (staged) sh_stage_load(sh, n, s_sh_rows);
    float gdx = 0.f, gdy = 0.f, gdz = 0.f;          // colour -> view direction -> position (SH only)
    if (i < n) {
        // activations and colour
        const float gop = g_opac[i];
        const float op_in = opacity[i];
        float go = gop;
        if (opacity_is_logit) {
            const float o = sigmoidf(op_in);
            go = gop * o * (1.f - o);
        }
        emit(g_opacity + i, go, acc);
        float pre[3] = {feat0[i * feat_stride + 0], feat0[i * feat_stride + 1], feat0[i * feat_stride + 2]};
        float Y[kShRest];
        float dirx = 0.f, diry = 0.f, dirz = 0.f, inv_len = 0.f;
        const int terms = sh_terms(sh.degree);
        float* row = nullptr;
        if (kSh) {
            row = staged ? s_sh_rows + threadIdx.x * kShRowFloats : const_cast<float*>(sh.rest + i * sh.stride);
            dirx = xyz[i * 3 + 0] - sh.cx; diry = xyz[i * 3 + 1] - sh.cy; dirz = xyz[i * 3 + 2] - sh.cz;
            inv_len = 1.0f / fmaxf(sqrtf(dirx * dirx + diry * diry + dirz * dirz), 1e-12f);
            dirx *= inv_len; diry *= inv_len; dirz *= inv_len;
            sh_basis(sh.degree, dirx, diry, dirz, Y);
            for (int k = 0; k < terms; ++k) {
                pre[0] = fmaf(Y[k], row[k * 3 + 0], pre[0]);
                pre[1] = fmaf(Y[k], row[k * 3 + 1], pre[1]);
                pre[2] = fmaf(Y[k], row[k * 3 + 2], pre[2]);
            }
        }
        float gpre[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float col = sigmoidf(pre[c]);
            gpre[c] = g_colors[i * 3 + c] * col * (1.f - col);
            emit(g_feat0 + i * g_feat_stride + c, gpre[c], acc);
        }
        if (kSh) {
            float sk[kShRest];
            float* grow = staged ? row : sh.g_rest + i * sh.g_stride;      // staged: overwrite the row in place
            for (int k = 0; k < kShRest; ++k) {
                const bool used = k < terms;
                sk[k] = used ? gpre[0] * row[k * 3 + 0] + gpre[1] * row[k * 3 + 1] + gpre[2] * row[k * 3 + 2] : 0.f;
                // the staged row is private scratch (added to global memory by sh_stage_store)
                emit(grow + k * 3 + 0, used ? gpre[0] * Y[k] : 0.f, acc && !staged);
                emit(grow + k * 3 + 1, used ? gpre[1] * Y[k] : 0.f, acc && !staged);
                emit(grow + k * 3 + 2, used ? gpre[2] * Y[k] : 0.f, acc && !staged);
            }
            float gx, gy, gz;
            sh_basis_grad(sh.degree, dirx, diry, dirz, sk, gx, gy, gz);
            const float dot = gx * dirx + gy * diry + gz * dirz;            // through the normalisation
            gdx = (gx - dirx * dot) * inv_len;
            gdy = (gy - diry * dot) * inv_len;
            gdz = (gz - dirz * dot) * inv_len;
        }
    }
    if (staged) sh_stage_store(sh, n, s_sh_rows, acc);
    if (i >= n) return;

    const float2 gm = g_means2d[i];
    if (stats.grad_norm != nullptr) {
        const bool v = stats.vis[i] != 0;
        const float gn = v ? sqrtf(gm.x * gm.x + gm.y * gm.y) : 0.f, c = v ? 1.0f : 0.f, r = v ? stats.radii[i] : 0.f;
        if (acc) {
            if (v) {
                stats.grad_norm[i] += gn;
                stats.count[i] += c;
                stats.max_radii[i] = fmaxf(stats.max_radii[i], r);
            }
        } else {                                    // first view of a step: initialises the statistics as well
            stats.grad_norm[i] = gn;
            stats.count[i] = c;
            stats.max_radii[i] = r;
        }
    }
    const float4 gq = g_conics[i];
    const float gz_in = g_depths[i];

    const bool any_geo = (gm.x != 0.f) || (gm.y != 0.f) || (gq.x != 0.f) || (gq.y != 0.f) || (gq.z != 0.f) ||
                         (gq.w != 0.f) || (gz_in != 0.f);
    if (!any_geo) {
        // nothing reached this splat's geometry through the rasteriser
        if (acc) {
            if (kSh) { g_xyz[i * 3 + 0] += gdx; g_xyz[i * 3 + 1] += gdy; g_xyz[i * 3 + 2] += gdz; }
            return;
        }
        g_xyz[i * 3 + 0] = gdx; g_xyz[i * 3 + 1] = gdy; g_xyz[i * 3 + 2] = gdz;
        if (kParamMode) {
            g_scaling[i * 3 + 0] = 0.f; g_scaling[i * 3 + 1] = 0.f; g_scaling[i * 3 + 2] = 0.f;
            g_rotation[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
#pragma unroll
            for (int k = 0; k < 9; ++k) g_cov3d[i * 9 + k] = 0.f;
        }
        return;
    }

    Splat3D g;
    if (kParamMode) {
        covariance_from_params(scaling_log, rotation, i, g);
    } else {
#pragma unroll
        for (int k = 0; k < 9; ++k) g.S[k] = cov3d[i * 9 + k];
    }
    const float x0 = xyz[i * 3 + 0], x1 = xyz[i * 3 + 1], x2 = xyz[i * 3 + 2];
    Proj p;
    project_point(cam, x0, x1, x2, g.S, 0.f, 1e30f, p);

    // conic = inv(cov2d):  g_cov2d = -Q^T gQ Q^T
    float G2[4];
    {
        const float Q[4] = {p.q00, p.q01, p.q10, p.q11};
        const float gQ[4] = {gq.x, gq.y, gq.z, gq.w};
        // T = Q^T gQ
        const float t00 = Q[0] * gQ[0] + Q[2] * gQ[2];
        const float t01 = Q[0] * gQ[1] + Q[2] * gQ[3];
        const float t10 = Q[1] * gQ[0] + Q[3] * gQ[2];
        const float t11 = Q[1] * gQ[1] + Q[3] * gQ[3];
        // G2 = -(T Q^T)
        G2[0] = -(t00 * Q[0] + t01 * Q[1]);
        G2[1] = -(t00 * Q[2] + t01 * Q[3]);
        G2[2] = -(t10 * Q[0] + t11 * Q[1]);
        G2[3] = -(t10 * Q[2] + t11 * Q[3]);
    }
    const float J[6] = {p.J00, 0.f, p.J02, 0.f, p.J11, p.J12};
    // gM = J^T G2 J   (3x3)
    float gM[9];
    {
        float GJ[6];  // G2 J (2x3)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            GJ[c] = G2[0] * J[c] + G2[1] * J[3 + c];
            GJ[3 + c] = G2[2] * J[c] + G2[3] * J[3 + c];
        }
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) gM[r * 3 + c] = J[r] * GJ[c] + J[3 + r] * GJ[3 + c];
    }
    // gJ = G2 J M^T + G2^T J M   (2x3)
    float gJ[6];
    {
        float JMt[6], JM[6];
        const float* M = p.M;
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                JMt[r * 3 + c] = J[r * 3 + 0] * M[c * 3 + 0] + J[r * 3 + 1] * M[c * 3 + 1] + J[r * 3 + 2] * M[c * 3 + 2];
                JM[r * 3 + c] = J[r * 3 + 0] * M[0 * 3 + c] + J[r * 3 + 1] * M[1 * 3 + c] + J[r * 3 + 2] * M[2 * 3 + c];
            }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            gJ[c] = G2[0] * JMt[c] + G2[1] * JMt[3 + c] + G2[0] * JM[c] + G2[2] * JM[3 + c];
            gJ[3 + c] = G2[2] * JMt[c] + G2[3] * JMt[3 + c] + G2[1] * JM[c] + G2[3] * JM[3 + c];
        }
    }
    const float iz = p.iz, iz2 = iz * iz, iz3 = iz2 * iz;
    const float fx = cam.fx, fy = cam.fy;
    const float gX = gm.x * fx * iz + gJ[2] * (-fx * iz2);
    const float gY = gm.y * (-fy * iz) + gJ[5] * (fy * iz2);
    const float gZ = gz_in + gm.x * (-fx * p.X * iz2) + gm.y * (fy * p.Y * iz2) + gJ[0] * (-fx * iz2) +
                     gJ[2] * (2.f * fx * p.X * iz3) + gJ[4] * (fy * iz2) + gJ[5] * (-2.f * fy * p.Y * iz3);
    emit(g_xyz + i * 3 + 0, cam.r[0] * gX + cam.r[3] * gY + cam.r[6] * gZ + gdx, acc);
    emit(g_xyz + i * 3 + 1, cam.r[1] * gX + cam.r[4] * gY + cam.r[7] * gZ + gdy, acc);
    emit(g_xyz + i * 3 + 2, cam.r[2] * gX + cam.r[5] * gY + cam.r[8] * gZ + gdz, acc);

    // gSigma = Rv^T gM Rv
    float gS[9];
    {
        float tmp[9];
        mat3_mul_at(cam.r, gM, tmp);
        mat3_mul(tmp, cam.r, gS);
    }
    if (!kParamMode) {
#pragma unroll
        for (int k = 0; k < 9; ++k) emit(g_cov3d + i * 9 + k, gS[k], acc);
        return;
    }
    // Sigma = R D R^T, D = diag(sig^2)
    const float* R = g.R;
    float gsym[9];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) gsym[r * 3 + c] = gS[r * 3 + c] + gS[c * 3 + r];
    float gR[9];
    {
        float tmp[9];
        mat3_mul(gsym, R, tmp);   // (gS + gS^T) R
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) gR[r * 3 + c] = tmp[r * 3 + c] * g.sig[c] * g.sig[c];
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        // (R^T gS R)_kk
        float racc = 0.f;
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) racc += R[r * 3 + k] * gS[r * 3 + c] * R[c * 3 + k];
        // d(sig^2)/d(log-scale) = 2 sig^2
        emit(g_scaling + i * 3 + k, 2.f * racc * g.sig[k] * g.sig[k], acc);
    }
    const float w = g.qh[0], x = g.qh[1], y = g.qh[2], z = g.qh[3];
    const float gw = 2.f * (-z * gR[1] + y * gR[2] + z * gR[3] - x * gR[5] - y * gR[6] + x * gR[7]);
    const float gx = 2.f * (y * gR[1] + z * gR[2] + y * gR[3] - 2.f * x * gR[4] - w * gR[5] + z * gR[6] + w * gR[7] - 2.f * x * gR[8]);
    const float gy = 2.f * (-2.f * y * gR[0] + x * gR[1] + w * gR[2] + x * gR[3] + z * gR[5] - w * gR[6] + z * gR[7] - 2.f * y * gR[8]);
    const float gzq = 2.f * (-2.f * z * gR[0] - w * gR[1] + x * gR[2] + w * gR[3] - 2.f * z * gR[4] + y * gR[5] + x * gR[6] + y * gR[7]);
    const float dotp = w * gw + x * gx + y * gy + z * gzq;
    const float inv = 1.0f / g.qn;
    float4 gq_out = make_float4((gw - w * dotp) * inv, (gx - x * dotp) * inv, (gy - y * dotp) * inv, (gzq - z * dotp) * inv);
    if (acc) {
        const float4 old = g_rotation[i];
        gq_out.x += old.x; gq_out.y += old.y; gq_out.z += old.z; gq_out.w += old.w;
    }
    g_rotation[i] = gq_out;
}

static ShParams make_sh(const float* rest, int64_t stride, float* g_rest, int64_t g_stride, int degree, const float* camera_host) {
    ShParams sh;
    sh.rest = rest; sh.stride = stride; sh.g_rest = g_rest; sh.g_stride = g_stride; sh.degree = degree;
    // contiguous [n,15,3] rows on both sides and 16-byte aligned bases: stage through shared memory
    sh.staged = degree > 0 && stride == kShRowFloats && (g_rest == nullptr || g_stride == kShRowFloats) &&
                ((uintptr_t)rest % 16 == 0) && ((uintptr_t)g_rest % 16 == 0);
    sh.cx = camera_host[16]; sh.cy = camera_host[17]; sh.cz = camera_host[18];
    return sh;
}

}  // namespace gs

using namespace gs;

extern "C" int gs_project_fwd(int64_t n, const float* xyz, const float* scaling_log, const float* rotation,
                              const float* cov3d, const float* opacity, int32_t opacity_is_logit,
                              const float* feat0, int64_t feat_stride, const float* sh_rest, int64_t sh_rest_stride,
                              int32_t sh_degree, const float* camera_host,
                              int32_t img_w, int32_t img_h, int32_t tile_size, float radius_min, float radius_max,
                              float* means2d, float* depths, float* conics, float* radii, float* colors,
                              float* opacities, uint8_t* vis, int32_t* tiles_touched, uint16_t* tile_rect,
                              uint32_t* depth_keys, float* splat_rec, void* stream) {
    GS_REQUIRE(n >= 0, "n < 0");
    GS_REQUIRE(camera_host != nullptr, "camera_host is NULL");
    GS_REQUIRE(img_w > 0 && img_h > 0, "image size must be positive");
    GS_REQUIRE(tile_size >= 1 && tile_size <= 4096, "tile_size out of range");
    GS_REQUIRE((img_w + tile_size - 1) / tile_size <= 65535 && (img_h + tile_size - 1) / tile_size <= 65535,
               "image too large for uint16 tile rects");
    if (n == 0) return GS_OK;
    const bool param_mode = scaling_log != nullptr && rotation != nullptr;
    GS_REQUIRE(param_mode || cov3d != nullptr, "need (scaling_log, rotation) or cov3d");
    GS_REQUIRE(xyz && opacity && feat0 && means2d && depths && conics && radii && colors && opacities && vis &&
                   tiles_touched && tile_rect && depth_keys && splat_rec, "NULL array argument");
    GS_REQUIRE(sh_degree >= 0 && sh_degree <= 3, "sh_degree must be 0..3");
    GS_REQUIRE(sh_degree == 0 || sh_rest != nullptr, "sh_degree > 0 needs sh_rest");
    DeviceGuard guard(xyz);
    const Camera cam = camera_from_host(camera_host);
    const ShParams sh = make_sh(sh_rest, sh_rest_stride, nullptr, 0, sh_degree, camera_host);
    const int threads = 256;
    const unsigned blocks = (unsigned)((n + threads - 1) / threads);
    const size_t smem = (sh.degree > 0 && sh.staged) ? (size_t)threads * kShRowFloats * sizeof(float) : 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (param_mode) {
#define GS_LAUNCH_FWD(PM, SH)                                                                                          \
    project_fwd_kernel<PM, SH><<<blocks, threads, smem, st>>>(                                                       \
        n, xyz, scaling_log, rotation, cov3d, opacity, opacity_is_logit, feat0, feat_stride, sh, cam, img_w, img_h, tile_size, \
        radius_min, radius_max, (float2*)means2d, depths, (float4*)conics, radii, colors, opacities, vis,             \
        tiles_touched, (ushort4*)tile_rect, depth_keys, (float4*)splat_rec)
        if (sh.degree > 0) GS_LAUNCH_FWD(true, true); else GS_LAUNCH_FWD(true, false);
    } else {
        if (sh.degree > 0) GS_LAUNCH_FWD(false, true); else GS_LAUNCH_FWD(false, false);
#undef GS_LAUNCH_FWD
    }
    GS_CUDA_TRY(cudaGetLastError());
    count_launches(1);
    return GS_OK;
}

extern "C" int gs_project_bwd(int64_t n, const float* xyz, const float* scaling_log, const float* rotation,
                              const float* cov3d, const float* opacity, int32_t opacity_is_logit,
                              const float* feat0, int64_t feat_stride, const float* sh_rest, int64_t sh_rest_stride,
                              int32_t sh_degree, const float* camera_host,
                              const float* g_means2d, const float* g_conics, const float* g_depths,
                              const float* g_colors, const float* g_opacities, float* g_xyz, float* g_scaling_log,
                              float* g_rotation, float* g_cov3d, float* g_opacity, float* g_feat0,
                              int64_t g_feat_stride, float* g_sh_rest, int64_t g_sh_rest_stride, int32_t accumulate,
                              const float* stat_radii, const uint8_t* stat_vis, float* stat_grad_norm,
                              float* stat_count, float* stat_max_radii, void* stream) {
    GS_REQUIRE(n >= 0, "n < 0");
    GS_REQUIRE(camera_host != nullptr, "camera_host is NULL");
    if (n == 0) return GS_OK;
    const bool param_mode = scaling_log != nullptr && rotation != nullptr;
    GS_REQUIRE(param_mode || cov3d != nullptr, "need (scaling_log, rotation) or cov3d");
    GS_REQUIRE(!param_mode || (g_scaling_log && g_rotation), "parameter mode needs g_scaling_log and g_rotation");
    GS_REQUIRE(param_mode || g_cov3d, "covariance mode needs g_cov3d");
    GS_REQUIRE(xyz && opacity && feat0 && g_means2d && g_conics && g_depths && g_colors && g_opacities && g_xyz &&
                   g_opacity && g_feat0, "NULL array argument");
    GS_REQUIRE(sh_degree >= 0 && sh_degree <= 3, "sh_degree must be 0..3");
    GS_REQUIRE(sh_degree == 0 || (sh_rest != nullptr && g_sh_rest != nullptr), "sh_degree > 0 needs sh_rest and g_sh_rest");
    const bool want_stats = stat_grad_norm != nullptr;
    GS_REQUIRE(!want_stats || (stat_radii && stat_vis && stat_count && stat_max_radii),
               "densification statistics need radii, vis, grad_norm, count and max_radii together");
    const DensifyStats stats = {stat_radii, stat_vis, want_stats ? stat_grad_norm : nullptr, stat_count, stat_max_radii};
    DeviceGuard guard(xyz);
    const Camera cam = camera_from_host(camera_host);
    const ShParams sh = make_sh(sh_rest, sh_rest_stride, g_sh_rest, g_sh_rest_stride, sh_degree, camera_host);
    const int threads = 256;
    const unsigned blocks = (unsigned)((n + threads - 1) / threads);
    const size_t smem = (sh.degree > 0 && sh.staged) ? (size_t)threads * kShRowFloats * sizeof(float) : 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (param_mode) {
#define GS_LAUNCH_BWD(PM, SH)                                                                                          \
    GS_CUDA_TRY(launch_pdl(project_bwd_kernel<PM, SH>, dim3(blocks), dim3(threads), smem, st,                          \
        n, xyz, scaling_log, rotation, cov3d, opacity, opacity_is_logit, feat0, feat_stride, sh, cam,                  \
        (const float2*)g_means2d, (const float4*)g_conics, g_depths, g_colors, g_opacities, g_xyz, g_scaling_log,      \
        (float4*)g_rotation, g_cov3d, g_opacity, g_feat0, g_feat_stride, accumulate, stats))
        if (sh.degree > 0) GS_LAUNCH_BWD(true, true); else GS_LAUNCH_BWD(true, false);
    } else {
        if (sh.degree > 0) GS_LAUNCH_BWD(false, true); else GS_LAUNCH_BWD(false, false);
#undef GS_LAUNCH_BWD
    }
    GS_CUDA_TRY(cudaGetLastError());
    count_launches(1);
    return GS_OK;
}
