// spsg_raycast.cu -- hand-written sm_100a kernels + C ABI of the SPSG-semantic raycaster.
//
// Functional contract: reference torch/utils/raycast_rgbd/raycast_rgbd_cuda_kernel.cu (cited per function
// as kernel.cu:<line>) evaluated with the exact fp32 operation order of its sm_100 SASS (SURVEY.md
// section 3.5; re-derived from `cuobjdump -sass oracle/_ref/*.so`).  Every rounding that can change a hit
// mask is spelled with __f*_rn intrinsics so nvcc can neither contract nor re-associate it.
//
// Design (not a port): see DESIGN.md.  In short
//   * the march replays the reference's `ray += inc` running sum bit-exactly but jumps over samples
//     that are provably invalid (outside the grid, or inside an empty 8^3 brick) with a closed form
//     of the fp32 recurrence that is exact inside one binade;
//   * the 8 corner indices / SDF values of a sample are fetched as two rounds of independent loads
//     instead of 16 dependent ones;
//   * the backward is a deterministic per-voxel gather (no float atomics, no 524288-block launch,
//     no 164 MB memsets).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "spsg_raycast.h"

namespace {

constexpr int kBrickLog2 = 3;
constexpr int kBrick = 1 << kBrickLog2;
constexpr float kBoxEps = 1.0f / 64.0f;  // shrink of skip boxes; >> every fp32 error term

thread_local char g_err[512] = "";

int fail(int code, const char *msg) {
    snprintf(g_err, sizeof(g_err), "%s", msg);
    return code;
}
int fail_cuda(cudaError_t e, const char *where) {
    snprintf(g_err, sizeof(g_err), "%s: %s", where, cudaGetErrorString(e));
    return SPSG_ERR_CUDA;
}
#define CUDA_TRY(x)                                      \
    do {                                                 \
        cudaError_t e_ = (x);                            \
        if (e_ != cudaSuccess) return fail_cuda(e_, #x); \
    } while (0)

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct Layout {
    int nbx, nby, nbz;
    size_t brick_off, brick_bytes;
    size_t hits_off, hits_bytes;
    size_t total;
};

Layout make_layout(const spsg_raycast_params *p) {
    Layout L;
    L.nbx = (p->dimx + kBrick - 1) >> kBrickLog2;
    L.nby = (p->dimy + kBrick - 1) >> kBrickLog2;
    L.nbz = (p->dimz + kBrick - 1) >> kBrickLog2;
    const int F = p->views_per_chunk > 0 ? p->views_per_chunk : 1;
    L.brick_off = 0;
    L.brick_bytes = align_up((size_t)p->num_chunks * L.nbx * L.nby * L.nbz, 256);
    L.hits_off = L.brick_off + L.brick_bytes;
    L.hits_bytes = align_up((size_t)p->num_chunks * F * p->width * p->height * sizeof(int32_t), 256);
    L.total = L.hits_off + L.hits_bytes;
    return L;
}

// ---------------------------------------------------------------------------------------------
// exact fp32 building blocks
// ---------------------------------------------------------------------------------------------

// make_int3(pos + make_float3(sign(pos))*0.5f)  (kernel.cu:89; cutil_math.h:31,179).  The reference
// SASS is trunc(fma(float(sign(t)), 0.5, t)); t + copysign(0.5, t) gives the same integer for every
// t (sign*0.5 is exact; for t == +-0 both truncate to 0).
__device__ __forceinline__ int round_voxel(float t) {
    return __float2int_rz(__fadd_rn(t, copysignf(0.5f, t)));
}

struct Ray {
    float camx, camy, camz;
    float dx, dy, dz;
    float d2r, t0, t1;
};

// kernel.cu:287-293 + :72-85 + :194-197, cutil_math.h:1207, cuda_SimpleMatrixUtil.h:888-907.
__device__ __forceinline__ Ray setup_ray(const float *__restrict__ M, const float *__restrict__ K, unsigned ux,
                                         unsigned uy, float dmin, float dmax) {
    const float fx = __ldg(K + 0), fy = __ldg(K + 1), mx = __ldg(K + 2), my = __ldg(K + 3);
    const float xn = __fdiv_rn(__fadd_rn((float)ux, -mx), fx);
    const float yn = __fdiv_rn(__fadd_rn((float)uy, -my), fy);
    const float zc = __fadd_rn(__fadd_rn(dmax, -dmin), dmin);
    const float vx = __fmul_rn(xn, zc), vy = __fmul_rn(yn, zc);
    float r = rsqrtf(__fmaf_rn(zc, zc, __fmaf_rn(vx, vx, __fmul_rn(vy, vy))));
    const float cx = __fmul_rn(vx, r), cy = __fmul_rn(vy, r), cz = __fmul_rn(r, zc);
    float m[12];
#pragma unroll
    for (int i = 0; i < 12; i++) m[i] = __ldg(M + i);
    Ray o;
    o.camx = __fadd_rn(m[3], __fmaf_rn(0.0f, m[2], __fmaf_rn(0.0f, m[0], __fmul_rn(0.0f, m[1]))));
    o.camy = __fadd_rn(m[7], __fmaf_rn(0.0f, m[6], __fmaf_rn(0.0f, m[4], __fmul_rn(0.0f, m[5]))));
    o.camz = __fadd_rn(m[11], __fmaf_rn(0.0f, m[10], __fmaf_rn(0.0f, m[8], __fmul_rn(0.0f, m[9]))));
    const float wx = __fmaf_rn(0.0f, m[3], __fmaf_rn(m[2], cz, __fmaf_rn(m[0], cx, __fmul_rn(m[1], cy))));
    const float wy = __fmaf_rn(0.0f, m[7], __fmaf_rn(m[6], cz, __fmaf_rn(m[4], cx, __fmul_rn(m[5], cy))));
    const float wz = __fmaf_rn(0.0f, m[11], __fmaf_rn(m[10], cz, __fmaf_rn(m[8], cx, __fmul_rn(m[9], cy))));
    r = rsqrtf(__fmaf_rn(wz, wz, __fmaf_rn(wx, wx, __fmul_rn(wy, wy))));
    o.dx = __fmul_rn(wx, r);
    o.dy = __fmul_rn(wy, r);
    o.dz = __fmul_rn(wz, r);
    o.d2r = __frcp_rn(cz);
    o.t0 = __fmul_rn(o.d2r, dmin);
    o.t1 = __fmul_rn(o.d2r, dmax);
    return o;
}

// Advance the reference's running sum `ray = ray + inc` (kernel.cu:257,260) by `want` >= 1 steps, or by
// fewer (>= 1) when the closed form would leave the current binade.  Bit-exact: inside [2^e, 2^(e+1))
// every partial sum is a multiple of u = 2^(e-23), so fl(s + inc) = s + d with d = inc rounded to the
// u grid -- a constant as long as inc is not an exact tie between two grid points -- and s + j*d is
// representable, so one fma reproduces j sequential adds.  Anything irregular falls back to real adds.
__device__ __forceinline__ float advance_ray(float ray, float inc, int want) {
    if (want > 2) {
        const float lo = __uint_as_float(__float_as_uint(ray) & 0x7f800000u);  // 2^e <= ray
        const float u = __fmul_rn(lo, 1.1920928955078125e-07f);                // 2^(e-23)
        const float d = __fadd_rn(__fadd_rn(lo, inc), -lo);                    // inc on the u grid
        const float rem = __fadd_rn(inc, -d);                                  // exact remainder
        const bool regular = (lo >= 1.0f) && (lo <= 8388608.0f) && (inc > 0.0f) && (inc <= 0.25f * lo) &&
                             (d > 0.0f) && (__fmul_rn(fabsf(rem), 2.0f) != u);
        if (regular) {
            // all partial sums must stay below 2^(e+1) - inc so that every add rounds on the u grid
            const float room = __fadd_rn(__fadd_rn(__fmul_rn(lo, 2.0f), -__fmul_rn(inc, 2.0f)), -ray);
            const int jmax = (room > 0.0f) ? __float2int_rd(__fdiv_rn(room, d)) : 0;
            const int j = min(want, jmax);
            if (j >= 1) return __fmaf_rn((float)j, d, ray);
        } else {
            for (int k = 0; k < want; k++) ray = __fadd_rn(ray, inc);
            return ray;
        }
        return __fadd_rn(ray, inc);
    }
    ray = __fadd_rn(ray, inc);
    if (want == 2) ray = __fadd_rn(ray, inc);
    return ray;
}

struct Volume {
    const int32_t *__restrict__ index;  // this chunk's slice of sparse_mapping
    const float *__restrict__ sdf;      // vals_sdf
    int dimx, dimy, dimz;
};

__device__ __forceinline__ bool in_grid(const Volume &v, int x, int y, int z) {
    return (x | y | z) >= 0 && x < v.dimx && y < v.dimy && z < v.dimz;
}

// trilinearInterpolationSimpleFastFast (kernel.cu:120-156) without the payload.  Exact corner
// coordinates, weights, product order and accumulation order of the reference SASS.
template <bool kNearest>
__device__ __forceinline__ bool sample_sdf(const Volume &v, float px, float py, float pz, float &dist, int &nearest) {
    const float qx = __fadd_rn(px, -0.5f), qy = __fadd_rn(py, -0.5f), qz = __fadd_rn(pz, -0.5f);
    const int x0 = round_voxel(qx), y0 = round_voxel(qy), z0 = round_voxel(qz);
    const int x1 = round_voxel(__fadd_rn(qx, 1.0f)), y1 = round_voxel(__fadd_rn(qy, 1.0f)),
              z1 = round_voxel(__fadd_rn(qz, 1.0f));
    if (kNearest) {
        const int nx = round_voxel(px), ny = round_voxel(py), nz = round_voxel(pz);
        nearest = in_grid(v, nx, ny, nz) ? __ldg(v.index + ((size_t)nz * v.dimy + ny) * v.dimx + nx) : -1;
    }
    if (!(in_grid(v, x0, y0, z0) && in_grid(v, x1, y1, z1))) return false;
    const int r00 = (z0 * v.dimy + y0) * v.dimx, r10 = (z0 * v.dimy + y1) * v.dimx;
    const int r01 = (z1 * v.dimy + y0) * v.dimx, r11 = (z1 * v.dimy + y1) * v.dimx;
    const int i000 = __ldg(v.index + r00 + x0), i100 = __ldg(v.index + r00 + x1);
    const int i010 = __ldg(v.index + r10 + x0), i110 = __ldg(v.index + r10 + x1);
    const int i001 = __ldg(v.index + r01 + x0), i101 = __ldg(v.index + r01 + x1);
    const int i011 = __ldg(v.index + r11 + x0), i111 = __ldg(v.index + r11 + x1);
    if ((i000 | i100 | i010 | i110 | i001 | i101 | i011 | i111) < 0) return false;
    const float v000 = __ldg(v.sdf + i000), v100 = __ldg(v.sdf + i100), v010 = __ldg(v.sdf + i010),
                v001 = __ldg(v.sdf + i001), v110 = __ldg(v.sdf + i110), v011 = __ldg(v.sdf + i011),
                v101 = __ldg(v.sdf + i101), v111 = __ldg(v.sdf + i111);
    const float wx = __fadd_rn(px, -floorf(px)), wy = __fadd_rn(py, -floorf(py)), wz = __fadd_rn(pz, -floorf(pz));
    const float ax = __fadd_rn(1.0f, -wx), ay = __fadd_rn(1.0f, -wy), az = __fadd_rn(1.0f, -wz);
    const float axay = __fmul_rn(ax, ay), wxay = __fmul_rn(wx, ay), axwy = __fmul_rn(ax, wy), wxwy = __fmul_rn(wx, wy);
    float d = __fmaf_rn(v000, __fmul_rn(axay, az), 0.0f);
    d = __fmaf_rn(v100, __fmul_rn(wxay, az), d);
    d = __fmaf_rn(v010, __fmul_rn(axwy, az), d);
    d = __fmaf_rn(v001, __fmul_rn(axay, wz), d);
    d = __fmaf_rn(v110, __fmul_rn(wxwy, az), d);
    d = __fmaf_rn(v011, __fmul_rn(axwy, wz), d);
    d = __fmaf_rn(v101, __fmul_rn(wxay, wz), d);
    d = __fmaf_rn(v111, __fmul_rn(wxwy, wz), d);
    dist = d;
    return true;
}

// Slab test of the ray against [lo, hi]^3-style box; returns parameter interval.
__device__ __forceinline__ void slab(float o, float d, float lo, float hi, float &tin, float &tout) {
    if (d != 0.0f) {
        const float inv = __frcp_rn(d);
        const float a = (lo - o) * inv, b = (hi - o) * inv;
        tin = fmaxf(tin, fminf(a, b));
        tout = fminf(tout, fmaxf(a, b));
    } else if (o < lo || o > hi) {
        tin = __int_as_float(0x7f800000);
        tout = -__int_as_float(0x7f800000);
    }
}

struct ForwardArgs {
    const int32_t *sparse_mapping;
    const float *vals_sdf, *vals_color, *vals_normal, *vals_semantic;
    const float *view_matrix, *intrinsics;
    float *image_color, *image_depth, *image_normal, *image_semantic;
    int32_t *mapping3dto2d, *mapping3dto2d_num;
    const uint8_t *bricks;
    int32_t *hits;
    int width, height;
    float depth_min, depth_max, thresh, inc;
    int dimx, dimy, dimz;
    int nbx, nby, nbz;
    int views, max_pixels;
    long long num_locs;
    unsigned flags;
};

constexpr int kTileW = 16, kTileH = 8;  // pixels per CTA: 4 warps of 8x4 pixels

// One thread per ray.  kernel.cu:265-297 (init + ray), :190-263 (march), :166-187 (regula falsi),
// :215-249 (hit write-out + voxel->pixel registration).
__global__ void __launch_bounds__(kTileW *kTileH) raycast_forward_kernel(const ForwardArgs a) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned ux = blockIdx.x * kTileW + (warp & 1) * 8 + (lane & 7);
    const unsigned uy = blockIdx.y * kTileH + (warp >> 1) * 4 + (lane >> 3);
    const int img = blockIdx.z;
    if (ux >= (unsigned)a.width || uy >= (unsigned)a.height) return;
    const int chunk = img / a.views, view = img - chunk * a.views;
    const unsigned pix = uy * a.width + ux;
    const size_t gpix = (size_t)img * a.width * a.height + pix;

    const Ray r = setup_ray(a.view_matrix + (size_t)img * 16, a.intrinsics + (size_t)img * 4, ux, uy, a.depth_min,
                            a.depth_max);
    Volume v;
    v.index = a.sparse_mapping + (size_t)chunk * a.dimz * a.dimy * a.dimx;
    v.sdf = a.vals_sdf;
    v.dimx = a.dimx; v.dimy = a.dimy; v.dimz = a.dimz;
    const uint8_t *__restrict__ bricks = a.bricks + (size_t)chunk * a.nbz * a.nby * a.nbx;
    const bool clip = !(a.flags & SPSG_FLAG_NO_CLIP);
    const bool skip = !(a.flags & SPSG_FLAG_NO_BRICK_SKIP);

    float ray = r.t0, t_end = r.t1;
    if (clip) {
        // Samples are valid only for p in (0, dim-1) on every axis (all 8 corners inside the grid).
        float tin = -__int_as_float(0x7f800000), tout = __int_as_float(0x7f800000);
        slab(r.camx, r.dx, -kBoxEps, (float)(a.dimx - 1) + kBoxEps, tin, tout);
        slab(r.camy, r.dy, -kBoxEps, (float)(a.dimy - 1) + kBoxEps, tin, tout);
        slab(r.camz, r.dz, -kBoxEps, (float)(a.dimz - 1) + kBoxEps, tin, tout);
        const float margin = 0.0625f;
        if (!(tin <= tout)) {
            t_end = ray;  // misses the grid: nothing to march
        } else {
            t_end = fminf(t_end, tout + margin);
            // jump to the last sample at or before tin - margin
            while (ray < tin - margin - a.inc && ray < t_end) {
                const float ahead = __fdiv_rn(tin - margin - ray, a.inc);
                const int want = max(1, min(__float2int_rd(ahead) - 1, 1 << 22));
                ray = advance_ray(ray, a.inc, want);
            }
        }
    }

    float last_sdf = 0.0f, last_alpha = 0.0f;
    bool last_ok = false;
    int hit = -1;
    float depth = 0.0f;

    while (ray < t_end) {  // kernel.cu:200
        const float px = __fmaf_rn(r.dx, ray, r.camx), py = __fmaf_rn(r.dy, ray, r.camy),
                    pz = __fmaf_rn(r.dz, ray, r.camz);
        if (skip) {
            // Brick of floor(p).  If p is at least kBoxEps inside an empty (or out-of-grid) brick on every axis,
            // the sample's corner (0,0,0) lies in that brick and is absent: the sample is invalid, and so is
            // every later sample until the ray leaves the shrunken brick.
            const float fx = floorf(px), fy = floorf(py), fz = floorf(pz);
            const int bx = __float2int_rd(px) >> kBrickLog2, by = __float2int_rd(py) >> kBrickLog2,
                      bz = __float2int_rd(pz) >> kBrickLog2;
            bool empty = true;
            if ((bx | by | bz) >= 0 && bx < a.nbx && by < a.nby && bz < a.nbz)
                empty = bricks[(bz * a.nby + by) * a.nbx + bx] == 0;
            if (empty) {
                const float lox = (float)(bx << kBrickLog2) + kBoxEps, hix = (float)((bx + 1) << kBrickLog2) - kBoxEps;
                const float loy = (float)(by << kBrickLog2) + kBoxEps, hiy = (float)((by + 1) << kBrickLog2) - kBoxEps;
                const float loz = (float)(bz << kBrickLog2) + kBoxEps, hiz = (float)((bz + 1) << kBrickLog2) - kBoxEps;
                (void)fx; (void)fy; (void)fz;
                if (px >= lox && px <= hix && py >= loy && py <= hiy && pz >= loz && pz <= hiz) {
                    float tin = -__int_as_float(0x7f800000), tout = __int_as_float(0x7f800000);
                    slab(r.camx, r.dx, lox, hix, tin, tout);
                    slab(r.camy, r.dy, loy, hiy, tin, tout);
                    slab(r.camz, r.dz, loz, hiz, tin, tout);
                    int want = 1;
                    if (tout > ray) want = max(1, min(__float2int_rd(__fdiv_rn(tout - ray, a.inc)) + 1, 1 << 22));
                    last_ok = false;  // kernel.cu:259
                    ray = advance_ray(ray, a.inc, want);
                    continue;
                }
            }
        }
        float dist;
        int unused;
        if (sample_sdf<false>(v, px, py, pz, dist, unused)) {
            if (last_ok && ((last_sdf > 0.0f && dist < 0.0f) || (last_sdf < 0.0f && dist > 0.0f))) {  // :205
                // findIntersectionBisection (:166-187)
                float ta = last_alpha, da = last_sdf, tb = ray, db = dist, c = 0.0f;
                bool ok = true;
                int nearest = -1;
#pragma unroll 1
                for (int k = 0; k < 3; k++) {
                    c = __fmaf_rn(__fadd_rn(tb, -ta), __fdiv_rn(da, __fadd_rn(da, -db)), ta);  // :161
                    float dc;
                    if (!sample_sdf<true>(v, __fmaf_rn(r.dx, c, r.camx), __fmaf_rn(r.dy, c, r.camy),
                                          __fmaf_rn(r.dz, c, r.camz), dc, nearest)) {
                        ok = false;
                        break;
                    }
                    if (__fmul_rn(da, dc) > 0.0f) { ta = c; da = dc; } else { tb = c; db = dc; }  // :180-181
                }
                if (ok && fabsf(__fadd_rn(last_sdf, -dist)) < a.thresh && fabsf(dist) < a.thresh) {  // :211-213
                    depth = __fdiv_rn(c, r.d2r);                                                     // :215
                    hit = nearest;  // == round(cam + alpha*dir), :241-242 (same fma as the last refinement point)
                    break;
                }
            }
            last_sdf = dist; last_alpha = ray; last_ok = true;  // :254-256
        } else {
            last_ok = false;  // :259
        }
        ray = __fadd_rn(ray, a.inc);  // :257,:260
    }

    // write-out (kernel.cu:276-285 init, :217-239 hit)
    const float ninf = __int_as_float(0xff800000);
    float *oc = a.image_color + gpix * 3, *on = a.image_normal + gpix * 3, *os = a.image_semantic + gpix * 14;
    if (hit >= 0) {
        const float *c = a.vals_color + (size_t)hit * 3, *n = a.vals_normal + (size_t)hit * 3,
                    *s = a.vals_semantic + (size_t)hit * 14;
        oc[0] = __ldg(c + 0); oc[1] = __ldg(c + 1); oc[2] = __ldg(c + 2);
        const float n0 = __ldg(n + 0), n1 = __ldg(n + 1), n2 = __ldg(n + 2);
        const bool zero_normal = (n0 == 0.0f && n1 == 0.0f && n2 == 0.0f);  // :220
        on[0] = zero_normal ? ninf : n0; on[1] = zero_normal ? ninf : n1; on[2] = zero_normal ? ninf : n2;
        a.image_depth[gpix] = depth;
#pragma unroll
        for (int k = 0; k < 14; k++) os[k] = __ldg(s + k);
        const size_t row = (size_t)view * (size_t)a.num_locs + (size_t)hit;
        const int offset = atomicAdd(a.mapping3dto2d_num + row, 1);                            // :244
        if (offset < a.max_pixels) a.mapping3dto2d[row * a.max_pixels + offset] = (int)pix;  // :245-247
    } else {
        oc[0] = ninf; oc[1] = ninf; oc[2] = ninf;
        on[0] = ninf; on[1] = ninf; on[2] = ninf;
        a.image_depth[gpix] = ninf;
#pragma unroll
        for (int k = 0; k < 14; k++) os[k] = ninf;
    }
    if (a.hits) a.hits[gpix] = hit;
}

// construct_dense_sparse_mapping_kernel (kernel.cu:346-362) + brick marking + counter reset.
template <bool kWriteIndex>
__global__ void __launch_bounds__(256) index_kernel(const longlong4 *__restrict__ locs, long long n,
                                                    int32_t *__restrict__ sparse_mapping, uint8_t *__restrict__ bricks,
                                                    int32_t *__restrict__ num, int views, int dimz, int dimy, int dimx,
                                                    int nbz, int nby, int nbx) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const longlong4 l = locs[i];  // (z, y, x, chunk)
    const long long z = l.x, y = l.y, x = l.z, b = l.w;
    if (kWriteIndex) sparse_mapping[((b * dimz + z) * dimy + y) * dimx + x] = (int32_t)i;
    if (bricks)
        bricks[((b * nbz + (z >> kBrickLog2)) * nby + (y >> kBrickLog2)) * nbx + (x >> kBrickLog2)] = 1;
    if (num)
        for (int f = 0; f < views; f++) num[(long long)f * n + i] = 0;
}

// Deterministic backward (replaces kernel.cu:365-423): one warp per 32 dense cells; for every present
// voxel the warp gathers, lane == channel, the gradients of its registered pixels in registration
// order and writes the per-view means.  Present voxels nobody hit are written as zeros, so no memset
// of d_* is needed.
struct BackwardArgs {
    const float *grad_color, *grad_depth, *grad_normal, *grad_semantic;
    const int32_t *sparse_mapping, *mapping3dto2d, *mapping3dto2d_num;
    float *d_color, *d_depth, *d_normal, *d_semantic;
    int width, height;
    long long cells_per_chunk;
    int num_chunks, views, max_pixels;
    long long num_locs;
};

__device__ __forceinline__ float load_grad(const BackwardArgs &a, int lane, size_t gpix) {
    // lane: 0-2 colour, 3 depth, 4-6 normal, 7-20 semantic
    if (lane < 3) return __ldg(a.grad_color + gpix * 3 + lane);
    if (lane == 3) return __ldg(a.grad_depth + gpix);
    if (lane < 7) return __ldg(a.grad_normal + gpix * 3 + (lane - 4));
    return __ldg(a.grad_semantic + gpix * 14 + (lane - 7));
}

__global__ void __launch_bounds__(256) raycast_backward_kernel(const BackwardArgs a) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long total = a.cells_per_chunk * a.num_chunks;
    const long long cell = warp * 32 + lane;
    int idx = -1;
    if (cell < total) idx = __ldg(a.sparse_mapping + cell);
    unsigned present = __ballot_sync(0xffffffffu, idx >= 0);
    if (!present) return;
    const int my_chunk = (int)(cell / a.cells_per_chunk);
    const size_t P = (size_t)a.width * a.height;
    while (present) {
        const int src = __ffs(present) - 1;
        present &= present - 1;
        const int vidx = __shfl_sync(0xffffffffu, idx, src);
        const int chunk = __shfl_sync(0xffffffffu, my_chunk, src);
        float acc = 0.0f;
        for (int f = 0; f < a.views; f++) {
            const size_t row = (size_t)f * a.num_locs + vidx;
            const int num = __ldg(a.mapping3dto2d_num + row);
            if (num <= 0) continue;
            const int cnt = min(num, a.max_pixels);
            const float fcnt = (float)cnt;
            const size_t img_base = ((size_t)chunk * a.views + f) * P;
            float sum = 0.0f;
            for (int t0 = 0; t0 < cnt; t0 += 32) {
                const int my_pix = (t0 + lane < cnt) ? __ldg(a.mapping3dto2d + row * a.max_pixels + t0 + lane) : 0;
                const int m = min(32, cnt - t0);
                for (int t = 0; t < m; t++) {
                    const int pixel = __shfl_sync(0xffffffffu, my_pix, t);
                    if (lane < SPSG_GRAD_CHANNELS)
                        sum = __fadd_rn(sum, __fdiv_rn(load_grad(a, lane, img_base + pixel), fcnt));  // :398-418
                }
            }
            acc += sum;
        }
        if (lane < 3) a.d_color[(size_t)vidx * 3 + lane] = acc;
        else if (lane == 3) a.d_depth[vidx] = acc;
        else if (lane < 7) a.d_normal[(size_t)vidx * 3 + (lane - 4)] = acc;
        else if (lane < SPSG_GRAD_CHANNELS) a.d_semantic[(size_t)vidx * 14 + (lane - 7)] = acc;
    }
}

// raycast_occ_cuda_kernel (kernel.cu:320-344) + traverseOccGrid (:301-318).
struct OccArgs {
    const uint8_t *occ3d;
    uint8_t *occ2d;
    const float *view_matrix, *intrinsics;
    int width, height;
    float depth_min, depth_max, inc;
    int dimx, dimy, dimz;
    unsigned flags;
};

__global__ void __launch_bounds__(kTileW *kTileH) raycast_occ_kernel(const OccArgs a) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned ux = blockIdx.x * kTileW + (warp & 1) * 8 + (lane & 7);
    const unsigned uy = blockIdx.y * kTileH + (warp >> 1) * 4 + (lane >> 3);
    const int img = blockIdx.z;
    if (ux >= (unsigned)a.width || uy >= (unsigned)a.height) return;
    const Ray r = setup_ray(a.view_matrix + (size_t)img * 16, a.intrinsics + (size_t)img * 4, ux, uy, a.depth_min,
                            a.depth_max);
    const uint8_t *__restrict__ occ = a.occ3d + (size_t)img * a.dimz * a.dimy * a.dimx;
    float ray = r.t0, t_end = r.t1;
    if (!(a.flags & SPSG_FLAG_NO_CLIP)) {
        // nearest voxel is inside the grid only for p in (-0.5, dim-0.5)
        float tin = -__int_as_float(0x7f800000), tout = __int_as_float(0x7f800000);
        slab(r.camx, r.dx, -0.5f - kBoxEps, (float)a.dimx - 0.5f + kBoxEps, tin, tout);
        slab(r.camy, r.dy, -0.5f - kBoxEps, (float)a.dimy - 0.5f + kBoxEps, tin, tout);
        slab(r.camz, r.dz, -0.5f - kBoxEps, (float)a.dimz - 0.5f + kBoxEps, tin, tout);
        const float margin = 0.0625f;
        if (!(tin <= tout)) {
            t_end = ray;
        } else {
            t_end = fminf(t_end, tout + margin);
            while (ray < tin - margin - a.inc && ray < t_end) {
                const float ahead = __fdiv_rn(tin - margin - ray, a.inc);
                const int want = max(1, min(__float2int_rd(ahead) - 1, 1 << 22));
                ray = advance_ray(ray, a.inc, want);
            }
        }
    }
    uint8_t out = 0;  // :334
    while (ray < t_end) {
        const int x = round_voxel(__fmaf_rn(r.dx, ray, r.camx)), y = round_voxel(__fmaf_rn(r.dy, ray, r.camy)),
                  z = round_voxel(__fmaf_rn(r.dz, ray, r.camz));
        if ((x | y | z) >= 0 && x < a.dimx && y < a.dimy && z < a.dimz &&
            occ[((size_t)z * a.dimy + y) * a.dimx + x] != 0) {  // :310-313
            out = 1;
            break;
        }
        ray = __fadd_rn(ray, a.inc);  // :316
    }
    a.occ2d[(size_t)img * a.width * a.height + uy * a.width + ux] = out;
}

int check_params(const spsg_raycast_params *p) {
    if (!p) return fail(SPSG_ERR_INVALID_ARGUMENT, "params is NULL");
    if (p->width <= 0 || p->height <= 0) return fail(SPSG_ERR_INVALID_ARGUMENT, "width/height must be positive");
    if (p->dimx <= 0 || p->dimy <= 0 || p->dimz <= 0) return fail(SPSG_ERR_INVALID_ARGUMENT, "dims must be positive");
    if (p->num_chunks <= 0) return fail(SPSG_ERR_INVALID_ARGUMENT, "num_chunks must be positive");
    if (p->views_per_chunk <= 0) return fail(SPSG_ERR_INVALID_ARGUMENT, "views_per_chunk must be >= 1");
    if (p->num_locs < 0) return fail(SPSG_ERR_INVALID_ARGUMENT, "num_locs must be >= 0");
    if ((long long)p->num_chunks * p->dimx * p->dimy * p->dimz >= (1ll << 31))
        return fail(SPSG_ERR_INVALID_ARGUMENT, "dense grid exceeds 32-bit indexing (reference limit, SURVEY 3.5)");
    if ((long long)p->num_chunks * p->views_per_chunk * p->width * p->height * 14 >= (1ll << 32))
        return fail(SPSG_ERR_INVALID_ARGUMENT, "image batch exceeds 32-bit indexing (reference limit, SURVEY 3.5)");
    return SPSG_OK;
}

int launch_forward(const spsg_raycast_params *p, bool build_index, int32_t *sparse_mapping, const int64_t *locs,
                   const float *vals_sdf, const float *vals_color, const float *vals_normal,
                   const float *vals_semantic, const float *view_matrix, const float *intrinsics, float *image_color,
                   float *image_depth, float *image_normal, float *image_semantic, int32_t *mapping3dto2d,
                   int32_t *mapping3dto2d_num, void *workspace, size_t workspace_bytes, cudaStream_t st) {
    if (int rc = check_params(p)) return rc;
    if (p->max_pixels_per_voxel <= 0) return fail(SPSG_ERR_INVALID_ARGUMENT, "max_pixels_per_voxel must be positive");
    if (!sparse_mapping || !view_matrix || !intrinsics || !image_color || !image_depth || !image_normal ||
        !image_semantic || !mapping3dto2d || !mapping3dto2d_num)
        return fail(SPSG_ERR_INVALID_ARGUMENT, "NULL tensor pointer");
    if (p->num_locs > 0 && (!locs || !vals_sdf || !vals_color || !vals_normal || !vals_semantic))
        return fail(SPSG_ERR_INVALID_ARGUMENT, "NULL voxel tensor pointer");
    const Layout L = make_layout(p);
    if (!workspace || workspace_bytes < L.total) return fail(SPSG_ERR_WORKSPACE_TOO_SMALL, "workspace too small");
    uint8_t *ws = (uint8_t *)workspace;
    uint8_t *bricks = ws + L.brick_off;
    const size_t cells = (size_t)p->num_chunks * p->dimz * p->dimy * p->dimx;

    if (build_index) CUDA_TRY(cudaMemsetAsync(sparse_mapping, 0xff, cells * sizeof(int32_t), st));  // kernel.cu:515
    CUDA_TRY(cudaMemsetAsync(bricks, 0, L.brick_bytes, st));
    if (p->num_locs > 0) {
        const unsigned blocks = (unsigned)((p->num_locs + 255) / 256);
        if (build_index)
            index_kernel<true><<<blocks, 256, 0, st>>>((const longlong4 *)locs, p->num_locs, sparse_mapping, bricks,
                                                       mapping3dto2d_num, p->views_per_chunk, p->dimz, p->dimy,
                                                       p->dimx, L.nbz, L.nby, L.nbx);
        else
            index_kernel<false><<<blocks, 256, 0, st>>>((const longlong4 *)locs, p->num_locs, sparse_mapping, bricks,
                                                        mapping3dto2d_num, p->views_per_chunk, p->dimz, p->dimy,
                                                        p->dimx, L.nbz, L.nby, L.nbx);
        CUDA_TRY(cudaGetLastError());
    }
    ForwardArgs a;
    a.sparse_mapping = sparse_mapping;
    a.vals_sdf = vals_sdf; a.vals_color = vals_color; a.vals_normal = vals_normal; a.vals_semantic = vals_semantic;
    a.view_matrix = view_matrix; a.intrinsics = intrinsics;
    a.image_color = image_color; a.image_depth = image_depth; a.image_normal = image_normal;
    a.image_semantic = image_semantic;
    a.mapping3dto2d = mapping3dto2d; a.mapping3dto2d_num = mapping3dto2d_num;
    a.bricks = bricks;
    a.hits = (p->flags & SPSG_FLAG_RECORD_HITS) ? (int32_t *)(ws + L.hits_off) : nullptr;
    a.width = p->width; a.height = p->height;
    a.depth_min = p->depth_min; a.depth_max = p->depth_max; a.thresh = p->thresh_sample_dist; a.inc = p->ray_increment;
    a.dimx = p->dimx; a.dimy = p->dimy; a.dimz = p->dimz;
    a.nbx = L.nbx; a.nby = L.nby; a.nbz = L.nbz;
    a.views = p->views_per_chunk; a.max_pixels = p->max_pixels_per_voxel;
    a.num_locs = p->num_locs;
    a.flags = p->flags;
    const dim3 grid((p->width + kTileW - 1) / kTileW, (p->height + kTileH - 1) / kTileH,
                    p->num_chunks * p->views_per_chunk);
    raycast_forward_kernel<<<grid, kTileW * kTileH, 0, st>>>(a);
    CUDA_TRY(cudaGetLastError());
    return SPSG_OK;
}

}  // namespace

extern "C" {

const char *spsg_version(void) { return "spsg_raycast_b200 0.1 (sm_100a)"; }
const char *spsg_last_error(void) { return g_err; }

size_t spsg_workspace_bytes(const spsg_raycast_params *p) {
    if (check_params(p)) return 0;
    return make_layout(p).total;
}

int spsg_build_index(const int64_t *locs, int64_t num_locs, int32_t *sparse_mapping, int32_t num_chunks, int32_t dimz,
                     int32_t dimy, int32_t dimx, void *stream) {
    if (!sparse_mapping || (num_locs > 0 && !locs)) return fail(SPSG_ERR_INVALID_ARGUMENT, "NULL tensor pointer");
    if (num_chunks <= 0 || dimz <= 0 || dimy <= 0 || dimx <= 0 || num_locs < 0)
        return fail(SPSG_ERR_INVALID_ARGUMENT, "bad sizes");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t cells = (size_t)num_chunks * dimz * dimy * dimx;
    CUDA_TRY(cudaMemsetAsync(sparse_mapping, 0xff, cells * sizeof(int32_t), st));
    if (num_locs > 0) {
        index_kernel<true><<<(unsigned)((num_locs + 255) / 256), 256, 0, st>>>(
            (const longlong4 *)locs, num_locs, sparse_mapping, nullptr, nullptr, 0, dimz, dimy, dimx, 0, 0, 0);
        CUDA_TRY(cudaGetLastError());
    }
    return SPSG_OK;
}

int spsg_raycast_forward(const spsg_raycast_params *p, const int32_t *sparse_mapping, const int64_t *locs,
                         const float *vals_sdf, const float *vals_color, const float *vals_normal,
                         const float *vals_semantic, const float *view_matrix, const float *intrinsics,
                         float *image_color, float *image_depth, float *image_normal, float *image_semantic,
                         int32_t *mapping3dto2d, int32_t *mapping3dto2d_num, void *workspace, size_t workspace_bytes,
                         void *stream) {
    return launch_forward(p, false, const_cast<int32_t *>(sparse_mapping), locs, vals_sdf, vals_color, vals_normal,
                          vals_semantic, view_matrix, intrinsics, image_color, image_depth, image_normal,
                          image_semantic, mapping3dto2d, mapping3dto2d_num, workspace, workspace_bytes,
                          (cudaStream_t)stream);
}

int spsg_raycast_forward_indexed(const spsg_raycast_params *p, int32_t *sparse_mapping, const int64_t *locs,
                                 const float *vals_sdf, const float *vals_color, const float *vals_normal,
                                 const float *vals_semantic, const float *view_matrix, const float *intrinsics,
                                 float *image_color, float *image_depth, float *image_normal, float *image_semantic,
                                 int32_t *mapping3dto2d, int32_t *mapping3dto2d_num, void *workspace,
                                 size_t workspace_bytes, void *stream) {
    return launch_forward(p, true, sparse_mapping, locs, vals_sdf, vals_color, vals_normal, vals_semantic, view_matrix,
                          intrinsics, image_color, image_depth, image_normal, image_semantic, mapping3dto2d,
                          mapping3dto2d_num, workspace, workspace_bytes, (cudaStream_t)stream);
}

int spsg_raycast_backward(const spsg_raycast_params *p, const float *grad_color, const float *grad_depth,
                          const float *grad_normal, const float *grad_semantic, const int32_t *sparse_mapping,
                          const int32_t *mapping3dto2d, const int32_t *mapping3dto2d_num, float *d_color,
                          float *d_depth, float *d_normal, float *d_semantic, void *stream) {
    if (int rc = check_params(p)) return rc;
    if (!grad_color || !grad_depth || !grad_normal || !grad_semantic || !sparse_mapping || !mapping3dto2d ||
        !mapping3dto2d_num)
        return fail(SPSG_ERR_INVALID_ARGUMENT, "NULL tensor pointer");
    if (p->num_locs == 0) return SPSG_OK;
    if (!d_color || !d_depth || !d_normal || !d_semantic) return fail(SPSG_ERR_INVALID_ARGUMENT, "NULL gradient pointer");
    BackwardArgs a;
    a.grad_color = grad_color; a.grad_depth = grad_depth; a.grad_normal = grad_normal; a.grad_semantic = grad_semantic;
    a.sparse_mapping = sparse_mapping; a.mapping3dto2d = mapping3dto2d; a.mapping3dto2d_num = mapping3dto2d_num;
    a.d_color = d_color; a.d_depth = d_depth; a.d_normal = d_normal; a.d_semantic = d_semantic;
    a.width = p->width; a.height = p->height;
    a.cells_per_chunk = (long long)p->dimz * p->dimy * p->dimx;
    a.num_chunks = p->num_chunks; a.views = p->views_per_chunk; a.max_pixels = p->max_pixels_per_voxel;
    a.num_locs = p->num_locs;
    cudaStream_t st = (cudaStream_t)stream;
    const long long warps = (a.cells_per_chunk * a.num_chunks + 31) / 32;
    const unsigned blocks = (unsigned)((warps + 7) / 8);
    raycast_backward_kernel<<<blocks, 256, 0, st>>>(a);
    CUDA_TRY(cudaGetLastError());
    return SPSG_OK;
}

int spsg_raycast_occ(const spsg_raycast_params *p, const uint8_t *occ3d, uint8_t *occ2d, const float *view_matrix,
                     const float *intrinsics, void *stream) {
    if (int rc = check_params(p)) return rc;
    if (!occ3d || !occ2d || !view_matrix || !intrinsics) return fail(SPSG_ERR_INVALID_ARGUMENT, "NULL tensor pointer");
    OccArgs a;
    a.occ3d = occ3d; a.occ2d = occ2d; a.view_matrix = view_matrix; a.intrinsics = intrinsics;
    a.width = p->width; a.height = p->height;
    a.depth_min = p->depth_min; a.depth_max = p->depth_max; a.inc = p->ray_increment;
    a.dimx = p->dimx; a.dimy = p->dimy; a.dimz = p->dimz;
    a.flags = p->flags;
    const dim3 grid((p->width + kTileW - 1) / kTileW, (p->height + kTileH - 1) / kTileH, p->num_chunks);
    raycast_occ_kernel<<<grid, kTileW * kTileH, 0, (cudaStream_t)stream>>>(a);
    CUDA_TRY(cudaGetLastError());
    return SPSG_OK;
}

int spsg_raycast_forward_loss(const spsg_raycast_params *, int32_t *, const int64_t *, const float *, const float *,
                              const float *, const float *, const float *, const float *, float *, float *, float *,
                              float *, int32_t *, int32_t *, const spsg_loss_targets *, float *, void *, size_t,
                              void *) {
    return fail(SPSG_ERR_INVALID_ARGUMENT, "spsg_raycast_forward_loss: not built yet");
}

int spsg_raycast_backward_loss(const spsg_raycast_params *, const float *, const float *, const float *,
                               const spsg_loss_targets *, const float *, float, const int32_t *, const int32_t *,
                               const int32_t *, float *, float *, float *, float *, void *) {
    return fail(SPSG_ERR_INVALID_ARGUMENT, "spsg_raycast_backward_loss: not built yet");
}

}  // extern "C"
