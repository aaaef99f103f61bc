// spsg_raycast.cu -- hand-written sm_100a kernels + C ABI of the SPSG-semantic raycaster.
//
// Functional contract: reference torch/utils/raycast_rgbd/raycast_rgbd_cuda_kernel.cu (cited per function
// as kernel.cu:<line>) evaluated with the exact fp32 operation order of its sm_100 SASS (SURVEY.md
// section 3.5; re-derived from `cuobjdump -sass oracle/_ref/*.so`).  Every rounding that can change a hit
// mask is spelled with __f*_rn intrinsics so nvcc can neither contract nor re-associate it.
//
// Design (not a port): see DESIGN.md.  In short
//   * the march replays the reference's `ray += inc` running sum bit-exactly but jumps over samples whose
//     outcome is known without evaluating them: samples outside the grid, samples inside an aligned 4/8/16/32-
//     voxel region without any valid sample cell ("empty"), and samples inside a region whose valid cells all
//     have 8 strictly positive (or all strictly negative) corners ("sign-uniform": no crossing can start there);
//     jumps use a closed form of the fp32 recurrence that is exact inside one binade;
//   * one byte per 4^3 block (region kind + size, staged in shared memory) and one byte per cell (class of the
//     cell's 8 corners) answer "does this sample need arithmetic"; SDF values come from a dense fp32 brick
//     (NaN = absent voxel), so a sample that does need arithmetic costs 8 independent loads, not 16 dependent ones;
//   * the regula-falsi refinement is deferred until every lane of the warp has found its crossing;
//   * rendered pixels are staged in shared memory and written with coalesced 128-bit stores;
//   * the forward appends every (voxel, view) pair that received a pixel to a list, so the backward is one launch:
//     a deterministic per-voxel gather over that list (no float atomics for one view per chunk, no 524288-block
//     launch, no 164 MB memsets), optionally fused with the 2D losses.
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <vector>

#include "spsg_raycast.h"

namespace {

constexpr int kFineLog2 = 2;             // finest region: 4^3 voxels ("block")
constexpr int kFine = 1 << kFineLog2;
constexpr int kSuper = 8;                // hierarchy kernel handles 8^3 blocks = 32^3 voxels per CTA
constexpr float kBoxEps = 1.0f / 64.0f;  // shrink of skip regions; >> every fp32 error term
constexpr float kFracGuard = 1.0f / 256.0f;  // fast corner path needs frac(p) in [guard, 1-guard]
constexpr int kMaxFastDim = 8192;        // fast corner path proven for coordinates < 2^13
constexpr int kLossSlots = 64;           // copies of the loss accumulators (spreads atomic contention)

// block map byte = kind << 3 | level; level k >= 1: the aligned region of edge 2^(k+1) voxels around the block
enum { kKindSurface = 0, kKindEmpty = 1, kKindPos = 2, kKindNeg = 3 };

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

// Optional per-kernel timing (bench.py's roofline leg): CUDA event pairs recorded on the launch stream around the
// two dominant kernels.  Off by default; not usable during stream capture.
struct EventPair { cudaEvent_t a, b; };
bool g_timing = false;
std::vector<EventPair> g_ev[2];  // 0 = raycast_forward_kernel, 1 = backward_gather_kernel
std::mutex g_timing_mu;

struct ScopedKernelTimer {
    int which; cudaStream_t st; EventPair ev; bool on;
    ScopedKernelTimer(int w, cudaStream_t s) : which(w), st(s), on(g_timing) {
        if (on) {
            on = cudaEventCreate(&ev.a) == cudaSuccess && cudaEventCreate(&ev.b) == cudaSuccess;
            if (on) cudaEventRecord(ev.a, st);
        }
    }
    ~ScopedKernelTimer() {
        if (on) {
            cudaEventRecord(ev.b, st);
            std::lock_guard<std::mutex> lk(g_timing_mu);
            g_ev[which].push_back(ev);
        }
    }
};

// Workspace layout (caller-owned scratch, see spsg_workspace_bytes).  Written by the forward; the list, and the
// voxel->pixel tables it indexes, are what the backward of the same call pair reads.
struct Layout {
    int nbx, nby, nbz;              // 4^3 blocks per axis
    size_t bpc;                     // block-map bytes per chunk (nbx*nby*nbz rounded up to 16)
    size_t dense_off, dense_bytes;  // f32 [B][Dz][Dy][Dx], NaN = absent
    int wpr;                        // 32-cell words per x row of the cell-class bit planes
    size_t vpc;                     // uint2 words per chunk of the cell-class map (Dz*Dy*wpr rounded up to even)
    size_t vbit_off, vbit_bytes;    // uint2 [B][vpc]: bit x of (.x, .y) = class of the sample cell whose corner (0,0,0)
                                    // is the voxel: 00 invalid, 10 positive, 01 negative, 11 mixed
    size_t bmap_off, bmap_bytes;    // u8  [B][bpc] block map
    size_t marks_off, marks_bytes;  // u8  [B][bpc]: region bits of each 4^3 block (1 positive, 2 negative, 4 mixed cell)
    size_t zero_off, zero_bytes;    // everything from here to the list is cleared by the fill kernel of every forward
    size_t head_off;                // int32 list counter (256 B)
    size_t tiles_off, tiles_bytes;  // int32 [B] dynamic tile counters of the forward
    size_t arrive_off, arrive_bytes;  // int32 [B][super blocks]: classifier warps done per 32^3 super block
    size_t loss_off, loss_bytes;    // double[kLossSlots][8] loss accumulators
    size_t list_off, list_bytes;    // int2 (voxel, image) per (voxel, view) pair that received a pixel
    size_t hits_off, hits_bytes;    // optional int32 per-pixel hit voxel
    size_t total;
};

Layout make_layout(const spsg_raycast_params *p) {
    Layout L;
    L.nbx = (p->dimx + kFine - 1) >> kFineLog2;
    L.nby = (p->dimy + kFine - 1) >> kFineLog2;
    L.nbz = (p->dimz + kFine - 1) >> kFineLog2;
    L.bpc = align_up((size_t)L.nbx * L.nby * L.nbz, 16);
    const int F = p->views_per_chunk > 0 ? p->views_per_chunk : 1;
    const size_t cells = (size_t)p->num_chunks * p->dimz * p->dimy * p->dimx;
    size_t off = 0;
    L.dense_off = off;
    L.dense_bytes = align_up(cells * sizeof(float), 256);
    off += L.dense_bytes;
    L.wpr = (p->dimx + 31) / 32;
    L.vpc = align_up((size_t)p->dimz * p->dimy * L.wpr, 2);
    L.vbit_off = off;
    L.vbit_bytes = align_up((size_t)p->num_chunks * L.vpc * sizeof(uint2), 256);
    off += L.vbit_bytes;
    L.bmap_off = off;
    L.bmap_bytes = align_up((size_t)p->num_chunks * L.bpc, 256);
    off += L.bmap_bytes;
    L.marks_off = off;
    L.marks_bytes = align_up((size_t)p->num_chunks * L.bpc, 256);
    off += L.marks_bytes;
    L.zero_off = off;
    L.head_off = off;
    off += 256;
    L.tiles_off = off;
    L.tiles_bytes = align_up((size_t)p->num_chunks * sizeof(int32_t), 256);
    off += L.tiles_bytes;
    L.arrive_off = off;
    L.arrive_bytes = align_up((size_t)p->num_chunks * ((L.nbz + 7) / 8) * ((L.nby + 7) / 8) * L.wpr * sizeof(int32_t), 256);
    off += L.arrive_bytes;
    L.loss_off = off;
    L.loss_bytes = align_up((size_t)kLossSlots * 8 * sizeof(double), 256);
    off += L.loss_bytes;
    L.zero_bytes = off - L.zero_off;
    L.list_off = off;
    L.list_bytes = align_up((size_t)(p->num_locs > 0 ? p->num_locs : 0) * F * 2 * sizeof(int32_t), 256);
    off += L.list_bytes;
    L.hits_off = off;
    L.hits_bytes = align_up((size_t)p->num_chunks * F * p->width * p->height * sizeof(int32_t), 256);
    off += L.hits_bytes;
    L.total = off;
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

__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

struct Ray {
    float camx, camy, camz;
    float dx, dy, dz;
    float d2r, t0, t1;
};

// kernel.cu:287-293 + :72-85 + :194-197, cutil_math.h:1207, cuda_SimpleMatrixUtil.h:888-907.
__device__ __forceinline__ Ray setup_ray(const float *__restrict__ M, const float *__restrict__ K, unsigned ux,
                                         unsigned uy, float dmin, float dmax) {
    const float4 k4 = __ldg(reinterpret_cast<const float4 *>(K));  // fx, fy, mx, my
    const float xn = __fdiv_rn(__fadd_rn((float)ux, -k4.z), k4.x);
    const float yn = __fdiv_rn(__fadd_rn((float)uy, -k4.w), k4.y);
    const float zc = __fadd_rn(__fadd_rn(dmax, -dmin), dmin);
    const float vx = __fmul_rn(xn, zc), vy = __fmul_rn(yn, zc);
    float r = rsqrtf(__fmaf_rn(zc, zc, __fmaf_rn(vx, vx, __fmul_rn(vy, vy))));
    const float cx = __fmul_rn(vx, r), cy = __fmul_rn(vy, r), cz = __fmul_rn(r, zc);
    const float4 r0 = __ldg(reinterpret_cast<const float4 *>(M)), r1 = __ldg(reinterpret_cast<const float4 *>(M) + 1),
                 r2 = __ldg(reinterpret_cast<const float4 *>(M) + 2);
    Ray o;
    o.camx = __fadd_rn(r0.w, __fmaf_rn(0.0f, r0.z, __fmaf_rn(0.0f, r0.x, __fmul_rn(0.0f, r0.y))));
    o.camy = __fadd_rn(r1.w, __fmaf_rn(0.0f, r1.z, __fmaf_rn(0.0f, r1.x, __fmul_rn(0.0f, r1.y))));
    o.camz = __fadd_rn(r2.w, __fmaf_rn(0.0f, r2.z, __fmaf_rn(0.0f, r2.x, __fmul_rn(0.0f, r2.y))));
    const float wx = __fmaf_rn(0.0f, r0.w, __fmaf_rn(r0.z, cz, __fmaf_rn(r0.x, cx, __fmul_rn(r0.y, cy))));
    const float wy = __fmaf_rn(0.0f, r1.w, __fmaf_rn(r1.z, cz, __fmaf_rn(r1.x, cx, __fmul_rn(r1.y, cy))));
    const float wz = __fmaf_rn(0.0f, r2.w, __fmaf_rn(r2.z, cz, __fmaf_rn(r2.x, cx, __fmul_rn(r2.y, cy))));
    r = rsqrtf(__fmaf_rn(wz, wz, __fmaf_rn(wx, wx, __fmul_rn(wy, wy))));
    o.dx = __fmul_rn(wx, r);
    o.dy = __fmul_rn(wy, r);
    o.dz = __fmul_rn(wz, r);
    o.d2r = __frcp_rn(cz);
    o.t0 = __fmul_rn(o.d2r, dmin);
    o.t1 = __fmul_rn(o.d2r, dmax);
    return o;
}

// The reference's running sum `ray = ray + inc` (kernel.cu:257,260), advanced by many steps at once.
// Bit-exact: inside a binade [2^e, 2^(e+1)) every partial sum is a multiple of u = 2^(e-23), so
// fl(s + inc) = s + d with d = inc rounded to the u grid -- a constant as long as inc is not an exact tie
// between two grid points -- and s + j*d is representable, so one fma reproduces j sequential adds as long
// as every partial sum stays below 2^(e+1) - inc.  Anything irregular falls back to real adds.
struct Stepper {
    float inc, inv_inc;
    float lo, hi, lim, d, inv_d;  // current binade [lo, hi = 2lo); closed form usable while ray < lim
    bool regular;

    __device__ __forceinline__ void init(float inc_) {
        inc = inc_;
        inv_inc = rcp_approx(inc_);
        lo = 0.0f; hi = 0.0f; lim = 0.0f; d = inc_; inv_d = inv_inc; regular = false;
    }
    __device__ __forceinline__ void rebin(float ray) {
        lo = __uint_as_float(__float_as_uint(ray) & 0x7f800000u);      // 2^e <= ray
        hi = __fmul_rn(lo, 2.0f);
        const float u = __fmul_rn(lo, 1.1920928955078125e-07f);         // 2^(e-23)
        d = __fadd_rn(__fadd_rn(lo, inc), -lo);                         // inc on the u grid
        const float rem = __fadd_rn(inc, -d);                           // exact remainder
        regular = (lo >= 1.0f) && (lo <= 8388608.0f) && (inc > 0.0f) && (inc <= 0.25f * lo) && (d > 0.0f) &&
                  (__fmul_rn(fabsf(rem), 2.0f) != u);
        lim = __fadd_rn(hi, -__fmul_rn(inc, 2.0f));                    // partial sums must stay below 2lo - inc
        inv_d = rcp_approx(d);
    }
    // exactly n >= 1 steps of `ray = ray + inc`
    __device__ __forceinline__ float advance(float ray, int n) {
        for (;;) {
            if (n <= 2) {
                ray = __fadd_rn(ray, inc);
                if (n == 2) ray = __fadd_rn(ray, inc);
                return ray;
            }
            if (!(ray >= lo && ray < hi)) rebin(ray);
            int j = 0;
            if (regular) {
                // floor((lim - ray)/d) computed approximately; the slack inc + d in `lim` dwarfs the error
                const float room = lim - ray;
                j = (room > 0.0f) ? min(n, __float2int_rd(room * inv_d)) : 0;
            }
            if (j >= 1) {
                ray = __fmaf_rn((float)j, d, ray);
                n -= j;
                if (n == 0) return ray;
            } else {  // top of the binade (the add that crosses it rounds on the next grid), or an irregular binade
                ray = __fadd_rn(ray, inc);
                n -= 1;
            }
        }
    }
};

// The same recurrence with the per-binade constants (they depend on inc only) tabulated once per CTA in shared
// memory: entry e describes the binade [2^e, 2^(e+1)) as (d, 1/d, lim, -); lim = -inf marks a binade where the closed
// form is not usable (entry 32 serves every ray parameter outside [1, 2^32)).
constexpr int kStepEntries = 33;

__device__ __forceinline__ void step_table_fill(float4 *table, int e, float inc) {
    float4 t = make_float4(inc, 0.0f, -CUDART_INF_F, 0.0f);
    if (e < 32) {
        const float lo = __uint_as_float((unsigned)(e + 127) << 23), hi = __fmul_rn(lo, 2.0f);
        const float u = __fmul_rn(lo, 1.1920928955078125e-07f);  // 2^(e-23)
        const float d = __fadd_rn(__fadd_rn(lo, inc), -lo);       // inc on the u grid
        const float rem = __fadd_rn(inc, -d);                     // exact remainder
        const bool regular = (lo <= 8388608.0f) && (inc > 0.0f) && (inc <= 0.25f * lo) && (d > 0.0f) &&
                             (__fmul_rn(fabsf(rem), 2.0f) != u);
        if (regular) t = make_float4(d, rcp_approx(d), __fadd_rn(hi, -__fmul_rn(inc, 2.0f)), 0.0f);
    }
    table[e] = t;
}

// exactly n >= 1 steps of `ray = ray + inc` (ray >= 0)
__device__ __forceinline__ float step_advance(const float4 *table, float inc, float ray, int n) {
    for (;;) {
        if (n <= 2) {
            ray = __fadd_rn(ray, inc);
            if (n == 2) ray = __fadd_rn(ray, inc);
            return ray;
        }
        const unsigned e = (__float_as_uint(ray) >> 23) - 127u;
        const float4 t = table[min(e, 32u)];
        // floor((lim - ray)/d) computed approximately; the slack inc + d in `lim` dwarfs the error
        const float room = t.z - ray;
        const int j = (room > 0.0f) ? min(n, __float2int_rd(room * t.y)) : 0;
        if (j >= 1) {
            ray = __fmaf_rn((float)j, t.x, ray);
            n -= j;
            if (n == 0) return ray;
        }
        // one real add: the next step of an irregular binade, or the one that crosses the top of this binade
        ray = __fadd_rn(ray, inc);
        if (--n == 0) return ray;
    }
}

struct Volume {
    const int32_t *__restrict__ index;  // this chunk's slice of sparse_mapping
    const float *__restrict__ sdf;      // vals_sdf
    const float *__restrict__ dense;    // this chunk's slice of the dense SDF brick (NaN = absent)
    int dimx, dimy, dimz;
    float guard;                        // fast corner path needs frac(p) in [guard, 1 - guard] (see frac_guard)
};

__device__ __forceinline__ bool in_grid(const Volume &v, int x, int y, int z) {
    return (x | y | z) >= 0 && x < v.dimx && y < v.dimy && z < v.dimz;
}

// trilinear weights and accumulation in the reference's exact product / fma order (kernel.cu:132-153).
__device__ __forceinline__ float trilerp(float wx, float wy, float wz, float v000, float v100, float v010, float v001,
                                         float v110, float v011, float v101, float v111) {
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
    return d;
}

// trilinearInterpolationSimpleFastFast (kernel.cu:120-156) without the payload: the exact, fully general
// evaluation (corner coordinates rounded like the reference, index -> value double gather).
__device__ __noinline__ bool sample_sdf_exact(const Volume &v, float px, float py, float pz, float &dist) {
    const float qx = __fadd_rn(px, -0.5f), qy = __fadd_rn(py, -0.5f), qz = __fadd_rn(pz, -0.5f);
    const int x0 = round_voxel(qx), y0 = round_voxel(qy), z0 = round_voxel(qz);
    const int x1 = round_voxel(__fadd_rn(qx, 1.0f)), y1 = round_voxel(__fadd_rn(qy, 1.0f)),
              z1 = round_voxel(__fadd_rn(qz, 1.0f));
    if (!(in_grid(v, x0, y0, z0) && in_grid(v, x1, y1, z1))) return false;
    const int r00 = (z0 * v.dimy + y0) * v.dimx, r10 = (z0 * v.dimy + y1) * v.dimx;
    const int r01 = (z1 * v.dimy + y0) * v.dimx, r11 = (z1 * v.dimy + y1) * v.dimx;
    const int i000 = __ldg(v.index + r00 + x0), i100 = __ldg(v.index + r00 + x1);
    const int i010 = __ldg(v.index + r10 + x0), i110 = __ldg(v.index + r10 + x1);
    const int i001 = __ldg(v.index + r01 + x0), i101 = __ldg(v.index + r01 + x1);
    const int i011 = __ldg(v.index + r11 + x0), i111 = __ldg(v.index + r11 + x1);
    if ((i000 | i100 | i010 | i110 | i001 | i101 | i011 | i111) < 0) return false;
    const float wx = __fadd_rn(px, -floorf(px)), wy = __fadd_rn(py, -floorf(py)), wz = __fadd_rn(pz, -floorf(pz));
    dist = trilerp(wx, wy, wz, __ldg(v.sdf + i000), __ldg(v.sdf + i100), __ldg(v.sdf + i010), __ldg(v.sdf + i001),
                   __ldg(v.sdf + i110), __ldg(v.sdf + i011), __ldg(v.sdf + i101), __ldg(v.sdf + i111));
    return true;
}

// Same result as sample_sdf_exact.  Fast path: when frac(p) is at least `guard` away from 0 and 1 on every axis and
// 0 <= floor(p), floor(p)+1 < dim, the reference's rounded corner coordinates are exactly floor(p) and floor(p)+1
// ("corner coordinates"), and the 8 values come straight from the dense brick where an absent corner is
// NaN, which the fma chain propagates: valid <=> dist is not NaN.  (A present voxel holding NaN, or inf * 0, makes the
// reference's sample "valid with NaN distance", which can never satisfy the sign test and leaves the same march state
// as an invalid sample -- observationally identical.)
//
// The guard.  For p >= 1 (below 2^23) q = p - 0.5 and q + 0.5 = p are exact in fp32, so corner 0 is trunc(p) = floor(p)
// whatever frac(p) is; corner 1 = trunc(fl(fl(q + 1) + 0.5)) accumulates at most two roundings of at most ulp(2p), so it
// is floor(p) + 1 as soon as frac(p) is 8 ulp(p) away from 0 and 1: guard = 8 ulp(largest coordinate).  In the first
// voxel layer (p < 1 on some axis) q is negative and p - 0.5 is no longer exact: there the guard is kFracGuard.
__device__ __forceinline__ float frac_guard(float guard, int ix, int iy, int iz) {
    return (((ix - 1) | (iy - 1) | (iz - 1)) < 0) ? kFracGuard : guard;
}

__device__ __forceinline__ float sample_dense(const Volume &v, int ix, int iy, int iz, float wx, float wy, float wz) {
    const float *__restrict__ b = v.dense + ((size_t)iz * v.dimy + iy) * v.dimx + ix;
    const int sy = v.dimx, sz = v.dimx * v.dimy;
    const float v000 = __ldg(b), v100 = __ldg(b + 1), v010 = __ldg(b + sy), v110 = __ldg(b + sy + 1);
    const float v001 = __ldg(b + sz), v101 = __ldg(b + sz + 1), v011 = __ldg(b + sz + sy), v111 = __ldg(b + sz + sy + 1);
    return trilerp(wx, wy, wz, v000, v100, v010, v001, v110, v011, v101, v111);
}

__device__ __forceinline__ bool sample_sdf(const Volume &v, bool fast_ok, float px, float py, float pz, float &dist) {
    const float fx = floorf(px), fy = floorf(py), fz = floorf(pz);
    const float wx = __fadd_rn(px, -fx), wy = __fadd_rn(py, -fy), wz = __fadd_rn(pz, -fz);
    const int ix = __float2int_rz(fx), iy = __float2int_rz(fy), iz = __float2int_rz(fz);
    const float g = frac_guard(v.guard, ix, iy, iz);
    const bool fast = fast_ok && fminf(wx, fminf(wy, wz)) >= g && fmaxf(wx, fmaxf(wy, wz)) <= 1.0f - g &&
                      (ix | iy | iz) >= 0 && ix + 1 < v.dimx && iy + 1 < v.dimy && iz + 1 < v.dimz;
    if (fast) {
        dist = sample_dense(v, ix, iy, iz, wx, wy, wz);
        return dist == dist;
    }
    return sample_sdf_exact(v, px, py, pz, dist);
}

// ---------------------------------------------------------------------------------------------
// per-call preparation: fill, index + dense brick, cell classes, block map
// ---------------------------------------------------------------------------------------------

// One launch instead of the reference's memsets (kernel.cu:475,483,515 and, when the gradient buffers are handed to
// the forward, :557-560): up to kFillRegions word-filled regions.
constexpr int kFillRegions = 7;
struct FillArgs {
    uint32_t *ptr[kFillRegions];
    size_t words[kFillRegions];
    uint32_t value[kFillRegions];
};

__global__ void __launch_bounds__(256) fill_kernel(const FillArgs a) {
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
#pragma unroll
    for (int r = 0; r < kFillRegions; r++) {
        uint32_t *p = a.ptr[r];
        const size_t n = a.words[r];
        if (!p || n == 0) continue;
        const uint32_t v = a.value[r];
        if ((reinterpret_cast<uintptr_t>(p) & 15u) == 0) {
            const size_t n4 = n >> 2;
            uint4 *p4 = reinterpret_cast<uint4 *>(p);
            for (size_t i = tid; i < n4; i += stride) p4[i] = make_uint4(v, v, v, v);
            for (size_t i = (n4 << 2) + tid; i < n; i += stride) p[i] = v;
        } else {
            for (size_t i = tid; i < n; i += stride) p[i] = v;
        }
    }
}

// construct_dense_sparse_mapping_kernel (kernel.cu:346-362) + dense SDF scatter + voxel->pixel counter reset,
// one pass over locs.
template <bool kWriteIndex>
__global__ void __launch_bounds__(256) index_kernel(const longlong4 *__restrict__ locs, long long n,
                                                    int32_t *__restrict__ sparse_mapping,
                                                    const float *__restrict__ vals_sdf, float *__restrict__ dense,
                                                    int32_t *__restrict__ num, int views, int dimz, int dimy,
                                                    int dimx) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const longlong4 l = locs[i];  // (z, y, x, chunk)
    const long long z = l.x, y = l.y, x = l.z, b = l.w;
    const long long cell = ((b * dimz + z) * dimy + y) * dimx + x;
    if (kWriteIndex) sparse_mapping[cell] = (int32_t)i;
    if (dense) dense[cell] = __ldg(vals_sdf + i);
    if (num)
        for (int f = 0; f < views; f++) num[(long long)f * n + i] = 0;
}

// Block map.  bits of a region = OR over its cells of {1: positive cell, 2: negative cell, 4: mixed cell}.  A region
// is sign-uniform when it has no mixed cell and not both signs; empty when it has no valid cell at all.  Every 4^3
// block gets the largest aligned region (edge 4, 8, 16, 32 = level 1..4) around it that is uniform:
//   empty     if that region is empty, or no larger than the largest empty region around the block (one jump);
//   positive / negative otherwise (two events: jump to the region's last sample, then step out);
//   surface   (byte 0) if even the block itself is not uniform: samples there are classified cell by cell.
__host__ __device__ constexpr uint8_t block_map_byte(int r1, int r2, int r3, int r4) {
    const int r[5] = {0, r1, r2, r3, r4};
    int lu = 0, le = 0;
    for (int l = 1; l <= 4; l++) {
        if (!(r[l] & 4) && (r[l] & 3) != 3) lu = l;
        if (r[l] == 0) le = l;
    }
    if (lu == 0) return 0;
    const int kind = (le == lu) ? kKindEmpty : (r[lu] & 1) ? kKindPos : kKindNeg;
    return (uint8_t)((kind << 3) | lu);
}

// the same function as a table over the four 3-bit region words (r1 | r2 << 3 | r3 << 6 | r4 << 9)
struct BlockLut { uint8_t v[4096]; };
constexpr BlockLut make_block_lut() {
    BlockLut t{};
    for (int i = 0; i < 4096; i++) t.v[i] = block_map_byte(i & 7, (i >> 3) & 7, (i >> 6) & 7, (i >> 9) & 7);
    return t;
}
__device__ const BlockLut kBlockLut = make_block_lut();

// Cell classes.  For the cell c = (x, y, z) look at the 8 voxels (x..x+1, y..y+1, z..z+1), the corners of every
// sample whose corner (0,0,0) is c (kernel.cu:131-153):
//   invalid  some corner absent or outside the grid: such a sample is invalid;
//   positive all present and in (kTiny, kHuge): the sample is valid and its trilinear value is > 0 -- every weight is
//            >= 0, they sum to ~1 so one is >= 1/8, and products with values above kTiny cannot underflow;
//   negative likewise with all corners in (-kHuge, -kTiny): value < 0;
//   mixed    all present, anything else: the value has to be computed.
// Two bit planes per 32 cells of an x row (see Layout).  One warp per (4 y) x (4 z) x (32 x) slab, i.e. per run of
// eight 4^3 blocks: it reads the 5 x 5 voxel rows once (all loads in flight together), emits the 16 class words and
// -- being the only writer of those blocks -- their region bits (block holds a positive / negative / mixed cell).
// grid = (ceil(nby*wpr / 4), nbz, B), block = 128.
constexpr float kTiny = 1e-30f, kHuge = 3e38f;

__global__ void __launch_bounds__(128) cell_class_kernel(const float *__restrict__ dense, uint2 *__restrict__ vbits,
                                                         size_t vpc, uint8_t *__restrict__ marks,
                                                         int dimz, int dimy, int dimx, int wpr, int nby, int nbx,
                                                         size_t bpc, uint8_t *__restrict__ bmap, int32_t *__restrict__ arrive,
                                                         int nbz) {
    const unsigned kFull = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int w = blockIdx.x * 4 + (threadIdx.x >> 5);  // (block row in y, xw) of this warp's slab
    if (w >= nby * wpr) return;
    const int yb = w / wpr, xw = w - yb * wpr;
    const int zb = blockIdx.y, chunk = blockIdx.z;
    const int y0 = yb * kFine, z0 = zb * kFine;
    const float *__restrict__ base = dense + (size_t)chunk * dimz * dimy * dimx;
    const int x = xw * 32 + lane, xn = xw * 32 + 32;
    // voxel rows (y0..y0+4, z0..z0+4): value of this lane's voxel, and of the first voxel of the next word for lane 0
    float val[5][5], nxt[5][5];
#pragma unroll
    for (int dz = 0; dz < 5; dz++)
#pragma unroll
        for (int dy = 0; dy < 5; dy++) {
            const int y = y0 + dy, z = z0 + dz;
            const bool row = y < dimy && z < dimz;
            const size_t o = ((size_t)z * dimy + y) * dimx;
            val[dz][dy] = (row && x < dimx) ? __ldg(base + o + x) : CUDART_NAN_F;
            nxt[dz][dy] = (row && lane == 0 && xn < dimx) ? __ldg(base + o + xn) : CUDART_NAN_F;
        }
    // per voxel row: present / positive-class / negative-class masks over x, shifted so that bit x also covers x+1
    unsigned pres[5][5], posm[5][5], negm[5][5];
#pragma unroll
    for (int dz = 0; dz < 5; dz++)
#pragma unroll
        for (int dy = 0; dy < 5; dy++) {
            const float a = val[dz][dy], n = nxt[dz][dy];
            const unsigned p = __ballot_sync(kFull, a == a), pp = __ballot_sync(kFull, a > kTiny && a < kHuge),
                           pn = __ballot_sync(kFull, a < -kTiny && a > -kHuge);
            // lane 0 holds the next word's first voxel
            const unsigned np = __shfl_sync(kFull, (unsigned)(n == n), 0), npp = __shfl_sync(kFull, (unsigned)(n > kTiny && n < kHuge), 0),
                           npn = __shfl_sync(kFull, (unsigned)(n < -kTiny && n > -kHuge), 0);
            pres[dz][dy] = p & ((p >> 1) | (np << 31));
            posm[dz][dy] = pp & ((pp >> 1) | (npp << 31));
            negm[dz][dy] = pn & ((pn >> 1) | (npn << 31));
        }
    unsigned any_pos = 0u, any_neg = 0u, any_mix = 0u;
#pragma unroll
    for (int dz = 0; dz < 4; dz++)
#pragma unroll
        for (int dy = 0; dy < 4; dy++) {
            const int y = y0 + dy, z = z0 + dz;
            const unsigned v = pres[dz][dy] & pres[dz][dy + 1] & pres[dz + 1][dy] & pres[dz + 1][dy + 1];
            const unsigned vp = posm[dz][dy] & posm[dz][dy + 1] & posm[dz + 1][dy] & posm[dz + 1][dy + 1];
            const unsigned vn = negm[dz][dy] & negm[dz][dy + 1] & negm[dz + 1][dy] & negm[dz + 1][dy + 1];
            any_pos |= vp; any_neg |= vn; any_mix |= v & ~vp & ~vn;
            if (lane == 0 && y < dimy && z < dimz)
                vbits[(size_t)chunk * vpc + ((size_t)z * dimy + y) * wpr + xw] = make_uint2(v & ~vn, v & ~vp);
        }
    if (lane < 8) {  // one block per lane: region bits 1 = holds a positive cell, 2 = negative, 4 = mixed
        const int bx = xw * 8 + lane;
        if (bx < nbx)
            marks[(size_t)chunk * bpc + ((size_t)zb * nby + yb) * nbx + bx] =
                (uint8_t)((((any_pos >> (4 * lane)) & 0xfu) ? 1 : 0) | (((any_neg >> (4 * lane)) & 0xfu) ? 2 : 0) |
                          (((any_mix >> (4 * lane)) & 0xfu) ? 4 : 0));
    }
    // ---- block map of the 32^3 super block (8 x 8 slabs of this x word) by whichever of its warps finishes last
    const int sby = (nby + 7) >> 3, sbz = (nbz + 7) >> 3;
    const int sy = yb >> 3, sz = zb >> 3;
    const int rows_y = min(8, nby - sy * 8), rows_z = min(8, nbz - sz * 8);
    __threadfence();  // this warp's region bits are visible before it is counted
    int prev = 0;
    if (lane == 0) prev = atomicAdd(arrive + ((size_t)chunk * sbz + sz) * sby * wpr + (size_t)sy * wpr + xw, 1);
    prev = __shfl_sync(kFull, prev, 0);
    if (prev != rows_y * rows_z - 1) return;
    __threadfence();
    // lane = zl * 4 + (yl >> 1) owns the two slab rows (zl, yl), (zl, yl + 1), yl even: eight region bytes each
    const int zl = lane >> 2, yl = (lane & 3) * 2;
    uint32_t rb[2][2] = {{0u, 0u}, {0u, 0u}};  // [row][x half]: four blocks per word
    const int nx = min(8, nbx - xw * 8);
#pragma unroll
    for (int r = 0; r < 2; r++) {
        const int gy = sy * 8 + yl + r, gz = sz * 8 + zl;
        if (gy < nby && gz < nbz) {
            const uint8_t *src = marks + (size_t)chunk * bpc + ((size_t)gz * nby + gy) * nbx + xw * 8;
            if ((nbx & 7) == 0) {  // rows are 8-byte aligned: one load
                const uint2 v2 = __ldcg(reinterpret_cast<const uint2 *>(src));
                rb[r][0] = v2.x; rb[r][1] = v2.y;
            } else {
                for (int k = 0; k < nx; k++) rb[r][k >> 2] |= (uint32_t)__ldcg(src + k) << (8 * (k & 3));
            }
        }
    }
    // region bits per level, byte-parallel.  8^3: x pairs, the lane's two rows, z neighbour (lane ^ 4)
    auto xpair = [](uint32_t v) { const uint32_t t = v | ((v >> 8) & 0x00ff00ffu); return (t & 0x00ff00ffu) | ((t & 0x00ff00ffu) << 8); };
    uint32_t r8[2];
#pragma unroll
    for (int h = 0; h < 2; h++) {
        uint32_t v = xpair(rb[0][h] | rb[1][h]);
        v |= __shfl_xor_sync(kFull, v, 4);
        r8[h] = v;  // every byte: bits of the 8^3 region of that block
    }
    // 16^3: x quad (all four bytes of a half), y quad (lane ^ 1), z quad (lane ^ 4 already in r8, plus lane ^ 8)
    uint32_t r16[2];
#pragma unroll
    for (int h = 0; h < 2; h++) {
        uint32_t v = r8[h];
        v |= v >> 16; v |= v >> 8; v &= 0xffu;
        v |= __shfl_xor_sync(kFull, v, 1);
        v |= __shfl_xor_sync(kFull, v, 8);
        r16[h] = v;  // one byte: bits of the 16^3 region of this half
    }
    // 32^3: both halves, all lanes
    uint32_t r32 = r16[0] | r16[1];
    r32 |= __shfl_xor_sync(kFull, r32, 2);
    r32 |= __shfl_xor_sync(kFull, r32, 16);
    const uint8_t *__restrict__ lut = kBlockLut.v;
#pragma unroll
    for (int r = 0; r < 2; r++) {
        const int gy = sy * 8 + yl + r, gz = sz * 8 + zl;
        if (gy < nby && gz < nbz) {
            uint8_t *dst = bmap + (size_t)chunk * bpc + ((size_t)gz * nby + gy) * nbx + xw * 8;
            uint32_t out[2] = {0u, 0u};
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const int h = k >> 2, sh = 8 * (k & 3);
                const uint32_t byte = __ldg(lut + (((rb[r][h] >> sh) & 7u) | (((r8[h] >> sh) & 7u) << 3) | ((r16[h] & 7u) << 6) | ((r32 & 7u) << 9)));
                out[h] |= byte << sh;
            }
            if ((nbx & 7) == 0) {
                *reinterpret_cast<uint2 *>(dst) = make_uint2(out[0], out[1]);
            } else {
                for (int k = 0; k < nx; k++) dst[k] = (uint8_t)(out[k >> 2] >> (8 * (k & 3)));
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------

#ifdef SPSG_STATS
// development build only (-DSPSG_STATS): event counters of the march, read back with spsg_debug_stats()
__device__ unsigned long long g_stats[48];
__device__ int g_tile_stats[8192][8];  // per tile: total, setup, march, refine, epilogue cycles, iterations, smid, start
#define STAT_MAX(k, v) atomicMax(&g_stats[k], (unsigned long long)(v))
#define STAT_ADD(k, v) atomicAdd(&g_stats[k], (unsigned long long)(v))
#if SPSG_STATS == 1
#define EVT_ADD(k, v) STAT_ADD(k, v)  // per-event counters (slow: perturbs timing)
#else
#define EVT_ADD(k, v)
#endif
#else
#define STAT_ADD(k, v)
#define STAT_MAX(k, v)
#endif

struct LossArgs {
    const float *target_depth, *target_color, *weight_color;
    const uint8_t *target_label;
    const float *class_weight;
    float voxelsize;
    double *accum;  // [0]=sum|d-t| [1]=#depth [2]=sum|c-t| [3]=#colour elems [4]=sum w*nll [5]=sum w
};

struct ForwardArgs {
    const int32_t *sparse_mapping;
    const float *vals_sdf, *vals_color, *vals_normal, *vals_semantic;
    const float *view_matrix, *intrinsics;
    float *image_color, *image_depth, *image_normal, *image_semantic;
    int32_t *mapping3dto2d, *mapping3dto2d_num;
    const float *dense;
    const uint2 *vbits;   // [B][vpc]
    const uint8_t *bmap;  // [B][bpc]
    size_t vpc, bpc;
    int wpr;
    int maps_in_smem;
    int32_t *tile_counter;  // [B], zeroed per call
    int32_t *list_count;
    int2 *list;
    int32_t *hits;
    int width, height;
    float depth_min, depth_max, thresh, inc;
    int dimx, dimy, dimz;
    int nbx, nby, nbz;
    int num_chunks, views, max_pixels;
    long long num_locs;
    unsigned flags;
    int vec_ok;  // image rows 16-byte aligned: float4 write-out allowed
    float guard; // see frac_guard
    LossArgs loss;
};

constexpr int kTileW = 16, kTileH = 8;  // pixels per CTA of the occupancy kernel: 4 warps of 8x4 pixels
constexpr int kTilePix = kTileW * kTileH;

constexpr int kWarpW = 8, kWarpH = 4;                   // pixels per warp tile
constexpr int kFwdWarps = 24;                           // warps of the persistent forward CTA (one CTA per SM)
constexpr int kFwdWarpsLarge = 28;                      // ... for launches with many tiles per SM (more latency hiding, a few spills)
constexpr int kFwdThreads = kFwdWarps * 32;
constexpr int kStageFloats = 14 * 32;                   // per-warp write-out staging: the widest channel group
__host__ __device__ constexpr size_t fwd_smem_fixed(int warps) { return 128 + (size_t)warps * kStageFloats * sizeof(float); }
constexpr size_t kFwdSmemFixed = fwd_smem_fixed(kFwdWarpsLarge);  // residency test uses the larger CTA
constexpr size_t kFwdSmemMax = 232448 - 1024;           // 227 KB opt-in limit per CTA, minus the static shared memory

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- TMA bulk copy global -> shared, completion on an mbarrier (sm_90+ PTX)
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void *dst, const void *src, unsigned bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned phase) {
    unsigned done;
    do {
        asm volatile(
            "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(phase)
            : "memory");
    } while (!done);
}

// Write-out of one channel group of a warp's 8x4 pixel tile: smem [kWarpH][kWarpW*C] -> global rows, by the warp.
template <int C>
__device__ __forceinline__ void store_warp_tile(const float *__restrict__ s, float *__restrict__ g, int img, int x0,
                                                int y0, int width, int height, bool vec, int lane) {
    const int rows = min(kWarpH, height - y0), cols = min(kWarpW, width - x0);
    constexpr int kRow = kWarpW * C;
    if (rows <= 0 || cols <= 0) return;
    if (vec && cols == kWarpW) {
        constexpr int kVecRow = kRow / 4;
        for (int e = lane; e < rows * kVecRow; e += 32) {
            const int r = e / kVecRow, k = e - r * kVecRow;
            float4 *dst = reinterpret_cast<float4 *>(g + ((size_t)(img * height + y0 + r) * width + x0) * C) + k;
            __stcs(dst, reinterpret_cast<const float4 *>(s + r * kRow)[k]);
        }
    } else {
        const int n = cols * C;
        for (int e = lane; e < rows * kRow; e += 32) {
            const int r = e / kRow, k = e - r * kRow;
            if (k < n) __stcs(g + ((size_t)(img * height + y0 + r) * width + x0) * C + k, s[r * kRow + k]);
        }
    }
}

// Persistent forward: one CTA per SM, one thread per ray, one 8x4-pixel tile per warp at a time.
// kernel.cu:265-297 (init + ray), :190-263 (march), :166-187 (regula falsi), :215-249 (hit write-out + voxel->pixel
// registration); kLoss adds the 2D losses (train.py:635-638, loss.py:246-257, train.py:744-746) to the epilogue.
// CTA i works on chunk i % B (then i % B + gridDim, ...): the chunk's cell-class bit planes and block map are pulled
// into shared memory once by TMA bulk copies, so the march's "does this sample need arithmetic" lookups never leave
// the SM; tiles of the chunk's images are dealt to warps first statically, then from a global counter.
template <bool kLoss, bool kSmemMaps, int kWarps>
__global__ void __launch_bounds__(kWarps * 32, 1) raycast_forward_kernel(const ForwardArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ float4 s_steps[kStepEntries];
    uint64_t *mbar = reinterpret_cast<uint64_t *>(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float *stage = reinterpret_cast<float *>(smem + 128) + warp * kStageFloats;
    uint2 *s_vbits = reinterpret_cast<uint2 *>(smem + fwd_smem_fixed(kWarps));
    uint8_t *s_bmap = reinterpret_cast<uint8_t *>(s_vbits + a.vpc);
    const unsigned kFull = 0xffffffffu;
    const float kInf = CUDART_INF_F;

    if (threadIdx.x < kStepEntries) step_table_fill(s_steps, threadIdx.x, a.inc);
    if (kSmemMaps && threadIdx.x == 0) mbar_init(mbar, 1);
    __syncthreads();
    unsigned phase = 0;
    const float inv_inc = rcp_approx(a.inc);

    const size_t cells = (size_t)a.dimz * a.dimy * a.dimx;
    const bool clip = !(a.flags & SPSG_FLAG_NO_CLIP);
    const bool skip = !(a.flags & SPSG_FLAG_NO_BRICK_SKIP);
    const bool fast_ok = max(a.dimx, max(a.dimy, a.dimz)) <= kMaxFastDim;
    const bool skip_ok = skip && fast_ok;
    // warp tiles are numbered so that the four tiles of a 16x8 pixel block are consecutive
    const int tiles_x = (a.width + kWarpW - 1) / kWarpW, tiles_y = (a.height + kWarpH - 1) / kWarpH;
    const int blocks_x = (tiles_x + 1) >> 1, blocks_y = (tiles_y + 1) >> 1;
    const int tiles_per_image = blocks_x * blocks_y * 4;
    const int total_tiles = tiles_per_image * a.views;

    for (int chunk = blockIdx.x % a.num_chunks; chunk < a.num_chunks; chunk += gridDim.x) {
        if (kSmemMaps) {
            if (threadIdx.x == 0) {  // cell classes + block map: TMA bulk copies, completion on the mbarrier
                const unsigned vb_bytes = (unsigned)(a.vpc * sizeof(uint2)), bm_bytes = (unsigned)a.bpc;
                mbar_expect_tx(mbar, vb_bytes + bm_bytes);
                const uint8_t *src = reinterpret_cast<const uint8_t *>(a.vbits + (size_t)chunk * a.vpc);
                uint8_t *dst = reinterpret_cast<uint8_t *>(s_vbits);
                for (unsigned o = 0; o < vb_bytes; o += 32768u) bulk_copy_g2s(dst + o, src + o, min(32768u, vb_bytes - o), mbar);
                bulk_copy_g2s(s_bmap, a.bmap + (size_t)chunk * a.bpc, bm_bytes, mbar);
            }
        }
        const uint2 *vbits = kSmemMaps ? s_vbits : a.vbits + (size_t)chunk * a.vpc;
        const uint8_t *bmap = kSmemMaps ? s_bmap : a.bmap + (size_t)chunk * a.bpc;
        Volume v;
        v.index = a.sparse_mapping + (size_t)chunk * cells;
        v.sdf = a.vals_sdf;
        v.dense = a.dense + (size_t)chunk * cells;
        v.dimx = a.dimx; v.dimy = a.dimy; v.dimz = a.dimz;
        v.guard = a.guard;

        // CTAs that share this chunk: ranks 0..group-1.  First round static and contiguous per CTA, then dynamic.
        const int nb = a.num_chunks;
        const int group = ((int)gridDim.x - 1 - (int)(blockIdx.x % nb)) / nb + 1, rank = blockIdx.x / nb;
        const int static_tiles = min(total_tiles, group * kWarps);
        const int per = static_tiles / group, extra = static_tiles - per * group;
        const int my_first = rank * per + min(rank, extra), my_count = per + (rank < extra ? 1 : 0);
        int tile = warp < my_count ? my_first + warp : total_tiles;
        int32_t *counter = a.tile_counter + chunk;
        if (tile >= total_tiles && static_tiles < total_tiles) {
            int t = 0;
            if (lane == 0) t = static_tiles + atomicAdd(counter, 1);
            tile = __shfl_sync(kFull, t, 0);
        }

        // Per-lane ray of the warp's current tile: set up (and clipped against the grid) before the chunk's maps are needed,
        // so that the first tile's set-up overlaps the TMA copies.
        struct TileRay {
            Ray r;
            float invx, invy, invz, kx, ky, kz;
            int sxm, sym, szm;
            float ray, t_end;
            int jump_cap;
            unsigned pix;
            size_t gpix;
            bool active, inside;
            int img, view, wx0, wy0;
#ifdef SPSG_STATS
            long long clk0;
#endif
        };
        auto prepare = [&](int tile, TileRay &q) {
            const int view = tile / tiles_per_image, tt = tile - view * tiles_per_image;
            const int blk = tt >> 2, sub = tt & 3;
            const int by = blk / blocks_x, bx = blk - by * blocks_x;
            const int wx0 = (bx * 2 + (sub & 1)) * kWarpW, wy0 = (by * 2 + (sub >> 1)) * kWarpH;
            const int img = chunk * a.views + view;
            q.view = view; q.img = img; q.wx0 = wx0; q.wy0 = wy0;
            q.inside = wx0 < a.width && wy0 < a.height;
            if (q.inside) {
                const unsigned ux = wx0 + (lane & 7), uy = wy0 + (lane >> 3);
                const bool active = ux < (unsigned)a.width && uy < (unsigned)a.height;
                const unsigned pix = uy * a.width + ux;
                const size_t gpix = (size_t)img * a.width * a.height + pix;

#ifdef SPSG_STATS
                q.clk0 = clock64();
#endif
                // Lanes outside the image run the same loops below with an exhausted ray.
                const Ray r = setup_ray(a.view_matrix + (size_t)img * 16, a.intrinsics + (size_t)img * 4,
                                        active ? ux : 0u, active ? uy : 0u, a.depth_min, a.depth_max);
                // approximate reciprocals are only used to size jumps; every margin below dwarfs their error
                const float invx = r.dx != 0.0f ? rcp_approx(r.dx) : 0.0f, invy = r.dy != 0.0f ? rcp_approx(r.dy) : 0.0f,
                            invz = r.dz != 0.0f ? rcp_approx(r.dz) : 0.0f;
                // exit-plane constants of the region jumps: t = (face -+ kBoxEps - cam) / dir, +inf for an axis-parallel ray
                const float kx = r.dx != 0.0f ? ((r.dx > 0.0f ? -kBoxEps : kBoxEps) - r.camx) * invx : kInf;
                const float ky = r.dy != 0.0f ? ((r.dy > 0.0f ? -kBoxEps : kBoxEps) - r.camy) * invy : kInf;
                const float kz = r.dz != 0.0f ? ((r.dz > 0.0f ? -kBoxEps : kBoxEps) - r.camz) * invz : kInf;
                const int sxm = r.dx > 0.0f ? -1 : 0, sym = r.dy > 0.0f ? -1 : 0, szm = r.dz > 0.0f ? -1 : 0;
                float ray = r.t0, t_end = active ? r.t1 : -kInf;
                // A closed-form jump of j steps lands within j * ulp(ray) / 2 of ray + j * inc (Stepper): cap j so that
                // this drift stays below kBoxEps / 4, far inside the kBoxEps the skip regions are shrunk by.
                int jump_cap = 1 << 22;
                {
                    const float top = fmaxf(fabsf(r.t1), 1.0f);
                    const float ulp = __uint_as_float(__float_as_uint(top) & 0x7f800000u) * 1.1920928955078125e-07f;
                    const float cap = (0.5f * kBoxEps) / ulp;
                    jump_cap = cap < 4194304.0f ? max(1, __float2int_rd(cap)) : (1 << 22);
                }
                if (clip && active) {
                    // Samples are valid only for p in (0, dim-1) on every axis (all 8 corners inside the grid).
                    float tin = -kInf, tout = kInf;
#define SPSG_SLAB(o, d, inv, lo, hi)                                        \
    if ((d) != 0.0f) {                                                      \
        const float ta_ = ((lo) - (o)) * (inv), tb_ = ((hi) - (o)) * (inv); \
        tin = fmaxf(tin, fminf(ta_, tb_));                                  \
        tout = fminf(tout, fmaxf(ta_, tb_));                                \
    } else if ((o) < (lo) || (o) > (hi)) {                                  \
        tin = kInf;                                                         \
        tout = -kInf;                                                       \
    }
                    SPSG_SLAB(r.camx, r.dx, invx, -kBoxEps, (float)(a.dimx - 1) + kBoxEps)
                    SPSG_SLAB(r.camy, r.dy, invy, -kBoxEps, (float)(a.dimy - 1) + kBoxEps)
                    SPSG_SLAB(r.camz, r.dz, invz, -kBoxEps, (float)(a.dimz - 1) + kBoxEps)
#undef SPSG_SLAB
                    const float margin = 0.0625f;
                    if (!(tin <= tout)) {
                        t_end = -kInf;  // misses the grid: nothing to march
                    } else {
                        t_end = fminf(t_end, tout + margin);
                        // jump to (at most) the last sample before tin - margin
                        while (ray < tin - margin - a.inc && ray < t_end) {
                            const int want = max(1, min(__float2int_rd((tin - margin - ray) * inv_inc) - 1, jump_cap));
                            ray = step_advance(s_steps, a.inc, ray, want);
                        }
                    }
                }
                q.r = r; q.invx = invx; q.invy = invy; q.invz = invz; q.kx = kx; q.ky = ky; q.kz = kz;
                q.sxm = sxm; q.sym = sym; q.szm = szm; q.ray = ray; q.t_end = t_end; q.jump_cap = jump_cap;
                q.pix = pix; q.gpix = gpix; q.active = active;
            }
        };
        TileRay q;
        q.inside = false;
        if (tile < total_tiles) prepare(tile, q);
        if (kSmemMaps) mbar_wait(mbar, phase);  // the chunk's class planes and block map have landed

        while (tile < total_tiles) {
            int next = total_tiles;
            if (lane == 0 && static_tiles < total_tiles) next = static_tiles + atomicAdd(counter, 1);  // prefetched
            if (q.inside) {
                const Ray r = q.r;
                const float invx = q.invx, invy = q.invy, invz = q.invz, kx = q.kx, ky = q.ky, kz = q.kz;
                const int sxm = q.sxm, sym = q.sym, szm = q.szm, jump_cap = q.jump_cap;
                float ray = q.ray;
                const float t_end = q.t_end;
                const unsigned pix = q.pix;
                const size_t gpix = q.gpix;
                const bool active = q.active;
                const int img = q.img, view = q.view, wx0 = q.wx0, wy0 = q.wy0;
                int hit = -1;
                float depth = 0.0f;
#ifdef SPSG_STATS
                const long long clk0 = q.clk0;
                long long clk_march = 0, clk_refine = 0;
                int my_iters = 0;
                const long long clk1 = clock64(), clk1b = clk1;
#endif

                // last valid sample (kernel.cu:64-69).  "No valid last sample" is encoded as last_sdf == 0: a last value of
                // +-0 can never satisfy the strict sign test (:205) either, so the two are indistinguishable.
                // last_lazy: last_sdf is only a +-1 placeholder carrying the sign the cell class guarantees; the value
                // is computed if and when a crossing needs it.
                float last_sdf = 0.0f, last_alpha = 0.0f;
                bool last_lazy = false;
                float dist = 0.0f;
                enum { kMarch = 0, kCross = 1, kDone = 2 };
                int state = kMarch;

                for (;;) {
#ifdef SPSG_STATS
                    const long long clk_a = clock64();
#endif
                    // ---- march.  The loop is warp-synchronous: all lanes take part in every vote and every iteration
                    // handles one event per marching lane, so diverged lanes re-join at the bottom of each iteration
                    // instead of running their iterations one group after the other.  An event is either a jump over
                    // samples whose outcome is known from the block map, or one sample.
                    while (__any_sync(kFull, state == kMarch)) {
#ifdef SPSG_STATS
                        my_iters++;
#endif
                        if (state == kMarch) {
                            if (!(ray < t_end)) {  // kernel.cu:200
                                state = kDone;
                            } else {
                                enum { kActExact = 0, kActDense = 1, kActInvalid = 2, kActSign = 3, kActJumpEmpty = 4, kActJumpSame = 5 };
                                int act = kActExact, nadv = 1;
                                float sgn = 0.0f, wx = 0.0f, wy = 0.0f, wz = 0.0f;
                                const float px = __fmaf_rn(r.dx, ray, r.camx), py = __fmaf_rn(r.dy, ray, r.camy),
                                            pz = __fmaf_rn(r.dz, ray, r.camz);
                                // floor: one conversion on the address path (exact for every in-grid p; an out-of-range p
                                // saturates and fails the bounds test below)
                                const int ix = __float2int_rd(px), iy = __float2int_rd(py), iz = __float2int_rd(pz);
                                const float fx = (float)ix, fy = (float)iy, fz = (float)iz;
                                if (skip_ok && (unsigned)ix < (unsigned)a.dimx && (unsigned)iy < (unsigned)a.dimy &&
                                    (unsigned)iz < (unsigned)a.dimz) {
                                    // block map and cell class are fetched together (independent shared-memory addresses)
                                    const int b = bmap[((iz >> kFineLog2) * a.nby + (iy >> kFineLog2)) * a.nbx + (ix >> kFineLog2)];
                                    const uint2 word = vbits[(iz * a.dimy + iy) * a.wpr + (ix >> 5)];
                                    wx = __fadd_rn(px, -fx); wy = __fadd_rn(py, -fy); wz = __fadd_rn(pz, -fz);
                                    const float wlo = fminf(wx, fminf(wy, wz)), whi = fmaxf(wx, fmaxf(wy, wz));
                                    if (b != 0 && wlo >= kBoxEps && whi <= 1.0f - kBoxEps) {
                                        // p is inside an aligned uniform region of edge `size`, at least kBoxEps away from
                                        // every cell face and hence from the region's faces: corner (0,0,0) of this sample
                                        // and of every later one up to the region's (shrunk) exit lies in the region.
                                        const int kind = b >> 3, size = 2 << (b & 7), mask = ~(size - 1);
                                        // a sign-uniform region cannot be jumped while the last valid sample has the
                                        // other sign: its first valid sample would be a crossing
                                        const bool opposite = (kind == kKindPos && last_sdf < 0.0f) ||
                                                              (kind == kKindNeg && last_sdf > 0.0f);
                                        if (!opposite) {
                                            // exit face per axis = origin + (dir > 0 ? size : 0), shrunk by kBoxEps (folded
                                            // into kx/ky/kz together with the camera position)
                                            const float tx_ = __fmaf_rn((float)((ix & mask) + (size & sxm)), invx, kx);
                                            const float ty_ = __fmaf_rn((float)((iy & mask) + (size & sym)), invy, ky);
                                            const float tz_ = __fmaf_rn((float)((iz & mask) + (size & szm)), invz, kz);
                                            const float tout = fminf(tx_, fminf(ty_, tz_));
                                            // steps to the first sample beyond the region's exit
                                            const int n = max(1, min(__float2int_rd((tout - ray) * inv_inc) + 1, jump_cap));
                                            if (kind == kKindEmpty) {
                                                // every sample before that one is invalid (kernel.cu:131,259)
                                                act = kActJumpEmpty; nadv = n;
                                            } else if (n >= 2) {
                                                // Samples up to the last one inside are invalid or share the region's sign,
                                                // and so does the last valid one before them: no crossing.  Land on the
                                                // last one inside; it is classified by its own cell and leaves the march
                                                // state exactly as the reference's sample-by-sample walk would.
                                                act = kActJumpSame; nadv = n - 1;
                                            }
                                        }
                                    }
                                    if (act == kActExact) {
                                        // One sample, decided by its own cell.  With frac(p) clear of the cell faces the
                                        // reference's corners are exactly floor(p) + {0,1}; the cell's class says whether
                                        // all 8 are present and whether they share a sign.
                                        const float g = frac_guard(v.guard, ix, iy, iz);
                                        if (wlo >= g && whi <= 1.0f - g) {
                                            const unsigned ca = (word.x >> (ix & 31)) & 1u, cb = (word.y >> (ix & 31)) & 1u;
                                            if ((ca | cb) == 0u) {
                                                act = kActInvalid;
                                            } else {
                                                act = kActDense;
                                                if ((ca & cb) == 0u) {
                                                    sgn = ca ? 1.0f : -1.0f;
                                                    // opposite strict signs <=> last_sdf * (+-1) < 0 (last_sdf is never NaN)
                                                    if (!(__fmul_rn(last_sdf, sgn) < 0.0f)) act = kActSign;
                                                }
                                            }
                                        }
                                    }
                                }
                                if (act <= kActDense) {
                                    bool valid;
                                    if (act == kActDense) {
                                        dist = sample_dense(v, ix, iy, iz, wx, wy, wz);
                                        valid = dist == dist;
                                    } else {
                                        valid = sample_sdf(v, fast_ok, px, py, pz, dist);  // the reference's exact corner arithmetic
                                    }
                                    if (valid && ((last_sdf > 0.0f && dist < 0.0f) || (last_sdf < 0.0f && dist > 0.0f))) {  // :205
                                        state = kCross;
                                    } else {
                                        last_sdf = valid ? dist : 0.0f; last_alpha = ray; last_lazy = false;  // :254-256 / :259
                                    }
                                } else if (act == kActSign) {
                                    last_sdf = sgn; last_alpha = ray; last_lazy = true;  // :254-256
                                } else if (act != kActJumpSame) {
                                    last_sdf = 0.0f;  // :259 (invalid sample, or a run of them)
                                }
                                if (state == kMarch) ray = step_advance(s_steps, a.inc, ray, nadv);  // :257,:260
                            }
                        }
                    }
#ifdef SPSG_STATS
                    const long long clk_b = clock64();
                    clk_march += clk_b - clk_a;
#endif
                    // ---- refinement round: every lane is either waiting with a crossing or finished
                    if (!__any_sync(kFull, state == kCross)) break;
                    if (state == kCross) {
                        if (last_lazy) {  // the crossing needs the previous sample's value after all (its class says it is valid)
                            float dl = last_sdf;
                            if (sample_sdf(v, fast_ok, __fmaf_rn(r.dx, last_alpha, r.camx), __fmaf_rn(r.dy, last_alpha, r.camy),
                                           __fmaf_rn(r.dz, last_alpha, r.camz), dl))
                                last_sdf = dl;
                            last_lazy = false;
                        }
                        // findIntersectionBisection (:166-187)
                        float ta = last_alpha, da = last_sdf, tb = ray, db = dist, c = 0.0f;
                        float cx = 0.0f, cy = 0.0f, cz = 0.0f;
                        bool ok = true;
#pragma unroll 1
                        for (int k = 0; k < 3; k++) {
                            c = __fmaf_rn(__fadd_rn(tb, -ta), __fdiv_rn(da, __fadd_rn(da, -db)), ta);  // :161
                            cx = __fmaf_rn(r.dx, c, r.camx);
                            cy = __fmaf_rn(r.dy, c, r.camy);
                            cz = __fmaf_rn(r.dz, c, r.camz);
                            float dc;
                            if (!sample_sdf(v, fast_ok, cx, cy, cz, dc)) {
                                ok = false;
                                break;
                            }
                            if (__fmul_rn(da, dc) > 0.0f) { ta = c; da = dc; } else { tb = c; db = dc; }  // :180-181
                        }
                        if (ok && fabsf(__fadd_rn(last_sdf, -dist)) < a.thresh && fabsf(dist) < a.thresh) {  // :211-213
                            depth = __fdiv_rn(c, r.d2r);                                                     // :215
                            // payload voxel = nearest voxel of the last refinement point (:129) == hit voxel
                            // round(cam + alpha*dir) (:241-242, same fma).  It is one of the 8 present corners; if rounding
                            // ever says otherwise the reference reads stale registers -- we keep marching instead.
                            const int nx = round_voxel(cx), ny = round_voxel(cy), nz = round_voxel(cz);
                            hit = in_grid(v, nx, ny, nz) ? __ldg(v.index + ((size_t)nz * v.dimy + ny) * v.dimx + nx) : -1;
                        }
                        if (hit >= 0) {
                            state = kDone;
                        } else {
                            last_sdf = dist; last_alpha = ray; last_lazy = false;  // :254-256
                            ray = __fadd_rn(ray, a.inc);                           // :257
                            state = kMarch;
                        }
                    }
#ifdef SPSG_STATS
                    clk_refine += clock64() - clk_b;
#endif
                }
#ifdef SPSG_STATS
                const long long clk2 = clock64();
#endif

                // ---- write-out (kernel.cu:276-285 init, :217-239 hit) through shared memory
                const float ninf = __int_as_float(0xff800000);
                float col0 = ninf, col1 = ninf, col2 = ninf, dep = ninf;
                float sem[14];
#pragma unroll
                for (int k = 0; k < 14; k++) sem[k] = ninf;
                float n0 = ninf, n1 = ninf, n2 = ninf;
                bool first = false;
                if (hit >= 0) {
                    const float *c = a.vals_color + (size_t)hit * 3, *n = a.vals_normal + (size_t)hit * 3;
                    col0 = __ldg(c + 0); col1 = __ldg(c + 1); col2 = __ldg(c + 2);
                    const float m0 = __ldg(n + 0), m1 = __ldg(n + 1), m2 = __ldg(n + 2);
                    if (!(m0 == 0.0f && m1 == 0.0f && m2 == 0.0f)) { n0 = m0; n1 = m1; n2 = m2; }  // :220
                    dep = depth;
                    const float2 *s2 = reinterpret_cast<const float2 *>(a.vals_semantic + (size_t)hit * 14);
#pragma unroll
                    for (int k = 0; k < 7; k++) {
                        const float2 t2 = __ldg(s2 + k);
                        sem[2 * k] = t2.x; sem[2 * k + 1] = t2.y;
                    }
                }
                {
                    // voxel -> pixel registration (:244-247), one atomic per distinct voxel of the warp: lanes that hit
                    // the same voxel take consecutive slots from a single atomicAdd (the reference's slot order is the
                    // arbitrary order of its per-pixel atomics)
                    const unsigned peers = __match_any_sync(kFull, hit);
                    if (hit >= 0) {
                        const int leader = __ffs(peers) - 1;
                        const size_t row = (size_t)view * (size_t)a.num_locs + (size_t)hit;
                        int base = 0;
                        if (lane == leader) base = atomicAdd(a.mapping3dto2d_num + row, __popc(peers));
                        base = __shfl_sync(peers, base, leader);
                        const int offset = base + __popc(peers & ((1u << lane) - 1));
                        if (offset < a.max_pixels) a.mapping3dto2d[row * a.max_pixels + offset] = (int)pix;
                        first = offset == 0;
                    }
                }
#ifdef SPSG_STATS
                // force the payload + atomic results before reading the clock
                const long long clk_e1 = (col0 != 12345.0f && sem[13] != 12345.0f && !(first && dep == 54321.0f)) ? clock64() : 0;
#endif
                {
                    // the first pixel of a (voxel, view) pair appends the pair to the backward's work list (one atomic
                    // per warp)
                    const unsigned m = __ballot_sync(kFull, first);
#ifdef SPSG_NO_LIST
                    if (false) {
#else
                    if (m) {
#endif
                        int base = 0;
                        if (lane == 0) base = atomicAdd(a.list_count, __popc(m));
                        base = __shfl_sync(kFull, base, 0);
                        if (first) a.list[base + __popc(m & ((1u << lane) - 1))] = make_int2(hit, img);
                    }
                }
                if (a.hits && active) a.hits[gpix] = hit;
#ifdef SPSG_STATS
                const long long clk_e2 = clock64();
#endif

                if (kLoss) {
                    float acc[6] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
                    if (hit >= 0) {
                        const LossArgs &L = a.loss;
                        if (L.target_depth) {  // train.py:635-638
                            const float t = __ldg(L.target_depth + gpix);
                            if (t != 0.0f) { acc[0] = fabsf(__fmul_rn(depth, L.voxelsize) - t); acc[1] = 1.0f; }
                        }
                        if (L.target_color) {  // loss.py:246-257
                            const float w = L.weight_color ? __ldg(L.weight_color + gpix) : 1.0f;
                            const float *t = L.target_color + gpix * 3;
                            acc[2] = fabsf(__fadd_rn(__fmul_rn(col0, w), -__fmul_rn(__ldg(t + 0), w))) +
                                     fabsf(__fadd_rn(__fmul_rn(col1, w), -__fmul_rn(__ldg(t + 1), w))) +
                                     fabsf(__fadd_rn(__fmul_rn(col2, w), -__fmul_rn(__ldg(t + 2), w)));
                            acc[3] = 3.0f;
                        }
                        if (L.target_label) {  // train.py:744-746
                            const int y = L.target_label[gpix];
                            if (y < 14 && sem[0] != ninf) {  // valid = (label < 14) & (logit0 != -inf), train.py:744
                                float m = sem[0];
#pragma unroll
                                for (int k = 1; k < 14; k++) m = fmaxf(m, sem[k]);
                                float s = 0.0f, ly = 0.0f;
#pragma unroll
                                for (int k = 0; k < 14; k++) {
                                    s += expf(sem[k] - m);
                                    if (k == y) ly = sem[k];
                                }
                                const float w = L.class_weight ? __ldg(L.class_weight + y) : 1.0f;
                                acc[4] = w * (logf(s) + m - ly);
                                acc[5] = w;
                            }
                        }
                    }
                    float mine = 0.0f;
#pragma unroll
                    for (int k = 0; k < 6; k++) {
                        const float t = warp_sum(acc[k]);
                        if (lane == k) mine = t;
                    }
                    // one double atomic per warp and term, spread over kLossSlots copies of the accumulators
                    const unsigned slot = ((unsigned)tile * 7u + (unsigned)img * 11u) % kLossSlots;
                    if (lane < 6 && mine != 0.0f) atomicAdd(a.loss.accum + slot * 8 + lane, (double)mine);
                }
                // the staging buffer is reused: semantic first, then colour + normal + depth
                const bool vec = a.vec_ok != 0;
#pragma unroll
                for (int k = 0; k < 7; k++) reinterpret_cast<float2 *>(stage + lane * 14)[k] = make_float2(sem[2 * k], sem[2 * k + 1]);
                __syncwarp();
                store_warp_tile<14>(stage, a.image_semantic, img, wx0, wy0, a.width, a.height, vec, lane);
                __syncwarp();
                float *s_col = stage, *s_nrm = stage + 96, *s_dep = stage + 192;
                s_col[lane * 3 + 0] = col0; s_col[lane * 3 + 1] = col1; s_col[lane * 3 + 2] = col2;
                s_nrm[lane * 3 + 0] = n0; s_nrm[lane * 3 + 1] = n1; s_nrm[lane * 3 + 2] = n2;
                s_dep[lane] = dep;
                __syncwarp();
                store_warp_tile<3>(s_col, a.image_color, img, wx0, wy0, a.width, a.height, vec, lane);
                store_warp_tile<3>(s_nrm, a.image_normal, img, wx0, wy0, a.width, a.height, vec, lane);
                store_warp_tile<1>(s_dep, a.image_depth, img, wx0, wy0, a.width, a.height, vec, lane);
                __syncwarp();
#ifdef SPSG_STATS
                if (lane == 0) {
                    const long long clk3 = clock64();
                    STAT_ADD(16, clk1 - clk0); STAT_MAX(17, clk1 - clk0);        // setup + clip
                    STAT_ADD(28, clk1b - clk1); STAT_MAX(29, clk1b - clk1);      // wait for the maps
                    STAT_ADD(18, clk_march); STAT_MAX(19, clk_march);            // march
                    STAT_ADD(20, clk_refine); STAT_MAX(21, clk_refine);          // refinement
                    STAT_ADD(22, clk3 - clk2); STAT_MAX(23, clk3 - clk2);        // epilogue
                    STAT_ADD(40, clk_e1 - clk2); STAT_MAX(41, clk_e1 - clk2);    // payload + registration atomics
                    STAT_ADD(42, clk_e2 - clk_e1); STAT_MAX(43, clk_e2 - clk_e1);  // list append
                    STAT_ADD(44, clk3 - clk_e2); STAT_MAX(45, clk3 - clk_e2);    // staging + stores
                    STAT_ADD(24, clk3 - clk0); STAT_MAX(25, clk3 - clk0);        // whole tile
                    STAT_MAX(26, my_iters);
                    if (tile < 8192) {
                        unsigned smid;
                        asm("mov.u32 %0, %%smid;" : "=r"(smid));
                        int *ts = g_tile_stats[tile];
                        ts[0] = (int)(clk3 - clk0); ts[1] = (int)(clk1 - clk0); ts[2] = (int)clk_march; ts[3] = (int)clk_refine;
                        ts[4] = (int)(clk3 - clk2); ts[5] = my_iters; ts[6] = (int)smid; ts[7] = (int)(clk0 & 0x7fffffff);
                    }
                    STAT_ADD(27, my_iters);
                    int bucket = 0;
                    for (int t = my_iters; t > 8; t >>= 1) bucket++;
                    STAT_ADD(32 + min(bucket, 9), 1);
                }
#endif
            }
            tile = __shfl_sync(kFull, next, 0);
            if (tile < total_tiles) prepare(tile, q);
        }
        if (kSmemMaps) {
            phase ^= 1u;
            if (chunk + (int)gridDim.x < a.num_chunks) {
                __syncthreads();  // all warps are done reading the maps before the next chunk's copy overwrites them
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            }
        }
    }
}


// loss_out[0..3] = depth, colour, semantic, weighted total; [4..6] = normalisers the backward needs.
__global__ void __launch_bounds__(32) finalize_loss_kernel(const double *__restrict__ acc, float *__restrict__ out,
                                                           float w_depth, float w_color, float w_sem, int has_depth,
                                                           int has_color, int has_sem) {
    // one warp: lane l sums slots l, l + 32, ...; xor-shuffle tree over the lanes (fixed order: deterministic)
    const int lane = threadIdx.x;
    double t[6] = {0, 0, 0, 0, 0, 0};
    for (int s = lane; s < kLossSlots; s += 32)
#pragma unroll
        for (int k = 0; k < 6; k++) t[k] += acc[s * 8 + k];
#pragma unroll
    for (int k = 0; k < 6; k++)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t[k] += __shfl_xor_sync(0xffffffffu, t[k], o);
    if (lane == 0) {
        const float ld = has_depth ? (float)(t[0] / t[1]) : 0.0f;  // mean over an empty set is NaN, like torch.mean
        const float lc = has_color ? (float)(t[2] / t[3]) : 0.0f;
        const float ls = has_sem ? (float)(t[4] / t[5]) : 0.0f;
        out[0] = ld; out[1] = lc; out[2] = ls;
        out[3] = w_depth * ld + w_color * lc + w_sem * ls;
        out[4] = (float)t[1]; out[5] = (float)t[3]; out[6] = (float)t[5];
        out[7] = 0.0f;
    }
}

// ---------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------

struct BackwardArgs {
    const float *grad_color, *grad_depth, *grad_normal, *grad_semantic;  // plain variant
    const float *image_color, *image_depth, *image_semantic;             // fused-loss variant
    LossArgs loss;
    const float *loss_out;
    float w_depth, w_color, w_sem;
    const float *grad_scale;  // device scalar or NULL (= 1)
    const int32_t *mapping3dto2d, *mapping3dto2d_num;
    float *d_color, *d_depth, *d_normal, *d_semantic;
    const int32_t *list_count;
    const int2 *list;
    int width, height;
    int views, max_pixels;
    long long num_locs;
    int zero_blocks;  // leading CTAs of the launch that clear gradient rows instead of gathering
    int vec4_ok;      // mapping3dto2d rows are 16-byte aligned
};

// Clears the 21 gradient slots of voxels [0, N) (replaces the 4 whole-buffer memsets of kernel.cu:557-560): a warp
// takes 32 consecutive voxels, so every store instruction writes one contiguous run of the AoS arrays.
// kSkipHit (one view per chunk): rows of voxels that received pixels are left to the gather, which overwrites them.
template <bool kSkipHit>
__device__ __forceinline__ void zero_rows(const BackwardArgs &a, long long first_warp, long long num_warps) {
    const int lane = threadIdx.x & 31;
    for (long long base = first_warp * 32; base < a.num_locs; base += num_warps * 32) {
        const long long i = base + lane;
        const bool keep = kSkipHit && i < a.num_locs && __ldg(a.mapping3dto2d_num + i) > 0;
        const unsigned kept = __ballot_sync(0xffffffffu, keep);
        const int n = (int)min((long long)32, a.num_locs - base);
        float2 *s = reinterpret_cast<float2 *>(a.d_semantic + (size_t)base * 14);
        for (int e = lane; e < n * 7; e += 32)
            if (!((kept >> (e / 7)) & 1u)) s[e] = make_float2(0.0f, 0.0f);
        float *c = a.d_color + (size_t)base * 3, *nm = a.d_normal + (size_t)base * 3;
        for (int e = lane; e < n * 3; e += 32)
            if (!((kept >> (e / 3)) & 1u)) {
                c[e] = 0.0f;
                nm[e] = 0.0f;
            }
        if (lane < n && !keep) a.d_depth[i] = 0.0f;
    }
}

__global__ void __launch_bounds__(256) backward_zero_kernel(const BackwardArgs a) {
    zero_rows<false>(a, ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, ((long long)gridDim.x * blockDim.x) >> 5);
}

// Upstream gradient of one pixel, all 21 channels (g[0..13] semantic, [14..16] colour, [17] depth->sdf, [18..20] normal).
// Plain variant: read from the four gradient images.  Fused variant: recomputed from the rendering and the targets
// of the 2D losses.
// per-term factors of the fused variant: weight * upstream scale / normaliser (loss_out[4..6]), hoisted out of the pixels
struct FusedCoef { float sem, col, dep; };

__device__ __forceinline__ FusedCoef fused_coef(const BackwardArgs &a) {
    FusedCoef c;
    const float scale = a.grad_scale ? __ldg(a.grad_scale) : 1.0f;
    c.dep = a.w_depth * a.loss.voxelsize * scale / a.loss_out[4];
    c.col = a.w_color * scale / a.loss_out[5];
    c.sem = a.w_sem * scale / a.loss_out[6];
    return c;
}

template <bool kFused>
__device__ __forceinline__ void pixel_grads(const BackwardArgs &a, const FusedCoef &fc, unsigned gpix, float (&g)[21]) {
    if (!kFused) {
        const float2 *s2 = reinterpret_cast<const float2 *>(a.grad_semantic + (size_t)gpix * 14);
#pragma unroll
        for (int k = 0; k < 7; k++) {
            const float2 t = __ldg(s2 + k);
            g[2 * k] = t.x; g[2 * k + 1] = t.y;
        }
        const float *c = a.grad_color + (size_t)gpix * 3, *n = a.grad_normal + (size_t)gpix * 3;
        g[14] = __ldg(c); g[15] = __ldg(c + 1); g[16] = __ldg(c + 2);
        g[17] = __ldg(a.grad_depth + gpix);
        g[18] = __ldg(n); g[19] = __ldg(n + 1); g[20] = __ldg(n + 2);
    } else {
        const LossArgs &L = a.loss;
#pragma unroll
        for (int k = 0; k < 21; k++) g[k] = 0.0f;
        // semantic: w[y] * (softmax - onehot) / sum_w   (d/dlogits of F.cross_entropy(..., weight), train.py:745)
        const int y = L.target_label ? (int)L.target_label[gpix] : 14;
        if (y < 14) {
            float l[14];
            const float2 *s2 = reinterpret_cast<const float2 *>(a.image_semantic + (size_t)gpix * 14);
#pragma unroll
            for (int k = 0; k < 7; k++) {
                const float2 t = __ldg(s2 + k);
                l[2 * k] = t.x; l[2 * k + 1] = t.y;
            }
            if (l[0] != -CUDART_INF_F) {  // valid = (label < 14) & (logit0 != -inf), train.py:744
                float m = l[0];
#pragma unroll
                for (int k = 1; k < 14; k++) m = fmaxf(m, l[k]);
                float sum = 0.0f;
#pragma unroll
                for (int k = 0; k < 14; k++) {
                    l[k] = expf(l[k] - m);
                    sum += l[k];
                }
                const float w = L.class_weight ? __ldg(L.class_weight + y) : 1.0f;
                const float f = fc.sem * w, inv_sum = 1.0f / sum;
#pragma unroll
                for (int k = 0; k < 14; k++) g[k] = f * (l[k] * inv_sum - (k == y ? 1.0f : 0.0f));
            }
        }
        if (L.target_color) {  // d/dc mean|c*w - t*w|  (loss.py:246-257)
            const float w = L.weight_color ? __ldg(L.weight_color + gpix) : 1.0f;
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const float c = __ldg(a.image_color + (size_t)gpix * 3 + k);
                const float d = __fadd_rn(__fmul_rn(c, w), -__fmul_rn(__ldg(L.target_color + (size_t)gpix * 3 + k), w));
                if (c != -CUDART_INF_F) g[14 + k] = fc.col * w * (float)((d > 0.0f) - (d < 0.0f));  // valid = != -inf
            }
        }
        if (L.target_depth) {  // d/ddepth mean|depth*voxelsize - t|  (train.py:635-638)
            const float t = __ldg(L.target_depth + gpix);
            const float r = __ldg(a.image_depth + gpix);
            if (t != 0.0f && r != -CUDART_INF_F) {
                const float d = __fmul_rn(r, L.voxelsize) - t;
                g[17] = fc.dep * (float)((d > 0.0f) - (d < 0.0f));
            }
        }
    }
}

// The gather (kernel.cu:391-419 turned inside out).  Work items are the (voxel, view) pairs the forward listed; one
// warp per item.  Lane k fetches the k-th registered pixel id (one coalesced load) and that pixel's 21 upstream
// gradients (independent 8- and 4-byte loads), parks them in shared memory, and lane c then adds column c in
// registration order as grad / count (kernel.cu:398-418) -- a fixed order, so with one view per chunk the result is
// deterministic and written with plain stores.  With several views per chunk the per-view means of a voxel are summed
// with float atomics onto rows the zero kernel cleared (kAtomic).
constexpr int kGatherWarps = 8;

template <bool kFused, bool kAtomic>
__global__ void __launch_bounds__(kGatherWarps * 32) backward_gather_kernel(const BackwardArgs a) {
    if ((int)blockIdx.x < a.zero_blocks) {
        zero_rows<true>(a, ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, ((long long)a.zero_blocks * blockDim.x) >> 5);
        return;
    }
    // half a warp per item: 16 lanes cover the typical pixel count of a voxel, the two halves work on different items
    __shared__ float s_g[kGatherWarps * 2][16 * 21];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int hl = lane & 15, half = lane >> 4;
    const unsigned hmask = 0xffffu << (16 * half);
    float *tile = s_g[warp * 2 + half];
    const int groups_total = (int)(gridDim.x - a.zero_blocks) * kGatherWarps * 2;
    const int count = *a.list_count;
    const unsigned P = (unsigned)(a.width * a.height);
    int item = (((int)blockIdx.x - a.zero_blocks) * kGatherWarps + warp) * 2 + half;
    int2 e = item < count ? a.list[item] : make_int2(0, 0);
    FusedCoef fc = {0.0f, 0.0f, 0.0f};
    if (kFused) fc = fused_coef(a);
    while (item < count) {
        const int idx = e.x, img = e.y;
        const int next_item = item + groups_total;
        if (next_item < count) e = a.list[next_item];  // prefetch the next pair
        const size_t row = (size_t)(img % a.views) * a.num_locs + idx;
        const int32_t *prow = a.mapping3dto2d + row * a.max_pixels;
        const int cnt = min(max(__ldg(a.mapping3dto2d_num + row), 0), a.max_pixels);  // kernel.cu:392-393
        const unsigned pixbase = (unsigned)img * P;  // global pixel index < 2^32 / 14 (check_params)
        const float inv = __frcp_rn((float)max(cnt, 1));
        float acc0 = 0.0f, acc1 = 0.0f;  // channels hl and 16 + hl
        for (int k0 = 0; k0 < cnt; k0 += 16) {
            const int m = min(16, cnt - k0);
            if (hl < m) {
                float g[21];
                pixel_grads<kFused>(a, fc, pixbase + (unsigned)__ldg(prow + k0 + hl), g);
#pragma unroll
                for (int c = 0; c < 21; c++) tile[hl * 21 + c] = g[c];
            }
            __syncwarp(hmask);
            for (int r = 0; r < m; r++) {
                acc0 = __fmaf_rn(tile[r * 21 + hl], inv, acc0);
                if (hl < 5) acc1 = __fmaf_rn(tile[r * 21 + 16 + hl], inv, acc1);
            }
            __syncwarp(hmask);
        }
        // channel c -> destination: 0-13 semantic, 14-16 colour, 17 depth->sdf, 18-20 normal
        float *d0 = hl < 14 ? a.d_semantic + (size_t)idx * 14 + hl : a.d_color + (size_t)idx * 3 + (hl - 14);
        float *d1 = nullptr;
        if (hl == 0) d1 = a.d_color + (size_t)idx * 3 + 2;
        else if (hl == 1) d1 = a.d_depth + idx;
        else if (hl < 5) d1 = a.d_normal + (size_t)idx * 3 + (hl - 2);
        if (kAtomic) {
            atomicAdd(d0, acc0);
            if (d1) atomicAdd(d1, acc1);
        } else {
            *d0 = acc0;
            if (d1) *d1 = acc1;
        }
        item = next_item;
    }
}

// ---------------------------------------------------------------------------------------------
// the 2D losses as stand-alone image-space ops (the reference's own boundary: loss.compute_2dcolor_loss and the
// inline expressions of train.py:635-638, 744-746 applied to rendered images)
// ---------------------------------------------------------------------------------------------

struct Losses2DArgs {
    const float *image_color, *image_depth, *image_semantic;
    LossArgs loss;
    long long num_pixels;
};

// one thread per pixel: the same six sums the fused forward accumulates in its epilogue
__global__ void __launch_bounds__(256) losses2d_forward_kernel(const Losses2DArgs a) {
    const float ninf = -CUDART_INF_F;
    float acc[6] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
    const LossArgs &L = a.loss;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < a.num_pixels; p += (long long)gridDim.x * blockDim.x) {
        if (L.target_depth) {  // train.py:635-638
            const float r = __ldg(a.image_depth + p), t = __ldg(L.target_depth + p);
            if (r != ninf && t != 0.0f) { acc[0] += fabsf(__fmul_rn(r, L.voxelsize) - t); acc[1] += 1.0f; }
        }
        if (L.target_color) {  // loss.py:246-257 (valid is per element)
            const float w = L.weight_color ? __ldg(L.weight_color + p) : 1.0f;
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const float c = __ldg(a.image_color + p * 3 + k);
                if (c != ninf) {
                    acc[2] += fabsf(__fadd_rn(__fmul_rn(c, w), -__fmul_rn(__ldg(L.target_color + p * 3 + k), w)));
                    acc[3] += 1.0f;
                }
            }
        }
        if (L.target_label) {  // train.py:744-746
            const int y = L.target_label[p];
            if (y < 14) {
                float l[14];
                const float2 *s2 = reinterpret_cast<const float2 *>(a.image_semantic + p * 14);
#pragma unroll
                for (int k = 0; k < 7; k++) {
                    const float2 t2 = __ldg(s2 + k);
                    l[2 * k] = t2.x; l[2 * k + 1] = t2.y;
                }
                if (l[0] != ninf) {
                    float m = l[0];
#pragma unroll
                    for (int k = 1; k < 14; k++) m = fmaxf(m, l[k]);
                    float sum = 0.0f, ly = 0.0f;
#pragma unroll
                    for (int k = 0; k < 14; k++) {
                        sum += expf(l[k] - m);
                        if (k == y) ly = l[k];
                    }
                    const float w = L.class_weight ? __ldg(L.class_weight + y) : 1.0f;
                    acc[4] += w * (logf(sum) + m - ly);
                    acc[5] += w;
                }
            }
        }
    }
    const int lane = threadIdx.x & 31;
    float mine = 0.0f;
#pragma unroll
    for (int k = 0; k < 6; k++) {
        const float t = warp_sum(acc[k]);
        if (lane == k) mine = t;
    }
    const unsigned slot = (blockIdx.x * 8u + (threadIdx.x >> 5)) % kLossSlots;
    if (lane < 6 && mine != 0.0f) atomicAdd(L.accum + slot * 8 + lane, (double)mine);
}

// gradient images of the weighted total w.r.t. the renderings (zero where a pixel is not part of a term)
__global__ void __launch_bounds__(256) losses2d_backward_kernel(const BackwardArgs a, long long num_pixels, float *d_color,
                                                               float *d_depth, float *d_semantic) {
    const FusedCoef fc = fused_coef(a);
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < num_pixels; p += (long long)gridDim.x * blockDim.x) {
        float g[21];
        pixel_grads<true>(a, fc, (unsigned)p, g);
        if (d_semantic) {
#pragma unroll
            for (int k = 0; k < 7; k++) reinterpret_cast<float2 *>(d_semantic + p * 14)[k] = make_float2(g[2 * k], g[2 * k + 1]);
        }
        if (d_color) { d_color[p * 3] = g[14]; d_color[p * 3 + 1] = g[15]; d_color[p * 3 + 2] = g[16]; }
        if (d_depth) d_depth[p] = g[17];
    }
}

// ---------------------------------------------------------------------------------------------
// occupancy raycast
// ---------------------------------------------------------------------------------------------

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

__global__ void __launch_bounds__(kTilePix) raycast_occ_kernel(const OccArgs a) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned ux = blockIdx.x * kTileW + (warp & 1) * 8 + (lane & 7);
    const unsigned uy = blockIdx.y * kTileH + (warp >> 1) * 4 + (lane >> 3);
    const int img = blockIdx.z;
    if (ux >= (unsigned)a.width || uy >= (unsigned)a.height) return;
    const Ray r = setup_ray(a.view_matrix + (size_t)img * 16, a.intrinsics + (size_t)img * 4, ux, uy, a.depth_min,
                            a.depth_max);
    const uint8_t *__restrict__ occ = a.occ3d + (size_t)img * a.dimz * a.dimy * a.dimx;
    Stepper step;
    step.init(a.inc);
    float ray = r.t0, t_end = r.t1;
    if (!(a.flags & SPSG_FLAG_NO_CLIP)) {
        // nearest voxel is inside the grid only for p in (-0.5, dim-0.5)
        const float kInf = CUDART_INF_F;
        float tin = -kInf, tout = kInf;
        const float o[3] = {r.camx, r.camy, r.camz}, d[3] = {r.dx, r.dy, r.dz};
        const float hi[3] = {(float)a.dimx - 0.5f + kBoxEps, (float)a.dimy - 0.5f + kBoxEps, (float)a.dimz - 0.5f + kBoxEps};
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const float lo = -0.5f - kBoxEps;
            if (d[k] != 0.0f) {
                const float inv = rcp_approx(d[k]);
                const float ta = (lo - o[k]) * inv, tb = (hi[k] - o[k]) * inv;
                tin = fmaxf(tin, fminf(ta, tb));
                tout = fminf(tout, fmaxf(ta, tb));
            } else if (o[k] < lo || o[k] > hi[k]) {
                tin = kInf; tout = -kInf;
            }
        }
        const float margin = 0.0625f;
        if (!(tin <= tout)) {
            t_end = ray;
        } else {
            t_end = fminf(t_end, tout + margin);
            while (ray < tin - margin - a.inc && ray < t_end) {
                const int want = max(1, min(__float2int_rd((tin - margin - ray) * step.inv_inc) - 1, 1 << 22));
                ray = step.advance(ray, want);
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

// ---------------------------------------------------------------------------------------------
// per-voxel normals of the sparse SDF (producer of the raycaster's vals_normals)
// ---------------------------------------------------------------------------------------------

// loss.compute_normals_sparse (loss.py:285-306) = compute_normals_dense (:261-267) on the scattered volume + gather +
// per-chunk rotation + -normalize, as one gather kernel over the voxel index:
//   g   = central differences of the SDF at the voxel, absent neighbours count as 0 (the dense volume is zero-filled,
//         loss.py:288-289), g = 0 on the volume border (the -inf padding that is then zeroed, :293-296);
//   m   = R_chunk * g            (transform[b, :3, :3], :299-301; identity without a transform)
//   out = -m / max(|m|, 1e-5)    (F.normalize(p=2, eps=1e-5), :305)
struct NormalsArgs {
    const longlong4 *locs;
    const float *sdf;
    const float *transform;  // (B,4,4) row-major or NULL
    const int32_t *index;    // (B,Dz,Dy,Dx) voxel -> row, -1 = absent
    const float *grad_out;   // backward: dL/d out (N,3)
    float *out;              // forward: normals (N,3); backward pass 1: u = dL/dg (N,3)
    float *d_sdf;            // backward pass 2: (N,1)
    long long n;
    int dimx, dimy, dimz;
};

__device__ __forceinline__ float sdf_at(const NormalsArgs &a, size_t chunk_base, int x, int y, int z) {
    const int i = __ldg(a.index + chunk_base + ((size_t)z * a.dimy + y) * a.dimx + x);
    return i >= 0 ? __ldg(a.sdf + i) : 0.0f;
}

__device__ __forceinline__ bool normals_gradient(const NormalsArgs &a, const longlong4 l, float &gx, float &gy, float &gz) {
    const int z = (int)l.x, y = (int)l.y, x = (int)l.z;
    gx = gy = gz = 0.0f;
    if (x < 1 || y < 1 || z < 1 || x > a.dimx - 2 || y > a.dimy - 2 || z > a.dimz - 2) return false;  // border: zero
    const size_t base = (size_t)l.w * a.dimz * a.dimy * a.dimx;
    gx = sdf_at(a, base, x + 1, y, z) - sdf_at(a, base, x - 1, y, z);
    gy = sdf_at(a, base, x, y + 1, z) - sdf_at(a, base, x, y - 1, z);
    gz = sdf_at(a, base, x, y, z + 1) - sdf_at(a, base, x, y, z - 1);
    return true;
}

__device__ __forceinline__ void load_rotation(const NormalsArgs &a, long long chunk, float (&R)[9]) {
    if (a.transform) {
        const float *t = a.transform + chunk * 16;
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
            for (int c = 0; c < 3; c++) R[r * 3 + c] = __ldg(t + r * 4 + c);
    } else {
#pragma unroll
        for (int k = 0; k < 9; k++) R[k] = (k % 4 == 0) ? 1.0f : 0.0f;
    }
}

__global__ void __launch_bounds__(256) normals_forward_kernel(const NormalsArgs a) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    const longlong4 l = a.locs[i];
    float gx, gy, gz, R[9];
    normals_gradient(a, l, gx, gy, gz);
    load_rotation(a, l.w, R);
    const float mx = R[0] * gx + R[1] * gy + R[2] * gz, my = R[3] * gx + R[4] * gy + R[5] * gz,
                mz = R[6] * gx + R[7] * gy + R[8] * gz;
    const float inv = 1.0f / fmaxf(sqrtf(mx * mx + my * my + mz * mz), 1e-5f);
    a.out[i * 3 + 0] = -(mx * inv);
    a.out[i * 3 + 1] = -(my * inv);
    a.out[i * 3 + 2] = -(mz * inv);
}

// backward pass 1: u = dL/dg per voxel (zero on the border), through -normalize and the rotation
__global__ void __launch_bounds__(256) normals_backward_u_kernel(const NormalsArgs a) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    const longlong4 l = a.locs[i];
    float gx, gy, gz, R[9];
    float ux = 0.0f, uy = 0.0f, uz = 0.0f;
    if (normals_gradient(a, l, gx, gy, gz)) {
        load_rotation(a, l.w, R);
        const float mx = R[0] * gx + R[1] * gy + R[2] * gz, my = R[3] * gx + R[4] * gy + R[5] * gz,
                    mz = R[6] * gx + R[7] * gy + R[8] * gz;
        const float len = sqrtf(mx * mx + my * my + mz * mz);
        const float qx = __ldg(a.grad_out + i * 3 + 0), qy = __ldg(a.grad_out + i * 3 + 1), qz = __ldg(a.grad_out + i * 3 + 2);
        float dmx, dmy, dmz;  // dL/dm for out = -m / max(len, eps)
        if (len > 1e-5f) {
            const float inv = 1.0f / len;
            const float hx = mx * inv, hy = my * inv, hz = mz * inv, dot = hx * qx + hy * qy + hz * qz;
            dmx = -(qx - hx * dot) * inv; dmy = -(qy - hy * dot) * inv; dmz = -(qz - hz * dot) * inv;
        } else {
            dmx = -qx * 1e5f; dmy = -qy * 1e5f; dmz = -qz * 1e5f;
        }
        ux = R[0] * dmx + R[3] * dmy + R[6] * dmz;  // R^T
        uy = R[1] * dmx + R[4] * dmy + R[7] * dmz;
        uz = R[2] * dmx + R[5] * dmy + R[8] * dmz;
    }
    a.out[i * 3 + 0] = ux; a.out[i * 3 + 1] = uy; a.out[i * 3 + 2] = uz;
}

// backward pass 2: the SDF value of voxel j enters g of its six neighbours with weight +-1 -- a gather, no atomics
__global__ void __launch_bounds__(256) normals_backward_gather_kernel(const NormalsArgs a) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= a.n) return;
    const longlong4 l = a.locs[j];
    const int z = (int)l.x, y = (int)l.y, x = (int)l.z;
    const size_t base = (size_t)l.w * a.dimz * a.dimy * a.dimx;
    const float *u = a.out;
    auto at = [&](int xx, int yy, int zz, int comp) -> float {
        if (xx < 0 || yy < 0 || zz < 0 || xx >= a.dimx || yy >= a.dimy || zz >= a.dimz) return 0.0f;
        const int i = __ldg(a.index + base + ((size_t)zz * a.dimy + yy) * a.dimx + xx);
        return i >= 0 ? __ldg(u + (size_t)i * 3 + comp) : 0.0f;
    };
    a.d_sdf[j] = (at(x - 1, y, z, 0) - at(x + 1, y, z, 0)) + (at(x, y - 1, z, 1) - at(x, y + 1, z, 1)) +
                 (at(x, y, z - 1, 2) - at(x, y, z + 1, 2));
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------

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
    if ((long long)p->num_locs * p->views_per_chunk >= (1ll << 31))
        return fail(SPSG_ERR_INVALID_ARGUMENT, "num_locs * views_per_chunk exceeds 32-bit indexing");
    return SPSG_OK;
}

bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cached[dev] = sms > 0 ? sms : 148;
    }
    return cached[dev];
}

LossArgs make_loss_args(const spsg_loss_targets *t, double *accum) {
    LossArgs L;
    memset(&L, 0, sizeof(L));
    if (t) {
        L.target_depth = t->target_depth; L.target_color = t->target_color; L.weight_color = t->weight_color;
        L.target_label = t->target_label; L.class_weight = t->class_weight; L.voxelsize = t->voxelsize;
    }
    L.accum = accum;
    return L;
}

int launch_forward(const spsg_raycast_params *p, bool build_index, int32_t *sparse_mapping, const int64_t *locs,
                   const float *vals_sdf, const float *vals_color, const float *vals_normal,
                   const float *vals_semantic, const float *view_matrix, const float *intrinsics, float *image_color,
                   float *image_depth, float *image_normal, float *image_semantic, int32_t *mapping3dto2d,
                   int32_t *mapping3dto2d_num, const spsg_loss_targets *targets, float *loss_out,
                   const spsg_grad_buffers *clear_grads, void *workspace, size_t workspace_bytes, cudaStream_t st) {
    if (int rc = check_params(p)) return rc;
    if (p->max_pixels_per_voxel <= 0) return fail(SPSG_ERR_INVALID_ARGUMENT, "max_pixels_per_voxel must be positive");
    if (!sparse_mapping || !view_matrix || !intrinsics || !image_color || !image_depth || !image_normal ||
        !image_semantic || !mapping3dto2d || !mapping3dto2d_num)
        return fail(SPSG_ERR_INVALID_ARGUMENT, "NULL tensor pointer");
    if (p->num_locs > 0 && (!locs || !vals_sdf || !vals_color || !vals_normal || !vals_semantic))
        return fail(SPSG_ERR_INVALID_ARGUMENT, "NULL voxel tensor pointer");
    if (!aligned16(view_matrix) || !aligned16(intrinsics))
        return fail(SPSG_ERR_INVALID_ARGUMENT, "view_matrix / intrinsics must be 16-byte aligned");
    if (p->num_locs > 0 && ((reinterpret_cast<uintptr_t>(vals_semantic) & 7u) || (reinterpret_cast<uintptr_t>(locs) & 15u)))
        return fail(SPSG_ERR_INVALID_ARGUMENT, "vals_semantic must be 8-byte and locs 16-byte aligned");
    if (reinterpret_cast<uintptr_t>(sparse_mapping) & 3u) return fail(SPSG_ERR_INVALID_ARGUMENT, "sparse_mapping must be 4-byte aligned");
    if (targets && !loss_out) return fail(SPSG_ERR_INVALID_ARGUMENT, "loss_out is NULL");
    const Layout L = make_layout(p);
    if (!workspace || workspace_bytes < L.total) return fail(SPSG_ERR_WORKSPACE_TOO_SMALL, "workspace too small");
    if (reinterpret_cast<uintptr_t>(workspace) & 255u) return fail(SPSG_ERR_INVALID_ARGUMENT, "workspace must be 256-byte aligned");
    uint8_t *ws = (uint8_t *)workspace;
    float *dense = (float *)(ws + L.dense_off);
    uint2 *vbits = (uint2 *)(ws + L.vbit_off);
    uint8_t *bmap = ws + L.bmap_off, *marks = ws + L.marks_off;
    double *accum = (double *)(ws + L.loss_off);
    const size_t cells = (size_t)p->num_chunks * p->dimz * p->dimy * p->dimx;
    const int sms = sm_count();

    {
        // sparse_mapping := -1 (kernel.cu:515), dense brick := NaN (0xffffffff: every voxel absent), block marks / list
        // counter / loss accumulators := 0 -- one launch
        FillArgs f;
        memset(&f, 0, sizeof(f));
        f.ptr[0] = build_index ? (uint32_t *)sparse_mapping : nullptr; f.words[0] = cells; f.value[0] = 0xffffffffu;
        f.ptr[1] = (uint32_t *)dense; f.words[1] = L.dense_bytes / 4; f.value[1] = 0xffffffffu;
        f.ptr[2] = (uint32_t *)(ws + L.zero_off); f.words[2] = L.zero_bytes / 4; f.value[2] = 0u;
        size_t grad_words = 0;
        if (clear_grads && p->num_locs > 0) {  // rows [0, N) of the backward's outputs (kernel.cu:557-560)
            if (!clear_grads->d_color || !clear_grads->d_depth || !clear_grads->d_normal || !clear_grads->d_semantic)
                return fail(SPSG_ERR_INVALID_ARGUMENT, "NULL gradient pointer in clear_grads");
            const size_t n = (size_t)p->num_locs;
            f.ptr[3] = (uint32_t *)clear_grads->d_semantic; f.words[3] = n * 14;
            f.ptr[4] = (uint32_t *)clear_grads->d_color; f.words[4] = n * 3;
            f.ptr[5] = (uint32_t *)clear_grads->d_normal; f.words[5] = n * 3;
            f.ptr[6] = (uint32_t *)clear_grads->d_depth; f.words[6] = n;
            grad_words = n * 21;
        }
        const size_t vecs = (cells * (build_index ? 2 : 1) + L.zero_bytes / 4 + grad_words) / 4;
        const unsigned blocks = (unsigned)std::min<size_t>((vecs + 1023) / 1024 + 1, (size_t)sms * 8);
        fill_kernel<<<blocks, 256, 0, st>>>(f);
        CUDA_TRY(cudaGetLastError());
    }
    if (p->num_locs > 0) {
        const unsigned blocks = (unsigned)((p->num_locs + 255) / 256);
        if (build_index)
            index_kernel<true><<<blocks, 256, 0, st>>>((const longlong4 *)locs, p->num_locs, sparse_mapping, vals_sdf,
                                                       dense, mapping3dto2d_num, p->views_per_chunk, p->dimz, p->dimy,
                                                       p->dimx);
        else
            index_kernel<false><<<blocks, 256, 0, st>>>((const longlong4 *)locs, p->num_locs, sparse_mapping, vals_sdf,
                                                        dense, mapping3dto2d_num, p->views_per_chunk, p->dimz, p->dimy,
                                                        p->dimx);
        CUDA_TRY(cudaGetLastError());
    }
    ForwardArgs a;
    memset(&a, 0, sizeof(a));
    // shared-memory residency of one chunk's maps: class bit planes + block map
    const size_t map_bytes = L.vpc * sizeof(uint2) + L.bpc;
    a.maps_in_smem = kFwdSmemFixed + map_bytes <= kFwdSmemMax;
    {
        const dim3 cgrid((unsigned)((L.nby * L.wpr + 3) / 4), (unsigned)L.nbz, (unsigned)p->num_chunks);
        cell_class_kernel<<<cgrid, 128, 0, st>>>(dense, vbits, L.vpc, marks, p->dimz, p->dimy, p->dimx, L.wpr, L.nby, L.nbx, L.bpc,
                                                 bmap, (int32_t *)(ws + L.arrive_off), L.nbz);
        CUDA_TRY(cudaGetLastError());
    }
    a.sparse_mapping = sparse_mapping;
    a.vals_sdf = vals_sdf; a.vals_color = vals_color; a.vals_normal = vals_normal; a.vals_semantic = vals_semantic;
    a.view_matrix = view_matrix; a.intrinsics = intrinsics;
    a.image_color = image_color; a.image_depth = image_depth; a.image_normal = image_normal;
    a.image_semantic = image_semantic;
    a.mapping3dto2d = mapping3dto2d; a.mapping3dto2d_num = mapping3dto2d_num;
    a.dense = dense; a.vbits = vbits; a.bmap = bmap; a.vpc = L.vpc; a.bpc = L.bpc; a.wpr = L.wpr;
    a.tile_counter = (int32_t *)(ws + L.tiles_off);
    a.num_chunks = p->num_chunks;
    a.list_count = (int32_t *)(ws + L.head_off);
    a.list = (int2 *)(ws + L.list_off);
    a.hits = (p->flags & SPSG_FLAG_RECORD_HITS) ? (int32_t *)(ws + L.hits_off) : nullptr;
    a.width = p->width; a.height = p->height;
    a.depth_min = p->depth_min; a.depth_max = p->depth_max; a.thresh = p->thresh_sample_dist; a.inc = p->ray_increment;
    a.dimx = p->dimx; a.dimy = p->dimy; a.dimz = p->dimz;
    a.nbx = L.nbx; a.nby = L.nby; a.nbz = L.nbz;
    a.views = p->views_per_chunk; a.max_pixels = p->max_pixels_per_voxel;
    a.num_locs = p->num_locs;
    a.flags = p->flags;
    a.vec_ok = (p->width % 4 == 0) && aligned16(image_color) && aligned16(image_depth) && aligned16(image_normal) &&
               aligned16(image_semantic);
    {
        int k = 0;
        while ((1 << k) < std::max(p->dimx, std::max(p->dimy, p->dimz))) k++;
        a.guard = std::min(kFracGuard, std::max(ldexpf(1.0f, k - 21), ldexpf(1.0f, -18)));  // 8 ulp of the largest coordinate
    }
    a.loss = make_loss_args(targets, accum);
    // persistent: one CTA per SM (fewer when there is less than one tile per warp)
    const long long tiles_x = (p->width + kWarpW - 1) / kWarpW, tiles_y = (p->height + kWarpH - 1) / kWarpH;
    const long long all_tiles = ((tiles_x + 1) / 2) * ((tiles_y + 1) / 2) * 4 * p->views_per_chunk * p->num_chunks;
    const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>(sms, all_tiles));
    // CTA size: 24 warps when every warp gets about one tile (latency-bound), 28 when there are many tiles per SM
    const bool large = all_tiles >= (long long)sms * 4 * kFwdWarpsLarge;
    const int warps = large ? kFwdWarpsLarge : kFwdWarps;
    const size_t dyn = fwd_smem_fixed(warps) + (a.maps_in_smem ? map_bytes : 0);
    {
        static std::mutex mu;
        static bool configured[64] = {false};
        int dev = 0;
        CUDA_TRY(cudaGetDevice(&dev));
        std::lock_guard<std::mutex> lk(mu);
        if (dev < 0 || dev >= 64 || !configured[dev]) {
#define SPSG_SET_SMEM(L, M, W) CUDA_TRY(cudaFuncSetAttribute(raycast_forward_kernel<L, M, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFwdSmemMax))
            SPSG_SET_SMEM(true, true, kFwdWarps); SPSG_SET_SMEM(false, true, kFwdWarps);
            SPSG_SET_SMEM(true, false, kFwdWarps); SPSG_SET_SMEM(false, false, kFwdWarps);
            SPSG_SET_SMEM(true, true, kFwdWarpsLarge); SPSG_SET_SMEM(false, true, kFwdWarpsLarge);
            SPSG_SET_SMEM(true, false, kFwdWarpsLarge); SPSG_SET_SMEM(false, false, kFwdWarpsLarge);
#undef SPSG_SET_SMEM
            if (dev >= 0 && dev < 64) configured[dev] = true;
        }
    }
    {
        ScopedKernelTimer timer(0, st);
#define SPSG_LAUNCH(L, M)                                                                                         \
    do {                                                                                                          \
        if (large) raycast_forward_kernel<L, M, kFwdWarpsLarge><<<grid, kFwdWarpsLarge * 32, dyn, st>>>(a);      \
        else raycast_forward_kernel<L, M, kFwdWarps><<<grid, kFwdWarps * 32, dyn, st>>>(a);                      \
    } while (0)
        if (targets) {
            if (a.maps_in_smem) SPSG_LAUNCH(true, true); else SPSG_LAUNCH(true, false);
        } else {
            if (a.maps_in_smem) SPSG_LAUNCH(false, true); else SPSG_LAUNCH(false, false);
        }
#undef SPSG_LAUNCH
    }
    CUDA_TRY(cudaGetLastError());
    if (targets) {
        finalize_loss_kernel<<<1, 32, 0, st>>>(accum, loss_out, targets->weight_depth, targets->weight_color_loss,
                                              targets->weight_semantic, targets->target_depth != nullptr,
                                              targets->target_color != nullptr, targets->target_label != nullptr);
        CUDA_TRY(cudaGetLastError());
    }
    return SPSG_OK;
}

int launch_backward(const spsg_raycast_params *p, bool fused, const float *g_or_img_color, const float *g_or_img_depth,
                    const float *grad_normal, const float *g_or_img_semantic, const spsg_loss_targets *targets,
                    const float *loss_out, const float *grad_scale, const int32_t *sparse_mapping,
                    const int32_t *mapping3dto2d, const int32_t *mapping3dto2d_num, float *d_color, float *d_depth,
                    float *d_normal, float *d_semantic, void *workspace, size_t workspace_bytes, cudaStream_t st) {
    if (int rc = check_params(p)) return rc;
    if (!g_or_img_color || !g_or_img_depth || !g_or_img_semantic || (!fused && !grad_normal) || !sparse_mapping ||
        !mapping3dto2d || !mapping3dto2d_num)
        return fail(SPSG_ERR_INVALID_ARGUMENT, "NULL tensor pointer");
    if (fused && (!targets || !loss_out)) return fail(SPSG_ERR_INVALID_ARGUMENT, "NULL loss targets");
    if (p->max_pixels_per_voxel <= 0) return fail(SPSG_ERR_INVALID_ARGUMENT, "max_pixels_per_voxel must be positive");
    if (p->num_locs == 0) return SPSG_OK;
    if (!d_color || !d_depth || !d_normal || !d_semantic) return fail(SPSG_ERR_INVALID_ARGUMENT, "NULL gradient pointer");
    if (reinterpret_cast<uintptr_t>(d_semantic) & 7u) return fail(SPSG_ERR_INVALID_ARGUMENT, "d_semantic must be 8-byte aligned");
    const Layout L = make_layout(p);
    if (!workspace || workspace_bytes < L.total) return fail(SPSG_ERR_WORKSPACE_TOO_SMALL, "workspace too small");
    uint8_t *ws = (uint8_t *)workspace;
    BackwardArgs a;
    memset(&a, 0, sizeof(a));
    if (fused) {
        a.image_color = g_or_img_color; a.image_depth = g_or_img_depth; a.image_semantic = g_or_img_semantic;
        a.loss = make_loss_args(targets, nullptr);
        a.loss_out = loss_out;
        a.w_depth = targets->weight_depth; a.w_color = targets->weight_color_loss; a.w_sem = targets->weight_semantic;
        a.grad_scale = grad_scale;
    } else {
        a.grad_color = g_or_img_color; a.grad_depth = g_or_img_depth; a.grad_normal = grad_normal;
        a.grad_semantic = g_or_img_semantic;
    }
    a.mapping3dto2d = mapping3dto2d; a.mapping3dto2d_num = mapping3dto2d_num;
    a.d_color = d_color; a.d_depth = d_depth; a.d_normal = d_normal; a.d_semantic = d_semantic;
    a.list_count = (const int32_t *)(ws + L.head_off);
    a.list = (const int2 *)(ws + L.list_off);
    a.width = p->width; a.height = p->height;
    a.views = p->views_per_chunk; a.max_pixels = p->max_pixels_per_voxel;
    a.num_locs = p->num_locs;
    a.vec4_ok = (p->max_pixels_per_voxel % 4 == 0) && aligned16(mapping3dto2d);
    const int sms = sm_count();
    // half a warp per listed (voxel, view) pair; the list length is only known on the device: size for its bound N * F
    const long long max_items = p->num_locs * p->views_per_chunk;
    const unsigned gather_blocks = (unsigned)std::max<long long>(1, std::min<long long>((max_items + 2 * kGatherWarps - 1) / (2 * kGatherWarps), (long long)sms * 8));
    const unsigned zero_blocks = (unsigned)std::min<long long>((p->num_locs + 255) / 256, (long long)sms * 4);
    const bool cleared = (p->flags & SPSG_FLAG_GRADS_CLEARED) != 0;  // the forward's fill pass cleared rows [0, N)
    if (cleared) {
        a.zero_blocks = 0;
        ScopedKernelTimer timer(1, st);
        if (p->views_per_chunk == 1) {
            if (fused) backward_gather_kernel<true, false><<<gather_blocks, kGatherWarps * 32, 0, st>>>(a);
            else backward_gather_kernel<false, false><<<gather_blocks, kGatherWarps * 32, 0, st>>>(a);
        } else {
            if (fused) backward_gather_kernel<true, true><<<gather_blocks, kGatherWarps * 32, 0, st>>>(a);
            else backward_gather_kernel<false, true><<<gather_blocks, kGatherWarps * 32, 0, st>>>(a);
        }
    } else if (p->views_per_chunk == 1) {
        // one launch: leading CTAs clear the rows of voxels nothing hit, the rest gather (plain stores)
        a.zero_blocks = (int)zero_blocks;
        ScopedKernelTimer timer(1, st);
        if (fused) backward_gather_kernel<true, false><<<zero_blocks + gather_blocks, kGatherWarps * 32, 0, st>>>(a);
        else backward_gather_kernel<false, false><<<zero_blocks + gather_blocks, kGatherWarps * 32, 0, st>>>(a);
    } else {
        a.zero_blocks = 0;
        backward_zero_kernel<<<zero_blocks, 256, 0, st>>>(a);
        CUDA_TRY(cudaGetLastError());
        ScopedKernelTimer timer(1, st);
        if (fused) backward_gather_kernel<true, true><<<gather_blocks, kGatherWarps * 32, 0, st>>>(a);
        else backward_gather_kernel<false, true><<<gather_blocks, kGatherWarps * 32, 0, st>>>(a);
    }
    CUDA_TRY(cudaGetLastError());
    return SPSG_OK;
}

}  // namespace

// error plumbing for the other translation units of the library (csrc/spsg_internal.h)
int spsg_internal_fail(int code, const char *msg) { return fail(code, msg); }
int spsg_internal_fail_cuda(cudaError_t e, const char *where) { return fail_cuda(e, where); }

extern "C" {

#ifdef SPSG_STATS
SPSG_API int spsg_debug_tile_stats(int *out) {
    return cudaMemcpyFromSymbol(out, g_tile_stats, sizeof(int) * 8192 * 8) == cudaSuccess ? SPSG_OK : SPSG_ERR_CUDA;
}

SPSG_API int spsg_debug_stats(unsigned long long *out, int reset) {
    if (cudaMemcpyFromSymbol(out, g_stats, sizeof(unsigned long long) * 48) != cudaSuccess) return SPSG_ERR_CUDA;
    if (reset) {
        unsigned long long z[48] = {0};
        cudaMemcpyToSymbol(g_stats, z, sizeof(z));
    }
    return SPSG_OK;
}
#endif

const char *spsg_version(void) { return "spsg_raycast_b200 0.3 (sm_100a)"; }
const char *spsg_last_error(void) { return g_err; }

void spsg_timing_enable(int on) {
    std::lock_guard<std::mutex> lk(g_timing_mu);
    g_timing = on != 0;
}

int spsg_timing_read(int which, double *total_ms, int *launches) {
    if (which < 0 || which > 1 || !total_ms || !launches) return fail(SPSG_ERR_INVALID_ARGUMENT, "bad timing query");
    std::vector<EventPair> evs;
    {
        std::lock_guard<std::mutex> lk(g_timing_mu);
        evs.swap(g_ev[which]);
    }
    *total_ms = 0.0;
    *launches = 0;
    for (EventPair &e : evs) {
        float ms = 0.0f;
        if (cudaEventSynchronize(e.b) == cudaSuccess && cudaEventElapsedTime(&ms, e.a, e.b) == cudaSuccess) {
            *total_ms += ms;
            *launches += 1;
        }
        cudaEventDestroy(e.a);
        cudaEventDestroy(e.b);
    }
    return SPSG_OK;
}

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
            (const longlong4 *)locs, num_locs, sparse_mapping, nullptr, nullptr, nullptr, 0, dimz, dimy, dimx);
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
                          image_semantic, mapping3dto2d, mapping3dto2d_num, nullptr, nullptr, nullptr, workspace,
                          workspace_bytes, (cudaStream_t)stream);
}

int spsg_raycast_forward_indexed(const spsg_raycast_params *p, int32_t *sparse_mapping, const int64_t *locs,
                                 const float *vals_sdf, const float *vals_color, const float *vals_normal,
                                 const float *vals_semantic, const float *view_matrix, const float *intrinsics,
                                 float *image_color, float *image_depth, float *image_normal, float *image_semantic,
                                 int32_t *mapping3dto2d, int32_t *mapping3dto2d_num,
                                 const spsg_grad_buffers *clear_grads, void *workspace, size_t workspace_bytes,
                                 void *stream) {
    return launch_forward(p, true, sparse_mapping, locs, vals_sdf, vals_color, vals_normal, vals_semantic, view_matrix,
                          intrinsics, image_color, image_depth, image_normal, image_semantic, mapping3dto2d,
                          mapping3dto2d_num, nullptr, nullptr, clear_grads, workspace, workspace_bytes,
                          (cudaStream_t)stream);
}

int spsg_raycast_backward(const spsg_raycast_params *p, const float *grad_color, const float *grad_depth,
                          const float *grad_normal, const float *grad_semantic, const int32_t *sparse_mapping,
                          const int32_t *mapping3dto2d, const int32_t *mapping3dto2d_num, float *d_color,
                          float *d_depth, float *d_normal, float *d_semantic, void *workspace, size_t workspace_bytes,
                          void *stream) {
    return launch_backward(p, false, grad_color, grad_depth, grad_normal, grad_semantic, nullptr, nullptr, nullptr,
                           sparse_mapping, mapping3dto2d, mapping3dto2d_num, d_color, d_depth, d_normal, d_semantic,
                           workspace, workspace_bytes, (cudaStream_t)stream);
}

int spsg_raycast_occ(const spsg_raycast_params *p, const uint8_t *occ3d, uint8_t *occ2d, const float *view_matrix,
                     const float *intrinsics, void *stream) {
    if (int rc = check_params(p)) return rc;
    if (!occ3d || !occ2d || !view_matrix || !intrinsics) return fail(SPSG_ERR_INVALID_ARGUMENT, "NULL tensor pointer");
    if (!aligned16(view_matrix) || !aligned16(intrinsics))
        return fail(SPSG_ERR_INVALID_ARGUMENT, "view_matrix / intrinsics must be 16-byte aligned");
    OccArgs a;
    a.occ3d = occ3d; a.occ2d = occ2d; a.view_matrix = view_matrix; a.intrinsics = intrinsics;
    a.width = p->width; a.height = p->height;
    a.depth_min = p->depth_min; a.depth_max = p->depth_max; a.inc = p->ray_increment;
    a.dimx = p->dimx; a.dimy = p->dimy; a.dimz = p->dimz;
    a.flags = p->flags;
    const dim3 grid((p->width + kTileW - 1) / kTileW, (p->height + kTileH - 1) / kTileH, p->num_chunks);
    raycast_occ_kernel<<<grid, kTilePix, 0, (cudaStream_t)stream>>>(a);
    CUDA_TRY(cudaGetLastError());
    return SPSG_OK;
}

int spsg_raycast_forward_loss(const spsg_raycast_params *p, int32_t *sparse_mapping, const int64_t *locs,
                              const float *vals_sdf, const float *vals_color, const float *vals_normal,
                              const float *vals_semantic, const float *view_matrix, const float *intrinsics,
                              float *image_color, float *image_depth, float *image_normal, float *image_semantic,
                              int32_t *mapping3dto2d, int32_t *mapping3dto2d_num, const spsg_loss_targets *t,
                              float *loss_out, const spsg_grad_buffers *clear_grads, void *workspace,
                              size_t workspace_bytes, void *stream) {
    if (!t) return fail(SPSG_ERR_INVALID_ARGUMENT, "loss targets are NULL");
    return launch_forward(p, true, sparse_mapping, locs, vals_sdf, vals_color, vals_normal, vals_semantic, view_matrix,
                          intrinsics, image_color, image_depth, image_normal, image_semantic, mapping3dto2d,
                          mapping3dto2d_num, t, loss_out, clear_grads, workspace, workspace_bytes, (cudaStream_t)stream);
}

int spsg_raycast_backward_loss(const spsg_raycast_params *p, const float *image_color, const float *image_depth,
                               const float *image_semantic, const spsg_loss_targets *t, const float *loss_out,
                               const float *grad_scale, const int32_t *sparse_mapping, const int32_t *mapping3dto2d,
                               const int32_t *mapping3dto2d_num, float *d_color, float *d_depth, float *d_normal,
                               float *d_semantic, void *workspace, size_t workspace_bytes, void *stream) {
    return launch_backward(p, true, image_color, image_depth, nullptr, image_semantic, t, loss_out, grad_scale,
                           sparse_mapping, mapping3dto2d, mapping3dto2d_num, d_color, d_depth, d_normal, d_semantic,
                           workspace, workspace_bytes, (cudaStream_t)stream);
}

static int normals_check(const int64_t *locs, int64_t n, const float *sdf, const int32_t *index, int32_t num_chunks,
                         int32_t dimz, int32_t dimy, int32_t dimx) {
    if (n < 0 || num_chunks <= 0 || dimz <= 0 || dimy <= 0 || dimx <= 0) return fail(SPSG_ERR_INVALID_ARGUMENT, "bad sizes");
    if ((long long)num_chunks * dimz * dimy * dimx >= (1ll << 31)) return fail(SPSG_ERR_INVALID_ARGUMENT, "dense grid exceeds 32-bit indexing");
    if (!index || (n > 0 && (!locs || !sdf))) return fail(SPSG_ERR_INVALID_ARGUMENT, "NULL tensor pointer");
    if (reinterpret_cast<uintptr_t>(locs) & 15u) return fail(SPSG_ERR_INVALID_ARGUMENT, "locs must be 16-byte aligned");
    return SPSG_OK;
}

int spsg_normals_forward(const int64_t *locs, int64_t num_locs, const float *vals_sdf, const float *transform,
                         int32_t *index, int32_t num_chunks, int32_t dimz, int32_t dimy, int32_t dimx, float *normals,
                         void *stream) {
    if (int rc = normals_check(locs, num_locs, vals_sdf, index, num_chunks, dimz, dimy, dimx)) return rc;
    if (num_locs > 0 && !normals) return fail(SPSG_ERR_INVALID_ARGUMENT, "NULL output pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (int rc = spsg_build_index(locs, num_locs, index, num_chunks, dimz, dimy, dimx, stream)) return rc;
    if (num_locs == 0) return SPSG_OK;
    NormalsArgs a;
    memset(&a, 0, sizeof(a));
    a.locs = (const longlong4 *)locs; a.sdf = vals_sdf; a.transform = transform; a.index = index; a.out = normals;
    a.n = num_locs; a.dimx = dimx; a.dimy = dimy; a.dimz = dimz;
    normals_forward_kernel<<<(unsigned)((num_locs + 255) / 256), 256, 0, st>>>(a);
    CUDA_TRY(cudaGetLastError());
    return SPSG_OK;
}

int spsg_normals_backward(const int64_t *locs, int64_t num_locs, const float *vals_sdf, const float *transform,
                          const int32_t *index, int32_t num_chunks, int32_t dimz, int32_t dimy, int32_t dimx,
                          const float *grad_normals, float *scratch_u, float *d_sdf, void *stream) {
    if (int rc = normals_check(locs, num_locs, vals_sdf, index, num_chunks, dimz, dimy, dimx)) return rc;
    if (num_locs == 0) return SPSG_OK;
    if (!grad_normals || !scratch_u || !d_sdf) return fail(SPSG_ERR_INVALID_ARGUMENT, "NULL gradient pointer");
    cudaStream_t st = (cudaStream_t)stream;
    NormalsArgs a;
    memset(&a, 0, sizeof(a));
    a.locs = (const longlong4 *)locs; a.sdf = vals_sdf; a.transform = transform; a.index = index;
    a.grad_out = grad_normals; a.out = scratch_u; a.d_sdf = d_sdf;
    a.n = num_locs; a.dimx = dimx; a.dimy = dimy; a.dimz = dimz;
    const unsigned blocks = (unsigned)((num_locs + 255) / 256);
    normals_backward_u_kernel<<<blocks, 256, 0, st>>>(a);
    CUDA_TRY(cudaGetLastError());
    normals_backward_gather_kernel<<<blocks, 256, 0, st>>>(a);
    CUDA_TRY(cudaGetLastError());
    return SPSG_OK;
}

int spsg_losses2d_forward(const spsg_loss_targets *t, const float *image_color, const float *image_depth,
                          const float *image_semantic, int64_t num_pixels, float *loss_out, void *scratch,
                          size_t scratch_bytes, void *stream) {
    if (!t || !loss_out) return fail(SPSG_ERR_INVALID_ARGUMENT, "NULL loss targets / loss_out");
    if (num_pixels < 0 || num_pixels * 14 >= (1ll << 32)) return fail(SPSG_ERR_INVALID_ARGUMENT, "bad pixel count");
    if ((t->target_depth && !image_depth) || (t->target_color && !image_color) || (t->target_label && !image_semantic))
        return fail(SPSG_ERR_INVALID_ARGUMENT, "a target is set but its rendering is NULL");
    if (image_semantic && (reinterpret_cast<uintptr_t>(image_semantic) & 7u)) return fail(SPSG_ERR_INVALID_ARGUMENT, "image_semantic must be 8-byte aligned");
    const size_t need = (size_t)kLossSlots * 8 * sizeof(double);
    if (!scratch || scratch_bytes < need || (reinterpret_cast<uintptr_t>(scratch) & 7u)) return fail(SPSG_ERR_WORKSPACE_TOO_SMALL, "scratch too small");
    cudaStream_t st = (cudaStream_t)stream;
    CUDA_TRY(cudaMemsetAsync(scratch, 0, need, st));
    Losses2DArgs a;
    memset(&a, 0, sizeof(a));
    a.image_color = image_color; a.image_depth = image_depth; a.image_semantic = image_semantic;
    a.loss = make_loss_args(t, (double *)scratch);
    a.num_pixels = num_pixels;
    if (num_pixels > 0) {
        const unsigned blocks = (unsigned)std::min<long long>((num_pixels + 255) / 256, (long long)sm_count() * 8);
        losses2d_forward_kernel<<<blocks, 256, 0, st>>>(a);
        CUDA_TRY(cudaGetLastError());
    }
    finalize_loss_kernel<<<1, 32, 0, st>>>((const double *)scratch, loss_out, t->weight_depth, t->weight_color_loss,
                                           t->weight_semantic, t->target_depth != nullptr, t->target_color != nullptr,
                                           t->target_label != nullptr);
    CUDA_TRY(cudaGetLastError());
    return SPSG_OK;
}

int spsg_losses2d_backward(const spsg_loss_targets *t, const float *image_color, const float *image_depth,
                           const float *image_semantic, int64_t num_pixels, const float *loss_out,
                           const float *grad_scale, float *d_color, float *d_depth, float *d_semantic, void *stream) {
    if (!t || !loss_out) return fail(SPSG_ERR_INVALID_ARGUMENT, "NULL loss targets / loss_out");
    if (num_pixels < 0 || num_pixels * 14 >= (1ll << 32)) return fail(SPSG_ERR_INVALID_ARGUMENT, "bad pixel count");
    if ((t->target_depth && !image_depth) || (t->target_color && !image_color) || (t->target_label && !image_semantic))
        return fail(SPSG_ERR_INVALID_ARGUMENT, "a target is set but its rendering is NULL");
    if (d_semantic && (reinterpret_cast<uintptr_t>(d_semantic) & 7u)) return fail(SPSG_ERR_INVALID_ARGUMENT, "d_semantic must be 8-byte aligned");
    if (num_pixels == 0) return SPSG_OK;
    BackwardArgs a;
    memset(&a, 0, sizeof(a));
    a.image_color = image_color; a.image_depth = image_depth; a.image_semantic = image_semantic;
    a.loss = make_loss_args(t, nullptr);
    a.loss_out = loss_out;
    a.w_depth = t->weight_depth; a.w_color = t->weight_color_loss; a.w_sem = t->weight_semantic;
    a.grad_scale = grad_scale;
    const unsigned blocks = (unsigned)std::min<long long>((num_pixels + 255) / 256, (long long)sm_count() * 8);
    losses2d_backward_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(a, num_pixels, d_color, d_depth, d_semantic);
    CUDA_TRY(cudaGetLastError());
    return SPSG_OK;
}

}  // extern "C"
