// spsg_fp32.cuh -- exact fp32 building blocks: voxel rounding, ray set-up, the step recurrence, trilinear sampling.
// Fragment of libspsg_raycast.so: included by spsg_raycast.cu INSIDE its anonymous namespace, in the order listed there
// (one translation unit; every device function is inlined into the kernels that use it).
#pragma once

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
// The per-binade constants (they depend on inc only) are tabulated once per CTA in shared memory: entry e describes the binade [2^e, 2^(e+1)) as (d, 1/d, lim, top); lim = -inf marks a binade where the closed
// form is not usable (entry 32 serves every ray parameter outside [1, 2^32); its "top" is 1).
constexpr int kStepEntries = 33;

__device__ __forceinline__ void step_table_fill(float4 *table, int e, float inc) {
    float4 t = make_float4(inc, 0.0f, -CUDART_INF_F, 1.0f);
    if (e < 32) {
        const float lo = __uint_as_float((unsigned)(e + 127) << 23), hi = __fmul_rn(lo, 2.0f);
        t.w = hi;
        const float u = __fmul_rn(lo, 1.1920928955078125e-07f);  // 2^(e-23)
        const float d = __fadd_rn(__fadd_rn(lo, inc), -lo);       // inc on the u grid
        const float rem = __fadd_rn(inc, -d);                     // exact remainder
        const bool regular = (lo <= 8388608.0f) && (inc > 0.0f) && (inc <= 0.25f * lo) && (d > 0.0f) &&
                             (__fmul_rn(fabsf(rem), 2.0f) != u);
        if (regular) t = make_float4(d, rcp_approx(d), __fadd_rn(hi, -__fmul_rn(inc, 2.0f)), hi);
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
        // real adds: the steps of an irregular binade, or the last ones below the top of this binade and the one that
        // crosses it (no table look-up until the ray is in the next binade)
        do {
            ray = __fadd_rn(ray, inc);
            --n;
        } while (n > 0 && ray < t.w);
        if (n == 0) return ray;
    }
}

struct Volume {
    const int32_t *__restrict__ index;  // sparse_mapping (all chunks)
    const float *__restrict__ sdf;      // vals_sdf
    const float *__restrict__ dense;    // dense SDF brick (all chunks; NaN = absent)
    unsigned cell0;                     // first cell of this chunk: chunk * dimz * dimy * dimx (all chunks < 2^31 cells)
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
// evaluation (corner coordinates rounded like the reference, index -> value double gather).  Out of line and by value:
// nothing the march keeps in registers has its address taken for this rare path.
struct SdfSample {
    float dist;
    bool valid;
};
__device__ __noinline__ SdfSample sample_sdf_exact(const int32_t *__restrict__ index, const float *__restrict__ sdf,
                                                   int dimx, int dimy, int dimz, float px, float py, float pz) {
    SdfSample out = {0.0f, false};
    const float qx = __fadd_rn(px, -0.5f), qy = __fadd_rn(py, -0.5f), qz = __fadd_rn(pz, -0.5f);
    const int x0 = round_voxel(qx), y0 = round_voxel(qy), z0 = round_voxel(qz);
    const int x1 = round_voxel(__fadd_rn(qx, 1.0f)), y1 = round_voxel(__fadd_rn(qy, 1.0f)),
              z1 = round_voxel(__fadd_rn(qz, 1.0f));
    if ((x0 | y0 | z0 | x1 | y1 | z1) < 0 || x0 >= dimx || x1 >= dimx || y0 >= dimy || y1 >= dimy || z0 >= dimz ||
        z1 >= dimz)
        return out;
    const int r00 = (z0 * dimy + y0) * dimx, r10 = (z0 * dimy + y1) * dimx;
    const int r01 = (z1 * dimy + y0) * dimx, r11 = (z1 * dimy + y1) * dimx;
    const int i000 = __ldg(index + r00 + x0), i100 = __ldg(index + r00 + x1);
    const int i010 = __ldg(index + r10 + x0), i110 = __ldg(index + r10 + x1);
    const int i001 = __ldg(index + r01 + x0), i101 = __ldg(index + r01 + x1);
    const int i011 = __ldg(index + r11 + x0), i111 = __ldg(index + r11 + x1);
    if ((i000 | i100 | i010 | i110 | i001 | i101 | i011 | i111) < 0) return out;
    const float wx = __fadd_rn(px, -floorf(px)), wy = __fadd_rn(py, -floorf(py)), wz = __fadd_rn(pz, -floorf(pz));
    out.dist = trilerp(wx, wy, wz, __ldg(sdf + i000), __ldg(sdf + i100), __ldg(sdf + i010), __ldg(sdf + i001),
                       __ldg(sdf + i110), __ldg(sdf + i011), __ldg(sdf + i101), __ldg(sdf + i111));
    out.valid = true;
    return out;
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
    // 32-bit offsets from the uniform base pointer: the whole brick is below 2^31 cells (checked on the host)
    const unsigned sy = (unsigned)v.dimx, sz = (unsigned)(v.dimx * v.dimy);
    const unsigned o00 = v.cell0 + ((unsigned)iz * (unsigned)v.dimy + (unsigned)iy) * sy + (unsigned)ix;
    const unsigned o10 = o00 + sy, o01 = o00 + sz, o11 = o01 + sy;
    const float *__restrict__ b00 = v.dense + o00, *__restrict__ b10 = v.dense + o10;
    const float *__restrict__ b01 = v.dense + o01, *__restrict__ b11 = v.dense + o11;
    const float v000 = __ldg(b00), v100 = __ldg(b00 + 1), v010 = __ldg(b10), v110 = __ldg(b10 + 1);
    const float v001 = __ldg(b01), v101 = __ldg(b01 + 1), v011 = __ldg(b11), v111 = __ldg(b11 + 1);
    return trilerp(wx, wy, wz, v000, v100, v010, v001, v110, v011, v101, v111);
}

// One sample at p: dist when all 8 corners are present, NaN otherwise.  (NaN never satisfies a sign test, so callers
// treat "valid with a NaN value" and "invalid" alike -- see the note above.)
__device__ __forceinline__ float sample_sdf(const Volume &v, bool fast_ok, float px, float py, float pz) {
    const float fx = floorf(px), fy = floorf(py), fz = floorf(pz);
    const float wx = __fadd_rn(px, -fx), wy = __fadd_rn(py, -fy), wz = __fadd_rn(pz, -fz);
    const int ix = __float2int_rz(fx), iy = __float2int_rz(fy), iz = __float2int_rz(fz);
    const float g = frac_guard(v.guard, ix, iy, iz);
    const bool fast = fast_ok && fminf(wx, fminf(wy, wz)) >= g && fmaxf(wx, fmaxf(wy, wz)) <= 1.0f - g &&
                      (ix | iy | iz) >= 0 && ix + 1 < v.dimx && iy + 1 < v.dimy && iz + 1 < v.dimz;
    if (fast) return sample_dense(v, ix, iy, iz, wx, wy, wz);
    const SdfSample s = sample_sdf_exact(v.index + v.cell0, v.sdf, v.dimx, v.dimy, v.dimz, px, py, pz);
    return s.valid ? s.dist : __int_as_float(0x7fc00000);
}
