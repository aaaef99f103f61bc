// spsg_occ.cuh -- the occupancy raycast.
// Fragment of libspsg_raycast.so: included by spsg_raycast.cu INSIDE its anonymous namespace, in the order listed there
// (one translation unit; every device function is inlined into the kernels that use it).
#pragma once

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
    __shared__ float4 s_steps[kStepEntries];
    if (threadIdx.x < kStepEntries) step_table_fill(s_steps, threadIdx.x, a.inc);
    __syncthreads();
    if (ux >= (unsigned)a.width || uy >= (unsigned)a.height) return;
    const Ray r = setup_ray(a.view_matrix + (size_t)img * 16, a.intrinsics + (size_t)img * 4, ux, uy, a.depth_min,
                            a.depth_max);
    const uint8_t *__restrict__ occ = a.occ3d + (size_t)img * a.dimz * a.dimy * a.dimx;
    const float inv_inc = rcp_approx(a.inc);
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
                const int want = max(1, min(__float2int_rd((tin - margin - ray) * inv_inc) - 1, 1 << 22));
                ray = step_advance(s_steps, a.inc, ray, want);
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
