// spsg_prep.cuh -- per-call preparation kernels: fill, voxel index + dense brick, cell classes + block map.
// Fragment of libspsg_raycast.so: included by spsg_raycast.cu INSIDE its anonymous namespace, in the order listed there
// (one translation unit; every device function is inlined into the kernels that use it).
#pragma once

// ---------------------------------------------------------------------------------------------
// per-call preparation: fill, index + dense brick, cell classes, block map
// ---------------------------------------------------------------------------------------------

// One launch instead of the reference's memsets (kernel.cu:475,483,515): up to kFillRegions word-filled regions.
// (The gradient rows of kernel.cu:557-560 are zeroed by the forward kernel's warps, see raycast_forward_kernel.)
constexpr int kFillRegions = 3;
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
// The same pass over rows that arrive as one uint32 linear cell index each (spsg_pack_locs_host: what this kernel derives
// from the int64 row anyway; 0xffffffff = row outside the grid): 4 instead of 32 bytes per voxel read here -- and, before
// that, carried over PCIe.
__global__ void __launch_bounds__(256) index_packed_kernel(const uint32_t *__restrict__ cells, long long n,
                                                           int32_t *__restrict__ sparse_mapping,
                                                           const float *__restrict__ vals_sdf, float *__restrict__ dense,
                                                           int32_t *__restrict__ num, int views, unsigned long long grid_cells) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (num)
        for (int f = 0; f < views; f++) num[(long long)f * n + i] = 0;
    const unsigned cell = __ldg(cells + i);
    if ((unsigned long long)cell >= grid_cells) return;  // the sentinel, or an index outside this grid
    sparse_mapping[cell] = (int32_t)i;
    if (dense) dense[cell] = __ldg(vals_sdf + i);
}

template <bool kWriteIndex>
__global__ void __launch_bounds__(256) index_kernel(const longlong4 *__restrict__ locs, long long n,
                                                    int32_t *__restrict__ sparse_mapping,
                                                    const float *__restrict__ vals_sdf, float *__restrict__ dense,
                                                    int32_t *__restrict__ num, int views, int dimz, int dimy,
                                                    int dimx, int num_chunks) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (num)
        for (int f = 0; f < views; f++) num[(long long)f * n + i] = 0;
    const longlong4 l = locs[i];  // (z, y, x, chunk)
    const long long z = l.x, y = l.y, x = l.z, b = l.w;
    // A row outside the grid is skipped (the reference would write out of bounds, kernel.cu:355-359).  Duplicate rows are
    // not supported: like the reference's scatter, the last writer wins, independently for the index and the brick.
    if ((unsigned long long)z >= (unsigned long long)dimz || (unsigned long long)y >= (unsigned long long)dimy ||
        (unsigned long long)x >= (unsigned long long)dimx || b < 0 || b >= num_chunks)
        return;
    const long long cell = ((b * dimz + z) * dimy + y) * dimx + x;
    if (kWriteIndex) sparse_mapping[cell] = (int32_t)i;
    if (dense) dense[cell] = __ldg(vals_sdf + i);
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
    // 32-bit cell offsets from the brick's base (the whole brick is below 2^31 cells, checked on the host)
    const unsigned plane = (unsigned)dimy * (unsigned)dimx;
    const unsigned x = (unsigned)(xw * 32 + lane);
    const bool in_x = x < (unsigned)dimx, in_x1 = x + 1u < (unsigned)dimx;
    const unsigned o00 = ((unsigned)chunk * (unsigned)dimz + (unsigned)z0) * plane + (unsigned)y0 * (unsigned)dimx + x;
    // Per voxel row (y0..y0+4, z0..z0+4): every lane loads its voxel and the next one in x (two coalesced loads), so
    // "voxels x and x+1 are both present / positive-class / negative-class" is a lane-local test and one ballot per mask.
    unsigned pres[5][5], posm[5][5], negm[5][5];
#pragma unroll
    for (int dz = 0; dz < 5; dz++)
#pragma unroll
        for (int dy = 0; dy < 5; dy++) {
            const bool row = y0 + dy < dimy && z0 + dz < dimz;
            const float *__restrict__ q = dense + (o00 + (unsigned)dz * plane + (unsigned)dy * (unsigned)dimx);
            const float a = (row && in_x) ? __ldg(q) : CUDART_NAN_F;
            const float n = (row && in_x1) ? __ldg(q + 1) : CUDART_NAN_F;
            pres[dz][dy] = __ballot_sync(kFull, a == a && n == n);
            posm[dz][dy] = __ballot_sync(kFull, a > kTiny && a < kHuge && n > kTiny && n < kHuge);
            negm[dz][dy] = __ballot_sync(kFull, a < -kTiny && a > -kHuge && n < -kTiny && n > -kHuge);
        }
    unsigned any_pos = 0u, any_neg = 0u, any_mix = 0u;
    const unsigned v00 = (unsigned)chunk * (unsigned)vpc + ((unsigned)z0 * (unsigned)dimy + (unsigned)y0) * (unsigned)wpr + (unsigned)xw;
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
                vbits[v00 + ((unsigned)dz * (unsigned)dimy + (unsigned)dy) * (unsigned)wpr] = make_uint2(v & ~vn, v & ~vp);
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
