// tiles.cu — moves between the dense column-major n x n GRM (what src/grm/calc.jl:123 allocates) and the packed
// 256 x 256 tiles the multi-GPU paths exchange.  All HBM-bound, all tiny next to the contraction:
//   unpack   gathered FP64 tiles of every rank -> dense matrix, both triangles, bit-identical mirror (calc.jl:128-129)
//   pack     dense upper triangle -> reduce-scatter layout (FP64 K-split)
//   scale    owned reduced FP64 tiles / d (calc.jl:207)
//   scatter  owned packed tiles -> their place in a dense matrix
#include "common.cuh"

namespace {

// block (32 x 8 threads) handles one 32 x 32 sub-block of one 256 x 256 tile: coalesced read of the packed tile,
// coalesced write of the upper element and — through a 32 x 33 shared tile — of its mirror
__global__ void __launch_bounds__(256)
unpack_tiles_kernel(const double *__restrict__ tiles_all, const int2 *__restrict__ tiles, const int32_t *__restrict__ slots,
                    int64_t n, double *__restrict__ G, int64_t ldg, int mirror) {
    __shared__ double t[32][33];
    const int2 tc = tiles[blockIdx.x];
    const double *src = tiles_all + static_cast<int64_t>(slots[blockIdx.x]) * 65536;
    const int sb = blockIdx.y;  // 8 x 8 sub-blocks
    const int r0 = (sb & 7) * 32, c0 = (sb >> 3) * 32;
    if (tc.x == tc.y && r0 > c0) return;  // strictly lower sub-block of a diagonal tile: written as a mirror
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int cl = c0 + ty + 8 * k, rl = r0 + tx;
        const int64_t row = tc.x + rl, col = tc.y + cl;
        double v = 0.0;
        if (row < n && col < n) {
            v = src[cl * 256 + rl];
            if (row <= col) G[row + col * ldg] = v;
        }
        t[ty + 8 * k][tx] = v;  // t[col_local][row_local]
    }
    if (!mirror) return;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        // mirror element (col, row) <- (row, col): consecutive threads walk along `col`
        const int rl = r0 + ty + 8 * k, cl = c0 + tx;
        const int64_t row = tc.x + rl, col = tc.y + cl;
        if (row < n && col < n && row < col) G[col + row * ldg] = t[tx][ty + 8 * k];
    }
}

__global__ void __launch_bounds__(256)
pack_tiles_f64_kernel(const double *__restrict__ G, int64_t n, int64_t ldg, const int2 *__restrict__ tiles,
                      const int32_t *__restrict__ slots, double *__restrict__ packed) {
    const int2 tc = tiles[blockIdx.x];
    double *dst = packed + static_cast<int64_t>(slots[blockIdx.x]) * 65536;
    for (int idx = threadIdx.x; idx < 65536; idx += 256) {
        const int cl = idx >> 8, rl = idx & 255;
        const int64_t row = tc.x + rl, col = tc.y + cl;
        dst[idx] = (row < n && col < n && row <= col) ? G[row + col * ldg] : 0.0;
    }
}

__global__ void __launch_bounds__(256)
scale_tiles_kernel(double *__restrict__ tilebuf, const int2 *__restrict__ tiles, int64_t n, double denom) {
    const int2 tc = tiles[blockIdx.x];
    double *tile = tilebuf + static_cast<int64_t>(blockIdx.x) * 65536;
    for (int idx = threadIdx.x; idx < 65536; idx += 256) {
        const int cl = idx >> 8, rl = idx & 255;
        const int64_t row = tc.x + rl, col = tc.y + cl;
        tile[idx] = (row < n && col < n && row <= col) ? tile[idx] / denom : 0.0;
    }
}

// owned tiles of a DENSE matrix: G[i,j] /= denom for i <= j inside the tile (sharded FP64 calls, rank / world without a
// communicator: the same quantity the int8 path returns)
__global__ void __launch_bounds__(256)
scale_dense_tiles_kernel(double *__restrict__ G, int64_t n, int64_t ldg, const int2 *__restrict__ tiles, double denom) {
    const int2 tc = tiles[blockIdx.x];
    for (int idx = threadIdx.x; idx < 65536; idx += 256) {
        const int cl = idx >> 8, rl = idx & 255;
        const int64_t row = tc.x + rl, col = tc.y + cl;
        if (row < n && col < n && row <= col) G[row + col * ldg] = G[row + col * ldg] / denom;
    }
}

// the rank's own tile list (256 x 256, tile coordinates in entries)
int32_t own_tiles(grm_cuda_ctx *ctx, int64_t n, int32_t rank, int32_t world, const int2 **d_tiles, int *count) {
    TileListCache &tc = ctx->tiles_f64_all;
    if (tc.n != n || tc.rank != rank || tc.world != world || tc.kind != ctx->tune.tile_band) {
        std::vector<int2> order;
        grm_tile_order(n, ctx->tune.tile_band, order);
        int64_t b, e;
        grm_tile_range(static_cast<int64_t>(order.size()), rank, world, &b, &e);
        std::vector<int2> tl;
        for (int64_t t = b; t < e; ++t) tl.push_back(make_int2(order[t].x * GRM_TILE, order[t].y * GRM_TILE));
        GRM_TRY(tc.buf.reserve(std::max<size_t>(tl.size(), 1) * sizeof(int2)));
        if (!tl.empty()) {
            GRM_CUDA_TRY(cudaMemcpyAsync(tc.buf.ptr, tl.data(), tl.size() * sizeof(int2), cudaMemcpyHostToDevice, ctx->stream));
            GRM_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        }
        tc.n = n;
        tc.rank = rank;
        tc.world = world;
        tc.kind = ctx->tune.tile_band;
        tc.count = static_cast<int32_t>(tl.size());
    }
    *d_tiles = tc.buf.as<int2>();
    *count = tc.count;
    return GRM_CUDA_OK;
}

// identity slots 0, 1, 2, ... (scatter of a rank's own packed tiles)
__global__ void iota_kernel(int32_t *out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = i;
}

}  // namespace

int32_t grm_unpack_tiles(grm_cuda_ctx *ctx, const double *d_tiles_all, int64_t n, int32_t world, double *d_G, int64_t ldg) {
    PackedPlanCache *pc;
    GRM_TRY(grm_packed_plan(ctx, n, world, 1, &pc));
    if (pc->total_tiles == 0) return GRM_CUDA_OK;
    const int2 *tiles = pc->buf.as<int2>();
    const int32_t *slots = reinterpret_cast<const int32_t *>(pc->buf.as<int2>() + pc->total_tiles);
    unpack_tiles_kernel<<<dim3(static_cast<unsigned>(pc->total_tiles), 64), 256, 0, ctx->stream>>>(d_tiles_all, tiles, slots,
                                                                                                 n, d_G, ldg, 1);
    GRM_CUDA_TRY(cudaGetLastError());
    ctx->launches++;
    return GRM_CUDA_OK;
}

int32_t grm_pack_tiles_f64(grm_cuda_ctx *ctx, const double *d_G, int64_t n, int64_t ldg, int32_t world, double *d_packed) {
    PackedPlanCache *pc;
    GRM_TRY(grm_packed_plan(ctx, n, world, 1, &pc));
    const int64_t tiles_max = grm_cuda_packed_tiles_max(n, world);
    if (static_cast<int64_t>(pc->total_tiles) != tiles_max * world)
        GRM_CUDA_TRY(cudaMemsetAsync(d_packed, 0, static_cast<size_t>(tiles_max) * world * 65536 * sizeof(double), ctx->stream));
    if (pc->total_tiles == 0) return GRM_CUDA_OK;
    const int2 *tiles = pc->buf.as<int2>();
    const int32_t *slots = reinterpret_cast<const int32_t *>(pc->buf.as<int2>() + pc->total_tiles);
    pack_tiles_f64_kernel<<<static_cast<unsigned>(pc->total_tiles), 256, 0, ctx->stream>>>(d_G, n, ldg, tiles, slots, d_packed);
    GRM_CUDA_TRY(cudaGetLastError());
    ctx->launches++;
    return GRM_CUDA_OK;
}

int32_t grm_scale_tiles_f64(grm_cuda_ctx *ctx, double *d_tiles, int64_t n, int32_t rank, int32_t world, double denom) {
    const int2 *tiles;
    int count;
    GRM_TRY(own_tiles(ctx, n, rank, world, &tiles, &count));
    if (count == 0) return GRM_CUDA_OK;
    scale_tiles_kernel<<<static_cast<unsigned>(count), 256, 0, ctx->stream>>>(d_tiles, tiles, n, denom);
    GRM_CUDA_TRY(cudaGetLastError());
    ctx->launches++;
    return GRM_CUDA_OK;
}

int32_t grm_scale_dense_tiles(grm_cuda_ctx *ctx, double *d_G, int64_t n, int64_t ldg, int32_t rank, int32_t world,
                              double denom) {
    const int2 *tiles;
    int count;
    GRM_TRY(own_tiles(ctx, n, rank, world, &tiles, &count));
    if (count == 0) return GRM_CUDA_OK;
    scale_dense_tiles_kernel<<<static_cast<unsigned>(count), 256, 0, ctx->stream>>>(d_G, n, ldg, tiles, denom);
    GRM_CUDA_TRY(cudaGetLastError());
    ctx->launches++;
    return GRM_CUDA_OK;
}

int32_t grm_scatter_tiles(grm_cuda_ctx *ctx, const double *d_tiles, int64_t n, int32_t rank, int32_t world, double *d_G,
                          int64_t ldg) {
    const int2 *tiles;
    int count;
    GRM_TRY(own_tiles(ctx, n, rank, world, &tiles, &count));
    if (count == 0) return GRM_CUDA_OK;
    GRM_TRY(ctx->misc2.reserve(static_cast<size_t>(count) * sizeof(int32_t)));
    iota_kernel<<<(count + 255) / 256, 256, 0, ctx->stream>>>(ctx->misc2.as<int32_t>(), count);
    GRM_CUDA_TRY(cudaGetLastError());
    unpack_tiles_kernel<<<dim3(static_cast<unsigned>(count), 64), 256, 0, ctx->stream>>>(d_tiles, tiles, ctx->misc2.as<int32_t>(),
                                                                                       n, d_G, ldg, 0);
    GRM_CUDA_TRY(cudaGetLastError());
    ctx->launches += 2;
    return GRM_CUDA_OK;
}
