// mesh_fft_xy.cuh -- x and y sweeps of the mesh FFT fused on whole z planes (unsharded plans, sm_100a).
//
// The five-sweep pipeline (mesh_fft_kernels.cuh) moves the half spectrum through L2 / HBM between the x and the y sweep,
// forward and inverse: two round trips of 8 M bytes that a plane-local 2-D transform does not need.  A z plane of the
// 256^2 mesh is 256 KB -- more than one CTA's shared memory, so one plane is transformed by a thread-block CLUSTER of C
// CTAs that exchange data through distributed shared memory:
//   forward   CTA c transforms rows y in [c LY/C, (c+1) LY/C) along x (R2C, the round-1 engine), stores every coefficient
//             straight into the shared memory of the CTA that owns its kx column (remote stores), cluster barrier, each CTA
//             transforms its LC/C columns along y and writes them out -- [z][y][kx], the layout the fused z sweep reads;
//   inverse   each CTA loads its columns, transforms along y, cluster barrier, fetches the other CTAs' columns of ITS rows
//             (remote loads into registers, cluster barrier, stores into the slots those CTAs have just given up), then the
//             C2R x transform of its rows and the real rows go out.
// Replaces cufftExecC2C's x / y passes of the reference (OrderParameterMeshGPU.cc:256, 322-325) exactly like the separate
// sweeps; same arithmetic per line (line_fft, r2c_pair / c2r_pair), so results agree with them to rounding.
#pragma once
#include "mesh_fft_kernels.cuh"
#ifdef __CUDACC__
#include <cooperative_groups.h>
#endif

namespace metad {
namespace fft {

template <int LC, int LY, int C, int NT_ = 512> struct XYPlan {
    static constexpr int NT = NT_;                      // threads per CTA
    static constexpr int XT = LC / kE;                  // threads per x line
    static constexpr int GX = NT / (kLines * XT);       // tiles of 16 rows transformed side by side along x
    static constexpr int YT = LY / kE;                  // threads per y line
    static constexpr int GY = NT / (kLines * YT);       // tiles of 16 columns transformed side by side along y
    static constexpr int NB = LC / C;                   // kx columns owned by one CTA
    static constexpr int ROWS = LY / C;                 // rows whose x transform one CTA performs
    static constexpr int XP = LC + 1, YP = LY + 1;      // pitches of an x tile row / a column line (odd: conflict-free)
    static constexpr int XCH = ((C - 1) * NB * ROWS + NT - 1) / NT;      // exchanged coefficients per thread (inverse)
    static constexpr size_t smem_bytes = sizeof(float2) * ((size_t)NB * YP + (size_t)GX * kLines * XP + 2 * LC + LY);
    static constexpr int MINB = (227 * 1024) / (int)smem_bytes >= 3 ? 3 : ((227 * 1024) / (int)smem_bytes >= 2 ? 2 : 1);    // CTAs per SM the shared memory allows
    static_assert(GX >= 1 && GY >= 1, "line too long for the CTA");
    static_assert(ROWS % (kLines * GX) == 0, "rows per CTA must be a multiple of the rows transformed per iteration");
    static_assert(NB % (kLines * GY) == 0, "columns per CTA must be a multiple of the columns transformed per iteration");
    static_assert((NB % 2) == 0, "two columns per 16-byte store");
};

#ifdef __CUDACC__
namespace cg = cooperative_groups;

// (32-bit density only: a plan that accumulates in 64 bits takes the separate sweeps)
template <int LC, int LY, int C, int NT>
__global__ void __launch_bounds__(NT, (XYPlan<LC, LY, C, NT>::MINB))
fft_xy_fwd_kernel(DensityIn in, const float2* __restrict__ g_twx /* 2 LC */, const float2* __restrict__ g_twy /* LY */, float2* __restrict__ out) {
    using P = XYPlan<LC, LY, C, NT>;
    extern __shared__ __align__(16) float2 smem[];
    float2* B = smem;                                        // [NB][YP]: my kx columns, all y
    float2* X = B + P::NB * P::YP;                           // [GX][kLines][XP]: x work tiles
    float2* s_twx = X + P::GX * kLines * P::XP;
    float2* s_twy = s_twx + 2 * LC;
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned crank = C > 1 ? cluster.block_rank() : 0u;
    const unsigned z = blockIdx.y;
    load_twiddles<2 * LC>(s_twx, g_twx);
    load_twiddles<LY>(s_twy, g_twy);
    pdl_wait(); pdl_trigger();
    const float inv_scale = __ldg(in.d_fx + 1);
    const float mean = (float)(in.d_sums[1] * in.inv_cells);
    constexpr int nthr = kLines * P::XT;                     // threads of one x group
    const int g = threadIdx.x / nthr, lt = threadIdx.x % nthr;
    float2* tile = X + g * kLines * P::XP;
    for (int it = 0; it < P::ROWS / (kLines * P::GX); ++it) {
        const unsigned y0 = crank * P::ROWS + (it * P::GX + g) * kLines;          // first row (y) of this group's tile
        const size_t row0 = (size_t)z * LY + y0;
        int4 vin[kE / 2];
#pragma unroll
        for (int q = 0; q < kE / 2; ++q) {
            const int idx = lt + q * nthr;
            vin[q] = __ldcs(reinterpret_cast<const int4*>(in.mesh + (row0 + idx / (LC / 2)) * LC) + idx % (LC / 2));
        }
        if (in.range_counter) {
            int m = 0;
#pragma unroll
            for (int q = 0; q < kE / 2; ++q) m = max(m, max(max(abs(vin[q].x), abs(vin[q].y)), max(abs(vin[q].z), abs(vin[q].w))));
            if (m > (1 << 30)) { atomicAdd(in.range_counter, 1u); if (in.h_range) *reinterpret_cast<volatile unsigned*>(in.h_range) = 1u; }
        }
#pragma unroll
        for (int q2 = 0; q2 < kE; ++q2) {
            const int idx = lt + (q2 >> 1) * nthr;
            const int w = idx / (LC / 2), l = 2 * (idx % (LC / 2)) + (q2 & 1);
            const int2 v = (q2 & 1) ? make_int2(vin[q2 >> 1].z, vin[q2 >> 1].w) : make_int2(vin[q2 >> 1].x, vin[q2 >> 1].y);
            float2 r = density_to_float(v, inv_scale);
            r.x -= mean; r.y -= mean;
            tile[LayoutRow::addr(w, l, LC)] = r;
        }
        {   // the accumulator is empty again for the next spread
            const int4 z4 = make_int4(0, 0, 0, 0);
#pragma unroll
            for (int q = 0; q < kE / 2; ++q) {
                const int idx = lt + q * nthr;
                in.zero[(row0 + idx / (LC / 2)) * (LC / 2) + idx % (LC / 2)] = z4;
            }
        }
        __syncthreads();
        line_fft<LC, -1, 2 * LC, LayoutRow>(tile, lt & (kLines - 1), lt / kLines, s_twx);     // block barriers inside: the groups run in lockstep
        for (int idx = lt; idx < kLines * (LC / 2 + 1); idx += nthr) {
            const int ww = idx & (kLines - 1), k = idx / kLines;
            r2c_pair(tile[LayoutRow::addr(ww, k, LC)], tile[LayoutRow::addr(ww, (LC - k) % LC, LC)], k, LC, s_twx[k]);
        }
        __syncthreads();
        // every coefficient goes to the CTA that owns its kx column: 16 consecutive y of one column are 128 contiguous bytes
        for (int idx = lt; idx < kLines * LC; idx += nthr) {
            const int ww = idx & (kLines - 1), l = idx / kLines;
            float2* dstB = C > 1 ? cluster.map_shared_rank(B, l / P::NB) : B;
            dstB[(l % P::NB) * P::YP + y0 + ww] = tile[LayoutRow::addr(ww, l, LC)];
        }
        __syncthreads();
    }
    if (C > 1) cluster.sync(); else __syncthreads();
    constexpr int nthy = kLines * P::YT;
    const int gy = threadIdx.x / nthy, lty = threadIdx.x % nthy;
    for (int it = 0; it < P::NB / (kLines * P::GY); ++it)
        line_fft<LY, -1, LY, LayoutRow>(B + (it * P::GY + gy) * kLines * P::YP, lty & (kLines - 1), lty / kLines, s_twy);
    // my kx range of every row: [z][y][kx]
    float2* dst = out + (size_t)z * LY * LC + crank * P::NB;
    for (int idx = threadIdx.x; idx < LY * (P::NB / 2); idx += P::NT) {
        const int y = idx / (P::NB / 2), c2 = idx % (P::NB / 2);
        const float2 a = B[(2 * c2) * P::YP + y], b = B[(2 * c2 + 1) * P::YP + y];
        *reinterpret_cast<float4*>(dst + (size_t)y * LC + 2 * c2) = make_float4(a.x, a.y, b.x, b.y);
    }
    if (C > 1) cluster.sync();       // nobody leaves while its shared memory may still be written (it is not, after the first barrier; cheap safety)
}

template <int LC, int LY, int C, int NT>
__global__ void __launch_bounds__(NT, (XYPlan<LC, LY, C, NT>::MINB))
fft_xy_inv_kernel(float2* buf /* in: [z][y][kx] spectrum of G after the inverse z sweep; out: the real rows, in place */,
                  const float2* __restrict__ g_twx, const float2* __restrict__ g_twy) {
    using P = XYPlan<LC, LY, C, NT>;
    extern __shared__ __align__(16) float2 smem[];
    float2* B = smem;
    float2* X = B + P::NB * P::YP;
    float2* s_twx = X + P::GX * kLines * P::XP;
    float2* s_twy = s_twx + 2 * LC;
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned crank = C > 1 ? cluster.block_rank() : 0u;
    const unsigned z = blockIdx.y;
    load_twiddles<2 * LC>(s_twx, g_twx);
    load_twiddles<LY>(s_twy, g_twy);
    pdl_wait(); pdl_trigger();
    float2* plane = buf + (size_t)z * LY * LC;
    {   // my kx columns of every row
        const float2* src = plane + crank * P::NB;
        for (int idx = threadIdx.x; idx < LY * (P::NB / 2); idx += P::NT) {
            const int y = idx / (P::NB / 2), c2 = idx % (P::NB / 2);
            const float4 v = *reinterpret_cast<const float4*>(src + (size_t)y * LC + 2 * c2);
            B[(2 * c2) * P::YP + y] = make_float2(v.x, v.y);
            B[(2 * c2 + 1) * P::YP + y] = make_float2(v.z, v.w);
        }
    }
    __syncthreads();
    constexpr int nthy = kLines * P::YT;
    const int gy = threadIdx.x / nthy, lty = threadIdx.x % nthy;
    for (int it = 0; it < P::NB / (kLines * P::GY); ++it)
        line_fft<LY, +1, LY, LayoutRow>(B + (it * P::GY + gy) * kLines * P::YP, lty & (kLines - 1), lty / kLines, s_twy);
    if constexpr (C > 1) {
        // Exchange: CTA c needs, for ITS rows, the columns of every other CTA q: B_q[line][c ROWS + r].  It fetches them into
        // registers, waits until everybody has fetched, and parks them in its own slots [q ROWS + r] (the rows of CTA q,
        // which q has fetched by then).  Afterwards coefficient kx = q NB + line of my row r sits at B[line][q ROWS + r] for
        // every q, my own columns included.  The first barrier also orders all loads of the plane before any row is written.
        cluster.sync();
        float2 xr[P::XCH > 0 ? P::XCH : 1];
#pragma unroll
        for (int k = 0; k < P::XCH; ++k) {
            const int idx = threadIdx.x + k * P::NT;                 // (qq, line, r), r fastest
            xr[k] = make_float2(0.f, 0.f);
            if (idx < (C - 1) * P::NB * P::ROWS) {
                const int r = idx % P::ROWS, line = (idx / P::ROWS) % P::NB, qq = idx / (P::ROWS * P::NB);
                const unsigned q = qq + (qq >= (int)crank ? 1 : 0);
                xr[k] = cluster.map_shared_rank(B, q)[line * P::YP + crank * P::ROWS + r];
            }
        }
        cluster.sync();
#pragma unroll
        for (int k = 0; k < P::XCH; ++k) {
            const int idx = threadIdx.x + k * P::NT;
            if (idx < (C - 1) * P::NB * P::ROWS) {
                const int r = idx % P::ROWS, line = (idx / P::ROWS) % P::NB, qq = idx / (P::ROWS * P::NB);
                const unsigned q = qq + (qq >= (int)crank ? 1 : 0);
                B[line * P::YP + q * P::ROWS + r] = xr[k];
            }
        }
    }
    __syncthreads();
    constexpr int nthr = kLines * P::XT;
    const int g = threadIdx.x / nthr, lt = threadIdx.x % nthr;
    float2* tile = X + g * kLines * P::XP;
    for (int it = 0; it < P::ROWS / (kLines * P::GX); ++it) {
        const unsigned r0 = (it * P::GX + g) * kLines;               // first of my rows in this tile
        for (int idx = lt; idx < kLines * LC; idx += nthr) {
            const int ww = idx & (kLines - 1), l = idx / kLines;
            tile[LayoutRow::addr(ww, l, LC)] = B[(l % P::NB) * P::YP + (l / P::NB) * P::ROWS + r0 + ww];
        }
        __syncthreads();
        for (int idx = lt; idx < kLines * (LC / 2 + 1); idx += nthr) {
            const int ww = idx & (kLines - 1), k = idx / kLines;
            c2r_pair(tile[LayoutRow::addr(ww, k, LC)], tile[LayoutRow::addr(ww, (LC - k) % LC, LC)], k, LC, s_twx[k]);
        }
        __syncthreads();
        line_fft<LC, +1, 2 * LC, LayoutRow>(tile, lt & (kLines - 1), lt / kLines, s_twx);
        float2* dst = plane + (size_t)(crank * P::ROWS + r0) * LC;
        for (int idx = lt; idx < kLines * LC; idx += nthr) {
            const int ww = idx / LC, l = idx % LC;
            dst[(size_t)ww * LC + l] = tile[LayoutRow::addr(ww, l, LC)];
        }
        __syncthreads();
    }
    if (C > 1) cluster.sync();
}
#endif  // __CUDACC__

}  // namespace fft
}  // namespace metad
