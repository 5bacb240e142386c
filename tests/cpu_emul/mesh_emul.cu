// mesh_emul.cu -- CPU emulation of the whole OrderParameterMesh device pipeline (host-only program, built with
// nvcc, runs without a GPU).  bin -> counting sort -> spread (per tile, 27 shift rounds) -> merge -> FFT sweeps
// -> gather, using the SAME __host__ __device__ bodies as the kernels in csrc/mesh_kernels.cuh and
// csrc/mesh_fft_kernels.cuh.  tests/test_mesh_emul.py compares the dump against the oracle.
#include <cstring>
#include <algorithm>
#include "../../metadynamics_plugin_b200/csrc/mesh_kernels.cuh"
#include "emul_fft.h"

using namespace metad::mesh;

static unsigned ilog2(unsigned n) { unsigned l = 0; while ((1u << l) < n) ++l; return l; }

// usage: mesh_emul nx ny nz Lx Ly Lz N_global bias lgT ntypes mode... in.bin out.bin
int main(int argc, char** argv) {
    if (argc < 13) { fprintf(stderr, "usage\n"); return 2; }
    Geom g; memset(&g, 0, sizeof g);
    g.nx = atoi(argv[1]); g.ny = atoi(argv[2]); g.nz = atoi(argv[3]);
    const double Ld[3] = {atof(argv[4]), atof(argv[5]), atof(argv[6])};
    const unsigned N_global = (unsigned)atol(argv[7]);
    const double bias = atof(argv[8]);
    g.lgT = atoi(argv[9]);
    const int ntypes = atoi(argv[10]);
    std::vector<float> mode(ntypes);
    for (int i = 0; i < ntypes; ++i) mode[i] = (float)atof(argv[11 + i]);
    const char* fin = argv[11 + ntypes];
    const char* fout = argv[12 + ntypes];
    g.lgx = ilog2(g.nx); g.lgy = ilog2(g.ny); g.lgz = ilog2(g.nz);
    g.ntx = g.nx >> g.lgT; g.nty = g.ny >> g.lgT; g.ntz = g.nz >> g.lgT;
    const unsigned n3[3] = {g.nx, g.ny, g.nz};
    for (int i = 0; i < 3; ++i) {
        g.L[i] = (float)Ld[i]; g.lo[i] = -(g.L[i] / 2.0f);
        g.dlo[i] = -Ld[i] / 2.0; g.dscale[i] = (double)n3[i] / Ld[i];
    }
    FILE* f = fopen(fin, "rb");
    fseek(f, 0, SEEK_END); const long bytes = ftell(f); fseek(f, 0, SEEK_SET);
    const unsigned N = (unsigned)(bytes / 16);
    std::vector<float4> postype(N);
    if (fread(postype.data(), 16, N, f) != N) return 2;
    fclose(f);
    const size_t M = (size_t)g.nx * g.ny * g.nz;

    // ---- bin (mesh_bin_kernel)
    std::vector<unsigned> keys(N), ranks(N), count(M, 0), start(M + 1), perm(N);
    double sums[2] = {0, 0};
    for (unsigned i = 0; i < N; ++i) {
        const float4 p = postype[i];
        const unsigned ix = cell_coord(p.x, g.lo[0], g.L[0], g.nx), iy = cell_coord(p.y, g.lo[1], g.L[1], g.ny),
                       iz = cell_coord(p.z, g.lo[2], g.L[2], g.nz);
        keys[i] = key_of(ix, iy, iz, g);
        ranks[i] = count[keys[i]]++;
        int t; memcpy(&t, &p.w, 4);
        sums[0] += (double)mode[t] * mode[t]; sums[1] += (double)mode[t];
    }
    // ---- scan
    unsigned run = 0;
    for (size_t c = 0; c < M; ++c) { start[c] = run; run += count[c]; }
    start[M] = run;
    // ---- reorder (emulate a scrambled rank order to exercise the deterministic selection: reverse ranks)
    std::vector<float4> sorted(N);
    for (unsigned i = 0; i < N; ++i) {
        float4 p = postype[i];
        int t; memcpy(&t, &p.w, 4);
        p.w = mode[t];
        const unsigned k = keys[i];
        const unsigned dst = start[k] + (count[k] - 1 - ranks[i]);
        sorted[dst] = p; perm[dst] = i;
    }
    // ---- spread (mesh_spread_kernel, thread per cell, 27 rounds)
    const unsigned T = 1u << g.lgT, P = T + 2, P3 = P * P * P, NC = T * T * T, ntiles = num_tiles(g);
    std::vector<float> scratch((size_t)ntiles * P3);
    std::vector<float> tile(P3);
    for (unsigned tile_id = 0; tile_id < ntiles; ++tile_id) {
        std::fill(tile.begin(), tile.end(), 0.f);
        for (unsigned lc = 0; lc < NC; ++lc) {
            const unsigned key = (tile_id << (3 * g.lgT)) + lc;
            unsigned ix, iy, iz; cell_of_key(key, g, ix, iy, iz);
            const unsigned lx = lc & (T - 1), ly = (lc >> g.lgT) & (T - 1), lz = lc >> (2 * g.lgT);
            const unsigned s = start[key], e = start[key + 1];
            float acc[27]; for (int r = 0; r < 27; ++r) acc[r] = 0.f;
            long long last = -1;
            for (unsigned it = s; it < e; ++it) {
                unsigned best = 0xffffffffu, bj = s;
                for (unsigned j = s; j < e; ++j) { const unsigned pj = perm[j]; if ((long long)pj > last && pj < best) { best = pj; bj = j; } }
                last = best;
                spread_accumulate(sorted[bj], ix, iy, iz, g, acc);
            }
            if (e > s)
                for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) for (int k = 0; k < 3; ++k)
                    tile[padded_index(lx, ly, lz, i, j, k, P)] += acc[(i * 3 + j) * 3 + k];
        }
        std::copy(tile.begin(), tile.end(), scratch.begin() + (size_t)tile_id * P3);
    }
    // ---- merge
    std::vector<float> rho(M), buf(M);
    const float mean = (float)(sums[1] / (double)M);
    for (size_t c = 0; c < M; ++c) {
        const unsigned x = (unsigned)(c & (g.nx - 1)), y = (unsigned)((c >> g.lgx) & (g.ny - 1)), z = (unsigned)(c >> (g.lgx + g.lgy));
        rho[c] = merge_cell(scratch.data(), x, y, z, g);
        buf[c] = rho[c] - mean;
    }
    // ---- FFT sweeps
    float2* b2 = reinterpret_cast<float2*>(buf.data());
    const unsigned nxh = g.nx / 2;
    const float inv_n = (float)(1.0 / (double)N_global);
    const float d = (float)(0.5 * sums[0] / (double)N_global / (double)N_global);
    double e = 0.0;
    DISPATCH(nxh, x_fwd<LL>(b2, g.ny * g.nz));
    DISPATCH(g.ny, (y_pass<LL, -1>(b2, nxh, g.nz)));
    DISPATCH(g.nz, e += z_plane0<LL>(b2, g.nx, g.ny, inv_n, d));
    DISPATCH(g.nz, e += z_fused<LL>(b2, g.nx, g.ny, inv_n, d));
    DISPATCH(g.ny, (y_pass<LL, +1>(b2, nxh, g.nz)));
    DISPATCH(nxh, x_inv<LL>(b2, g.ny * g.nz));
    const double cv = 0.5 * e;
    // ---- gather (mesh_gather_kernel)
    ForceParams fp; memset(&fp, 0, sizeof fp);
    fp.nb1[0] = (float)((double)g.nx / Ld[0]); fp.nb2[1] = (float)((double)g.ny / Ld[1]); fp.nb3[2] = (float)((double)g.nz / Ld[2]);
    fp.two_over_n = 2.0 / (double)N_global;
    std::vector<float4> force(N);
    for (unsigned tile_id = 0; tile_id < ntiles; ++tile_id) {
        const unsigned tx = tile_id % g.ntx, ty = (tile_id / g.ntx) % g.nty, tz = tile_id / (g.ntx * g.nty);
        for (unsigned i = 0; i < P3; ++i) {
            const unsigned px = i % P, py = (i / P) % P, pz = i / (P * P);
            const unsigned x = ((tx << g.lgT) + px + g.nx - 1) & (g.nx - 1), y = ((ty << g.lgT) + py + g.ny - 1) & (g.ny - 1),
                           z = ((tz << g.lgT) + pz + g.nz - 1) & (g.nz - 1);
            tile[i] = buf[(size_t)x + (size_t)g.nx * (y + (size_t)g.ny * z)];
        }
        const unsigned s = start[tile_id << (3 * g.lgT)], en = start[(tile_id + 1) << (3 * g.lgT)];
        for (unsigned j = s; j < en; ++j) {
            const float4 p = sorted[j];
            const unsigned ix = cell_coord(p.x, g.lo[0], g.L[0], g.nx), iy = cell_coord(p.y, g.lo[1], g.L[1], g.ny),
                           iz = cell_coord(p.z, g.lo[2], g.L[2], g.nz);
            force[perm[j]] = gather_force(p, ix, iy, iz, tile.data(), g, fp, bias);
        }
    }
    // ---- dump: cv, mode_sq, rho[M], inv[M], force[4N], cells[3N]
    f = fopen(fout, "wb");
    fwrite(&cv, 8, 1, f); fwrite(&sums[0], 8, 1, f);
    fwrite(rho.data(), 4, M, f); fwrite(buf.data(), 4, M, f); fwrite(force.data(), 16, N, f);
    std::vector<int> cells(3 * (size_t)N);
    for (unsigned i = 0; i < N; ++i) { unsigned ix, iy, iz; cell_of_key(keys[i], g, ix, iy, iz); cells[3 * i] = ix; cells[3 * i + 1] = iy; cells[3 * i + 2] = iz; }
    fwrite(cells.data(), 4, cells.size(), f);
    fclose(f);
    return 0;
}
