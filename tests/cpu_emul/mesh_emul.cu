// mesh_emul.cu -- CPU emulation of the whole OrderParameterMesh device pipeline (host-only program, built with
// nvcc, runs without a GPU).  bin -> counting sort (place + stable reorder) -> spread (thread per cell column,
// rolling accumulators, replica exchange) -> merge -> FFT sweeps -> gather, using the SAME __host__ __device__
// bodies as the kernels in csrc/mesh_kernels.cuh and csrc/mesh_fft_kernels.cuh, with the kernels' loop structure.
// tests/test_emulation.py compares the dump against the oracle.
#include <cstring>
#include <algorithm>
#include "../../metadynamics_plugin_b200/csrc/mesh_kernels.cuh"
#include "emul_fft.h"

using namespace metad::mesh;

// usage: mesh_emul nx ny nz Lx Ly Lz N_global bias lgT ntypes mode... in.bin out.bin
int main(int argc, char** argv) {
    if (argc < 13) { fprintf(stderr, "usage\n"); return 2; }
    Geom g; memset(&g, 0, sizeof g);
    const unsigned nx = atoi(argv[1]), ny = atoi(argv[2]), nz = atoi(argv[3]);
    const double Ld[3] = {atof(argv[4]), atof(argv[5]), atof(argv[6])};
    const unsigned N_global = (unsigned)atol(argv[7]);
    const double bias = atof(argv[8]);
    geom_set_dims(g, nx, ny, nz, atoi(argv[9]));
    const int ntypes = atoi(argv[10]);
    std::vector<float> mode(ntypes);
    for (int i = 0; i < ntypes; ++i) mode[i] = (float)atof(argv[11 + i]);
    const char* fin = argv[11 + ntypes];
    const char* fout = argv[12 + ntypes];
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
    std::vector<unsigned> keys(N), ranks(N), count(M, 0), start(M + 1), perm(N), slot(N), skey(N);
    double sums[2] = {0, 0};
    for (unsigned i = 0; i < N; ++i) {
        const float4 p = postype[i];
        const unsigned ix = cell_coord(p.x, g.lo[0], g.L[0], g.nx), iy = cell_coord(p.y, g.lo[1], g.L[1], g.ny),
                       iz = cell_coord(p.z, g.lo[2], g.L[2], g.nz);
        keys[i] = key_of(ix, iy, iz, g);
        unsigned cx, cy, cz; cell_of_key(keys[i], g, cx, cy, cz);
        if (cx != ix || cy != iy || cz != iz) { fprintf(stderr, "key round trip failed\n"); return 3; }
        ranks[i] = count[keys[i]]++;
        int t; memcpy(&t, &p.w, 4);
        sums[0] += (double)mode[t] * mode[t]; sums[1] += (double)mode[t];
    }
    // ---- scan
    unsigned run = 0;
    for (size_t c = 0; c < M; ++c) { start[c] = run; run += count[c]; }
    start[M] = run;
    // ---- place (mesh_place_kernel); the arrival rank of the device is arbitrary: emulate it reversed
    for (unsigned i = 0; i < N; ++i) slot[start[keys[i]] + (count[keys[i]] - 1 - ranks[i])] = i;
    // ---- stable reorder (mesh_reorder_kernel, one thread per slot)
    std::vector<float4> sorted(N);
    for (unsigned j = 0; j < N; ++j) {
        const unsigned i = slot[j], key = keys[i], s = start[key], e = start[key + 1];
        unsigned dst = s;
        for (unsigned m = s; m < e; ++m) dst += slot[m] < i ? 1u : 0u;
        float4 p = postype[i];
        int t; memcpy(&t, &p.w, 4);
        p.w = mode[t];
        sorted[dst] = p; perm[dst] = i; skey[dst] = key;
    }
    for (unsigned j = 1; j < N; ++j)
        if (skey[j] == skey[j - 1] && perm[j] < perm[j - 1]) { fprintf(stderr, "order inside a cell is not stable\n"); return 3; }
    // ---- spread (mesh_spread_kernel): thread per column, rolling accumulators, replica exchange
    const unsigned T = 1u << g.lgT, P = T + 2, PP = P * P, P3 = PP * P, NT = T * T, ntiles = num_tiles(g);
    const unsigned CAP = 8;     // tiny chunk capacity: exercises the multi-chunk path
    std::vector<float> scratch((size_t)ntiles * P3);
    std::vector<float> tile(P3);
    for (unsigned tile_id = 0; tile_id < ntiles; ++tile_id) {
        unsigned tx, ty, tz; tile_coords(tile_id, g, tx, ty, tz);
        std::vector<float> acc(NT * 27, 0.f);
        std::vector<float> wbuf(9 * CAP), rep(9 * PP);
        float* out = scratch.data() + (size_t)tile_id * P3;
        for (unsigned lz = 0; lz < T + 2; ++lz) {
            if (lz < T) {
                const unsigned key0 = (tile_id << (3 * g.lgT)) + lz * NT;
                const unsigned s_plane = start[key0], e_plane = start[key0 + NT];
                for (unsigned c0 = s_plane; c0 < e_plane; c0 += CAP) {
                    const unsigned c1 = std::min(c0 + CAP, e_plane);
                    for (unsigned j = c0; j < c1; ++j) {                       // phase 1
                        const unsigned local = skey[j] & (NT - 1);
                        float w[9];
                        spread_weights(sorted[j], (tx << g.lgT) + (local & (T - 1)), (ty << g.lgT) + (local >> g.lgT), (tz << g.lgT) + lz, g, w);
                        for (int c = 0; c < 9; ++c) wbuf[c * CAP + (j - c0)] = w[c];
                    }
                    for (unsigned tid = 0; tid < NT; ++tid) {                  // phase 2
                        const unsigned s = start[key0 + tid], e = start[key0 + tid + 1];
                        const unsigned a = std::max(s, c0), b = std::min(e, c1);
                        float (&ac)[27] = *reinterpret_cast<float (*)[27]>(&acc[tid * 27]);
                        for (unsigned j = a; j < b; ++j) {
                            float w[9]; for (int c = 0; c < 9; ++c) w[c] = wbuf[c * CAP + (j - c0)];
                            spread_accumulate9(w, ac);
                        }
                    }
                }
            }
            std::fill(rep.begin(), rep.end(), 1e30f);                          // poison: unwritten replicas must not be read
            for (unsigned tid = 0; tid < NT; ++tid) {
                const unsigned lx = tid & (T - 1), ly = tid >> g.lgT;
                for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) rep[replica_index(i * 3 + j, lx + i, ly + j, P)] = acc[tid * 27 + i * 3 + j];
            }
            for (unsigned idx = 0; idx < PP; ++idx) out[(size_t)lz * PP + idx] = reduce_replicas(rep.data(), idx % P, idx / P, T);
            for (unsigned tid = 0; tid < NT; ++tid)
                for (int r = 0; r < 9; ++r) { acc[tid * 27 + r] = acc[tid * 27 + 9 + r]; acc[tid * 27 + 9 + r] = acc[tid * 27 + 18 + r]; acc[tid * 27 + 18 + r] = 0.f; }
        }
    }
    // ---- merge (mesh_merge_kernel incl. its interior fast path)
    std::vector<float> rho(M), buf(M);
    const float mean = (float)(sums[1] / (double)M);
    for (size_t c = 0; c < M; ++c) {
        const unsigned x = (unsigned)(c & (g.nx - 1)), y = (unsigned)((c >> g.lgx) & (g.ny - 1)), z = (unsigned)(c >> (g.lgx + g.lgy));
        const unsigned lx = x & (T - 1), ly = y & (T - 1), lz = z & (T - 1);
        float v;
        if (lx != 0 && lx != T - 1 && ly != 0 && ly != T - 1 && lz != 0 && lz != T - 1) {
            const unsigned t = tile_index(x >> g.lgT, y >> g.lgT, z >> g.lgT, g);
            v = scratch[(size_t)t * P3 + ((lz + 1) * P + (ly + 1)) * P + (lx + 1)];
        } else {
            v = merge_cell(scratch.data(), x, y, z, g);
        }
        rho[c] = v;
        buf[c] = v - mean;
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
    const float scale = (float)(fp.two_over_n * bias);
    std::vector<float4> force(N);
    for (unsigned tile_id = 0; tile_id < ntiles; ++tile_id) {
        unsigned tx, ty, tz; tile_coords(tile_id, g, tx, ty, tz);
        for (unsigned row = 0; row < PP; ++row) {
            const unsigned py = row % P, pz = row / P;
            const unsigned y = ((ty << g.lgT) + py + g.ny - 1) & (g.ny - 1), z = ((tz << g.lgT) + pz + g.nz - 1) & (g.nz - 1);
            for (unsigned lane = 0; lane < P; ++lane) {
                const unsigned x = ((tx << g.lgT) + lane + g.nx - 1) & (g.nx - 1);
                tile[row * P + lane] = buf[(size_t)x + (size_t)g.nx * (y + (size_t)g.ny * z)];
            }
        }
        const unsigned s = start[tile_id << (3 * g.lgT)], en = start[(tile_id + 1) << (3 * g.lgT)];
        for (unsigned j = s; j < en; ++j) {
            const unsigned local = skey[j] & ((1u << (3 * g.lgT)) - 1);
            const unsigned lx = local & (T - 1), ly = (local >> g.lgT) & (T - 1), lz = local >> (2 * g.lgT);
            force[perm[j]] = gather_force(sorted[j], (tx << g.lgT) + lx, (ty << g.lgT) + ly, (tz << g.lgT) + lz, lx, ly, lz, tile.data(), g, fp, scale);
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
