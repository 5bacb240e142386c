// mesh_emul.cu -- CPU emulation of the whole OrderParameterMesh device pipeline (host-only program, built with
// nvcc, runs without a GPU).  tile order (bin -> scan -> place) -> fixed-point spread (CTA per tile, padded integer
// tile, direct path for drifted particles, flush) -> int-to-float + mean removal -> FFT sweeps -> gather, using the
// SAME __host__ __device__ bodies as the kernels in csrc/mesh_kernels.cuh and csrc/mesh_fft_kernels.cuh, with the
// kernels' loop structure.  tests/test_emulation.py compares the dump against the oracle.
#include <cstring>
#include <cmath>
#include <algorithm>
#include "../../metadynamics_plugin_b200/csrc/mesh_kernels.cuh"
#include "emul_fft.h"

using namespace metad::mesh;

// usage: mesh_emul nx ny nz Lx Ly Lz N_global bias lgT stale ntypes mode... in.bin out.bin
//   stale: the tile order is built from positions displaced by up to `stale` cells (the real positions are then
//   spread / gathered through that stale order; displacements > kHalo - 1 exercise the direct path)
static unsigned hash32(unsigned x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

int main(int argc, char** argv) {
    if (argc < 14) { fprintf(stderr, "usage\n"); return 2; }
    Geom g; memset(&g, 0, sizeof g);
    const unsigned nx = atoi(argv[1]), ny = atoi(argv[2]), nz = atoi(argv[3]);
    const double Ld[3] = {atof(argv[4]), atof(argv[5]), atof(argv[6])};
    const unsigned N_global = (unsigned)atol(argv[7]);
    const double bias = atof(argv[8]);
    geom_set_dims(g, nx, ny, nz, atoi(argv[9]));
    const double stale = atof(argv[10]);
    const int ntypes = atoi(argv[11]);
    std::vector<float> mode(ntypes);
    float amax = 0.f;
    for (int i = 0; i < ntypes; ++i) { mode[i] = (float)atof(argv[12 + i]); amax = std::max(amax, std::fabs(mode[i])); }
    const char* fin = argv[12 + ntypes];
    const char* fout = argv[13 + ntypes];
    // triclinic box: METAD_EMUL_TILT="xy,xz,yz" (the positions are then expected inside the sheared box)
    double tilt[3] = {0.0, 0.0, 0.0};
    if (const char* ts = getenv("METAD_EMUL_TILT")) sscanf(ts, "%lf,%lf,%lf", &tilt[0], &tilt[1], &tilt[2]);
    // METAD_EMUL_TILT_LITERAL=0: geometrically correct offsets instead of the reference's (Geom::tq)
    const bool literal = !(getenv("METAD_EMUL_TILT_LITERAL") && atoi(getenv("METAD_EMUL_TILT_LITERAL")) == 0);
    geom_set_box(g, Ld, tilt, literal);
    FILE* f = fopen(fin, "rb");
    fseek(f, 0, SEEK_END); const long bytes = ftell(f); fseek(f, 0, SEEK_SET);
    const unsigned N = (unsigned)(bytes / 16);
    std::vector<float4> postype(N);
    if (fread(postype.data(), 16, N, f) != N) return 2;
    fclose(f);
    const size_t M = (size_t)g.nx * g.ny * g.nz;
    const unsigned n3[3] = {g.nx, g.ny, g.nz};

    // ---- hot cell rule (magic truncation) against its reference form (C cast):
    // the input particles, positions within a few ulps of every cell boundary, and random positions
    unsigned long long cell_mismatch = 0;
    {
        auto check = [&](float x, int axis) {
            const unsigned nn = axis == 0 ? g.nx : (axis == 1 ? g.ny : g.nzg);
            if (cell_coord(x, axis, g) != cell_coord_ref(x, g.lo[axis], g.rcpL[axis], nn)) ++cell_mismatch;
        };
        for (unsigned i = 0; i < N; ++i) { check(postype[i].x, 0); check(postype[i].y, 1); check(postype[i].z, 2); }
        for (int axis = 0; axis < 3; ++axis) {
            const unsigned nn = axis == 0 ? g.nx : (axis == 1 ? g.ny : g.nzg);
            for (unsigned k = 0; k <= nn; ++k) {
                float x = (float)(-Ld[axis] / 2.0 + (double)k * Ld[axis] / (double)nn);
                float lo_side = x, hi_side = x;
                for (int u = 0; u < 24; ++u) {
                    check(lo_side, axis); check(hi_side, axis);
                    lo_side = nextafterf(lo_side, -1e30f); hi_side = nextafterf(hi_side, 1e30f);
                }
            }
            for (unsigned k = 0; k < 200000; ++k) {
                const float u = (hash32(k * 7u + axis + 1000u) >> 8) * (1.0f / 16777216.0f);
                check((u - 0.5f) * g.L[axis], axis);
            }
        }
    }
    // ---- tile order (mesh_bin_kernel / scan / mesh_place_kernel) from the displaced positions
    std::vector<unsigned> keys(N), ranks(N), count(M, 0), start(M + 1), perm(N);
    unsigned max_count = 0;
    for (unsigned i = 0; i < N; ++i) {
        float4 p = postype[i];
        float* c = &p.x;
        for (int d = 0; d < 3; ++d) {
            const float u = (hash32(i * 3u + d + 17u) >> 8) * (1.0f / 16777216.0f);
            float v = c[d] + (u - 0.5f) * 2.f * (float)(stale * Ld[d] / n3[d]);
            if (!g.tri) {
                if (v >= g.L[d] / 2.0f) v -= g.L[d];
                if (v < -(g.L[d] / 2.0f)) v += g.L[d];
            }
            c[d] = v;
        }
        if (g.tri && stale > 0.0) {          // wrap the displaced position back into the sheared box (BoxDim::wrap, z first)
            if (p.z >= g.L[2] / 2.0f) { p.z -= g.L[2]; p.y -= g.L[2] * (float)tilt[2]; p.x -= g.L[2] * (float)tilt[1]; }
            else if (p.z < -(g.L[2] / 2.0f)) { p.z += g.L[2]; p.y += g.L[2] * (float)tilt[2]; p.x += g.L[2] * (float)tilt[1]; }
            const float ys = p.y - (float)tilt[2] * p.z;
            if (ys >= g.L[1] / 2.0f) { p.y -= g.L[1]; p.x -= g.L[1] * (float)tilt[0]; }
            else if (ys < -(g.L[1] / 2.0f)) { p.y += g.L[1]; p.x += g.L[1] * (float)tilt[0]; }
            const float xs = p.x - (float)g.d_a * p.z - (float)tilt[0] * p.y;
            if (xs >= g.L[0] / 2.0f) p.x -= g.L[0];
            else if (xs < -(g.L[0] / 2.0f)) p.x += g.L[0];
        }
        const Cell cc = particle_cell(p, g);
        keys[i] = key_of(cc.ix, cc.iy, cc.iz, g);
        unsigned cx, cy, cz; cell_of_key(keys[i], g, cx, cy, cz);
        if ((int)cx != cc.ix || (int)cy != cc.iy || (int)cz != cc.iz) { fprintf(stderr, "key round trip failed\n"); return 3; }
        ranks[i] = count[keys[i]]++;
        max_count = std::max(max_count, ranks[i] + 1);
    }
    unsigned run = 0;
    for (size_t c = 0; c < M; ++c) { start[c] = run; run += count[c]; }
    start[M] = run;
    // the arrival rank of the device is arbitrary: emulate it reversed
    for (unsigned i = 0; i < N; ++i) perm[start[keys[i]] + (count[keys[i]] - 1 - ranks[i])] = i;
    const unsigned T = 1u << g.lgT, PX = T + 2 * kHaloX, PY = T + 2 * kHalo, PZ = PY, P3 = PX * PY * PZ, ntiles = num_tiles(g);
    std::vector<unsigned> tstart(ntiles + 1);
    for (unsigned t = 0; t <= ntiles; ++t) tstart[t] = start[(size_t)t << (3 * g.lgT)];
    // bank order inside every tile (mesh_bank_order_kernel): slot of the r-th particle of bank class b
    {
        std::vector<unsigned> order(N, 0xffffffffu);
        const unsigned cells = 1u << (3 * g.lgT);
        for (unsigned t = 0; t < ntiles; ++t) {
            const size_t key0 = (size_t)t << (3 * g.lgT);
            unsigned class_count[32] = {0};
            std::vector<unsigned> base(cells);
            for (unsigned c = 0; c < cells; ++c) {
                const unsigned b = g.lgT == 4 ? bank_class<4>(c) : bank_class<3>(c);
                base[c] = class_count[b];
                class_count[b] += start[key0 + c + 1] - start[key0 + c];
            }
            for (unsigned c = 0; c < cells; ++c) {
                const unsigned b = g.lgT == 4 ? bank_class<4>(c) : bank_class<3>(c);
                for (unsigned i = 0; i < start[key0 + c + 1] - start[key0 + c]; ++i) {
                    const unsigned slot = tstart[t] + bank_order_slot(base[c] + i, b, class_count);
                    if (slot >= tstart[t + 1] || order[slot] != 0xffffffffu) { fprintf(stderr, "bank order: bad slot\n"); return 3; }
                    order[slot] = perm[start[key0 + c] + i];
                }
            }
        }
        perm.swap(order);
    }
    const float scale = fx_scale_for(amax, amax * (float)max_count);
    const float inv_scale = 1.0f / scale;

    // ---- spread (mesh_spread_kernel): CTA per tile, integer padded tile, flush into the integer mesh
    std::vector<int> mesh_i(M, 0), tile(P3);
    std::vector<unsigned> cell_keys(N), cache_code_v(N);
    std::vector<float4> cache4(N);
    double sums[2] = {0, 0}, shift_err = 0.0, stencil_err = 0.0;
    unsigned strays = 0;
    for (unsigned tile_id = 0; tile_id < ntiles; ++tile_id) {
        unsigned tx, ty, tz; tile_coords(tile_id, g, tx, ty, tz);
        const int ox = (int)(tx << g.lgT) - kHaloX, oy = (int)(ty << g.lgT) - kHalo, oz = (int)(tz << g.lgT) - kHalo;
        std::fill(tile.begin(), tile.end(), 0);
        double sq = 0.0, s1 = 0.0;
        for (unsigned j = tstart[tile_id]; j < tstart[tile_id + 1]; ++j) {
            const unsigned n = perm[j];
            const float4 p = postype[n];
            int t; memcpy(&t, &p.w, 4);
            const float a = mode[t];
            Cell c = particle_cell(p, g);
            cell_keys[n] = key_of(c.ix, c.iy, c.iz, g);
            sq += (double)a * (double)a; s1 += (double)a;
            float w[9];
            float3 sh;
            if (g.tri) {
                // reference form of the stencil in a triclinic box: fractional coordinate (BoxDim::makeFraction) in fp64,
                // cell = floor, offset from the cell centre
                const double u[3] = {(double)p.x - (g.d_a * (double)p.z + g.d_xy * (double)p.y), (double)p.y - g.d_yz * (double)p.z, (double)p.z};
                int ci[3]; float so[3];
                const unsigned nn[3] = {g.nx, g.ny, g.nzg};
                for (int d = 0; d < 3; ++d) {
                    const double r = (u[d] - g.dlo[d]) * g.dscale[d];
                    const double fl = std::floor(r);
                    ci[d] = (int)(((long long)fl % (long long)nn[d] + nn[d]) % nn[d]);
                    so[d] = (float)(r - fl - 0.5);
                }
                c.ix = ci[0]; c.iy = ci[1]; c.iz = ci[2];
                sh = make_float3(so[0] + g.tq[0], so[1] + g.tq[1], so[2]);
            } else {
                sh = particle_shift(p, c, g);
                const float xyz[3] = {p.x, p.y, p.z};
                const int cxyz[3] = {c.ix, c.iy, c.iz}, rxyz[3] = {c.rx, c.ry, c.rz};
                for (int d = 0; d < 3; ++d)
                    shift_err = std::max(shift_err, (double)std::fabs(cell_shift(xyz[d], rxyz[d], d, g) - cell_shift_f64(xyz[d], cxyz[d], d, g)));
                particle_rebase(c, sh, g);       // stencil base = the cell the accurate offset points to (see mesh_kernels.cuh)
            }
            {   // the hot form used by the kernels must pick the same base (or the neighbour across a face it sits on) and the same offset
                Cell ch; float3 sh2;
                if (g.tri) particle_stencil<true>(p, g, ch, sh2);
                else particle_stencil<false>(p, g, ch, sh2);
                const int dc[3] = {ch.ix - c.ix, ch.iy - c.iy, ch.iz - c.iz};
                const float ds[3] = {sh2.x - sh.x, sh2.y - sh.y, sh2.z - sh.z};
                const int nn[3] = {(int)g.nx, (int)g.ny, (int)g.nz};
                for (int d = 0; d < 3; ++d) {
                    const int dd = ((dc[d] % nn[d]) + nn[d]) % nn[d];           // 0, 1 or n-1
                    const double err = dd == 0 ? std::fabs(ds[d]) : (dd == 1 ? std::fabs(ds[d] + 1.0f) : (dd == nn[d] - 1 ? std::fabs(ds[d] - 1.0f) : 1e9));
                    stencil_err = std::max(stencil_err, err);
                }
                c = ch; sh = sh2;
            }
            if (g.tri) spread_weights<true>(sh, a * scale, w);
            else spread_weights<false>(sh, a * scale, w);
            unsigned lx, ly, lz;
            const bool inside = padded_coords(c, ox, oy, oz, g, PX, PY, PZ, lx, ly, lz);
            cache4[j] = make_float4(sh.x, sh.y, sh.z, a);
            cache_code_v[j] = cache_code(lx, ly, lz, (unsigned)t, inside, true);
            if (inside) {
                int* base = tile.data() + ((lz - 1) * PY + (ly - 1)) * PX + (lx - 1);
                for (int k = 0; k < 3; ++k) for (int jj = 0; jj < 3; ++jj) {
                    const float wyz = f_mul(w[3 + jj], w[6 + k]);
                    for (int i = 0; i < 3; ++i) base[(k * PY + jj) * PX + i] += fx_round(w[i], wyz);
                }
            } else {
                ++strays;
                for (int k = 0; k < 3; ++k) for (int jj = 0; jj < 3; ++jj) for (int i = 0; i < 3; ++i) {
                    long long idx;
                    if (tap_index(c, i, jj, k, g, idx)) mesh_i[idx] += tap_value(w, i, jj, k);
                }
            }
        }
        sums[0] += sq; sums[1] += s1;
        // flush: row-wise bulk reductions (tile_row: start, wrap point)
        for (unsigned row = 0; row < PY * PZ; ++row) {
            TileRow r;
            if (!tile_row(ox, oy, oz, row % PY, row / PY, PX, g, r)) continue;
            int* dst = mesh_i.data() + ((size_t)r.z * g.ny + r.y) * g.nx;
            for (int i = 0; i < r.first; ++i) dst[r.x0 + i] += tile[row * PX + i];
            for (int i = r.first; i < (int)PX; ++i) dst[i - r.first] += tile[row * PX + i];
        }
    }
    // ---- x forward load phase (fft_x_fwd_kernel): int -> float, mean removal
    std::vector<float> rho(M), buf(M);
    const float mean = (float)(sums[1] * (1.0 / (double)M));
    for (size_t c = 0; c < M; c += 2) {
        const float2 r = density_to_float(make_int2(mesh_i[c], mesh_i[c + 1]), inv_scale);
        rho[c] = r.x; rho[c + 1] = r.y;
        buf[c] = r.x - mean; buf[c + 1] = r.y - mean;
    }
    // ---- FFT sweeps
    float2* b2 = reinterpret_cast<float2*>(buf.data());
    const unsigned nxh = g.nx / 2;
    const float inv_n = (float)(1.0 / (double)N_global);
    const float d = (float)(0.5 * sums[0] / (double)N_global / (double)N_global);
    double e = 0.0;
    DISPATCH(nxh, x_fwd<LL>(b2, g.ny * g.nz));
    DISPATCH(g.ny, (y_pass<LL, -1>(b2, nxh, g.nz)));
    // ConvParams::dc_restore: what the mean removal took out of f_0 goes back in before the convolution (literal triclinic offsets)
    const float dc = (g.tri && (g.tq[0] != 0.f || g.tq[1] != 0.f)) ? (float)((double)mean * (double)M * (double)inv_n) : 0.f;
    DISPATCH(g.nz, e += z_plane0<LL>(b2, g.nx, g.ny, inv_n, d, dc));
    DISPATCH(g.nz, e += z_fused<LL>(b2, g.nx, g.ny, inv_n, d));
    DISPATCH(g.ny, (y_pass<LL, +1>(b2, nxh, g.nz)));
    DISPATCH(nxh, x_inv<LL>(b2, g.ny * g.nz));
    const double cv = 0.5 * e;
    // ---- gather (mesh_gather_kernel)
    ForceParams fp; memset(&fp, 0, sizeof fp);
    {   // n_a b_a with the reciprocal lattice vectors b_a = (a_b x a_c) / V of the (sheared) box, as launch_gather (csrc/mesh.cu) sets them
        const double a1[3] = {Ld[0], 0.0, 0.0}, a2[3] = {tilt[0] * Ld[1], Ld[1], 0.0}, a3[3] = {tilt[1] * Ld[2], tilt[2] * Ld[2], Ld[2]};
        const double V = Ld[0] * Ld[1] * Ld[2];
        auto cross = [&](const double* u, const double* v, double nn, float* o) {
            o[0] = (float)(nn * (u[1] * v[2] - u[2] * v[1]) / V); o[1] = (float)(nn * (u[2] * v[0] - u[0] * v[2]) / V); o[2] = (float)(nn * (u[0] * v[1] - u[1] * v[0]) / V);
        };
        cross(a2, a3, (double)g.nx, fp.nb1); cross(a3, a1, (double)g.ny, fp.nb2); cross(a1, a2, (double)g.nz, fp.nb3);
    }
    fp.two_over_n = 2.0 / (double)N_global;
    const float fscale = (float)(fp.two_over_n * bias);
    std::vector<float4> force(N);
    std::vector<float> ftile(P3);
    for (unsigned tile_id = 0; tile_id < ntiles; ++tile_id) {
        unsigned tx, ty, tz; tile_coords(tile_id, g, tx, ty, tz);
        const int ox = (int)(tx << g.lgT) - kHaloX, oy = (int)(ty << g.lgT) - kHalo, oz = (int)(tz << g.lgT) - kHalo;
        for (unsigned row = 0; row < PY * PZ; ++row) {
            TileRow r;
            if (!tile_row(ox, oy, oz, row % PY, row / PY, PX, g, r)) continue;
            const float* src = buf.data() + ((size_t)r.z * g.ny + r.y) * g.nx;
            for (int i = 0; i < r.first; ++i) ftile[row * PX + i] = src[r.x0 + i];
            for (int i = r.first; i < (int)PX; ++i) ftile[row * PX + i] = src[i - r.first];
        }
        for (unsigned j = tstart[tile_id]; j < tstart[tile_id + 1]; ++j) {
            const unsigned n = perm[j];
            const float4 p = postype[n];
            const unsigned code = cache_code_v[j];
            const float4 q = cache4[j];
            Cell c;
            float3 sh_g;
            if (g.tri) particle_stencil<true>(p, g, c, sh_g);
            else particle_stencil<false>(p, g, c, sh_g);
            GatherWeights w;
            if (g.tri) gather_weights<true>(make_float3(q.x, q.y, q.z), w);
            else gather_weights<false>(make_float3(q.x, q.y, q.z), w);
            float Sx, Sy, Sz;
            if (code & kCacheInside) {
                const unsigned lx = code & 31u, ly = (code >> 5) & 31u, lz = (code >> 10) & 31u;
                gather_sums(ftile.data() + ((lz - 1) * PY + (ly - 1)) * PX + (lx - 1), PX, PX * PY, w.wx, w.wy, w.wz, w.dx, w.dy, w.dz, Sx, Sy, Sz);
            } else {        // gather_direct
                float t27[27];
                for (int k = 0; k < 3; ++k) for (int jj = 0; jj < 3; ++jj) for (int i = 0; i < 3; ++i) {
                    const unsigned x = (unsigned)(c.ix + i - 1) & (g.nx - 1), y = (unsigned)(c.iy + jj - 1) & (g.ny - 1), z = (unsigned)(c.iz + k - 1) & (g.nz - 1);
                    t27[(k * 3 + jj) * 3 + i] = buf[(size_t)x + (size_t)g.nx * (y + (size_t)g.ny * z)];
                }
                gather_sums(t27, 3, 9, w.wx, w.wy, w.wz, w.dx, w.dy, w.dz, Sx, Sy, Sz);
            }
            force[n] = force_from_sums(Sx, Sy, Sz, q.w, fp, fscale);
        }
    }
    // ---- dump: cv, mode_sq, shift_err, strays, scale, cell rule mismatches, rho[M], inv[M], force[4N], cells[3N]
    f = fopen(fout, "wb");
    const double dstrays = strays, dscale = scale, dmis = (double)cell_mismatch;
    fprintf(stderr, "shift_err %.3e stencil_err %.3e\n", shift_err, stencil_err);
    shift_err = std::max(shift_err, stencil_err);      // both against the 1e-7 bound of the test
    fwrite(&cv, 8, 1, f); fwrite(&sums[0], 8, 1, f); fwrite(&shift_err, 8, 1, f); fwrite(&dstrays, 8, 1, f); fwrite(&dscale, 8, 1, f);
    fwrite(&dmis, 8, 1, f);
    fwrite(rho.data(), 4, M, f); fwrite(buf.data(), 4, M, f); fwrite(force.data(), 16, N, f);
    std::vector<int> cells(3 * (size_t)N);
    for (unsigned i = 0; i < N; ++i) { unsigned ix, iy, iz; cell_of_key(cell_keys[i], g, ix, iy, iz); cells[3 * i] = ix; cells[3 * i + 1] = iy; cells[3 * i + 2] = iz; }
    fwrite(cells.data(), 4, cells.size(), f);
    fclose(f);
    return 0;
}
