// general_emul.cu -- CPU emulation of the general mesh path (csrc/mesh_general.cuh: any mesh size, one particle per
// thread, mixed-radix Stockham lines), host-only program built with nvcc, runs without a GPU.  Uses the SAME
// __host__ __device__ bodies as the kernels (particle_stencil<TRI, false>, spread_weights, tap_value, stage_output,
// line_of, conv_mode, gather_sums, force_from_sums) with the kernels' loop structure.  Same command line and dump
// format as mesh_emul (the tile size and the stale-order argument are ignored).
#include <cstring>
#include <cmath>
#include <vector>
#include <algorithm>
#include "../../metadynamics_plugin_b200/csrc/mesh_general.cuh"

using namespace metad::mesh;
using namespace metad::meshgen;

static void fft_axis(std::vector<float2>& data, unsigned nx, unsigned ny, unsigned nz, int axis, float sign) {
    LineMap lm; lm.nx = nx; lm.ny = ny; lm.nz = nz; lm.axis = axis;
    const unsigned n = axis == 0 ? nx : (axis == 1 ? ny : nz);
    if (n == 1) return;
    const Radices rad = factorize(n);
    std::vector<float2> tw(n), a(n), b(n);
    for (unsigned k = 0; k < n; ++k) {            // upload_twiddles (csrc/mesh.cu)
        const double ph = -2.0 * M_PI * (double)k / (double)n;
        tw[k] = make_float2((float)cos(ph), (float)sin(ph));
    }
    for (unsigned line = 0; line < line_count(lm); ++line) {        // gen_fft_kernel: one line per CTA
        size_t base, stride; unsigned nn;
        line_of(lm, line, base, stride, nn);
        for (unsigned i = 0; i < n; ++i) a[i] = data[base + i * stride];
        float2 *x = a.data(), *y = b.data();
        int Ns = 1;
        for (int s = 0; s < rad.count; ++s) {
            const int R = rad.r[s];
            for (unsigned w = 0; w < n; ++w) { int o; const float2 v = stage_output(x, tw.data(), (int)n, R, Ns, (int)w, sign, o); y[o] = v; }
            std::swap(x, y);
            Ns *= R;
        }
        for (unsigned i = 0; i < n; ++i) data[base + i * stride] = x[i];
    }
}

int main(int argc, char** argv) {
    if (argc < 14) { fprintf(stderr, "usage\n"); return 2; }
    Geom g;
    const unsigned nx = atoi(argv[1]), ny = atoi(argv[2]), nz = atoi(argv[3]);
    const double Ld[3] = {atof(argv[4]), atof(argv[5]), atof(argv[6])};
    const unsigned N_global = (unsigned)atol(argv[7]);
    const double bias = atof(argv[8]);
    geom_set_dims_general(g, nx, ny, nz);
    const int ntypes = atoi(argv[11]);
    std::vector<float> mode(ntypes);
    float amax = 0.f;
    for (int i = 0; i < ntypes; ++i) { mode[i] = (float)atof(argv[12 + i]); amax = std::max(amax, std::fabs(mode[i])); }
    const char* fin = argv[12 + ntypes];
    const char* fout = argv[13 + ntypes];
    double tilt[3] = {0.0, 0.0, 0.0};
    if (const char* ts = getenv("METAD_EMUL_TILT")) sscanf(ts, "%lf,%lf,%lf", &tilt[0], &tilt[1], &tilt[2]);
    const bool literal = !(getenv("METAD_EMUL_TILT_LITERAL") && atoi(getenv("METAD_EMUL_TILT_LITERAL")) == 0);
    geom_set_box(g, Ld, tilt, literal);
    FILE* f = fopen(fin, "rb");
    fseek(f, 0, SEEK_END); const long bytes = ftell(f); fseek(f, 0, SEEK_SET);
    const unsigned N = (unsigned)(bytes / 16);
    std::vector<float4> postype(N);
    if (fread(postype.data(), 16, N, f) != N) return 2;
    fclose(f);
    const size_t M = (size_t)nx * ny * nz;

    // ---- gen_spread_kernel
    const float scale = fx_scale_for(amax, amax);
    std::vector<long long> mesh64(M, 0);
    std::vector<int> cells(3 * (size_t)N);
    double sums[2] = {0.0, 0.0};
    for (unsigned n = 0; n < N; ++n) {
        const float4 p = postype[n];
        int t; memcpy(&t, &p.w, 4);
        const float a = mode[t];
        const Cell cf = particle_cell(p, g);
        cells[3 * (size_t)n] = cf.ix; cells[3 * (size_t)n + 1] = cf.iy; cells[3 * (size_t)n + 2] = cf.iz;
        Cell c; float3 sh;
        float w[9];
        if (g.tri) { particle_stencil<true, false>(p, g, c, sh); spread_weights<true>(sh, a * scale, w); }
        else { particle_stencil<false, false>(p, g, c, sh); spread_weights<false>(sh, a * scale, w); }
        for (int k = 0; k < 3; ++k) for (int j = 0; j < 3; ++j) for (int i = 0; i < 3; ++i)
            mesh64[tap_cell(c, i, j, k, g)] += (long long)tap_value(w, i, j, k);
        sums[0] += (double)a * (double)a; sums[1] += (double)a;
    }
    // ---- gen_density_kernel
    const double inv_cells = 1.0 / (double)M;
    const float inv_scale = 1.0f / scale;
    const float mean = (float)(sums[1] * inv_cells);
    std::vector<float> rho(M);
    std::vector<float2> spec(M);
    for (size_t c = 0; c < M; ++c) {
        const float r = (float)mesh64[c] * inv_scale;
        rho[c] = r;
        spec[c] = make_float2(r - mean, 0.f);
    }
    // ---- transforms, convolution, energy
    fft_axis(spec, nx, ny, nz, 0, -1.f); fft_axis(spec, nx, ny, nz, 1, -1.f); fft_axis(spec, nx, ny, nz, 2, -1.f);
    ConvGeom cg;
    cg.nx = nx; cg.ny = ny; cg.nz = nz;
    cg.inv_n = (float)(1.0 / (double)N_global);
    cg.d = (float)(0.5 * sums[0] / (double)N_global / (double)N_global);
    cg.dc = (g.tri && (g.tq[0] != 0.f || g.tq[1] != 0.f)) ? (float)((double)mean / inv_cells * (double)cg.inv_n) : 0.f;      // ConvParams::dc_restore
    double e = 0.0;
    for (size_t c = 0; c < M; ++c) { float val; unsigned kx, ky, kz; spec[c] = conv_mode(spec[c], c, cg, e, val, kx, ky, kz); }
    const double cv = 0.5 * e;
    fft_axis(spec, nx, ny, nz, 2, +1.f); fft_axis(spec, nx, ny, nz, 1, +1.f); fft_axis(spec, nx, ny, nz, 0, +1.f);
    // ---- gen_gather_kernel
    ForceParams fp; memset(&fp, 0, sizeof fp);
    {
        const double a1[3] = {Ld[0], 0.0, 0.0}, a2[3] = {tilt[0] * Ld[1], Ld[1], 0.0}, a3[3] = {tilt[1] * Ld[2], tilt[2] * Ld[2], Ld[2]};
        const double V = Ld[0] * Ld[1] * Ld[2];
        auto cross = [&](const double* u, const double* v, double nn, float* o) {
            o[0] = (float)(nn * (u[1] * v[2] - u[2] * v[1]) / V); o[1] = (float)(nn * (u[2] * v[0] - u[0] * v[2]) / V); o[2] = (float)(nn * (u[0] * v[1] - u[1] * v[0]) / V);
        };
        cross(a2, a3, (double)nx, fp.nb1); cross(a3, a1, (double)ny, fp.nb2); cross(a1, a2, (double)nz, fp.nb3);
    }
    fp.two_over_n = 2.0 / (double)N_global;
    const float fscale = (float)(fp.two_over_n * bias);
    std::vector<float4> force(N);
    for (unsigned n = 0; n < N; ++n) {
        const float4 p = postype[n];
        int t; memcpy(&t, &p.w, 4);
        Cell c; float3 sh;
        GatherWeights w;
        if (g.tri) { particle_stencil<true, false>(p, g, c, sh); gather_weights<true>(sh, w); }
        else { particle_stencil<false, false>(p, g, c, sh); gather_weights<false>(sh, w); }
        float t27[27];
        for (int k = 0; k < 3; ++k) for (int j = 0; j < 3; ++j) for (int i = 0; i < 3; ++i) t27[(k * 3 + j) * 3 + i] = spec[tap_cell(c, i, j, k, g)].x;
        float Sx, Sy, Sz;
        gather_sums(t27, 3, 9, w.wx, w.wy, w.wz, w.dx, w.dy, w.dz, Sx, Sy, Sz);
        force[n] = force_from_sums(Sx, Sy, Sz, mode[t], fp, fscale);
    }
    // ---- dump (format of mesh_emul): cv, mode_sq, shift_err, strays, scale, cell rule mismatches, rho[M], inv[M], force[4N], cells[3N]
    std::vector<float> inv(M);
    for (size_t c = 0; c < M; ++c) inv[c] = spec[c].x;
    f = fopen(fout, "wb");
    const double zero = 0.0, dscale = scale;
    fwrite(&cv, 8, 1, f); fwrite(&sums[0], 8, 1, f); fwrite(&zero, 8, 1, f); fwrite(&zero, 8, 1, f); fwrite(&dscale, 8, 1, f); fwrite(&zero, 8, 1, f);
    fwrite(rho.data(), 4, M, f); fwrite(inv.data(), 4, M, f); fwrite(force.data(), 16, N, f);
    fwrite(cells.data(), 4, cells.size(), f);
    fclose(f);
    return 0;
}
