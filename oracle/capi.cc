// capi.cc -- plain-C entry points over metad_oracle.hpp for ctypes (tests/, smoke(), bench cpu_baseline).
// TEST INFRASTRUCTURE ONLY; see the header of metad_oracle.hpp ("parity unpinned").
// Every entry point exists twice: suffix _f32 (S=float) and _f64 (S=double).  Arrays cross the
// boundary as double (outputs) / float (particle data, as HOOMD's fp32 position array).
#include <iostream>
#include "metad_oracle.hpp"

static std::ios_base::Init g_iostream_init;   // the toolchain links libstdc++ statically: make sure locales exist

using namespace oracle;

namespace {
template <class S> Box<S> mkbox(const double* b) { return Box<S>::make(b[0], b[1], b[2], b[3], b[4], b[5]); }

template <class S> struct MeshH { MeshCV<S> m; };

template <class S> void* mesh_create(unsigned nx, unsigned ny, unsigned nz, const double* mode, int ntypes,
                                     const double* box6, unsigned n_global) {
    std::vector<double> md(mode, mode + ntypes);
    auto* h = new MeshCV<S>(nx, ny, nz, md, mkbox<S>(box6));
    h->N_global = n_global;
    return h;
}
template <class S> void mesh_get(void* hv, int which, double* out) {
    auto* h = (MeshCV<S>*)hv;
    size_t M = h->M;
    switch (which) {
        case 0: for (size_t i = 0; i < M; ++i) out[i] = h->mesh[i].real(); break;
        case 1: for (size_t i = 0; i < M; ++i) { out[2 * i] = h->fourier[i].real(); out[2 * i + 1] = h->fourier[i].imag(); } break;
        case 2: for (size_t i = 0; i < M; ++i) { out[2 * i] = h->fourier_G[i].real(); out[2 * i + 1] = h->fourier_G[i].imag(); } break;
        case 3: for (size_t i = 0; i < M; ++i) out[i] = h->inv[i].real(); break;
        case 4: for (size_t i = 0; i < M; ++i) out[i] = h->interp[i]; break;
        case 5: for (size_t i = 0; i < M; ++i) out[i] = h->inv[i].imag(); break;
    }
}
template <class S> void mesh_forces(void* hv, const float* pt, unsigned N, double bias, double* out4) {
    auto* h = (MeshCV<S>*)hv;
    std::vector<S> f((size_t)4 * N);
    h->forces(pt, N, (S)bias, f.data());
    for (size_t i = 0; i < f.size(); ++i) out4[i] = f[i];
}
template <class S> Lamellar<S> mklam(unsigned n_global, const double* mode, int ntypes, const int* lv, int nw, const double* box6) {
    Lamellar<S> l;
    for (int i = 0; i < ntypes; ++i) l.mode.push_back((S)mode[i]);
    l.lattice.assign(lv, lv + 3 * nw);
    l.box = mkbox<S>(box6);
    l.N_global = n_global;
    return l;
}
template <class S> double lam_cv(const float* pt, unsigned N, unsigned n_global, const double* mode, int ntypes,
                                 const int* lv, int nw, const double* box6, double* modes_out) {
    auto l = mklam<S>(n_global, mode, ntypes, lv, nw, box6);
    double cv = l.compute_cv(pt, N);
    if (modes_out) for (int i = 0; i < 2 * nw; ++i) modes_out[i] = l.fourier[i];
    return cv;
}
template <class S> void lam_forces(const float* pt, unsigned N, unsigned n_global, const double* mode, int ntypes,
                                   const int* lv, int nw, const double* box6, double bias, double* out4) {
    auto l = mklam<S>(n_global, mode, ntypes, lv, nw, box6);
    std::vector<S> f((size_t)4 * N);
    l.forces(pt, N, (S)bias, f.data());
    for (size_t i = 0; i < f.size(); ++i) out4[i] = f[i];
}
template <class S> void* grid_create(int d, const double* cv_min, const double* cv_max, const unsigned* npts,
                                     const double* sigma, double W, double T_shift, double T, unsigned stride,
                                     int add_bias, int well_tempered) {
    auto* g = new MetaGrid<S>();
    for (int i = 0; i < d; ++i) {
        typename MetaGrid<S>::Var v;
        v.sigma = (S)sigma[i]; v.cv_min = (S)cv_min[i]; v.cv_max = (S)cv_max[i]; v.num_points = npts[i];
        v.name = "cv" + std::to_string(i);
        g->vars.push_back(v);
    }
    g->W = (S)W; g->T_shift = (S)T_shift; g->temp = (S)T; g->stride = stride;
    g->add_bias = add_bias != 0; g->well_tempered = well_tempered != 0;
    g->setup();
    return g;
}
template <class S> void grid_update(void* gv, unsigned timestep, const double* cur, double* bias_out) {
    auto* g = (MetaGrid<S>*)gv;
    std::vector<S> c(g->vars.size()), b;
    for (size_t i = 0; i < c.size(); ++i) c[i] = (S)cur[i];
    g->update(timestep, c, b);
    for (size_t i = 0; i < b.size(); ++i) bias_out[i] = b[i];
}
template <class S> void grid_get(void* gv, int which, double* out) {
    auto* g = (MetaGrid<S>*)gv;
    size_t G = g->grid.size();
    for (size_t i = 0; i < G; ++i) {
        switch (which) {
            case 0: out[i] = g->grid[i]; break;
            case 1: out[i] = g->grid_reweighted[i]; break;
            case 2: out[i] = g->grid_weight[i]; break;
            case 3: out[i] = g->sigma_grid[i]; break;
            case 4: out[i] = g->hist[i]; break;
            case 5: out[i] = g->hist_gauss[i]; break;
            case 6: out[i] = g->hist_delta[i]; break;
            case 7: out[i] = g->grid_delta[i]; break;
        }
    }
}
template <class S> void grid_scalars(void* gv, double* out4) {
    auto* g = (MetaGrid<S>*)gv;
    out4[0] = g->curr_bias_potential; out4[1] = g->curr_reweight; out4[2] = g->num_gaussians; out4[3] = g->n_out_of_bounds;
}
template <class S> double grid_interp(void* gv, const double* val, int reweight) {
    auto* g = (MetaGrid<S>*)gv;
    std::vector<S> c(g->vars.size());
    for (size_t i = 0; i < c.size(); ++i) c[i] = (S)val[i];
    return g->interpolate(c, reweight != 0);
}
template <class S> int grid_bin(void* gv, const double* val) {
    auto* g = (MetaGrid<S>*)gv;
    std::vector<S> c(g->vars.size());
    for (size_t i = 0; i < c.size(); ++i) c[i] = (S)val[i];
    unsigned idx = 0;
    return g->bin_of(c, idx) ? (int)idx : -1;
}
template <class S> double umb(int what, int kind, double cv0, double kappa, double width, double scale, double val, double bias_in) {
    UmbrellaParams<S> u; u.kind = kind; u.cv0 = (S)cv0; u.kappa = (S)kappa; u.width_flat = (S)width; u.scale = (S)scale;
    return what == 0 ? (double)umbrella_bias<S>(u, (S)val, (S)bias_in) : (double)umbrella_potential<S>(u, (S)val);
}
template <class S> void wte_scale_c(float* f4, float* t4, float* vir, unsigned pitch, unsigned N, double bias, double* ext6) {
    // arrays arrive as fp32 (HOOMD single-precision layout); computed in S, returned rounded to fp32
    std::vector<S> f(f4, f4 + (size_t)4 * N), t(t4, t4 + (size_t)4 * N), v(vir, vir + (size_t)6 * pitch);
    S e[6]; for (int i = 0; i < 6; ++i) e[i] = (S)ext6[i];
    wte_scale<S>(f.data(), t.data(), v.data(), pitch, N, (S)bias, e);
    for (size_t i = 0; i < f.size(); ++i) { f4[i] = (float)f[i]; t4[i] = (float)t[i]; }
    for (size_t i = 0; i < v.size(); ++i) vir[i] = (float)v[i];
    for (int i = 0; i < 6; ++i) ext6[i] = e[i];
}
}  // namespace

#define ORC_INSTANTIATE(SFX, S)                                                                                       \
    extern "C" {                                                                                                      \
    void* orc_mesh_create_##SFX(unsigned nx, unsigned ny, unsigned nz, const double* mode, int ntypes,               \
                                const double* box6, unsigned n_global) {                                              \
        return mesh_create<S>(nx, ny, nz, mode, ntypes, box6, n_global);                                              \
    }                                                                                                                 \
    void orc_mesh_destroy_##SFX(void* h) { delete (MeshCV<S>*)h; }                                                    \
    void orc_mesh_assign_##SFX(void* h, const float* pt, unsigned N) { ((MeshCV<S>*)h)->assign(pt, N); }              \
    void orc_mesh_update_##SFX(void* h) { ((MeshCV<S>*)h)->update_meshes(); }                                         \
    double orc_mesh_cv_##SFX(void* h) { auto* m = (MeshCV<S>*)h; m->cv = m->compute_cv(); return m->cv; }             \
    double orc_mesh_current_value_##SFX(void* h, const float* pt, unsigned N) {                                       \
        return ((MeshCV<S>*)h)->current_value(pt, N);                                                                 \
    }                                                                                                                 \
    void orc_mesh_forces_##SFX(void* h, const float* pt, unsigned N, double bias, double* out4) {                     \
        mesh_forces<S>(h, pt, N, bias, out4);                                                                         \
    }                                                                                                                 \
    void orc_mesh_get_##SFX(void* h, int which, double* out) { mesh_get<S>(h, which, out); }                          \
    void orc_mesh_cells_##SFX(void* h, int* out, unsigned N) {                                                        \
        auto* m = (MeshCV<S>*)h; for (size_t i = 0; i < (size_t)3 * N; ++i) out[i] = m->cells[i];                     \
    }                                                                                                                 \
    double orc_mesh_mode_sq_##SFX(void* h) { return ((MeshCV<S>*)h)->mode_sq; }                                       \
    void orc_mesh_set_literal_copysignf_##SFX(void* h, int on) { ((MeshCV<S>*)h)->literal_copysignf = on != 0; }      \
    void orc_mesh_set_literal_tilt_offset_##SFX(void* h, int on) { ((MeshCV<S>*)h)->literal_tilt_offset = on != 0; } \
    void orc_mesh_virial_##SFX(void* h, const double* table_d, unsigned n, double kmin, double kmax, int use_table,  \
                               double bias, double* out6) {                                                           \
        std::vector<S> t(table_d, table_d + n); S o[6];                                                               \
        ((MeshCV<S>*)h)->virial(t, (S)kmin, (S)kmax, use_table != 0, (S)bias, o); for (int i = 0; i < 6; ++i) out6[i] = o[i]; \
    }                                                                                                                 \
    void orc_mesh_qmax_##SFX(void* h, double* out4) {                                                                 \
        S o[4]; ((MeshCV<S>*)h)->qmax(o); for (int i = 0; i < 4; ++i) out4[i] = o[i];                                 \
    }                                                                                                                 \
    double orc_lamellar_cv_##SFX(const float* pt, unsigned N, unsigned n_global, const double* mode, int ntypes,     \
                                 const int* lv, int nw, const double* box6, double* modes_out) {                      \
        return lam_cv<S>(pt, N, n_global, mode, ntypes, lv, nw, box6, modes_out);                                     \
    }                                                                                                                 \
    void orc_lamellar_forces_##SFX(const float* pt, unsigned N, unsigned n_global, const double* mode, int ntypes,   \
                                   const int* lv, int nw, const double* box6, double bias, double* out4) {            \
        lam_forces<S>(pt, N, n_global, mode, ntypes, lv, nw, box6, bias, out4);                                       \
    }                                                                                                                 \
    void* orc_grid_create_##SFX(int d, const double* cv_min, const double* cv_max, const unsigned* npts,             \
                                const double* sigma, double W, double T_shift, double T, unsigned stride,             \
                                int add_bias, int well_tempered) {                                                    \
        return grid_create<S>(d, cv_min, cv_max, npts, sigma, W, T_shift, T, stride, add_bias, well_tempered);        \
    }                                                                                                                 \
    void orc_grid_destroy_##SFX(void* g) { delete (MetaGrid<S>*)g; }                                                  \
    void orc_grid_update_##SFX(void* g, unsigned timestep, const double* cur, double* bias_out) {                     \
        grid_update<S>(g, timestep, cur, bias_out);                                                                   \
    }                                                                                                                 \
    void orc_grid_update_deposit_##SFX(void* g, unsigned timestep, const double* cur) {                               \
        auto* m = (MetaGrid<S>*)g; std::vector<S> c(m->vars.size()); for (size_t i = 0; i < c.size(); ++i) c[i] = (S)cur[i]; \
        m->update_deposit(timestep, c);                                                                               \
    }                                                                                                                 \
    void orc_grid_update_merge_##SFX(void* g, unsigned timestep, const double* cur, double* bias_out) {               \
        auto* m = (MetaGrid<S>*)g; std::vector<S> c(m->vars.size()), b; for (size_t i = 0; i < c.size(); ++i) c[i] = (S)cur[i]; \
        m->update_merge(timestep, c, b); for (size_t i = 0; i < b.size(); ++i) bias_out[i] = b[i];                     \
    }                                                                                                                 \
    /* which: 0 grid_delta, 1 sigma_grid_delta, 2 hist_delta, 3 hist_gauss_delta; dir 0 = read into buf, 1 = write from buf */ \
    void orc_grid_delta_io_##SFX(void* g, int which, int dir, double* buf) {                                          \
        auto* m = (MetaGrid<S>*)g; const size_t G = m->grid.size();                                                   \
        for (size_t i = 0; i < G; ++i) {                                                                              \
            if (which == 0) { if (dir) m->grid_delta[i] = (S)buf[i]; else buf[i] = m->grid_delta[i]; }                \
            else if (which == 1) { if (dir) m->sigma_grid_delta[i] = (S)buf[i]; else buf[i] = m->sigma_grid_delta[i]; } \
            else if (which == 2) { if (dir) m->hist_delta[i] = (unsigned)buf[i]; else buf[i] = m->hist_delta[i]; }    \
            else { if (dir) m->hist_gauss_delta[i] = (unsigned)buf[i]; else buf[i] = m->hist_gauss_delta[i]; }        \
        }                                                                                                             \
    }                                                                                                                 \
    /* forces: ncv pointers (float4 arrays of N particles) or null */                                                 \
    void orc_grid_compute_sigma_##SFX(void* g, const float* const* forces, unsigned N, double sigma_g, double* sigma_inv_out) { \
        auto* m = (MetaGrid<S>*)g; std::vector<const float*> f(forces, forces + m->vars.size());                      \
        m->compute_sigma(f, N, (S)sigma_g);                                                                           \
        for (size_t i = 0; i < m->sigma_inv.size(); ++i) sigma_inv_out[i] = m->sigma_inv[i];                          \
    }                                                                                                                 \
    void orc_grid_get_sigma_inv_##SFX(void* g, double* si) {                                                          \
        auto* m = (MetaGrid<S>*)g; for (size_t i = 0; i < m->sigma_inv.size(); ++i) si[i] = m->sigma_inv[i];          \
    }                                                                                                                 \
    void orc_grid_set_sigma_inv_##SFX(void* g, const double* si) {                                                    \
        auto* m = (MetaGrid<S>*)g; for (size_t i = 0; i < m->sigma_inv.size(); ++i) m->sigma_inv[i] = (S)si[i];       \
    }                                                                                                                 \
    void orc_grid_get_##SFX(void* g, int which, double* out) { grid_get<S>(g, which, out); }                          \
    void orc_grid_scalars_##SFX(void* g, double* out4) { grid_scalars<S>(g, out4); }                                  \
    double orc_grid_interpolate_##SFX(void* g, const double* val, int reweight) { return grid_interp<S>(g, val, reweight); } \
    int orc_grid_bin_##SFX(void* g, const double* val) { return grid_bin<S>(g, val); }                                \
    void orc_grid_set_flags_##SFX(void* g, int add_bias, int well_tempered, unsigned stride) {                        \
        auto* m = (MetaGrid<S>*)g; m->add_bias = add_bias != 0; m->well_tempered = well_tempered != 0; m->stride = stride; \
    }                                                                                                                 \
    void orc_grid_reset_histogram_##SFX(void* g) { ((MetaGrid<S>*)g)->reset_histogram(); }                            \
    void orc_grid_write_##SFX(void* g, const char* fn, unsigned timestep) { ((MetaGrid<S>*)g)->write_grid(fn, timestep); } \
    int orc_grid_read_##SFX(void* g, const char* fn) {                                                                \
        try { ((MetaGrid<S>*)g)->read_grid(fn); } catch (...) { return -1; } return 0;                                \
    }                                                                                                                 \
    double orc_umbrella_##SFX(int what, int kind, double cv0, double kappa, double width, double scale, double val,  \
                              double bias_in) {                                                                       \
        return umb<S>(what, kind, cv0, kappa, width, scale, val, bias_in);                                            \
    }                                                                                                                 \
    double orc_wte_pe_##SFX(const float* nf4, unsigned N, double ext) { return wte_potential_energy<S>(nf4, N, (S)ext); } \
    void orc_wte_scale_##SFX(float* f4, float* t4, float* vir, unsigned pitch, unsigned N, double bias, double* ext6) { \
        wte_scale_c<S>(f4, t4, vir, pitch, N, bias, ext6);                                                            \
    }                                                                                                                 \
    double orc_aspect_value_##SFX(const double* box6, unsigned d1, unsigned d2) {                                     \
        return aspect_ratio_value<S>(mkbox<S>(box6), d1, d2);                                                         \
    }                                                                                                                 \
    void orc_aspect_virial_##SFX(const double* box6, unsigned d1, unsigned d2, double bias, double* out6) {           \
        S o[6]; aspect_ratio_virial<S>(mkbox<S>(box6), d1, d2, (S)bias, o); for (int i = 0; i < 6; ++i) out6[i] = o[i]; \
    }                                                                                                                 \
    double orc_density_value_##SFX(const double* box6, unsigned n) { return density_value<S>(mkbox<S>(box6), n); }    \
    void orc_density_virial_##SFX(const double* box6, unsigned n, double bias, double* out6) {                        \
        S o[6]; density_virial<S>(mkbox<S>(box6), n, (S)bias, o); for (int i = 0; i < 6; ++i) out6[i] = o[i];         \
    }                                                                                                                 \
    }

ORC_INSTANTIATE(f32, float)
ORC_INSTANTIATE(f64, double)

extern "C" {
// IndexGrid: pure integer, no precision suffix
unsigned orc_indexgrid_index(const unsigned* lengths, int d, const unsigned* coords) {
    IndexGrid g; g.setLengths(std::vector<unsigned>(lengths, lengths + d));
    return g.getIndex(std::vector<unsigned>(coords, coords + d));
}
void orc_indexgrid_coords(const unsigned* lengths, int d, unsigned idx, unsigned* coords) {
    IndexGrid g; g.setLengths(std::vector<unsigned>(lengths, lengths + d));
    std::vector<unsigned> c(d); g.getCoordinates(idx, c);
    for (int i = 0; i < d; ++i) coords[i] = c[i];
}
unsigned orc_indexgrid_num(const unsigned* lengths, int d) {
    IndexGrid g; g.setLengths(std::vector<unsigned>(lengths, lengths + d));
    return g.getNumElements();
}
// 3-D unnormalised DFT of a complex array (interleaved re,im doubles), dims (nz,ny,nx); for cross-checks
void orc_fft3d_f64(const double* in, double* out, unsigned nx, unsigned ny, unsigned nz, int sign) {
    size_t M = (size_t)nx * ny * nz;
    std::vector<std::complex<double>> a(M), b;
    for (size_t i = 0; i < M; ++i) a[i] = {in[2 * i], in[2 * i + 1]};
    fft3d(a, b, nx, ny, nz, sign);
    for (size_t i = 0; i < M; ++i) { out[2 * i] = b[i].real(); out[2 * i + 1] = b[i].imag(); }
}
void orc_fft3d_f32(const double* in, double* out, unsigned nx, unsigned ny, unsigned nz, int sign) {
    size_t M = (size_t)nx * ny * nz;
    std::vector<std::complex<float>> a(M), b;
    for (size_t i = 0; i < M; ++i) a[i] = {(float)in[2 * i], (float)in[2 * i + 1]};
    fft3d(a, b, nx, ny, nz, sign);
    for (size_t i = 0; i < M; ++i) { out[2 * i] = b[i].real(); out[2 * i + 1] = b[i].imag(); }
}
}
