// fft_emul.cu -- CPU emulation of the mesh FFT sweeps; see emul_fft.h.  Host-only program (nvcc, no GPU needed).
#include "emul_fft.h"

// usage: fft_emul nx ny nz N mode_sq stage in.bin out.bin
//   stage: 0 = forward only (x,y fwd; dumps packed half spectrum before z), 1 = full pipeline (dumps inverse mesh)
int main(int argc, char** argv) {
    if (argc < 9) { fprintf(stderr, "usage\n"); return 2; }
    unsigned nx = atoi(argv[1]), ny = atoi(argv[2]), nz = atoi(argv[3]);
    double N = atof(argv[4]), mode_sq = atof(argv[5]);
    int stage = atoi(argv[6]);
    size_t M = (size_t)nx * ny * nz;
    std::vector<float> mesh(M);
    FILE* f = fopen(argv[7], "rb");
    if (!f || fread(mesh.data(), 4, M, f) != M) { fprintf(stderr, "read error\n"); return 2; }
    fclose(f);
    float2* buf = reinterpret_cast<float2*>(mesh.data());
    const unsigned nxh = nx / 2;
    DISPATCH(nxh, x_fwd<LL>(buf, ny * nz));
    DISPATCH(ny, (y_pass<LL, -1>(buf, nxh, nz)));
    double e = 0.0;
    if (stage == 1) {
        float inv_n = (float)(1.0 / N);
        float d = (float)(0.5 * mode_sq / N / N);
        DISPATCH(nz, e += z_plane0<LL>(buf, nx, ny, inv_n, d));
        DISPATCH(nz, e += z_fused<LL>(buf, nx, ny, inv_n, d));
        DISPATCH(ny, (y_pass<LL, +1>(buf, nxh, nz)));
        DISPATCH(nxh, x_inv<LL>(buf, ny * nz));
    }
    f = fopen(argv[8], "wb");
    fwrite(mesh.data(), 4, M, f);
    double cv = 0.5 * e;
    fwrite(&cv, 8, 1, f);
    fclose(f);
    return 0;
}
