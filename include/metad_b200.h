/* metad_b200.h -- C ABI of the B200-native (sm_100a) CV + bias-force hot path.
 *
 * Drop-in boundary for jglaser/metadynamics-plugin's GPU kernel drivers.  Each entry point names the
 * reference interface it replaces (paths relative to the reference's metadynamics/ directory).  The
 * reference drivers are C++-linkage free functions taking HOOMD types (BoxDim, Index2D, GPUPartition,
 * CachedAllocator) and launching on the default stream; here everything is extern "C", plain pointers
 * and sizes, an explicit stream, and an int status (0 = ok, <0 = error; text via metad_last_error()).
 *
 * Conventions
 *   - d_* pointers are DEVICE pointers owned by the caller and valid for the duration of the call.
 *   - postype / force arrays use HOOMD's single-precision layout: float4 {x,y,z,w}; postype.w holds
 *     the integer type id as raw bits (__scalar_as_int), force.w the per-particle energy (always 0).
 *   - Scalars that the reference reads back to the host every step (CV value, bias factor dV/ds) live in
 *     device memory as double, so a whole step can be enqueued without a host round trip.
 *   - All calls are asynchronous with respect to the host unless stated otherwise.
 *   - There is no CPU fallback: without a CUDA device every call fails with METAD_ERR_CUDA.
 */
#ifndef METAD_B200_H
#define METAD_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* metad_stream_t; /* == cudaStream_t */

enum {
    METAD_OK = 0,
    METAD_ERR_INVALID = -1,     /* bad argument */
    METAD_ERR_CUDA = -2,        /* CUDA runtime error (see metad_last_error) */
    METAD_ERR_UNSUPPORTED = -3, /* valid in the reference, not implemented here (e.g. a mesh beyond 1024 points per direction) */
    METAD_ERR_STATE = -4        /* call order violated (e.g. forces before cv) */
};

/* HOOMD BoxDim flattened to a POD (hoomd/BoxDim.h; call sites OrderParameterMesh.cc:543,570-573,
 * LamellarOrderParameter.cc:151-159).  lo = -L/2.  Device code derives the single-precision members
 * exactly as a SINGLE_PRECISION HOOMD build would: L=(float)L, hi=L/2, lo=-hi. */
typedef struct metad_box {
    double L[3];    /* box lengths Lx, Ly, Lz */
    double tilt[3]; /* xy, xz, yz */
} metad_box;

int metad_version(void);
const char* metad_last_error(void);

/* ------------------------------------------------------------------------------------------------
 * LamellarOrderParameter
 * replaces gpu_calculate_fourier_modes (LamellarOrderParameterGPU.cuh:21-29, .cu:104-147) + the host sum
 * of LamellarOrderParameterGPU.cc:79-89, and gpu_compute_sq_forces (LamellarOrderParameterGPU.cuh:45-54).
 * ---------------------------------------------------------------------------------------------- */
typedef struct metad_lamellar metad_lamellar;

/* lattice_vectors: 3*n_wave ints (Miller indices); mode: ntypes per-type coefficients a(type). */
int metad_lamellar_create(metad_lamellar** out, int n_wave, const int* lattice_vectors, int ntypes,
                          const double* mode);
int metad_lamellar_destroy(metad_lamellar* p);

/* Local partial Fourier modes F_k = sum_j a_j (cos q_k.r_j, sin q_k.r_j) over the N local particles ->
 * d_modes[2*n_wave] (device, double).  If finalize != 0 also writes CV = (sum_k Re F_k)/N_global to
 * *d_cv (single-GPU case).  Multi-GPU: call with finalize=0, all-reduce d_modes, then
 * metad_lamellar_finalize (reference: MPI_Allreduce, LamellarOrderParameterGPU.cc:70-77). */
int metad_lamellar_modes(metad_lamellar* p, const float* d_postype, unsigned N, unsigned N_global,
                         const metad_box* global_box, double* d_modes, int finalize, double* d_cv,
                         metad_stream_t stream);
int metad_lamellar_finalize(metad_lamellar* p, const double* d_modes, unsigned N_global, double* d_cv,
                            metad_stream_t stream);
/* F_j = (bias/N_global) sum_k 2 a_j sin(q_k.r_j) q_k, force.w = 0; bias read from *d_bias (device). */
int metad_lamellar_forces(metad_lamellar* p, const float* d_postype, float* d_force, unsigned N,
                          unsigned N_global, const metad_box* global_box, const double* d_bias,
                          metad_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * OrderParameterMesh
 * one plan replaces the cuFFT plans + scratch of OrderParameterMeshGPU.cc:75-155; metad_mesh_cv replaces
 * gpu_bin_particles + gpu_assign_binned_particles_to_mesh + gpu_compute_mode_sq + cufftExecC2C (x2) +
 * gpu_update_meshes + gpu_compute_cv (OrderParameterMeshGPU.cuh:9-88; OrderParameterMeshGPU.cc:157-364,
 * 454-506); metad_mesh_forces replaces gpu_compute_forces (OrderParameterMeshGPU.cuh:51-62).
 * ---------------------------------------------------------------------------------------------- */
typedef struct metad_mesh metad_mesh;

/* nx,ny,nz: mesh points, 1 <= n <= 1024 each (the reference takes any size on one rank, OrderParameterMesh.cc:70-79);
 * powers of two with 32 <= nx <= 1024, 16 <= ny,nz <= 512 run through the tiled kernels, every other size through the
 * general path (csrc/mesh_general.cuh: same results, slower).  mode: ntypes per-type coefficients.  The box of every call
 * may be triclinic (metad_box.tilt, key 16 of metad_mesh_set). */
int metad_mesh_create(metad_mesh** out, unsigned nx, unsigned ny, unsigned nz, int ntypes, const double* mode);
int metad_mesh_destroy(metad_mesh* p);

/* getCurrentValue: assignParticles + updateMeshes + computeCV.  Writes the CV to *d_cv (device double).
 * Keeps the inverse-transformed mesh and the particle tile order for a following metad_mesh_forces.
 * The density is accumulated in 32-bit fixed point (bitwise independent of the particle order); the tile order
 * (a permutation of the particles, tile by tile) is rebuilt only every `period` calls or when particles drifted. */
int metad_mesh_cv(metad_mesh* p, const float* d_postype, unsigned N, unsigned N_global, const metad_box* box,
                  double* d_cv, metad_stream_t stream);
/* interpolateForces for the positions last passed to metad_mesh_cv; bias read from *d_bias. */
int metad_mesh_forces(metad_mesh* p, const float* d_postype, float* d_force, unsigned N, unsigned N_global,
                      const metad_box* box, const double* d_bias, metad_stream_t stream);

/* Kernel table of cv.mesh.set_kernel (OrderParameterMesh::setTable / setUseTable, OrderParameterMesh.cc:148-189): only the
 * derivative table dK enters the hot path, through the k-space virial (computeVirial, :970-1050).  n entries on
 * [k_min, k_max]; n == 0 only switches use_table.  With metad_mesh_set(p, 13, 1) the fused z sweep of metad_mesh_cv also
 * accumulates the virial sums and the arg-max of |f_k|^2 (computeQmax, :1108-1179); metad_mesh_get(p, 10, double[12])
 * returns them: virial xx, xy, xz, yy, yz, zz (to be multiplied by the bias factor), q_max x, y, z, sq_max, flat index, |f|^2. */
int metad_mesh_set_table(metad_mesh* p, const double* dK, unsigned n, double k_min, double k_max, int use_table);

/* Multi-GPU: z-slab decomposition (reference: HOOMD domain decomposition + CommunicatorGrid ghost exchange + dfft,
 * OrderParameterMesh.cc:231-315, 659-746).  Rank r of n_ranks owns the planes [r nz/P, (r+1) nz/P) and the particles
 * inside them.  The library provides the five compute stages of a step; the caller issues the collectives between
 * them (ops.MeshSlab does it with NCCL through torch.distributed):
 *
 *   metad_mesh_slab_spread      local density; d_sums[3] = local {sum a^2, sum a, #particles outside the slab};
 *                               d_ghost_send[2][ny*nx + 4] (int32) = two halo messages: the fixed-point density of plane
 *                               z0-1 (for rank r-1) and of plane z0+nz/P (for rank r+1), each followed by 4 ints of which
 *                               the first holds the bits of the sender's 1/scale
 *      -> all-reduce(d_sums), neighbour exchange of the two messages
 *   metad_mesh_slab_fft_x       adds d_ghost_recv[2][ny*nx + 4] ([0] from rank r-1, [1] from rank r+1; integer addition
 *                               when the scales agree, so the density equals the single-GPU one bit for bit), removes the
 *                               global mean, x FFT;
 *                               d_send = M/P/2 complex, already packed [dest rank][plane][y][kx in pencil]
 *      -> all-to-all (equal splits): d_pencil = [nz][ny][nx/2/P] complex, planes in rank order = global z order
 *   metad_mesh_slab_fft_yz      y FFT, fused z FFT + convolution + inverse z FFT, inverse y FFT on the pencil, in
 *                               place; *d_cv_partial = this rank's share of the CV
 *      -> all-reduce(d_cv_partial), all-to-all back (the pencil buffer is contiguous per destination rank)
 *   metad_mesh_slab_fft_x_inv   unpacks d_recv while transforming; copies the first and the last local plane of
 *                               Re IFFT(G) to d_planes_out[0], [1] (to send to rank r-1 / r+1)
 *      -> neighbour exchange: d_ghost_inv[0] = last plane of rank r-1, d_ghost_inv[1] = first plane of rank r+1
 *   metad_mesh_slab_forces      interpolateForces for the local particles
 * Requirements: nx, ny, nz powers of two (32 <= nx <= 1024, 16 <= ny,nz <= 512; the reference has the same rule under domain
 * decomposition, OrderParameterMesh.cc:70-79), n_ranks a power of two, nz/n_ranks >= 8, nx/2/n_ranks >= 16.  `global_box` is the global box. */
int metad_mesh_slab_create(metad_mesh** out, unsigned nx, unsigned ny, unsigned nz, unsigned n_ranks, unsigned rank, int ntypes,
                           const double* mode);
int metad_mesh_slab_spread(metad_mesh* p, const float* d_postype, unsigned N_local, const metad_box* global_box, double* d_sums,
                           int* d_ghost_send, metad_stream_t stream);
int metad_mesh_slab_fft_x(metad_mesh* p, const int* d_ghost_recv, const double* d_sums_global, float* d_send,
                          metad_stream_t stream);
int metad_mesh_slab_fft_yz(metad_mesh* p, float* d_pencil, const double* d_sums_global, unsigned N_global, double* d_cv_partial,
                           metad_stream_t stream);
int metad_mesh_slab_fft_x_inv(metad_mesh* p, const float* d_recv, float* d_planes_out /* [2][ny][nx] */, metad_stream_t stream);
int metad_mesh_slab_forces(metad_mesh* p, const float* d_ghost_inv, const float* d_postype, float* d_force, unsigned N_local,
                           unsigned N_global, const metad_box* global_box, const double* d_bias, metad_stream_t stream);

/* The same decomposition over PEER MEMORY (NVLink, CUDA IPC) -- no library collective inside a step.  Every rank owns an
 * arena (pencil, receive buffer, halo slots, per-rank scalar tables, barrier flags) that its peers map; the transposes of the
 * distributed FFT are fused into the FFT sweeps (the x forward pass stores each kx pencil straight into the owner's
 * arena, the inverse y pass each plane), halos and partial sums are pushed with plain P2P stores, and the ranks meet at
 * four flag barriers per step which also perform the tiny all-reduces in rank order (identical result on every rank).
 *   metad_mesh_slab_p2p_arena     allocate the arena, return its CUDA IPC handle (64 bytes) to be all-gathered by the caller
 *   metad_mesh_slab_p2p_connect   map the peers: handles = n_ranks x 64 bytes in rank order (own slot ignored)
 *   metad_mesh_slab_p2p_connect_local   same for "ranks" living in ONE process (tests): plans[] in rank order
 *   metad_mesh_slab_p2p_cv        stage = -1: the whole step (spread ... CV, ready for forces); every rank must call it
 *                                 the same number of times.  stage = 0..4: one stage without its barrier wait, for
 *                                 bulk-synchronous emulation of all ranks in one process (stage k for every rank, then k+1).
 *                                 *d_cv receives the global CV on every rank.
 *   metad_mesh_slab_p2p_forces    interpolateForces for the local particles (halo planes are already in the arena)
 * Up to 8 ranks (one NVSwitch domain).  metad_mesh_get(p, 6, unsigned[2]) reports a barrier time-out / misplaced particles. */
int metad_mesh_slab_p2p_arena(metad_mesh* p, void* handle_out /* 64 bytes */, unsigned long long* bytes_out);
int metad_mesh_slab_p2p_connect(metad_mesh* p, const void* handles);
int metad_mesh_slab_p2p_connect_local(metad_mesh* p, metad_mesh* const* plans);
int metad_mesh_slab_p2p_cv(metad_mesh* p, const float* d_postype, unsigned N_local, unsigned N_global, const metad_box* global_box,
                           double* d_cv, int stage, metad_stream_t stream);
int metad_mesh_slab_p2p_forces(metad_mesh* p, const float* d_postype, float* d_force, unsigned N_local, unsigned N_global,
                               const metad_box* global_box, const double* d_bias, metad_stream_t stream);

/* Introspection for parity tests (synchronous, copies to HOST buffers):
 *   which = 0: cell coordinates (ix,iy,iz) per particle as computed by the spread, int[3*N], input order (needs key 3)
 *           1: density mesh rho, float[nx*ny*nz], index x + nx*(y + ny*z) (needs key 1)
 *           2: Re(inverse mesh), float[nx*ny*nz]
 *           3: sum of mode^2 (double[1])
 *           4: per-stage milliseconds of the last cv + forces pair, float[8] (needs key 2): tile order (0 on calls that
 *              reuse it), spread, fft x fwd, fft y fwd, fft z fused, fft y inv, fft x inv, gather
 *           5: statistics, double[6]: rebuilds of the tile order so far; of the last spread: particles that took the
 *              direct path (drifted out of their padded tile), particles outside the slab, padded-tile cells past 1/8 of
 *              the fixed-point range; the fixed-point scale; calls since the last rebuild
 *           6: peer-memory mode, unsigned[2]: {a barrier timed out, particles outside their slab summed over the ranks}
 *           7: unsigned long long: CUDA-graph replays so far
 *           8: peer-memory step with profiling on (key 2), float[12]: milliseconds of spread, halo push, barrier 1, x forward,
 *              barrier 2, y forward, fused z, y inverse, barrier 3, x inverse, halo push, barrier 4 of the last step        */
int metad_mesh_get(metad_mesh* p, int which, void* h_out);
/* knobs: key 0 = rebuild period of the tile order in calls (default 32; value 0 = rebuild at the next call)
 *        key 1 = keep a copy of rho for metad_mesh_get(1)      key 2 = record per-stage CUDA events (profiling)
 *        key 3 = record the cell index of every particle for metad_mesh_get(0)
 *        key 4 = CUDA-graph replay: everything metad_mesh_cv / metad_mesh_slab_p2p_cv enqueue after the tile-order
 *                decision is captured once per argument signature (pointers, N, box, stream) and replayed with one launch
 *                (metad_mesh_get(p, 7, unsigned long long*) = replays so far)
 *        key 5 = peer-memory mode: 1 folds the inter-rank signal / wait into the producer / consumer kernels, 0 (default,
 *                measured faster) uses separate barrier launches
 *        key 6 = order of the particles inside a tile: 1 (default) bank order, 0 layer order (mesh_kernels.cuh); results
 *                do not depend on it
 *        key 7 = programmatic dependent launch of the per-step kernels: 1 (default) on, 0 off
 *        key 8 = peer-memory mode: halo push and the barrier after it in one launch: 1 on, 0 (default, measured faster) off
 *        key 9..15 = kernel variants for measurements (particle cache, tensor-map flush / gather, debug, q_max / virial
 *                epilogues (13), use_table (14), fused x+y sweeps (15)); see csrc/mesh.cu
 *        key 16 = triclinic boxes (box->tilt != 0): 1 (default) in-cell offsets exactly as the reference computes them --
 *                OrderParameterMesh.cc:571-573 / 806-808 take makeFraction(shift_cart + lo), which shears `lo` too, so every
 *                offset carries a constant ((xz - yz xy) Lz + xy Ly) nx / (2 Lx) along x and yz Lz ny / (2 Ly) along y and
 *                TSC weight beyond |x| = 3/2 is dropped; 0 = geometrically correct offsets (set before the first call) */
int metad_mesh_set(metad_mesh* p, int key, long value);

/* ------------------------------------------------------------------------------------------------
 * IntegratorMetaDynamics grid bias
 * metad_grid_step replaces, on the device and without host round trips, the root-rank body of
 * IntegratorMetaDynamics::updateBiasPotential (IntegratorMetaDynamics.cc:363-451): updateHistogram,
 * updateSigmaGrid, the well-tempered scale, gpu_update_grid (IntegratorMetaDynamics.cuh:1-11),
 * updateReweightedEstimator, the delta merge, biasPotentialDerivative and interpolateGrid.
 * ---------------------------------------------------------------------------------------------- */
typedef struct metad_grid metad_grid;

int metad_grid_create(metad_grid** out, int n_cv, const double* cv_min, const double* cv_max,
                      const unsigned* num_points, const double* sigma, double W, double T_shift, double T,
                      unsigned stride, int add_bias, int well_tempered);
int metad_grid_destroy(metad_grid* g);
/* d_cv_values: n_cv doubles (device); d_bias_out: n_cv doubles (device) = dV/ds_i. */
int metad_grid_step(metad_grid* g, unsigned timestep, const double* d_cv_values, double* d_bias_out,
                    metad_stream_t stream);
/* Multiple walkers (IntegratorMetaDynamics.cc:65-71, 392-410: MPI_Allreduce of the four delta arrays over the partition
 * communicator between the Gaussian deposit and the merge): the step in two halves.  On every step
 *   metad_grid_step_deposit   updateHistogram; on deposit steps also updateSigmaGrid, the well-tempered scale and the
 *                             Gaussian into grid_delta;
 *   [deposit steps only]      metad_grid_deltas_export -> all-reduce (sum) over the walkers -> metad_grid_deltas_import;
 *   metad_grid_step_merge     on deposit steps updateReweightedEstimator + the merge of the deltas; always the hand-off
 *                             (dV/ds_i, V(s), weight).
 * deposit + merge without an exchange in between == metad_grid_step.  The delta arrays travel as two device buffers:
 * double[2G] = grid_delta | sigma_grid_delta, unsigned[2G] = hist_delta | hist_gauss_delta. */
int metad_grid_step_deposit(metad_grid* g, unsigned timestep, const double* d_cv_values, metad_stream_t stream);
int metad_grid_step_merge(metad_grid* g, unsigned timestep, const double* d_cv_values, double* d_bias_out,
                          metad_stream_t stream);
int metad_grid_is_deposit_step(const metad_grid* g, unsigned timestep);
int metad_grid_deltas_export(metad_grid* g, double* d_out_2G, unsigned* d_out_u_2G, metad_stream_t stream);
int metad_grid_deltas_import(metad_grid* g, const double* d_in_2G, const unsigned* d_in_u_2G, metad_stream_t stream);
/* Adaptive Gaussians (setAdaptive, IntegratorMetaDynamics.cc:333-341): the inverse sigma matrix (n_cv x n_cv, row-major,
 * host) that computeSigma (:1205-1294) derives from the CV derivatives replaces diag(1/sigma_i); see metad_force_dot. */
int metad_grid_set_sigma_inv(metad_grid* g, const double* sigma_inv);
int metad_grid_set_flags(metad_grid* g, int add_bias, int well_tempered, unsigned stride);
int metad_grid_reset_histogram(metad_grid* g, metad_stream_t stream);
/* Synchronous host access for dump/restart (writeGrid/readGrid, IntegratorMetaDynamics.cc:831-1000).
 * which: 0 grid, 1 grid_reweighted, 2 grid_weight, 3 sigma_grid (double[G]); 4 hist, 5 hist_gauss (unsigned[G]) */
int metad_grid_download(metad_grid* g, int which, void* h_out);
int metad_grid_upload(metad_grid* g, int which, const void* h_in);
/* h_out4: curr_bias_potential, curr_reweight, num_gaussians, out-of-bounds warning count */
int metad_grid_scalars(metad_grid* g, double* h_out4);
int metad_grid_set_num_gaussians(metad_grid* g, unsigned n);
unsigned metad_grid_num_elements(const metad_grid* g);

/* ------------------------------------------------------------------------------------------------
 * CollectiveVariable umbrella (CollectiveVariable.cc:22-66, 68-106), evaluated on the device so the
 * CV value never has to visit the host: *d_bias_out = *d_bias_in + umbrella'(cv); optional energy.
 * kind: 0 none, 1 linear, 2 harmonic, 3 wall, 4 gaussian (CollectiveVariable.h:35-42)
 * ---------------------------------------------------------------------------------------------- */
int metad_umbrella_apply(int kind, double cv0, double kappa, double width_flat, double scale,
                         const double* d_cv, const double* d_bias_in, double* d_bias_out,
                         double* d_energy_out, metad_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Scalar all-reduce over peer memory (NVLink, CUDA IPC) for the sharded CVs
 * replaces the host-synchronous MPI_Allreduce of a few Scalars per step: the 2 n_q Fourier modes
 * (LamellarOrderParameterGPU.cc:70-77), the potential energy (WellTemperedEnsemble.cc:58-64, CollectiveWrapper.cc:63-69),
 * computeSigma's matrix (IntegratorMetaDynamics.cc:1265-1274).  One 256-thread kernel on the caller's stream, graph-capturable;
 * no library collective.  Set-up: every rank creates a handle, the 64-byte IPC handles are all-gathered once (any
 * out-of-band channel), every rank connects.
 * ---------------------------------------------------------------------------------------------- */
typedef struct metad_peer metad_peer;
int metad_peer_create(metad_peer** out, unsigned n_ranks, unsigned rank);           /* up to 8 ranks (one NVSwitch domain) */
int metad_peer_destroy(metad_peer* p);
int metad_peer_handle(metad_peer* p, void* handle_out64);
int metad_peer_connect(metad_peer* p, const void* handles /* n_ranks x 64 bytes, rank order */);
int metad_peer_connect_local(metad_peer* p, metad_peer* const* all /* all ranks of one process, rank order (tests) */);
/* d_data[0..n) (device doubles, n <= 32) <- sum over the ranks, added in rank order (identical bits on every rank).
 * phase: -1 in multi-process use; with connect_local 0 (publish; every rank first) then 1 (reduce). */
int metad_peer_allreduce_sum(metad_peer* p, double* d_data, unsigned n, int phase, metad_stream_t stream);
int metad_peer_status(metad_peer* p, unsigned* timed_out);

/* ------------------------------------------------------------------------------------------------
 * WellTemperedEnsemble
 * replaces gpu_reduce_potential_energy and gpu_scale_netforce (WellTemperedEnsemble.cuh:3-19).
 * ---------------------------------------------------------------------------------------------- */
/* *d_dst = value, asynchronously on `stream` (the value is a kernel argument: no staging buffer, no synchronisation);
 * how CVs that are host scalars (AspectRatio, Density) enter the device-resident step of IntegratorMetaDynamics::update */
int metad_set_double(double* d_dst, double value, metad_stream_t stream);
/* *d_out = scale * sum_n (f_i[n].xyz . f_j[n].xyz): the sums of products of CV derivatives of computeSigma
 * (IntegratorMetaDynamics.cc:1238-1247; scale = sigma_g^2), f = Scalar4 force arrays filled by computeDerivatives */
int metad_force_dot(const float* d_force_i, const float* d_force_j, unsigned N, double scale, double* d_out,
                    metad_stream_t stream);
/* *d_pe = sum_i net_force[i].w + external_energy */
int metad_wte_reduce(const float* d_net_force, unsigned N, double external_energy, double* d_pe,
                     metad_stream_t stream);
/* net_force.xyz, net_torque.xyz, six virial rows (pitch-strided) *= (1 + *d_bias) */
int metad_wte_scale(float* d_net_force, float* d_net_torque, float* d_net_virial, unsigned pitch, unsigned N,
                    const double* d_bias, metad_stream_t stream);

/* Stand-in for HOOMD's net-force summation (Integrator::computeNetForceGPU, not part of the plugin): the shim
 * integrator of the host layer uses it to build the net force the WellTemperedEnsemble CV reads.
 * net[i] = (init ? 0 : net[i]) + f[i] for all four components. */
int metad_accumulate_force(float* d_net_force, const float* d_force, unsigned N, int init, metad_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* METAD_B200_H */
