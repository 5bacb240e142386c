"""Multi-GPU drivers: one process per GPU, torch.distributed (NCCL) for the collectives, the C ABI for the compute.

LamellarSharded   any particle partition; one all-reduce of 2*n_wave doubles per step
                  (reference: MPI_Allreduce, LamellarOrderParameterGPU.cc:70-77).
MeshSlab          z-slab decomposition of the particle mesh (reference: HOOMD domain decomposition + CommunicatorGrid
                  ghost exchange + dfft, OrderParameterMesh.cc:231-315, 659-746): two neighbour plane exchanges,
                  two all-to-all transposes (slab <-> kx pencil) and two tiny all-reduces per step.
The bias grid is replicated: every rank applies the identical metad_grid_step after the CV all-reduce.

MeshSlabP2P       the same decomposition over peer memory (NVLink, CUDA IPC): no NCCL call inside a step -- the transposes
                  are fused into the FFT sweeps, halos / partial sums are P2P stores, ranks meet at flag barriers.
                  torch.distributed is only used once, to all-gather the 64-byte IPC handles.
The communication pattern is written against a small `Comm` interface so that the same driver code runs
  * over NCCL (TorchComm, one process per GPU),
  * over gloo on CPU tensors with a numpy stand-in for the local compute (tests/test_distributed_cpu.py), and
  * over LocalComm: all ranks emulated in one process, bulk-synchronously, on one GPU (tests/test_gpu_sharded.py).
"""
import ctypes as C

import numpy as np
import torch

from ._abi import Box, check, lib


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


# ---------------------------------------------------------------------------------------------------- communicators
class TorchComm:
    """torch.distributed process group (NCCL on GPUs, gloo on CPU)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.rank, self.size = dist.get_rank(group), dist.get_world_size(group)

    def all_reduce_sum(self, t):
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)

    def all_to_all(self, out, inp):
        self.dist.all_to_all_single(out, inp, group=self.group)

    def neighbour_exchange(self, send_down, send_up, recv_from_down, recv_from_up):
        """send_down -> rank-1, send_up -> rank+1 (periodic); receive the matching planes."""
        d = self.dist
        down, up = (self.rank - 1) % self.size, (self.rank + 1) % self.size
        if self.size == 1:
            recv_from_up.copy_(send_down)
            recv_from_down.copy_(send_up)
            return
        ops = [d.P2POp(d.isend, send_down, down, self.group), d.P2POp(d.isend, send_up, up, self.group),
               d.P2POp(d.irecv, recv_from_down, down, self.group), d.P2POp(d.irecv, recv_from_up, up, self.group)]
        if self.size == 2:
            # down == up: pair the two messages by tag order (both ranks post send_down first, then send_up):
            # what rank r sends "down" is what rank 1-r must receive "from up"
            ops = [d.P2POp(d.isend, send_down, down, self.group), d.P2POp(d.irecv, recv_from_up, up, self.group),
                   d.P2POp(d.isend, send_up, up, self.group), d.P2POp(d.irecv, recv_from_down, down, self.group)]
        for r in d.batch_isend_irecv(ops):
            r.wait()


class PeerAllReduce:
    """All-reduce of up to 32 doubles over peer memory (metad_peer_*, csrc/peer.cu): one small kernel on the current stream
    instead of a library collective -- the exchange of the Lamellar Fourier modes, of the WTE potential energy, of
    computeSigma's matrix.  torch.distributed is used once, to all-gather the 64-byte IPC handles."""

    def __init__(self, comm):
        self.h = C.c_void_p()
        check(lib.metad_peer_create(C.byref(self.h), comm.size, comm.rank))
        if comm.size > 1:
            handle = (C.c_ubyte * 64)()
            check(lib.metad_peer_handle(self.h, handle))
            mine = torch.tensor(list(handle), dtype=torch.uint8, device="cuda")
            allh = torch.empty(comm.size * 64, dtype=torch.uint8, device="cuda")
            comm.dist.all_gather_into_tensor(allh, mine, group=comm.group)
            buf = (C.c_ubyte * (64 * comm.size))(*allh.cpu().tolist())
            err = None
            try:
                check(lib.metad_peer_connect(self.h, buf))
            except Exception as e:
                err = e
            ok = torch.tensor([0 if err else 1], dtype=torch.int32, device="cuda")
            comm.dist.all_reduce(ok, op=comm.dist.ReduceOp.MIN, group=comm.group)
            if ok.item() == 0:
                raise RuntimeError("peer memory is not available on every rank (%s)" % (err or "another rank failed"))

    def __del__(self):
        if getattr(self, "h", None):
            lib.metad_peer_destroy(self.h)
            self.h = None

    def all_reduce_sum(self, t):
        assert t.is_cuda and t.dtype == torch.float64 and t.is_contiguous() and t.numel() <= 32
        check(lib.metad_peer_allreduce_sum(self.h, _ptr(t), t.numel(), -1, _stream()))

    def timed_out(self):
        out = C.c_uint(0)
        check(lib.metad_peer_status(self.h, C.byref(out)))
        return bool(out.value)


class PeerComm:
    """TorchComm whose small float64 all-reduces go over peer memory (PeerAllReduce); everything else is delegated."""

    def __init__(self, comm):
        self.comm, self.peer = comm, PeerAllReduce(comm)
        self.rank, self.size, self.dist, self.group = comm.rank, comm.size, comm.dist, comm.group

    def all_reduce_sum(self, t):
        if t.is_cuda and t.dtype == torch.float64 and t.numel() <= 32 and t.is_contiguous():
            self.peer.all_reduce_sum(t)
        else:
            self.comm.all_reduce_sum(t)

    def all_to_all(self, out, inp):
        self.comm.all_to_all(out, inp)

    def neighbour_exchange(self, *a):
        self.comm.neighbour_exchange(*a)


class LocalPeerGroup:
    """All ranks of a peer all-reduce in one process on one GPU (tests): publish for every rank, then reduce for every rank."""

    def __init__(self, size):
        self.size = size
        self.hs = [C.c_void_p() for _ in range(size)]
        for r, h in enumerate(self.hs):
            check(lib.metad_peer_create(C.byref(h), size, r))
        arr = (C.c_void_p * size)(*[h.value for h in self.hs])
        for h in self.hs:
            check(lib.metad_peer_connect_local(h, arr))

    def __del__(self):
        for h in getattr(self, "hs", []):
            lib.metad_peer_destroy(h)
        self.hs = []

    def all_reduce_sum(self, ts):
        if self.size == 1:
            check(lib.metad_peer_allreduce_sum(self.hs[0], _ptr(ts[0]), ts[0].numel(), -1, _stream()))
            return
        for phase in (0, 1):
            for h, t in zip(self.hs, ts):
                check(lib.metad_peer_allreduce_sum(h, _ptr(t), t.numel(), phase, _stream()))


class LocalComm:
    """All ranks in one process (bulk-synchronous emulation): every collective takes the list of per-rank tensors."""

    def __init__(self, size):
        self.size = size

    def all_reduce_sum(self, ts):
        s = torch.stack(list(ts)).sum(0)
        for t in ts:
            t.copy_(s)

    def all_to_all(self, outs, inps):
        P = self.size
        for r in range(P):
            o = outs[r].view(P, -1)
            for q in range(P):
                o[q].copy_(inps[q].view(P, -1)[r])

    def neighbour_exchange(self, send_down, send_up, recv_from_down, recv_from_up):
        P = self.size
        for r in range(P):
            recv_from_up[(r - 1) % P].copy_(send_down[r])
            recv_from_down[(r + 1) % P].copy_(send_up[r])


# ---------------------------------------------------------------------------------------------------- raw device buffers
class _DevicePtr:
    """__cuda_array_interface__ view of a raw device pointer handed over by the host classes (no copy)."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = dict(shape=(int(n),), typestr=typestr, data=(int(ptr), False), version=2)


def tensor_from_ptr(ptr, n, dtype):
    """torch tensor aliasing n elements of device memory at `ptr` (float64 or int32)."""
    return torch.as_tensor(_DevicePtr(ptr, n, {torch.float64: "<f8", torch.int32: "<i4"}[dtype]), device="cuda")


def walker_allreduce_hook(comm):
    """Callable for IntegratorMetaDynamics.walker_allreduce: sums the two delta buffers over the walkers of `comm` (one
    walker per process; reference: MPI_Allreduce over the partition communicator, IntegratorMetaDynamics.cc:401-408)."""
    def hook(ptr_d, n_d, ptr_u, n_u):
        comm.all_reduce_sum(tensor_from_ptr(ptr_d, n_d, torch.float64))
        comm.all_reduce_sum(tensor_from_ptr(ptr_u, n_u, torch.int32))
        torch.cuda.current_stream().synchronize()
    return hook


# ---------------------------------------------------------------------------------------------------- lamellar
class LamellarSharded:
    def __init__(self, comm, mode, lattice_vectors, lamellar_factory=None):
        """lamellar_factory: class with the interface of ops.Lamellar (tests substitute a CPU stand-in)."""
        if lamellar_factory is None:
            from .ops import Lamellar as lamellar_factory
        self.comm, self.lam = comm, lamellar_factory(mode, lattice_vectors)

    def compute_cv(self, postype_local, n_global, box):
        self.lam.compute_modes(postype_local, n_global, box, finalize=False)
        self.comm.all_reduce_sum(self.lam.modes)
        return self.lam.finalize(n_global)

    def forces(self, postype_local, n_global, box, bias, out=None):
        return self.lam.forces(postype_local, n_global, box, bias, out=out)


class WTESharded:
    """WellTemperedEnsemble over particle shards (any partition).  Reference: the local sum of net_force.w is reduced with
    one MPI_Allreduce of a double and the external energy is added once (WellTemperedEnsemble.cc:46-66); the scaling of
    net force / torque / virial by (1 + bias) is local to every rank (:135-188)."""

    def __init__(self, comm, reduce_fn=None, scale_fn=None):
        """reduce_fn / scale_fn: functions with the interface of ops.wte_reduce / ops.wte_scale (tests substitute CPU
        stand-ins)."""
        if reduce_fn is None or scale_fn is None:
            from . import ops
            reduce_fn, scale_fn = reduce_fn or ops.wte_reduce, scale_fn or ops.wte_scale
        self.comm, self.reduce_fn, self.scale_fn = comm, reduce_fn, scale_fn
        self.cv = None

    def compute_cv(self, net_force_local, external_energy=0.0):
        # every rank contributes its particles only; the external energy enters once, after the reduction
        self.cv = self.reduce_fn(net_force_local, 0.0, out=self.cv)
        self.comm.all_reduce_sum(self.cv)
        self.cv += float(external_energy)
        return self.cv

    def scale(self, net_force_local, net_torque_local, net_virial_local, pitch, bias):
        return self.scale_fn(net_force_local, net_torque_local, net_virial_local, pitch, bias)


class WalkerBias:
    """Multiple walkers sharing one bias potential (reference: setMultipleWalkers + the partition communicator,
    IntegratorMetaDynamics.cc:65-71, 392-410): every process is one walker with its own replica of the bias grid; on deposit
    steps the four delta arrays (Gaussian, sigma grid, histogram, Gaussian histogram) are summed over the walkers between
    the deposit and the merge, so all replicas stay identical.  `grid`: ops.BiasGrid, or any object with its step_deposit /
    is_deposit_step / deltas_export / deltas_import / step_merge interface (tests substitute a CPU stand-in)."""

    def __init__(self, comm, grid):
        self.comm, self.grid = comm, grid

    def step(self, timestep, cv_values):
        g = self.grid
        g.step_deposit(timestep, cv_values)
        if g.is_deposit_step(timestep):
            dd, du = g.deltas_export()
            self.comm.all_reduce_sum(dd)
            self.comm.all_reduce_sum(du)
            g.deltas_import(dd, du)
        return g.step_merge(timestep, cv_values)


# ---------------------------------------------------------------------------------------------------- mesh
def slab_of(z, L, nz, n_ranks):
    """Owner rank of a particle: the slab holding its global mesh plane (single-precision cell rule of the path)."""
    Lf = np.float32(L)
    Linv = np.float32(1) / Lf                                   # HOOMD's BoxDim::makeFraction multiplies by the stored 1/L
    f = (np.asarray(z, np.float32) - (-(Lf / np.float32(2)))) * Linv
    iz = (f * np.float32(nz)).astype(np.int64)
    iz[iz == nz] = 0
    return iz // (nz // n_ranks)


class MeshSlabRank:
    """The five compute stages of one rank (thin wrappers of metad_mesh_slab_*) and its communication buffers."""

    def __init__(self, nx, ny, nz, n_ranks, rank, mode, device="cuda"):
        m = np.ascontiguousarray(mode, dtype=np.float64)
        self.dims, self.P, self.rank = (nx, ny, nz), n_ranks, rank
        self.nzl, self.kxl = nz // n_ranks, nx // 2 // n_ranks
        self.h = C.c_void_p()
        check(lib.metad_mesh_slab_create(C.byref(self.h), nx, ny, nz, n_ranks, rank, len(m), m.ctypes.data_as(C.POINTER(C.c_double))))
        f32, f64 = dict(dtype=torch.float32, device=device), dict(dtype=torch.float64, device=device)
        i32 = dict(dtype=torch.int32, device=device)
        m_local = nx * ny * self.nzl
        self.sums = torch.zeros(3, **f64)
        # fixed-point density halo messages: ny*nx ints + 4 trailing ints carrying the sender's scale
        self.ghost_send = torch.zeros(2, ny * nx + 4, **i32)
        self.ghost_recv = torch.zeros(2, ny * nx + 4, **i32)
        self.send = torch.empty(m_local, **f32)            # packed half spectrum of the slab: [dest][plane][y][kx]
        self.pencil = torch.empty(m_local, **f32)          # [nz][ny][kxl] complex
        self.cv = torch.zeros(1, **f64)
        self.inv_send = torch.zeros(2, ny, nx, **f32)
        self.inv_ghost = torch.zeros(2, ny, nx, **f32)
        self._n = 0

    def __del__(self):
        if getattr(self, "h", None):
            lib.metad_mesh_destroy(self.h)
            self.h = None

    P2P_SEGMENTS = ("spread", "halo_push_rho", "barrier_1", "fft_x_fwd", "barrier_2", "fft_y_fwd", "fft_z_fused", "fft_y_inv",
                    "barrier_3", "fft_x_inv", "halo_push_inv", "barrier_4")

    def p2p_timings(self):
        """Milliseconds of the 12 segments of the last peer-memory step (profiling knob 2 on; barrier launches)."""
        out = np.empty(len(self.P2P_SEGMENTS), dtype=np.float32)
        check(lib.metad_mesh_get(self.h, 8, out.ctypes.data_as(C.c_void_p)))
        return dict(zip(self.P2P_SEGMENTS, (float(v) for v in out)))

    def set(self, key, value):
        check(lib.metad_mesh_set(self.h, int(key), int(value)))

    def stage_spread(self, postype, box):
        self._n = postype.shape[0]
        check(lib.metad_mesh_slab_spread(self.h, _ptr(postype), self._n, C.byref(box), _ptr(self.sums), _ptr(self.ghost_send), _stream()))

    def stage_fft_x(self):
        check(lib.metad_mesh_slab_fft_x(self.h, _ptr(self.ghost_recv), _ptr(self.sums), _ptr(self.send), _stream()))

    def stage_fft_yz(self, n_global):
        check(lib.metad_mesh_slab_fft_yz(self.h, _ptr(self.pencil), _ptr(self.sums), int(n_global), _ptr(self.cv), _stream()))

    def stage_fft_x_inv(self):
        check(lib.metad_mesh_slab_fft_x_inv(self.h, _ptr(self.send), _ptr(self.inv_send), _stream()))

    def stage_forces(self, postype, n_global, box, bias, out=None):
        if out is None:
            out = torch.empty_like(postype)
        check(lib.metad_mesh_slab_forces(self.h, _ptr(self.inv_ghost), _ptr(postype), _ptr(out), postype.shape[0], int(n_global),
                                         C.byref(box), _ptr(bias), _stream()))
        return out

    def cells(self):
        out = np.empty((self._n, 3), dtype=np.int32)
        check(lib.metad_mesh_get(self.h, 0, out.ctypes.data_as(C.c_void_p)))
        return out

    def stats(self):
        out = np.empty(6, dtype=np.float64)
        check(lib.metad_mesh_get(self.h, 5, out.ctypes.data_as(C.c_void_p)))
        return dict(rebuilds=int(out[0]), drifted=int(out[1]), outside_slab=int(out[2]), range_warnings=int(out[3]),
                    fx_scale=float(out[4]), calls_since_rebuild=int(out[5]))

    def local_mesh(self, which):
        nx, ny, _ = self.dims
        out = np.empty((self.nzl, ny, nx), dtype=np.float32)
        check(lib.metad_mesh_get(self.h, which, out.ctypes.data_as(C.c_void_p)))
        return out


class MeshSlab:
    """One rank of the sharded mesh CV, driving the stages and the collectives of `comm` (TorchComm)."""

    def __init__(self, comm, nx, ny, nz, mode, rank_factory=None):
        """rank_factory: class with the interface of MeshSlabRank (tests substitute a CPU stand-in)."""
        self.comm = comm
        self.r = (rank_factory or MeshSlabRank)(nx, ny, nz, comm.size, comm.rank, mode)

    def compute_cv(self, postype_local, n_global, box):
        r, c = self.r, self.comm
        r.stage_spread(postype_local, box)
        c.all_reduce_sum(r.sums)
        c.neighbour_exchange(r.ghost_send[0], r.ghost_send[1], r.ghost_recv[0], r.ghost_recv[1])
        r.stage_fft_x()
        c.all_to_all(r.pencil, r.send)
        r.stage_fft_yz(n_global)
        c.all_reduce_sum(r.cv)
        c.all_to_all(r.send, r.pencil)
        r.stage_fft_x_inv()
        # my first plane is rank-1's upper halo, my last plane rank+1's lower halo
        c.neighbour_exchange(r.inv_send[0], r.inv_send[1], r.inv_ghost[0], r.inv_ghost[1])
        return r.cv

    def forces(self, postype_local, n_global, box, bias, out=None):
        return self.r.stage_forces(postype_local, n_global, box, bias, out=out)


def mesh_slab_step_local(ranks, comm, postypes, n_global, box, bias):
    """Bulk-synchronous emulation of all ranks in one process (LocalComm): returns (cv tensors, forces per rank)."""
    for r, pt in zip(ranks, postypes):
        r.stage_spread(pt, box)
    comm.all_reduce_sum([r.sums for r in ranks])
    comm.neighbour_exchange([r.ghost_send[0] for r in ranks], [r.ghost_send[1] for r in ranks],
                            [r.ghost_recv[0] for r in ranks], [r.ghost_recv[1] for r in ranks])
    for r in ranks:
        r.stage_fft_x()
    comm.all_to_all([r.pencil for r in ranks], [r.send for r in ranks])
    for r in ranks:
        r.stage_fft_yz(n_global)
    comm.all_reduce_sum([r.cv for r in ranks])
    comm.all_to_all([r.send for r in ranks], [r.pencil for r in ranks])
    for r in ranks:
        r.stage_fft_x_inv()
    comm.neighbour_exchange([r.inv_send[0] for r in ranks], [r.inv_send[1] for r in ranks],
                            [r.inv_ghost[0] for r in ranks], [r.inv_ghost[1] for r in ranks])
    forces = [r.stage_forces(pt, n_global, box, bias) for r, pt in zip(ranks, postypes)]
    return [r.cv for r in ranks], forces


# ---------------------------------------------------------------------------------------------------- mesh over peer memory
class MeshSlabP2P:
    """One rank of the sharded mesh CV over peer memory (metad_mesh_slab_p2p_*)."""

    def __init__(self, comm, nx, ny, nz, mode):
        self.comm = comm
        self.r = MeshSlabRank(nx, ny, nz, comm.size, comm.rank, mode)
        handle = (C.c_ubyte * 64)()
        nbytes = C.c_ulonglong(0)
        check(lib.metad_mesh_slab_p2p_arena(self.r.h, handle, C.byref(nbytes)))
        mine = torch.tensor(list(handle), dtype=torch.uint8, device="cuda")
        allh = torch.empty(comm.size * 64, dtype=torch.uint8, device="cuda")
        comm.dist.all_gather_into_tensor(allh, mine, group=comm.group)
        buf = (C.c_ubyte * (64 * comm.size))(*allh.cpu().tolist())
        err = None
        try:
            check(lib.metad_mesh_slab_p2p_connect(self.r.h, buf))
        except Exception as e:              # e.g. no peer access between two devices: every rank must learn about it
            err = e
        ok = torch.tensor([0 if err else 1], dtype=torch.int32, device="cuda")
        comm.dist.all_reduce(ok, op=comm.dist.ReduceOp.MIN, group=comm.group)   # also: every rank has mapped its peers
        if ok.item() == 0:
            raise RuntimeError("peer-memory mode is not available on every rank (%s); use the staged MeshSlab (NCCL) driver" % (err or "another rank failed"))
        self.arena_bytes = int(nbytes.value)

    def compute_cv(self, postype_local, n_global, box):
        r = self.r
        r._n = postype_local.shape[0]
        check(lib.metad_mesh_slab_p2p_cv(r.h, _ptr(postype_local), r._n, int(n_global), C.byref(box), _ptr(r.cv), -1, _stream()))
        return r.cv

    def forces(self, postype_local, n_global, box, bias, out=None):
        if out is None:
            out = torch.empty_like(postype_local)
        check(lib.metad_mesh_slab_p2p_forces(self.r.h, _ptr(postype_local), _ptr(out), postype_local.shape[0], int(n_global),
                                             C.byref(box), _ptr(bias), _stream()))
        return out

    def status(self):
        out = np.zeros(2, dtype=np.uint32)
        check(lib.metad_mesh_get(self.r.h, 6, out.ctypes.data_as(C.c_void_p)))
        return dict(barrier_timeout=int(out[0]), outside_slab=int(out[1]))


def mesh_slab_p2p_step_local(ranks, postypes, n_global, box, bias):
    """All ranks in one process on one GPU: stage k for every rank, then stage k+1 (the launch order replaces the flag
    barriers).  ranks: MeshSlabRank objects already connected with connect_local()."""
    for stage in range(5):
        for r, pt in zip(ranks, postypes):
            r._n = pt.shape[0]
            check(lib.metad_mesh_slab_p2p_cv(r.h, _ptr(pt), pt.shape[0], int(n_global), C.byref(box), _ptr(r.cv), stage, _stream()))
    forces = []
    for r, pt in zip(ranks, postypes):
        out = torch.empty_like(pt)
        check(lib.metad_mesh_slab_p2p_forces(r.h, _ptr(pt), _ptr(out), pt.shape[0], int(n_global), C.byref(box), _ptr(bias), _stream()))
        forces.append(out)
    return [r.cv for r in ranks], forces


def connect_local(ranks):
    arr = (C.c_void_p * len(ranks))(*[r.h for r in ranks])
    for r in ranks:
        check(lib.metad_mesh_slab_p2p_connect_local(r.h, arr))
