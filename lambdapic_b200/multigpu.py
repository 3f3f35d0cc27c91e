"""Multi-GPU: one process per GPU, patches partitioned in static contiguous blocks (workloads.block_rank_map),
guard / current / particle exchange between ranks over NCCL point-to-point (NVLink 5 via NVSwitch).

Replaces core/mpi/mpi_manager.py + core/mpi/sync_{fields,particles}_{2d,3d}.c of the reference (MPI_Isend/Irecv per
(patch, boundary[, attribute])).  Here every phase is: ONE pack kernel per peer -> one ncclSend/ncclRecv pair per peer
on persistent device staging buffers -> ONE unpack kernel.  Order of exchanges inside a step is the reference's
(simulation/simulation.py:937-1130): E guards, B guards, [particles pushed], local current reduce then remote current
reduce, migration, B guards, E guards.

Two transports share the exchange plan and the pack / unpack kernels:

* :class:`NcclExchange` -- the production path.  NCCL lives INSIDE the library (csrc/comm.cu: ``lpic_comm_init``,
  ``lpic_halo_start/wait``, ``lpic_migrate_remote_start/wait``): a dedicated comm stream, CUDA events instead of host
  synchronisation, persistent staging buffers, counts exchanged device to device.  ``*_start`` returns at once and the
  intra-rank guard copy / current reduce / particle fill run between start and wait (simulation/simulation.py:948-952).
* :class:`RankProgram` -- the same step as a generator whose yields are the exchange phases, driven either in-process
  (all emulated ranks on ONE GPU, used by the single-GPU tests: no kernel ever waits on another rank) or by
  ``torch.distributed`` send/recv (``LPIC_HOST_EXCHANGE=1``).

The per-rank step is written once as a generator (:class:`RankProgram`) that yields its outgoing buffers and is
resumed with the incoming ones, so the same code runs under the NCCL driver (one rank per process) and under the
in-process driver used by the single-GPU tests (all ranks' programs advanced in lockstep, buffers handed over by
reference -- no kernel ever waits on another rank).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import check
from .engine import B_MASK, E_MASK, J_MASK
from .workloads import DIR2, DIR3


def _opp(dim, b):
    dirs = DIR3 if dim == 3 else DIR2
    s = dirs[b]
    return dirs.index((-s[0], -s[1], -s[2]))


def build_plan(grid):
    """Per-peer send/recv entry lists in the canonical order (ascending global index of the RECEIVING patch, then
    boundary at the receiver).  Returns (peers, send, recv) with send[r] / recv[r] = list of (local patch, boundary)."""
    nb = grid.neighbor_rank.shape[1]
    send, recv = {}, {}
    for p in range(grid.npatch):
        for b in range(nb):
            r = int(grid.neighbor_rank[p, b])
            if r < 0:
                continue
            ob = _opp(grid.dim, b)
            # we send to patch neighbor_index[p,b] on r, which sees us through its boundary ob
            send.setdefault(r, []).append(((int(grid.neighbor_index[p, b]), ob), (p, b)))
            # we receive from r for our own patch p through boundary b
            recv.setdefault(r, []).append(((int(grid.index[p]), b), (p, b)))
    peers = sorted(set(send) | set(recv))
    s = {r: [e for _, e in sorted(send.get(r, []))] for r in peers}
    v = {r: [e for _, e in sorted(recv.get(r, []))] for r in peers}
    return peers, s, v


def plan_view(patches, rank, nranks):
    """What build_plan needs, read from the host patch objects (their neighbour tables are the ones MovingWindow updates)."""
    import types
    return types.SimpleNamespace(dim=patches.dimension, npatch=patches.npatches, rank=rank, nranks=nranks,
                                 index=np.array([p.index for p in patches], dtype=np.int64),
                                 neighbor_index=np.stack([p.neighbor_index for p in patches]).astype(np.int64),
                                 neighbor_rank=np.stack([p.neighbor_rank for p in patches]).astype(np.int64))


def register_plan(eng, grid):
    """Build the exchange plan and hand it to the library (lpic_halo_plan); returns (peers, send, recv, nsend, nrecv)."""
    peers, send, recv = build_plan(grid)
    i64 = lambda a: np.ascontiguousarray(a, dtype=np.int64)  # noqa: E731
    nsend = i64([len(send[r]) for r in peers])
    nrecv = i64([len(recv[r]) for r in peers])
    sp = i64([e[0] for r in peers for e in send[r]])
    sb = i64([e[1] for r in peers for e in send[r]])
    rp = i64([e[0] for r in peers for e in recv[r]])
    rb = i64([e[1] for r in peers for e in recv[r]])
    P = lambda a: C.c_void_p(a.ctypes.data)  # noqa: E731
    check(eng.L.lpic_halo_plan(eng.ctx, len(peers), P(nsend), P(sp), P(sb), P(nrecv), P(rp), P(rb)))
    return peers, send, recv, nsend, nrecv


class NcclExchange:
    """Inter-rank exchange through the library's own NCCL transport (csrc/comm.cu).  `bcast(obj, root)` carries the NCCL
    unique id from rank 0 to the other ranks (torch.distributed / sim.mpi.comm); nothing else goes through the host."""

    def __init__(self, eng, grid, bcast):
        self.eng, self.grid, self.L = eng, grid, eng.L
        self.peers, self.send_entries, self.recv_entries, self.nsend, self.nrecv = register_plan(eng, grid)
        buf = C.create_string_buffer(128)
        if grid.rank == 0:
            check(self.L.lpic_comm_unique_id(buf))
        ident = bcast(bytes(buf.raw), 0)
        peer_rank = np.ascontiguousarray(self.peers if self.peers else [0], dtype=np.int64)
        check(self.L.lpic_comm_init(eng.ctx, C.create_string_buffer(ident, 128), int(grid.rank), int(grid.nranks),
                                    C.c_void_p(peer_rank.ctypes.data)))
        self.migrated = {}  # species -> per-patch extension of the migration this object has already completed

    @property
    def bytes_sent(self):
        return int(self.L.lpic_comm_bytes_sent(self.eng.ctx))

    def replan(self, grid):
        """The neighbour tables changed (MovingWindow shift): new exchange plan, staging re-sized; the NCCL communicator is
        kept (lpic_halo_plan + lpic_comm_update).  `grid`: anything with dim / npatch / index / neighbor_index /
        neighbor_rank, e.g. :func:`plan_view` of the patches."""
        self.grid = grid
        self.peers, self.send_entries, self.recv_entries, self.nsend, self.nrecv = register_plan(self.eng, grid)
        peer_rank = np.ascontiguousarray(self.peers if self.peers else [0], dtype=np.int64)
        check(self.L.lpic_comm_update(self.eng.ctx, C.c_void_p(peer_rank.ctypes.data)))

    # ---- fields ------------------------------------------------------------------------------------------------------
    def halo_start(self, mask, reduce):
        check(self.L.lpic_halo_start(self.eng.ctx, int(mask), int(reduce)))

    def halo_wait(self):
        check(self.L.lpic_halo_wait(self.eng.ctx))

    def sync_guard_fields(self, mask):
        self.halo_start(mask, 0)
        self.eng.sync_guard_fields(mask)   # intra-rank neighbours while the messages are on the wire
        self.halo_wait()

    def sync_currents(self):
        self.halo_start(J_MASK, 1)         # packs (and zeroes) the guards facing other ranks ...
        self.eng.sync_currents()           # ... the intra-rank reduce touches the others
        self.halo_wait()                   # remote contributions are added after the local ones, as the reference does

    # ---- particles ---------------------------------------------------------------------------------------------------
    def sync_particles(self, ispec):
        """Whole migration of one species: other ranks and intra-rank (core/mpi/sync_particles_3d.c + core/patch/patch.py:705-764)."""
        eng = self.eng
        ext = np.zeros(eng.npatch, dtype=np.int64)
        info = np.zeros(4, dtype=np.int64)
        r = self.L.lpic_migrate_remote_start(eng.ctx, ispec, 0, C.c_void_p(ext.ctypes.data), C.c_void_p(info.ctypes.data))
        moved = False
        if r == 1:  # some patch has to grow first
            moved = eng.extend(ispec, ext)
            r = self.L.lpic_migrate_remote_start(eng.ctx, ispec, 1, None, C.c_void_p(info.ctypes.data))
        check(r)
        check(self.L.lpic_migrate_remote_wait(eng.ctx, ispec))
        return dict(sent=int(info[0]), received=int(info[1]), to_extend=ext, extended=bool(ext.any()), moved=moved)

    # ---- one full step (same operator order as DeviceEngine.step) ------------------------------------------------------
    def step(self, dt, q, m, reverse_x, write_part=False, event_slot=None, marks=None):
        eng = self.eng
        mark = marks if marks is not None else (lambda name: None)
        eng.update_efield(0.5 * dt); mark("update E")
        self.sync_guard_fields(E_MASK); mark("sync E (local + NCCL)")
        eng.update_bfield(0.5 * dt); mark("update B")
        self.sync_guard_fields(B_MASK); mark("sync B (local + NCCL)")
        nbuf = [eng.sort(s, reverse_x[s]) for s in range(eng.nspec)]; mark("sort")
        eng.reset_currents()
        for s in range(eng.nspec):
            if event_slot is not None:
                eng.record_event(event_slot + 2 * s)
            eng.push_deposit(s, dt, q[s], m[s], write_part)
            if event_slot is not None:
                eng.record_event(event_slot + 2 * s + 1)
        mark("push + deposit")
        self.halo_start(J_MASK, 1)
        eng.sync_currents(); mark("current reduce (local, NCCL in flight)")
        mig = [self.sync_particles(s) for s in range(eng.nspec)]; mark("migration (local + NCCL)")
        self.halo_wait(); mark("current reduce (remote part)")
        eng.update_bfield(0.5 * dt); mark("update B")
        self.sync_guard_fields(B_MASK); mark("sync B (local + NCCL)")
        eng.update_efield(0.5 * dt); mark("update E")
        self.sync_guard_fields(E_MASK); mark("sync E (local + NCCL)")
        return nbuf, mig


class RankProgram:
    """The inter-rank part of one rank's PIC step.  `alloc(nwords)` returns a device fp64 buffer object exposing
    `.data_ptr()` (a torch CUDA tensor)."""

    def __init__(self, eng, grid, alloc):
        self.eng, self.grid, self.alloc = eng, grid, alloc
        self.L = eng.L
        self.peers, send, recv, nsend, nrecv = register_plan(eng, grid)
        self.send_entries, self.recv_entries = send, recv
        self.nsend, self.nrecv = nsend, nrecv
        self.send_words = [int(self.L.lpic_halo_words(eng.ctx, i, 0)) for i in range(len(self.peers))]
        self.recv_words = [int(self.L.lpic_halo_words(eng.ctx, i, 1)) for i in range(len(self.peers))]
        self._fbuf = {}
        self.bytes_sent = 0

    # ---- field halos -------------------------------------------------------------------------------------------
    def _field_buffers(self, nattr):
        if nattr not in self._fbuf:
            self._fbuf[nattr] = ([self.alloc(max(w * nattr, 1)) for w in self.send_words],
                                 [self.alloc(max(w * nattr, 1)) for w in self.recv_words])
        return self._fbuf[nattr]

    def exchange_fields(self, mask, reduce):
        """generator: pack -> yield sends -> unpack what came back"""
        nattr = bin(mask).count("1")
        sbufs, rbufs = self._field_buffers(nattr)
        for i in range(len(self.peers)):
            check(self.L.lpic_halo_pack(self.eng.ctx, i, mask, int(reduce), C.c_void_p(sbufs[i].data_ptr())))
        self.eng.sync()
        sends = {r: (sbufs[i], self.send_words[i] * nattr) for i, r in enumerate(self.peers)}
        recvs = {r: (rbufs[i], self.recv_words[i] * nattr) for i, r in enumerate(self.peers)}
        self.bytes_sent += 8 * sum(n for _, n in sends.values())
        got = yield ("f64", sends, recvs)
        ptrs = (C.c_void_p * max(len(self.peers), 1))(*[got[r].data_ptr() for r in self.peers])
        check(self.L.lpic_halo_unpack(self.eng.ctx, mask, int(reduce), ptrs))

    def sync_guard_fields(self, mask):
        self.eng.sync_guard_fields(mask)          # intra-rank neighbours
        yield from self.exchange_fields(mask, 0)  # inter-rank neighbours

    def sync_currents(self):
        self.eng.sync_currents()                  # local reduce first (simulation.py:1155-1176)
        yield from self.exchange_fields(J_MASK, 1)

    # ---- particles ---------------------------------------------------------------------------------------------
    def migrate_remote(self, ispec):
        eng, L = self.eng, self.L
        npeer = len(self.peers)
        ns_tot, nr_tot = int(self.nsend.sum()), int(self.nrecv.sum())
        send_counts = np.zeros(max(ns_tot, 1), dtype=np.int64)
        ndead = np.zeros(eng.npatch, dtype=np.int64)
        check(L.lpic_remote_migrate_prepare(eng.ctx, ispec, C.c_void_p(send_counts.ctypes.data), C.c_void_p(ndead.ctypes.data)))
        # 1. counts (core/mpi/sync_particles_3d.c:893-895: one MPI_LONG per (patch, boundary); here one vector per peer)
        cs, cr = {}, {}
        s0 = r0 = 0
        keep = []
        for i, r in enumerate(self.peers):
            sb, rb = self.alloc(max(int(self.nsend[i]), 1)), self.alloc(max(int(self.nrecv[i]), 1))
            _copy_in(sb, send_counts[s0:s0 + int(self.nsend[i])].astype(np.float64))
            cs[r], cr[r] = (sb, int(self.nsend[i])), (rb, int(self.nrecv[i]))
            keep.append((sb, rb))
            s0 += int(self.nsend[i])
        got = yield ("f64", cs, cr)
        recv_counts = np.zeros(max(nr_tot, 1), dtype=np.int64)
        for i, r in enumerate(self.peers):
            n = int(self.nrecv[i])
            recv_counts[r0:r0 + n] = np.rint(_copy_out(got[r], n)).astype(np.int64)
            r0 += n
        # 2. grow the arrays if the dead slots cannot take the newcomers (same rule as the intra-rank pass)
        incoming = np.zeros(eng.npatch, dtype=np.int64)
        k = 0
        for r in self.peers:
            for (p, b) in self.recv_entries[r]:
                incoming[p] += recv_counts[k]
                k += 1
        npart = eng.species[ispec].npart
        ext = np.where(incoming - ndead > 0, incoming - ndead + (npart * 0.25).astype(np.int64), 0)
        if ext.any():
            eng.extend(ispec, ext)
            check(L.lpic_remote_migrate_relist(eng.ctx, ispec))
        # 3. payload
        nw = int(L.lpic_particle_record_words(eng.ctx, ispec))
        ps, pr = {}, {}
        s0 = r0 = 0
        for i, r in enumerate(self.peers):
            nsnd = int(send_counts[s0:s0 + int(self.nsend[i])].sum())
            nrcv = int(recv_counts[r0:r0 + int(self.nrecv[i])].sum())
            sb, rb = self.alloc(max(nsnd * nw, 1)), self.alloc(max(nrcv * nw, 1))
            n_out = C.c_int64(0)
            check(L.lpic_remote_migrate_pack(eng.ctx, ispec, i, C.c_void_p(sb.data_ptr()), C.byref(n_out)))
            assert n_out.value == nsnd
            ps[r], pr[r] = (sb, nsnd * nw), (rb, nrcv * nw)
            s0 += int(self.nsend[i])
            r0 += int(self.nrecv[i])
        eng.sync()
        self.bytes_sent += 8 * sum(n for _, n in ps.values())
        got = yield ("f64", ps, pr)
        ptrs = (C.c_void_p * max(npeer, 1))(*[got[r].data_ptr() for r in self.peers])
        check(L.lpic_remote_migrate_unpack(eng.ctx, ispec, C.c_void_p(recv_counts.ctypes.data), ptrs))
        return dict(sent=int(send_counts[:ns_tot].sum()), received=int(recv_counts[:nr_tot].sum()), extended=bool(ext.any()))

    # ---- one full step (same operator order as DeviceEngine.step) -------------------------------------------------
    def step(self, dt, q, m, reverse_x, write_part=False, event_slot=None):
        eng = self.eng
        eng.update_efield(0.5 * dt)
        yield from self.sync_guard_fields(E_MASK)
        eng.update_bfield(0.5 * dt)
        yield from self.sync_guard_fields(B_MASK)
        nbuf = [eng.sort(s, reverse_x[s]) for s in range(eng.nspec)]
        eng.reset_currents()
        for s in range(eng.nspec):
            if event_slot is not None:
                eng.record_event(event_slot + 2 * s)
            eng.push_deposit(s, dt, q[s], m[s], write_part)
            if event_slot is not None:
                eng.record_event(event_slot + 2 * s + 1)
        yield from self.sync_currents()
        mig = []
        for s in range(eng.nspec):
            rem = yield from self.migrate_remote(s)   # remote before local (simulation.py:1067-1077)
            mig.append(rem)
        for s in range(eng.nspec):
            mig[s]["local"] = eng.sync_particles(s)
        eng.update_bfield(0.5 * dt)
        yield from self.sync_guard_fields(B_MASK)
        eng.update_efield(0.5 * dt)
        yield from self.sync_guard_fields(E_MASK)
        return nbuf, mig


def _copy_in(buf, values):
    import torch
    buf[:len(values)].copy_(torch.from_numpy(np.ascontiguousarray(values)))


def _copy_out(buf, n):
    return buf[:n].cpu().numpy()


def torch_alloc(device):
    import torch

    def alloc(nwords):
        return torch.empty(int(nwords), dtype=torch.float64, device=device)
    return alloc


# ---------------------------------------------------------------------------------------------------------------------
# drivers
# ---------------------------------------------------------------------------------------------------------------------
def drive_nccl(gen, rank):
    """Run one rank's program; every yield becomes one batch of ncclSend/ncclRecv (torch.distributed P2P)."""
    import torch
    import torch.distributed as dist
    try:
        req = next(gen)
        while True:
            _, sends, recvs = req
            ops = []
            for r in sorted(set(sends) | set(recvs)):
                # fixed order on both sides: lower rank sends first
                snd = dist.P2POp(dist.isend, sends[r][0][:max(sends[r][1], 1)], r) if r in sends and sends[r][1] > 0 else None
                rcv = dist.P2POp(dist.irecv, recvs[r][0][:max(recvs[r][1], 1)], r) if r in recvs and recvs[r][1] > 0 else None
                ops += [o for o in ((snd, rcv) if rank < r else (rcv, snd)) if o is not None]
            if ops:
                for w in dist.batch_isend_irecv(ops):
                    w.wait()
                torch.cuda.synchronize()
            req = gen.send({r: recvs[r][0] for r in recvs})
    except StopIteration as e:
        return e.value


def drive_in_process(gens):
    """Advance all ranks' programs in lockstep on ONE device: rank a's send buffer for b is copied into b's receive
    buffer between phases.  Used by the single-GPU tests; semantics are identical to the NCCL driver."""
    n = len(gens)
    results = [None] * n
    reqs = [None] * n
    alive = [True] * n
    for a in range(n):
        try:
            reqs[a] = next(gens[a])
        except StopIteration as e:
            results[a], alive[a] = e.value, False
    while any(alive):
        assert all(alive), "rank programs must exchange the same number of times"
        for a in range(n):
            _, sends, _ = reqs[a]
            for b, (buf, cnt) in sends.items():
                rbuf, rcnt = reqs[b][2][a]
                assert rcnt == cnt, f"rank {a}->{b}: sender has {cnt} words, receiver expects {rcnt}"
                if cnt:
                    rbuf[:cnt].copy_(buf[:cnt])
        import torch
        torch.cuda.synchronize()
        for a in range(n):
            try:
                reqs[a] = gens[a].send({r: reqs[a][2][r][0] for r in reqs[a][2]})
            except StopIteration as e:
                results[a], alive[a] = e.value, False
    return results


def _torch_bcast(obj, root):
    import torch.distributed as dist
    box = [obj]
    dist.broadcast_object_list(box, src=root)
    return box[0]


class HaloExchanger:
    """bench.py glue: owns this process's exchange object.  Default: the library's NCCL transport (:class:`NcclExchange`);
    ``LPIC_HOST_EXCHANGE=1``: the host-driven torch.distributed path (:class:`RankProgram` + :func:`drive_nccl`)."""

    def __init__(self, eng, grid):
        import os
        import torch
        self.rank = grid.rank
        self.host_driven = bool(os.environ.get("LPIC_HOST_EXCHANGE"))
        if self.host_driven:
            self.prog = RankProgram(eng, grid, torch_alloc(torch.device("cuda", torch.cuda.current_device())))
        else:
            self.prog = NcclExchange(eng, grid, _torch_bcast)

    def step(self, dt, q, m, reverse_x, event_slot=None, marks=None):
        if self.host_driven:
            return drive_nccl(self.prog.step(dt, q, m, reverse_x, event_slot=event_slot), self.rank)
        return self.prog.step(dt, q, m, reverse_x, event_slot=event_slot, marks=marks)


class MultiRankMPI:
    """``sim.mpi`` for several ranks (core/mpi/mpi_manager.py:9-298): same method names.  Field exchanges are truly
    asynchronous: ``*_start`` packs and posts the NCCL sends / receives on the library's comm stream and returns, the caller
    runs the intra-rank copy / reduce, ``*_wait`` makes the compute stream wait and unpacks.  ``sync_particles_start``
    performs the species' whole migration (other ranks AND intra-rank: the two share one classification pass and one list
    of dead slots), so the ``Patches.sync_particles`` that follows finds that species already done."""

    def __init__(self, sim, comm):
        self.sim, self.comm = sim, comm
        self.rank, self.size = comm.Get_rank(), comm.Get_size()
        self._xch = None

    @property
    def xch(self):
        eng = self.sim.bridge.engine
        if self._xch is None or self._xch.eng is not eng:
            self._xch = NcclExchange(eng, self.sim.grid, lambda obj, root: self.comm.bcast(obj, root=root))
            eng.comm_exchange = self._xch
        return self._xch

    @staticmethod
    def _mask(attrs):
        from ._lib import FIELD_ATTRS
        mask = 0
        for a in attrs:
            mask |= 1 << FIELD_ATTRS.index(a)
        return mask

    def sync_guard_fields_start(self, attrs):
        with self.sim.bridge.coherent():
            self.xch.halo_start(self._mask(attrs), 0)
        return "guards"

    def sync_guard_fields_wait(self, handle):
        with self.sim.bridge.coherent():
            self.xch.halo_wait()

    def sync_guard_fields(self, attrs):
        self.sync_guard_fields_wait(self.sync_guard_fields_start(attrs))

    def sync_currents_start(self):
        with self.sim.bridge.coherent():
            self.xch.halo_start(J_MASK, 1)
        return "currents"

    def sync_currents_wait(self, handle):
        with self.sim.bridge.coherent():
            self.xch.halo_wait()

    def sync_currents(self):
        self.sync_currents_wait(self.sync_currents_start())

    def sync_particles_start(self, ispec):
        with self.sim.bridge.coherent():
            rec = self.xch.sync_particles(ispec)
            self.xch.migrated[ispec] = rec["to_extend"]
        return "done"

    def sync_particles_wait(self, handle):
        pass

    def sync_particles(self, ispec):
        self.sync_particles_start(ispec)

    def replan(self):
        """Rebuild the exchange plan from the patches' current neighbour tables (after a MovingWindow shift)."""
        self.xch.replan(plan_view(self.sim.patches, self.rank, self.size))
