"""MovingWindow callback (reference: callback/utils.py:471-840): follow the laser along x by recycling one column of
patches at a time.  Same constructor, stage (`start`) and bookkeeping (total_shift, patch_this_shift, num_shifts) as the
reference; the shift itself works on the host mirrors of the DeviceBridge:

  non-shift steps  nothing crosses PCIe (`needs_host = False`: the bridge does not refresh the mirrors for this callback);
  shift steps      rotate ipatch_x / x0 / axes of the recycled column on the host objects, rebuild the neighbour tables,
                   re-load the recycled patches' particles from the density profile with the reference's generator
                   stream, recompute the sort origins.  Then, device-resident (DeviceBridge.recycle): the recycled
                   patches' fields and psi arrays are cleared ON the device, only their new particles are uploaded
                   (lpic_species_set_npart + lpic_upload_particles_patch) and the geometry is re-registered -- no mirror
                   is downloaded, no other patch is touched.  The first activation also drops the PMLX faces, which changes
                   the set of CPML instances: that one step takes the download -> re-register -> upload route.

Several ranks: patches keep their owner; the rotated column positions are allgathered (callback/utils.py:648-716), the
neighbour tables rebuilt and the inter-rank exchange plan replaced (`sim.mpi.replan()`: lpic_halo_plan + lpic_comm_update,
the NCCL communicator survives).
"""
from __future__ import annotations

from typing import Callable, Optional, Union

from .patch import load_patch_particles, loader_spacing
from .workloads import C_LIGHT


class MovingWindow:
    DEFAULT_STAGE = "start"
    needs_host = False  # the callback manages the mirrors itself, and only on shift steps

    def __init__(self, velocity: Union[float, Callable[[float], float]], start_time: Optional[float] = None,
                 inject_particles: bool = True, stop_inject_time: Optional[float] = None):
        self.stage = self.DEFAULT_STAGE
        self.velocity = velocity
        self.start_time = start_time
        self.inject_particles = inject_particles
        self.stop_inject_time = stop_inject_time
        self.total_shift = None
        self.patch_this_shift = None
        self.num_shifts: int = 0

    def __call__(self, sim):
        patch_Lx = sim.nx_per_patch * sim.dx
        if self.start_time is None:
            self.start_time = sim.Lx / C_LIGHT
        if self.total_shift is None:
            self.total_shift = patch_Lx
        if self.patch_this_shift is None:
            self.patch_this_shift = patch_Lx
        if sim.time < self.start_time:
            return
        drop_pmlx = self.num_shifts == 0 and any(m.axis == 0 for p in sim.patches for m in p.pml_boundary)

        current_velocity = self.velocity(sim.time) if callable(self.velocity) else self.velocity
        shift_amount = current_velocity * sim.dt
        self.total_shift += shift_amount
        self.patch_this_shift += shift_amount
        self.num_shifts += 1
        direction = 0
        if self.patch_this_shift >= patch_Lx:
            direction = 1
            self.patch_this_shift -= patch_Lx
        elif self.patch_this_shift <= -patch_Lx:
            direction = -1
            self.patch_this_shift += patch_Lx
        if not (drop_pmlx or direction):
            return

        br = sim.bridge
        was_resident = br.resident
        fast = was_resident and direction and not drop_pmlx  # device-side recycle: nothing but the new particles moves
        if was_resident and not fast:
            br.download()
            br.resident = False
        if drop_pmlx:  # the x faces stop absorbing once the window moves (callback/utils.py:545-552)
            for p in sim.patches:
                p.pml_boundary = [m for m in p.pml_boundary if m.axis != 0]
        new_patches = []
        if direction:
            new_patches = self._shift(sim, direction)
            sim.patches.geometry_version = getattr(sim.patches, "geometry_version", 0) + 1
            self._update_patch_info(sim)
            if fast:  # the host objects are stale while the device is authoritative: the id counters live in the engine
                where = {id(p): ip for ip, p in enumerate(sim.patches)}
                for p in new_patches:
                    for ispec, pt in enumerate(p.particles):
                        pt._npart_created = int(br.engine.npart_created[ispec][where[id(p)]])
            self._fill_particles(sim, new_patches)
            for p in new_patches:  # recycled patches start from vacuum fields
                for attr in p.fields.attrs:
                    getattr(p.fields, attr).fill(0.0)
                for m in p.pml_boundary:
                    for k, v in vars(m).items():
                        if k.startswith("psi"):
                            v.fill(0.0)
            for sorter in sim.sorter:
                sorter.generate_field_lists()
                sorter.generate_particle_lists()
        if fast:
            where = {id(p): ip for ip, p in enumerate(sim.patches)}
            self.last_shift_bytes = br.recycle([where[id(p)] for p in new_patches])
            return
        br.refresh_geometry(pml_changed=drop_pmlx)
        if was_resident:
            br.upload()
            br.resident = True

    @staticmethod
    def _shift(sim, direction):
        """callback/utils.py:591-646: the trailing column jumps ahead by Lx; every other column moves one index back."""
        last = sim.npatch_x - 1
        new_patches = []
        for p in sim.patches:
            if direction > 0 and p.ipatch_x == 0 or direction < 0 and p.ipatch_x == last:
                p.ipatch_x = last if direction > 0 else 0
                # in-place additions, as the reference does them: the axes are NOT recomputed from the new origin
                p.x0 += direction * sim.Lx
                p.xaxis += direction * sim.Lx
                p.fields.x0 += direction * sim.Lx
                p.fields.xaxis += direction * sim.Lx
                new_patches.append(p)
            else:
                p.ipatch_x -= direction
        return new_patches

    @staticmethod
    def _update_patch_info(sim):
        """callback/utils.py:648-716: every rank learns where every patch sits now (allgather of (position, index, rank)),
        then the neighbour index / local position / rank tables are rebuilt."""
        ps = sim.patches
        three = sim.dimension == 3
        local = [((p.ipatch_x, p.ipatch_y) + ((p.ipatch_z,) if three else ()), p.index, p.rank) for p in ps]
        index_map, rank_map = {}, {}
        for info in sim.mpi.comm.allgather(local):
            for pos, idx, r in info:
                index_map[pos] = idx
                rank_map[idx] = r
        if three:
            ps.init_rect_neighbor_index_3d(sim.npatch_x, sim.npatch_y, sim.npatch_z, boundary_conditions=sim.boundary_conditions,
                                           patch_index_map=index_map)
            ps.init_neighbor_ipatch_3d()
            ps.init_neighbor_rank_3d(rank_map)
        else:
            ps.init_rect_neighbor_index_2d(sim.npatch_x, sim.npatch_y, boundary_conditions=sim.boundary_conditions,
                                           patch_index_map=index_map)
            ps.init_neighbor_ipatch_2d()
            ps.init_neighbor_rank_2d(rank_map)
        if sim.mpi.size > 1:
            sim.mpi.replan()

    def _fill_particles(self, sim, new_patches):
        """callback/utils.py:718-840: every species with a density profile is re-initialised in the recycled patches
        (fresh ids continue the patch's counter) and loaded from one new generator per patch, spawned per species."""
        if not new_patches or not self.inject_particles:
            return
        if self.stop_inject_time is not None and sim.time >= self.stop_inject_time:
            return
        dim = sim.dimension
        from .patch import _node_profiles
        for ispec, s in enumerate(sim.species):
            if s.density is None:
                continue
            gens = sim.rand_gen.spawn(len(new_patches))
            d = loader_spacing(new_patches[0], dim)
            for k, p in enumerate(new_patches):
                _, ppc = _node_profiles(s, p, dim)
                part = p.particles[ispec]
                part.initialize(int(ppc.sum()))
                load_patch_particles(s, p, part, gens[k], dim, d)
        sim.update_lists()
