/*
 * lpic_b200.h -- C-ABI of the B200-native lambdaPIC inner loop (liblpic_b200.so).
 *
 * Plain pointers and sizes only: no CUDA, torch or Python types cross this boundary.  Every entry point
 * names the reference interface it replaces (paths relative to /root/reference/src/lambdapic/).  All
 * functions return 0 on success or a negative status; lpic_last_error() then holds a message (no
 * exceptions cross the boundary).  Kernels run on the context's own CUDA stream; calls that return
 * values to the host (download, sort's nbuf, migrate_count) synchronise that stream, the others do not.
 *
 * Data model (reference: core/fields.py:71-170, core/particles.py:8-217):
 *   fields    10 fp64 grids per patch  ex ey ez bx by bz jx jy jz rho, C-contiguous (NX,NY[,NZ]) with
 *             N = n + 2*n_guard and the reference's WRAPPED guard layout (logical index -k lives at N-k).
 *             Device arena = [attr][patch][NX*NY*NZ]; the host mirror handed to upload/download has the
 *             same shape, so one copy moves one attribute of every patch.
 *   particles fp64 x y z w ux uy uz inv_gamma ex_part..bz_part _id + uint8 is_dead per (species, patch).
 *             HOST side of every call: one array per attribute with the arena layout below (the reference's
 *             numpy arrays, core/particles.py:60-89).  Patch p owns the slots [off[p], off[p] + npart[p]) of a
 *             segment of physical size pcap[p] >= npart[p]; `npart` is the reference's capacity (alive + dead
 *             slots).  The six *_part arrays are optional.  DEVICE side (internal): x y z w ux uy uz inv_gamma
 *             are interleaved into one 64-byte record per slot, the others are plain arrays; the upload /
 *             download entries convert.
 */
#ifndef LPIC_B200_H
#define LPIC_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct lpic_ctx lpic_ctx;

/* attribute ids: order of Fields.attrs (core/fields.py:71-75) and ParticlesBase.attrs (core/particles.py:63-67) */
enum { LPIC_EX = 0, LPIC_EY, LPIC_EZ, LPIC_BX, LPIC_BY, LPIC_BZ, LPIC_JX, LPIC_JY, LPIC_JZ, LPIC_RHO, LPIC_NFIELD };
enum { LPIC_P_X = 0, LPIC_P_Y, LPIC_P_Z, LPIC_P_W, LPIC_P_UX, LPIC_P_UY, LPIC_P_UZ, LPIC_P_INV_GAMMA,
       LPIC_P_EX_PART, LPIC_P_EY_PART, LPIC_P_EZ_PART, LPIC_P_BX_PART, LPIC_P_BY_PART, LPIC_P_BZ_PART,
       LPIC_P_ID, LPIC_NPATTR, LPIC_P_IS_DEAD = LPIC_NPATTR };
/* sorter arrays (core/sort/particle_sort.py:20-160) */
enum { LPIC_SORT_BUCKET_COUNT = 0, LPIC_SORT_BOUND_MIN, LPIC_SORT_BOUND_MAX, LPIC_SORT_PARTICLE_INDEX };
/* flags of lpic_push_deposit */
enum { LPIC_PUSH_WRITE_PART = 1, LPIC_PUSH_SLOT_ORDER = 2 /* process slots in memory order (round-1 v1 kernel) */ };

const char *lpic_last_error(void);
int lpic_device_count(void);
/* version / build info: "lpic_b200 <ver> sm_100a" */
const char *lpic_version(void);

/* pinned host memory for the mirrors (falls back to malloc when no CUDA device is present) */
void *lpic_host_alloc(int64_t bytes);
void lpic_host_free(void *p);

/* ---- context ------------------------------------------------------------------------------------
 * Replaces the per-call pointer-table construction of every reference extension
 * (core/utils/cutils.h:32-100, core/pusher/unified/unified_pusher_3d.c:234-276). */
lpic_ctx *lpic_create(int dim, int64_t npatch, int64_t nx, int64_t ny, int64_t nz, int64_t n_guard,
                      double dx, double dy, double dz, int nspec, int device);
void lpic_destroy(lpic_ctx *ctx);
/* x0,y0,z0: patch origins (core/patch/patch.py:272-273); neighbor_ipatch: (npatch, 8|26) int64 in the
 * reference's Boundary2D/3D enum order (core/patch/patch.py:24-69), <0 = no local neighbour;
 * box: (npatch, 6) particle boxes xmin,xmax,ymin,ymax,zmin,zmax ALREADY widened by half a cell
 * (core/patch/sync_particles_3d.c:402-411); glob: global particle box (simulation/simulation.py:425-430);
 * rank: MPI-style rank used in new particle ids (core/particles.py:91-116); patch_index: Patch.index. */
int lpic_set_patch_geometry(lpic_ctx *ctx, const double *x0, const double *y0, const double *z0,
                            const int64_t *neighbor_ipatch, const double *box, const double *glob,
                            int64_t rank, const int64_t *patch_index);
int lpic_sync(lpic_ctx *ctx); /* cudaStreamSynchronize on the context stream */

/* ---- field mirrors ------------------------------------------------------------------------------ */
int64_t lpic_field_cells(const lpic_ctx *ctx);                                    /* NX*NY*NZ per patch */
/* host = base of a [LPIC_NFIELD][npatch][cells] fp64 arena; only attributes in attr_mask are copied */
int lpic_upload_fields(lpic_ctx *ctx, uint32_t attr_mask, const double *host);
int lpic_download_fields(lpic_ctx *ctx, uint32_t attr_mask, double *host);
/* per-patch pointer variant for callers that keep the reference's separately allocated numpy arrays */
int lpic_upload_field_ptrs(lpic_ctx *ctx, int attr, const double *const *host_ptrs);
int lpic_download_field_ptrs(lpic_ctx *ctx, int attr, double *const *host_ptrs);

/* ---- particle mirrors --------------------------------------------------------------------------- */
/* (re)allocate species `ispec` with logical capacities npart[npatch]; physical segment size is
 * max(npart*slack, npart+min_extra) rounded up to 32 slots.  with_part_fields!=0 also keeps ex_part..bz_part. */
int lpic_species_alloc(lpic_ctx *ctx, int ispec, const int64_t *npart, double slack, int64_t min_extra,
                       int with_part_fields);
int lpic_species_layout(const lpic_ctx *ctx, int ispec, int64_t *off, int64_t *pcap, int64_t *npart, int64_t *total);
/* host = base of an arena with the device's layout (total slots; fp64, or uint8 for LPIC_P_IS_DEAD) */
int lpic_upload_particles(lpic_ctx *ctx, int ispec, int attr, const void *host);
int lpic_download_particles(lpic_ctx *ctx, int ispec, int attr, void *host);
/* The eight per-step attributes (LPIC_P_X .. LPIC_P_INV_GAMMA, bit a of mask = attribute a) in ONE pass: host[a] = base of
 * the arena-layout array of attribute a (entries of attributes outside mask are not read).  On the device these eight form
 * 64-byte records; the host side keeps the reference's one-numpy-array-per-attribute layout
 * (core/particles.py:60-89), so the move is chunked through two staging buffers: the (un)packing kernel of one chunk runs
 * under the PCIe copies of the other.  This is what Simulation.run's entry / exit copies use. */
int lpic_upload_particle_records(lpic_ctx *ctx, int ispec, uint32_t mask, const double *const *host);
int lpic_download_particle_records(lpic_ctx *ctx, int ispec, uint32_t mask, double *const *host);
int lpic_upload_particle_ptrs(lpic_ctx *ctx, int ispec, int attr, const void *const *host_ptrs);
int lpic_download_particle_ptrs(lpic_ctx *ctx, int ispec, int attr, void *const *host_ptrs);
/* ParticlesBase.extend (core/particles.py:141-168) for every patch at once: appends ext[p] dead slots
 * (NaN attributes, w = 0, fresh ids starting at id_first[p]).  *relayout is set when a segment outgrew its
 * physical size and the arena was re-laid-out (host arena offsets must be re-read). */
int lpic_species_extend(lpic_ctx *ctx, int ispec, const int64_t *ext, const uint64_t *id_first, int *relayout);
/* Set the slot counts of every patch inside the EXISTING segments (npart[p] <= capacity[p], else error): used when the
 * host re-initialised the particles of some patches (ParticlesBase.initialize, core/particles.py:118-139, called by
 * MovingWindow._fill_particles, callback/utils.py:760-776) and uploads the new values next. */
int lpic_species_set_npart(lpic_ctx *ctx, int ispec, const int64_t *npart);

/* ---- Maxwell: core/maxwell/cpu.py:9-35 (2D), :83-112 (3D); facade core/maxwell/solver/solver.py:193-254 */
int lpic_update_efield(lpic_ctx *ctx, double dt);
int lpic_update_bfield(lpic_ctx *ctx, double dt);

/* ---- CPML open boundaries (core/boundary/cpml.py; selected per patch by core/maxwell/solver/solver.py:52-106).
 *      One instance = one PML face of one patch.  inst_slot = position in the patch's pml_boundary list (the psi
 *      corrections of a patch are applied in that order).  ranges: (ninst, 4) = efield_start, efield_end, bfield_start,
 *      bfield_end along the face's axis.  profiles: (ninst, 6, nmax) = kappa_e, sigma_e, a_e, kappa_b, sigma_b, a_b along
 *      that axis (kappa = 1, sigma = a = 0 outside the layer).  After this call lpic_update_efield / _bfield use the
 *      kappa-scaled update in the patches that own an instance and advance the psi currents (cpml.py:527-730).
 *      psi arena: [instance][E1, E2, B1, B2][nx*ny*nz]; E1/E2/B1/B2 in the reference's order per axis
 *      (x: psi_ey_x psi_ez_x psi_by_x psi_bz_x; y: psi_ex_y psi_ez_y psi_bx_y psi_bz_y; z: psi_ex_z psi_ey_z psi_bx_z psi_by_z). */
int lpic_pml_configure(lpic_ctx *ctx, int64_t ninst, const int64_t *inst_patch, const int64_t *inst_axis, const int64_t *inst_slot,
                       const int64_t *ranges, const double *profiles, int64_t nmax);
int64_t lpic_pml_psi_words(const lpic_ctx *ctx);
int lpic_pml_upload_psi(lpic_ctx *ctx, const double *host);
int lpic_pml_download_psi(lpic_ctx *ctx, double *host);

/* ---- laser antenna at xmin (callback/laser.py:17-77, 171-186, 218-241): rewrites B at the plane laserpos-1 of the listed
 *      (xmin edge) patches from per-patch source planes.  ranges: (n, 4) = iy_start, iy_end, iz_start, iz_end (interior
 *      minus the transverse PML); ey_src / ez_src: (n, NY[, NZ]) host arrays in the padded, wrapped layout of one x-plane. */
int lpic_laser_bfields(lpic_ctx *ctx, int64_t laserpos, int64_t n, const int64_t *patches, const int64_t *ranges,
                       const double *ey_src, const double *ez_src, double dt);

/* ---- guard cells: core/patch/sync_fields3d.c:350-620 / :84-348, sync_fields2d.c:150-255 / :43-148;
 *      facade core/patch/patch.py:670-703.  attr_mask: bit a = field attribute a. */
int lpic_sync_guard_fields(lpic_ctx *ctx, uint32_t attr_mask);
int lpic_sync_currents(lpic_ctx *ctx);
/* core/current/cpu3d.c:185-240 (reset_current_cpu_3d), cpu2d.c twin */
int lpic_reset_currents(lpic_ctx *ctx);

/* ---- particles ----------------------------------------------------------------------------------
 * unified_boris_pusher_cpu_{2d,3d}(particles_list, fields_list, npatches, dt, q, m)
 * (core/pusher/unified/unified_pusher_3d.c:219-436, unified_pusher_2d.c:157-365) */
int lpic_push_deposit(lpic_ctx *ctx, int ispec, double dt, double q, double m, int flags);
/* non-fused stages: interpolation_patches_* (core/interpolation/cpu3d.c:99-169), boris_push_patches and
 * push_position_patches_2d (core/pusher/cpu.py:10-90), current_deposition_cpu_* (core/current/cpu3d.c:118-184) */
int lpic_interpolate(lpic_ctx *ctx, int ispec);
int lpic_push_momentum(lpic_ctx *ctx, int ispec, double dt, double q, double m);
int lpic_push_position(lpic_ctx *ctx, int ispec, double dt);
int lpic_deposit(lpic_ctx *ctx, int ispec, double dt, double q);

/* ---- sort: sort_particles_patches_{2d,3d} (core/sort/cpu3d.c:214-299); x0s..: per-patch bucket origins
 *      (core/sort/particle_sort.py:331-333).  *nbuf_total receives the reference's return value. */
int lpic_sort(lpic_ctx *ctx, int ispec, int reverse_x, int64_t nxb, int64_t nyb, int64_t nzb,
              double dxb, double dyb, double dzb, const double *x0s, const double *y0s, const double *z0s,
              int64_t *nbuf_total);
/* which = LPIC_SORT_*; out: bucket arrays (npatch, nbin) int64, particle_index: arena layout int64 */
int lpic_sort_download(lpic_ctx *ctx, int ispec, int which, int64_t *out);
/* sums for the mirrored-order decision (core/sort/particle_sort.py:64-89): out[0]=sum w, out[1]=sum w*ux (alive) */
int lpic_weighted_drift(lpic_ctx *ctx, int ispec, double *out2);

/* ---- intra-rank migration: get_npart_to_extend_* / fill_particles_from_boundary_*
 *      (core/patch/sync_particles_3d.c:365-482 / :484-695; facade core/patch/patch.py:705-764) */
int lpic_migrate_count(lpic_ctx *ctx, int ispec, int64_t *npart_to_extend, int64_t *npart_incoming,
                       int64_t *npart_outgoing, int64_t *npart_alive);
int lpic_migrate_fill(lpic_ctx *ctx, int ispec);

/* ---- diagnostics (device-side reductions used by the bench and the energy-history test) ---------- */
int lpic_count_alive(lpic_ctx *ctx, int ispec, int64_t *out);
/* out[0] = sum w*(gamma-1) over alive particles of ispec (kinetic energy / (m c^2)) */
int lpic_kinetic_sum(lpic_ctx *ctx, int ispec, double *out);
/* out[0] = sum E^2, out[1] = sum B^2 over interior cells */
int lpic_field_energy_sums(lpic_ctx *ctx, double *out2);

/* ---- synthetic loader for the bench (uniform-in-cell positions as core/patch/cpu.py:66-99, thermal
 *      momenta with per-component sigma `uth`); deterministic in (seed, patch, slot) */
int lpic_species_init_uniform(lpic_ctx *ctx, int ispec, int64_t ppc, double weight, double uth, uint64_t seed);

/* ---- inter-rank exchange (replaces core/mpi/sync_fields{2d,3d}.c and core/mpi/sync_particles_{2d,3d}.c).
 *      The library packs ONE staging buffer per peer rank and unpacks what arrived; the transport is NCCL
 *      send/recv issued by the host on those device buffers (lambdapic_b200/multigpu.py).
 *      Plan: for each peer slot, nsend[slot] / nrecv[slot] entries (local patch, boundary) in the canonical order
 *      (ascending global index of the receiving patch, then boundary id at the receiver). */
int lpic_halo_plan(lpic_ctx *ctx, int npeers, const int64_t *nsend, const int64_t *send_patch, const int64_t *send_b,
                   const int64_t *nrecv, const int64_t *recv_patch, const int64_t *recv_b);
/* fp64 words per grid attribute exchanged with the peer (recv = 0: what we send, 1: what we receive) */
int64_t lpic_halo_words(lpic_ctx *ctx, int peer_slot, int recv);
/* reduce = 0: E/B guard copy, sends interior strips (core/mpi/sync_fields3d.c:883-996);
 * reduce = 1: J/rho, sends guard strips and zeroes them (:713-866).  Buffer layout [attribute][words]. */
int lpic_halo_pack(lpic_ctx *ctx, int peer_slot, uint32_t attr_mask, int reduce, double *dev_send);
int lpic_halo_unpack(lpic_ctx *ctx, uint32_t attr_mask, int reduce, const double *const *dev_recv_per_peer);
/* remote migration (core/mpi/sync_particles_3d.c:413-745): prepare counts the alive leavers of every send entry
 * and the dead slots per patch; pack writes them AoS (record = resident attributes) and marks them dead; unpack
 * fills the dead slots of each patch in ascending order with the arrivals ordered by (boundary, sender slot). */
int lpic_particle_record_words(lpic_ctx *ctx, int ispec);
int lpic_remote_migrate_prepare(lpic_ctx *ctx, int ispec, int64_t *send_counts, int64_t *ndead);
int lpic_remote_migrate_relist(lpic_ctx *ctx, int ispec);
int lpic_remote_migrate_pack(lpic_ctx *ctx, int ispec, int peer_slot, double *dev_send, int64_t *nparticles);
int lpic_remote_migrate_unpack(lpic_ctx *ctx, int ispec, const int64_t *recv_counts, const double *const *dev_recv_per_peer);

/* ---- measurement helpers: CUDA events on the context stream (bench.py), and the number of kernels this
 *      library has launched so far in the process */
int lpic_event_record(lpic_ctx *ctx, int slot);                              /* slot in [0, 4096) */
int lpic_event_elapsed_ms(lpic_ctx *ctx, int slot_a, int slot_b, double *ms); /* synchronises on slot_b */
int64_t lpic_launch_count(void);
/* MovingWindow recycle without a round trip of the whole state (callback/utils.py:591-840): the recycled patches' fields and psi
 * arrays are cleared on the device, and only their freshly loaded particles are uploaded (slots [off[p], off[p] + npart[p]) of
 * an arena with the device's layout, after lpic_species_set_npart). */
int lpic_zero_patches(lpic_ctx *ctx, int64_t n, const int64_t *patches);
int lpic_upload_particles_patch(lpic_ctx *ctx, int ispec, int attr, int64_t patch, const void *host_arena);
/* One interior z-plane per patch of the attributes in attr_mask (3D): kz[p] = plane index inside patch p (0 <= kz < nz) or -1
 * if the patch does not contain the plane.  host = [nattr][npatch][nx][ny] fp64.  Replaces the whole-patch copies behind
 * callback/utils.py:125-230 (get_fields_3d) and callback/hdf5.py:451-481 (SaveFieldsToHDF5(slice=...)). */
int lpic_download_field_slice(lpic_ctx *ctx, uint32_t attr_mask, const int64_t *kz, double *host);
/* ---- inter-rank transport inside the library (comm.cu): NCCL point-to-point on its own stream, ordered against the compute
 * stream with events.  Replaces core/mpi/mpi_manager.py:9-298 and the start/wait pairs of core/mpi/sync_fields{2,3}d.c
 * (:713-866 currents, :883-996 guards) and core/mpi/sync_particles_{2,3}d.c:413-745.  Call order: lpic_halo_plan ->
 * lpic_comm_unique_id on one rank, id passed to the others by the host (MPI_Bcast / torch.distributed) -> lpic_comm_init
 * on every rank (collective).  peer_rank[s] = rank of the plan's peer slot s. */
int lpic_comm_unique_id(void *id128);
int lpic_comm_nccl_version(void);
int lpic_comm_init(lpic_ctx *ctx, const void *id128, int rank, int nranks, const int64_t *peer_rank);
int64_t lpic_comm_bytes_sent(const lpic_ctx *ctx);
/* after lpic_halo_plan replaced the plan of a context that has a communicator (MovingWindow shift, callback/utils.py:648-716) */
int lpic_comm_update(lpic_ctx *ctx, const int64_t *peer_rank);
/* pack + ncclSend/ncclRecv per peer, asynchronous; run the intra-rank lpic_sync_guard_fields / lpic_sync_currents between
 * start and wait (simulation/simulation.py:948-952); wait = compute stream waits for the receive, then one unpack kernel */
int lpic_halo_start(lpic_ctx *ctx, uint32_t attr_mask, int reduce);
int lpic_halo_wait(lpic_ctx *ctx);
/* particles: classification, counts exchanged device to device, ONE host synchronisation for the message sizes, payload
 * exchange overlapped with the intra-rank fill.  Returns 1 when some patch has to grow first: call lpic_species_extend with
 * to_extend, then again with resume = 1.  info[4] = sent, received, largest per-patch remote / intra-rank arrival count.
 * The call pair performs the WHOLE migration of the species (other ranks and intra-rank). */
int lpic_migrate_remote_start(lpic_ctx *ctx, int ispec, int resume, int64_t *to_extend, int64_t *info);
int lpic_migrate_remote_wait(lpic_ctx *ctx, int ispec);
/* PCI bus id of the context's device, e.g. "0000:1b:00.0" (NUMA placement of the pinned host mirrors) */
int lpic_device_pci_bus_id(lpic_ctx *ctx, char *out, int len);
/* measured fp64 FMA throughput of the context's device in TFLOP/s (bench.py's roofline.fp64 co-bound; no reference counterpart) */
int lpic_fp64_peak(lpic_ctx *ctx, double *tflops);

void *lpic_stream(lpic_ctx *ctx); /* cudaStream_t of the context (for event timing by the host) */

#ifdef __cplusplus
}
#endif
#endif
