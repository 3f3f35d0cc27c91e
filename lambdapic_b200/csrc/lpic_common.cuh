// Shared definitions of the device library: context, geometry, index helpers.
// Layout conventions follow the reference (core/fields.py:24-26, core/utils/cutils.h:19-26): a grid is
// C-contiguous (NX,NY,NZ) with N = n + 2*ng and logical index i in [-ng, n+ng) stored at (i < 0 ? i + N : i).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "../../include/lpic_b200.h"

typedef int64_t i64;
typedef uint8_t u8;
typedef uint64_t u64;

#define LPIC_C_LIGHT 299792458.0           // core/utils/cutils.h:17
#define LPIC_ONE_THIRD 0.3333333333333333  // core/utils/cutils.h:18
#define LPIC_EPS0 8.8541878188e-12         // scipy.constants.epsilon_0 (CODATA 2022), core/maxwell/cpu.py:91

struct Geom {
    int dim, nb;         // 2|3, 8|26 boundaries
    int nx, ny, nz, ng;  // interior cells per patch, guard width
    int ngz;             // ng in 3D, 0 in 2D
    int NX, NY, NZ;      // padded sizes (NZ = 1 in 2D)
    int ncell;           // NX*NY*NZ
    int npatch;
    double dx, dy, dz;
};

// Everything the bucket key of a position depends on (lpic_sort's arguments); org = (3, npatch) bucket origins on the device
struct SortKeyParams {
    const double *org = nullptr;
    int npatch = 0, nxb = 0, nyb = 0, nzb = 0, nbin = 0, reverse_x = 0, dim = 0;
    double dxb = 0, dyb = 0, dzb = 0;
};

struct SortState {
    // Bucket keys left behind by the migration (migrate.cu): its classification pass reads every alive slot's position at
    // the end of step n, which is the position the sorter of step n+1 keys by -- except for the slots the migration itself
    // fills, and it keys those too.  With records a position-only pass pays for whole 64-byte DRAM accesses, so the sorter
    // reading 4 + 1 bytes per slot instead is worth a 4-byte array.  keys_valid: every alive slot's key is current for the
    // parameters in kp; anything else that moves particles or slots clears it.
    int *kcache = nullptr;
    i64 kcache_cap = 0;
    bool keys_written = false, keys_valid = false, have_kp = false;
    SortKeyParams kp;
    double *d_korg = nullptr, *h_korg = nullptr;  // this species' bucket origins (3, npatch), device and host copies
    i64 nxb = 0, nyb = 0, nzb = 0, nbin = 0;
    i64 *bucket_count = nullptr, *bound_min = nullptr, *bound_max = nullptr;  // (npatch, nbin) device
    int *pidx = nullptr;                                                       // arena-sized, pre-sort bucket of every slot
    i64 pidx_cap = 0;
    bool valid = false;  // bound_min/max describe the current slot order (set by sort, cleared by extend/fill)
};

struct Species {
    bool allocated = false;
    bool with_part = false;
    i64 total = 0;                  // physical slots in the arena
    i64 *h_off = nullptr, *h_pcap = nullptr, *h_npart = nullptr;  // host copies (npatch)
    i64 *d_off = nullptr, *d_npart = nullptr;                     // device copies
    i64 max_npart = 0;
    bool lists_valid = false;  // migrate.cu: la / lb / out / ndead of the last k_lists still describe the slots ...
    unsigned long long remote_epoch = ~0ull;  // host-driven inter-rank path: scratch epoch at which prepare / relist built the lists
    unsigned long long lists_epoch = 0;  // ... and nobody else used the shared scratch lists since (lpic_ctx::scratch_epoch)
    i64 max_incoming = -1;  // largest per-patch newcomer count of the last lpic_migrate_count (-1: unknown)
    // Particle layout.  The eight attributes every particle kernel reads or writes (x y z w | ux uy uz inv_gamma, enum order)
    // live in ONE arena of 64-byte records: a particle is two full 32-byte sectors wherever its slot is, so the fused kernel
    // costs the same whether the reference's slot order is still cell order or has decayed (SoA: 15 sectors per particle
    // then).  attr[a] of those eight points INTO the record arena (rec + a) and is indexed with stride pstride = 8; _id and
    // the optional *_part arrays are plain arrays (stride 1).  LPIC_PARTICLE_LAYOUT=soa at allocation keeps eight separate
    // arrays (pstride = 1) for A/B runs: every kernel indexes through the stride.
    double *attr[LPIC_NPATTR] = {nullptr};
    double *rec = nullptr;  // record arena (pstride == 8), owner of attr[0..7]
    int pstride = 1;
    u8 *dead = nullptr;
    SortState sort;
    // migration bookkeeping (device): per (patch, boundary) leaver counts, per patch dead counts
    i64 *d_out = nullptr, *d_ndead = nullptr, *d_incoming = nullptr, *d_extend = nullptr, *d_alive = nullptr;
};

#define LPIC_MAX_PEERS 32
// Inter-rank exchange plan: entries are (local patch, boundary) pairs whose neighbour lives on another rank, grouped
// by peer and ordered canonically (ascending global index of the RECEIVING patch, then boundary at the receiver) so
// that the sender's and the receiver's lists line up without any negotiation.
struct HaloPlan {
    int npeers = 0;
    i64 nsend_total = 0, nrecv_total = 0;
    i64 send_first[LPIC_MAX_PEERS + 1] = {0}, recv_first[LPIC_MAX_PEERS + 1] = {0};  // entry ranges per peer
    i64 send_words[LPIC_MAX_PEERS] = {0}, recv_words[LPIC_MAX_PEERS] = {0};          // fp64 words per grid attribute
    int *h_send_patch = nullptr, *h_send_b = nullptr, *h_recv_patch = nullptr, *h_recv_b = nullptr;
    int *d_send_patch = nullptr, *d_send_b = nullptr;
    int *d_recv_patch = nullptr, *d_recv_b = nullptr;  // receive entries in plan order (comm.cu: counts arrive in this order)
    i64 *d_send_woff = nullptr;                        // word offset of each send entry inside its peer's buffer
    int *d_recv_peer = nullptr;                        // (npatch, nb): peer slot the boundary receives from, or -1
    i64 *d_recv_woff = nullptr;                        // (npatch, nb): word offset inside that peer's buffer
    // particle migration (per call): send entry -> particle count / offset; (patch, boundary) -> count / offset
    i64 *h_mig_send_cnt = nullptr, *d_mig_send_cnt = nullptr, *d_mig_send_poff = nullptr;
    i64 *d_mig_recv_cnt = nullptr, *d_mig_recv_poff = nullptr, *d_mig_incoming = nullptr;
};

// CPML state (fields.cu): per-patch kappa profiles, per-instance psi arrays and coefficient profiles
struct PmlState {
    i64 ninst = 0, nmax = 0, ncint = 0;
    int nslot = 0;
    u8 *d_is_pml = nullptr;          // (npatch)
    double *d_kappa = nullptr;       // (npatch, 2 [e,b], 3 [axis], nmax)
    int *d_inst = nullptr;           // (ninst, 8): patch, axis, e_lo, e_hi, b_lo, b_hi, -, -   (grouped by slot)
    i64 slot_first[5] = {0, 0, 0, 0, 0};
    double *h_prof = nullptr;        // (ninst, 6, nmax) host copy, slot-grouped order
    double *d_coef = nullptr;        // (ninst, 4 [bE, cE, bB, cB], nmax) for the dt of the last call
    double coef_dt[2] = {-1.0, -1.0};
    double *d_psi = nullptr;         // (ninst, 4, ncint), CALLER's instance order
    i64 *h_order = nullptr, *d_order = nullptr;  // slot-grouped position -> caller's instance index
    int *h_axis = nullptr;
};

struct lpic_ctx {
    Geom g;
    int device = 0;
    int nspec = 0;
    cudaStream_t stream = nullptr;
    double *fields = nullptr;  // [LPIC_NFIELD][npatch][ncell]
    double *d_x0 = nullptr, *d_y0 = nullptr, *d_z0 = nullptr;
    double *h_x0 = nullptr, *h_y0 = nullptr, *h_z0 = nullptr;
    i64 *d_nbr = nullptr, *h_nbr = nullptr;  // (npatch, nb)
    double *d_box = nullptr;                 // (npatch, 6)
    double glob[6] = {0, 0, 0, 0, 0, 0};
    i64 rank = 0;
    i64 *h_patch_index = nullptr;
    Species *spec = nullptr;
    // scratch shared by sort / migration of all species (arena-sized, grown on demand)
    int *scr_a = nullptr, *scr_b = nullptr;  // int32 lists
    double *scr_buf = nullptr;               // staging for the sort's value move
    i64 scr_cap = 0;
    bool perm_attr_set = false;              // k_cell_perm's dynamic shared-memory limit raised on this context's device
    bool tile2d_attr_set = false;
    bool tile_attr_set = false;              // same for the tile kernels (push_tile.cu)
    int *d_tile_start = nullptr;             // (npatch, ntile + 1) first position of every tile in the cell-ordered permutation
    size_t tile_start_cap = 0;
    double *d_slice = nullptr;               // staging of lpic_download_field_slice
    int *d_slice_k = nullptr;
    size_t slice_cap = 0;
    i64 *d_ext_tab = nullptr;                // lpic_species_extend: ext / first id / new offsets (3 x npatch)
    double **d_ext_attrs = nullptr;          // ... and its table of attribute pointers
    int *d_sort_hist = nullptr;              // global histogram of lpic_sort when a patch has more bins than fit shared memory
    size_t sort_hist_cap = 0;
    int *d_laser_i = nullptr;                // staging of lpic_laser_bfields (patch list, ranges, source planes)
    double *d_laser_s = nullptr;
    size_t laser_cap_n = 0, laser_cap_words = 0;
    unsigned long long scratch_epoch = 0;    // bumped by every user of the scratch lists (lpic_ensure_scratch)
    double *d_sort_org = nullptr;            // (3, npatch) bucket origins
    i64 *d_tmp64 = nullptr;                  // small reductions (>= 8 + npatch words)
    double *d_tmpf = nullptr;
    struct PmlState *pml = nullptr;   // null: no open boundaries
    struct HaloPlan *halo = nullptr;  // inter-rank exchange plan (halo.cu), null on a single rank
    const i64 *comm_remote_in = nullptr;  // comm.cu: per-patch remote arrival counts of the exchange in flight
    struct CommState *comm = nullptr;  // NCCL communicator, comm stream, staging buffers (comm.cu), null until lpic_comm_init
    cudaEvent_t *events = nullptr;  // lazily created, 4096 slots
    struct XferState *xfer = nullptr;  // staging of the host <-> record moves (api.cu), null until first used
};

// Every entry point runs on its context's device whatever device is current in the calling thread (a process may hold
// contexts on several GPUs, and torch may have switched the current device between two calls).
struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    explicit DeviceGuard(const lpic_ctx *c) {
        if (c && cudaGetDevice(&prev) == cudaSuccess && prev != c->device) switched = cudaSetDevice(c->device) == cudaSuccess;
    }
    ~DeviceGuard() {
        if (switched) cudaSetDevice(prev);
    }
    DeviceGuard(const DeviceGuard &) = delete;
    DeviceGuard &operator=(const DeviceGuard &) = delete;
};

#define LPIC_NREC 8  // attributes 0..7 form the record
inline int attr_stride(const Species &sp, int a) { return a < LPIC_NREC ? sp.pstride : 1; }

inline double *field_ptr(const lpic_ctx *c, int attr) { return c->fields + (size_t)attr * c->g.npatch * c->g.ncell; }

void lpic_set_error(const char *fmt, ...);
#define CUDA_TRY(expr)                                                                            \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) {                                                                  \
            lpic_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return -1;                                                                            \
        }                                                                                         \
    } while (0)
#define KERNEL_CHECK() CUDA_TRY(cudaGetLastError())
#define REQUIRE(cond, ...)               \
    do {                                 \
        if (!(cond)) {                   \
            lpic_set_error(__VA_ARGS__); \
            return -2;                   \
        }                                \
    } while (0)

// boundary direction tables, enum order of core/patch/sync_fields3d.c:19-50 and sync_fields2d.c:18-28
__constant__ const signed char kDir3[26][3] = {
    {-1, 0, 0}, {1, 0, 0}, {0, -1, 0}, {0, 1, 0}, {0, 0, -1}, {0, 0, 1},
    {-1, -1, 0}, {-1, 1, 0}, {-1, 0, -1}, {-1, 0, 1}, {1, -1, 0}, {1, 1, 0}, {1, 0, -1}, {1, 0, 1},
    {0, -1, -1}, {0, -1, 1}, {0, 1, -1}, {0, 1, 1},
    {-1, -1, -1}, {-1, -1, 1}, {-1, 1, -1}, {-1, 1, 1}, {1, -1, -1}, {1, -1, 1}, {1, 1, -1}, {1, 1, 1}};
__constant__ const signed char kDir2[8][3] = {{-1, 0, 0}, {1, 0, 0}, {0, -1, 0}, {0, 1, 0},
                                              {-1, -1, 0}, {1, -1, 0}, {-1, 1, 0}, {1, 1, 0}};
// (sx+1) + 3*(sy+1) + 9*(sz+1) -> boundary id (-1 for the centre)
__constant__ const signed char kLut3[27] = {18, 14, 22, 8, 4, 12, 20, 16, 24, 6, 2, 10, 0, -1, 1, 7, 3, 11,
                                            19, 15, 23, 9, 5, 13, 21, 17, 25};
__constant__ const signed char kLut2[9] = {4, 2, 5, 0, -1, 1, 6, 3, 7};

static const signed char hDir3[26][3] = {
    {-1, 0, 0}, {1, 0, 0}, {0, -1, 0}, {0, 1, 0}, {0, 0, -1}, {0, 0, 1},
    {-1, -1, 0}, {-1, 1, 0}, {-1, 0, -1}, {-1, 0, 1}, {1, -1, 0}, {1, 1, 0}, {1, 0, -1}, {1, 0, 1},
    {0, -1, -1}, {0, -1, 1}, {0, 1, -1}, {0, 1, 1},
    {-1, -1, -1}, {-1, -1, 1}, {-1, 1, -1}, {-1, 1, 1}, {1, -1, -1}, {1, -1, 1}, {1, 1, -1}, {1, 1, 1}};
static const signed char hDir2[8][3] = {{-1, 0, 0}, {1, 0, 0}, {0, -1, 0}, {0, 1, 0},
                                        {-1, -1, 0}, {1, -1, 0}, {-1, 1, 0}, {1, 1, 0}};
static const signed char hLut3[27] = {18, 14, 22, 8, 4, 12, 20, 16, 24, 6, 2, 10, 0, -1, 1, 7, 3, 11,
                                      19, 15, 23, 9, 5, 13, 21, 17, 25};
static const signed char hLut2[9] = {4, 2, 5, 0, -1, 1, 6, 3, 7};

__host__ __device__ inline int dir_component(int dim, int b, int axis) {
#ifdef __CUDA_ARCH__
    return dim == 3 ? kDir3[b][axis] : kDir2[b][axis];
#else
    return dim == 3 ? hDir3[b][axis] : hDir2[b][axis];
#endif
}
__host__ __device__ inline int dir_lookup(int dim, int sx, int sy, int sz) {
#ifdef __CUDA_ARCH__
    return dim == 3 ? kLut3[(sx + 1) + 3 * (sy + 1) + 9 * (sz + 1)] : kLut2[(sx + 1) + 3 * (sy + 1)];
#else
    return dim == 3 ? hLut3[(sx + 1) + 3 * (sy + 1) + 9 * (sz + 1)] : hLut2[(sx + 1) + 3 * (sy + 1)];
#endif
}
__host__ __device__ inline int dir_opposite(int dim, int b) {
    return dir_lookup(dim, -dir_component(dim, b, 0), -dir_component(dim, b, 1), -dir_component(dim, b, 2));
}

__host__ __device__ inline int wrapneg(int i, int N) { return i < 0 ? i + N : i; }

// Bucket of a position (core/sort/cpu3d.c:8-60 calculate_cell_index, restated): bucket coordinates as doubles (floor of the
// reference's quotient); NaN compares false everywhere and ends up out of range / clamped to 0, like
// (npy_intp)floor(NaN) = INT64_MIN does on the host.  Single-bucket axes (ny_buckets = nz_buckets = 1, the default):
// floor(v / d) is 0 exactly when 0 <= v < d -- for doubles v < d the rounded quotient is at most 1 - 2^-53 < 1 -- so the
// reference's division is replaced by two comparisons there without changing any result.  ONE definition, used by the
// sorter and by the migration kernels that leave keys behind for it.
__device__ __forceinline__ int sort_bucket_key(const SortKeyParams &k, double x0, double y0, double z0, double px, double py, double pz) {
    const double fx = floor((px - x0) / k.dxb);
    const double vy = py - y0, vz = k.dim == 3 ? pz - z0 : 0.0;
    const double fy = k.nyb == 1 ? (vy >= 0.0 ? (vy < k.dyb ? 0.0 : 1.0) : -1.0) : floor(vy / k.dyb);
    const double fz = k.dim != 3 ? 0.0 : (k.nzb == 1 ? (vz >= 0.0 ? (vz < k.dzb ? 0.0 : 1.0) : -1.0) : floor(vz / k.dzb));
    const bool inx = fx >= 0.0 && fx < (double)k.nxb, iny = fy >= 0.0 && fy < (double)k.nyb, inz = fz >= 0.0 && fz < (double)k.nzb;
    if (k.reverse_x) {
        const int ix = inx ? (int)fx : (fx >= (double)k.nxb ? k.nxb - 1 : 0);
        const int iy = iny ? (int)fy : (fy >= (double)k.nyb ? k.nyb - 1 : 0);
        const int iz = inz ? (int)fz : (fz >= (double)k.nzb ? k.nzb - 1 : 0);
        return iz + iy * k.nzb + (k.nxb - 1 - ix) * k.nyb * k.nzb;
    }
    if (inx && iny && inz) return (int)fz + (int)fy * k.nzb + (int)fx * k.nyb * k.nzb;
    return k.nbin - 1;
}

// kernels' launch helpers
static inline unsigned div_up(i64 a, i64 b) { return (unsigned)((a + b - 1) / b); }

#include <atomic>
extern std::atomic<long long> g_lpic_launches;  // kernels launched by this library (bench.py's gpu_launches)
#define LAUNCHED(n) (g_lpic_launches += (n))

// entry points implemented per file
int lpic_ensure_scratch(lpic_ctx *c, i64 slots);
