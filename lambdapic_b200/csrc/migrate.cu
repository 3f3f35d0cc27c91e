// Intra-rank particle migration between patches, reproducing the reference's slot assignment bit-for-bit.
//
// Reference behaviour restated (not copied): core/patch/sync_particles_3d.c:79-193 (classification into the 26
// directions), :365-482 (counts, growth rule), :204-323 (leaver lists and the AoS buffer order), :349-363
// (periodic shift), :326-346 (kill + NaN), :484-695 (driver); 2D twin sync_particles_2d.c; facade patch.py:705-764.
//   incoming stream of patch p = for b in enum order (neighbour q = nbr[p][b] >= 0):
//                                    q's alive leavers towards opposite(b), ascending slot order
//   k-th incoming particle -> k-th dead slot of p (ascending).
// Device plan: (1) k_lists, one CTA per patch and ONE classification pass: leavers counted per direction and listed
// grouped by direction (stable), dead slots counted and listed; (2) k_plan, one thread per patch: incoming / alive /
// npart_to_extend; [host grows the arrays: lpic_species_extend -- only then are the lists rebuilt]; (3) k_fill, one
// thread per incoming particle copies every attribute; (4) k_mark kills the listed leavers (and NaNs the unfilled dead
// slots in 3D) from the lists, without classifying again.
#include <algorithm>
#include <vector>
#include "lpic_common.cuh"

namespace {

constexpr int T = 256;

struct MigArgs {
    int dim, nb, npatch, nattr;
    const i64 *off, *npart, *nbr;
    const double *box;  // (npatch, 6)
    double *x, *y, *z;
    u8 *dead;
    i64 *out, *ndead, *incoming, *extend, *alive;  // per patch (out: per patch x boundary)
    const i64 *remote_in;                          // per patch: arrivals from other ranks take the first dead slots (null: none)
    int *la, *lb;                                  // arena-sized int lists
    int *dirstart;                                 // (npatch, nb)
    SortKeyParams kp;  // the sorter's parameters and ...
    int *kcache;       // ... its key cache (SortState): every kernel that sees or sets a position leaves the bucket key (null: off)
    double *attrs[LPIC_NPATTR];
    int astride[LPIC_NPATTR];  // 8 for the attributes inside the record arena, else 1
    int ps;                    // stride of x, y, z
    int ia_x, ia_y, ia_z;
    double glob[6], cell[3];
};

__device__ __forceinline__ void load_position(const MigArgs &a, i64 ip, double &x, double &y, double &z) {
    if (a.ps == LPIC_NREC) {
        const double2 xy = *reinterpret_cast<const double2 *>(a.x + ip * LPIC_NREC);
        x = xy.x; y = xy.y;
    } else {
        x = a.x[ip]; y = a.y[ip];
    }
    z = a.dim == 3 ? a.z[ip * a.ps] : 0.0;
}
__device__ __forceinline__ int classify_position(const MigArgs &a, const double *bx, double x, double y, double z) {
    const int sx = x < bx[0] ? -1 : (x > bx[1] ? 1 : 0);
    const int sy = y < bx[2] ? -1 : (y > bx[3] ? 1 : 0);
    const int sz = a.dim == 3 ? (z < bx[4] ? -1 : (z > bx[5] ? 1 : 0)) : 0;
    return dir_lookup(a.dim, sx, sy, sz);  // -1 when inside
}
__device__ __forceinline__ int classify(const MigArgs &a, const double *bx, i64 ip) {
    double x, y, z;
    load_position(a, ip, x, y, z);
    return classify_position(a, bx, x, y, z);
}

__device__ __forceinline__ int warp_incl_sum(int v) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int n = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += n;
    }
    return v;
}
__device__ __forceinline__ int block_incl_sum(int v, int *sw, int &total) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    v = warp_incl_sum(v);
    __syncthreads();
    if (lane == 31) sw[w] = v;
    __syncthreads();
    int add = 0, tot = 0;
#pragma unroll
    for (int i = 0; i < T / 32; i++) {
        if (i < w) add += sw[i];
        tot += sw[i];
    }
    total = tot;
    return v + add;
}

__global__ void k_plan(MigArgs a) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= a.npatch) return;
    i64 incoming = 0;
    for (int b = 0; b < a.nb; b++) {
        const i64 q = a.nbr[(size_t)p * a.nb + b];
        if (q >= 0) incoming += a.out[(size_t)q * a.nb + dir_opposite(a.dim, b)];
    }
    const i64 npart = a.npart[p], ndead = a.ndead[p];
    a.incoming[p] = incoming;
    a.alive[p] = npart - ndead + incoming;
    // grow only when the dead slots cannot take the newcomers; a quarter of the capacity is added on top
    a.extend[p] = incoming - ndead > 0 ? incoming - ndead + (i64)((double)npart * 0.25) : 0;
}

// ONE classification pass per patch: leavers counted per direction (out), listed grouped by direction (stable, ascending
// slots) in lb[off ..]; dead slots counted (ndead) and listed ascending in la from the back.
__global__ void __launch_bounds__(T) k_lists(MigArgs a) {
    __shared__ int sw[T / 32];
    __shared__ int s_cur[32];
    __shared__ int s_cnt[32];
    __shared__ signed char s_code[T * 4];
    const int p = blockIdx.x, tid = threadIdx.x;
    const i64 off = a.off[p];
    const int np = (int)a.npart[p];
    const double *bx = a.box + 6 * (size_t)p;
    double kx0 = 0.0, ky0 = 0.0, kz0 = 0.0;
    if (a.kcache) { kx0 = a.kp.org[p]; ky0 = a.kp.org[a.kp.npatch + p]; kz0 = a.kp.org[2 * a.kp.npatch + p]; }
    if (tid < 32) s_cnt[tid] = 0;
    __syncthreads();
    int nl = 0, nd = 0;
    constexpr int IT = 4;  // consecutive slots per thread and scan step
    for (int base = 0; base < np; base += T * IT) {
        const int ip0 = base + tid * IT;
        bool isdead[IT], leaves[IT];
        int cl = 0, cd = 0;
        // classified with adjacent lanes on adjacent slots (coalesced for arrays and for 64-byte records alike), handed to the
        // thread that owns IT consecutive slots through shared memory: 0 = stays (or past the end), 1 = dead, 2 = leaves
        {
            const int wbase = base + (tid >> 5) * (32 * IT), lane = tid & 31;
#pragma unroll
            for (int m = 0; m < IT; m++) {
                const int ip = wbase + 32 * m + lane;
                signed char code = 0;
                if (ip < np) {
                    if (a.dead[off + ip] != 0) code = 1;
                    else {
                        double x, y, z;
                        load_position(a, off + ip, x, y, z);
                        const int b = classify_position(a, bx, x, y, z);
                        if (b >= 0) { code = 2; atomicAdd(&s_cnt[b], 1); }
                        // the position the next step's sorter keys by (the leavers' slots will be dead by then)
                        if (a.kcache) a.kcache[off + ip] = sort_bucket_key(a.kp, kx0, ky0, kz0, x, y, z);
                    }
                }
                s_code[(tid >> 5) * (32 * IT) + 32 * m + lane] = code;
            }
            __syncwarp();
#pragma unroll
            for (int j = 0; j < IT; j++) {
                const signed char code = s_code[tid * IT + j];
                isdead[j] = code == 1;
                leaves[j] = code == 2;
                cl += leaves[j];
                cd += isdead[j];
            }
            __syncwarp();
        }
        int tot;
        int pl = nl + block_incl_sum(cl, sw, tot) - cl;
        nl += tot;
        __syncthreads();
        int pd = nd + block_incl_sum(cd, sw, tot) - cd;
        nd += tot;
#pragma unroll
        for (int j = 0; j < IT; j++) {
            if (leaves[j]) a.la[off + pl++] = ip0 + j;
            if (isdead[j]) a.la[off + np - 1 - pd++] = ip0 + j;
        }
        __syncthreads();
    }
    __syncthreads();
    if (tid == 0) {
        a.ndead[p] = nd;
        int run = 0;
        for (int b = 0; b < a.nb; b++) {
            s_cur[b] = run;
            a.dirstart[(size_t)p * a.nb + b] = run;
            a.out[(size_t)p * a.nb + b] = s_cnt[b];
            run += s_cnt[b];
        }
    }
    __syncthreads();
    // stable grouping by direction: warps take turns, lanes of one direction get consecutive places
    for (int base = 0; base < nl; base += T) {
        const int i = base + tid;
        const bool act = i < nl;
        int slot = 0, b = -2 - (tid & 31);
        if (act) {
            slot = a.la[off + i];
            b = classify(a, bx, off + slot);
        }
        for (int w = 0; w < T / 32; w++) {
            if ((tid >> 5) == w) {
                const unsigned lane = tid & 31;
                const unsigned peers = __match_any_sync(0xffffffffu, b);
                const int leader = __ffs(peers) - 1;
                int start = 0;
                if (act && (int)lane == leader) { start = s_cur[b]; s_cur[b] = start + __popc(peers); }
                start = __shfl_sync(0xffffffffu, start, leader);
                if (act) a.lb[off + start + __popc(peers & ((1u << lane) - 1u))] = slot;
            }
            __syncthreads();
        }
    }
}

// one thread per incoming particle of patch p
__global__ void __launch_bounds__(T) k_fill(MigArgs a, int blocks_per_patch) {
    const int p = blockIdx.x / blocks_per_patch;
    const i64 k = (i64)(blockIdx.x - p * blocks_per_patch) * blockDim.x + threadIdx.x;
    const i64 skip = a.remote_in ? a.remote_in[p] : 0;  // dead slots already promised to arrivals from other ranks
    if (k >= a.incoming[p] || k + skip >= a.ndead[p]) return;
    // which boundary does the k-th newcomer arrive through?  (fill_boundary_particles_to_buffer, :302-323)
    i64 run = 0;
    i64 q = -1;
    int ob = 0;
    i64 r = 0;
    for (int b = 0; b < a.nb; b++) {
        const i64 qq = a.nbr[(size_t)p * a.nb + b];
        if (qq < 0) continue;
        const int o = dir_opposite(a.dim, b);
        const i64 cnt = a.out[(size_t)qq * a.nb + o];
        if (k < run + cnt) { q = qq; ob = o; r = k - run; break; }
        run += cnt;
    }
    if (q < 0) return;
    const i64 src = a.off[q] + a.lb[a.off[q] + a.dirstart[(size_t)q * a.nb + ob] + r];
    const i64 np = a.npart[p];
    const i64 dst = a.off[p] + a.la[a.off[p] + np - 1 - (k + skip)];
    const double *bx = a.box + 6 * (size_t)p;
    double kpx = 0.0, kpy = 0.0, kpz = 0.0;
    for (int t = 0; t < a.nattr; t++) {
        double v = a.attrs[t][src * a.astride[t]];
        const int d = t == a.ia_x ? 0 : (t == a.ia_y ? 1 : (t == a.ia_z ? 2 : -1));
        if (d >= 0) {  // handle_periodic, :349-363
            const double lo = a.glob[2 * d], hi = a.glob[2 * d + 1], L = hi - lo, c0 = v;
            if (c0 > hi && fabs(bx[2 * d] - lo) < a.cell[d]) v -= L;
            if (c0 < lo && fabs(bx[2 * d + 1] - hi) < a.cell[d]) v += L;
            if (d == 0) kpx = v; else if (d == 1) kpy = v; else kpz = v;
        }
        a.attrs[t][dst * a.astride[t]] = v;
    }
    a.dead[dst] = 0;
    if (a.kcache)  // the newcomer's bucket in its new patch, for the next step's sorter
        a.kcache[dst] = sort_bucket_key(a.kp, a.kp.org[p], a.kp.org[a.kp.npatch + p], a.kp.org[2 * a.kp.npatch + p], kpx, kpy, kpz);
}

// mark_out_of_bound_as_dead (:326-346): alive outside the box -> dead with NaN position; in 3D every dead slot's
// position is NaN'd, in 2D only the freshly killed ones (sync_particles_2d.c:185-202).  Works from the lists of k_lists
// (leavers in lb, dead slots in la; the first min(incoming, ndead) dead slots were just filled and are alive again), so
// nothing is classified a third time.  One CTA per patch.
__global__ void __launch_bounds__(T) k_mark(MigArgs a) {
    const int p = blockIdx.x;
    const i64 off = a.off[p];
    const int np = (int)a.npart[p];
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    int nl = 0;
    for (int b = 0; b < a.nb; b++) nl += (int)a.out[(size_t)p * a.nb + b];
    for (int i = threadIdx.x; i < nl; i += T) {
        const i64 ip = off + a.lb[off + i];
        a.dead[ip] = 1;
        if (a.ps == LPIC_NREC) *reinterpret_cast<double2 *>(a.x + ip * LPIC_NREC) = make_double2(nan, nan);
        else { a.x[ip] = nan; a.y[ip] = nan; }
        if (a.dim == 3) a.z[ip * a.ps] = nan;
    }
    if (a.dim != 3) return;
    const int nd = (int)a.ndead[p];
    const i64 arrivals = a.incoming[p] + (a.remote_in ? a.remote_in[p] : 0);
    const int filled = (int)(arrivals < nd ? arrivals : nd);
    for (int k = filled + threadIdx.x; k < nd; k += T) {
        const i64 ip = off + a.la[off + np - 1 - k];
        if (a.ps == LPIC_NREC) {
            // almost all of these slots have been dead (and NaN) for many steps: a 24-byte write into a 32-byte sector costs a
            // read-modify-write in L2 / DRAM, so look first (x y z share the sector) and write only what is not NaN yet
            double2 *r = reinterpret_cast<double2 *>(a.x + ip * LPIC_NREC);
            const double2 xy = r[0];
            const long long nb = 0x7ff8000000000000ll;  // (bit patterns, so that the result is the one an unconditional write gives)
            if (__double_as_longlong(xy.x) != nb || __double_as_longlong(xy.y) != nb) r[0] = make_double2(nan, nan);
            if (__double_as_longlong(a.z[ip * LPIC_NREC]) != nb) a.z[ip * LPIC_NREC] = nan;
        } else {
            a.x[ip] = nan;
            a.y[ip] = nan;
            a.z[ip] = nan;
        }
    }
}

int make_args(lpic_ctx *c, int ispec, MigArgs &a) {
    REQUIRE(ispec >= 0 && ispec < c->nspec && c->spec[ispec].allocated, "species %d not allocated", ispec);
    Species &sp = c->spec[ispec];
    const Geom &g = c->g;
    if (int r = lpic_ensure_scratch(c, sp.total)) return r;
    a.dim = g.dim; a.nb = g.nb; a.npatch = g.npatch;
    a.off = sp.d_off; a.npart = sp.d_npart; a.nbr = c->d_nbr; a.box = c->d_box;
    a.x = sp.attr[LPIC_P_X]; a.y = sp.attr[LPIC_P_Y]; a.z = sp.attr[LPIC_P_Z]; a.dead = sp.dead;
    a.out = sp.d_out; a.ndead = sp.d_ndead; a.incoming = sp.d_incoming; a.extend = sp.d_extend; a.alive = sp.d_alive;
    a.la = c->scr_a; a.lb = c->scr_b;
    a.remote_in = nullptr;
    {   // the sorter's key cache: on when the sorter has run (parameters known) and its array covers the arena
        SortState &st = sp.sort;
        a.kp = st.kp;
        a.kcache = (st.have_kp && st.kcache && st.kcache_cap >= sp.total) ? st.kcache : nullptr;
    }
    a.dirstart = (int *)(c->d_tmp64 + 64);  // npatch*nb ints = 13 npatch i64 words at most; d_tmp64 holds 64 + 16 npatch (lpic_create)
    a.nattr = 0; a.ia_x = a.ia_y = a.ia_z = -1;
    for (int t = 0; t < LPIC_NPATTR; t++) {
        if (!sp.attr[t]) continue;
        if (t == LPIC_P_X) a.ia_x = a.nattr;
        if (t == LPIC_P_Y) a.ia_y = a.nattr;
        if (t == LPIC_P_Z && g.dim == 3) a.ia_z = a.nattr;
        a.astride[a.nattr] = attr_stride(sp, t);
        a.attrs[a.nattr++] = sp.attr[t];
    }
    a.ps = sp.pstride;
    for (int i = 0; i < 6; i++) a.glob[i] = c->glob[i];
    a.cell[0] = g.dx; a.cell[1] = g.dy; a.cell[2] = g.dz;
    return 0;
}

// ---- inter-rank migration (core/mpi/sync_particles_3d.c:413-745 restated for packed per-peer buffers) -------------
struct PeerRecs {
    const double *p[LPIC_MAX_PEERS];
};

// grid: (chunks of the largest entry, entries of this peer).  Record layout AoS: buf[(poff + r) * nattr + t].
__global__ void __launch_bounds__(T) k_remote_pack(MigArgs a, const int *__restrict__ ent_patch, const int *__restrict__ ent_b,
                                                   const i64 *__restrict__ cnt, const i64 *__restrict__ poff, i64 first,
                                                   double *__restrict__ buf) {
    const i64 e = first + blockIdx.y;
    const i64 r = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= cnt[e]) return;
    const int p = ent_patch[e], b = ent_b[e];
    const i64 src = a.off[p] + a.lb[a.off[p] + a.dirstart[(size_t)p * a.nb + b] + r];
    double *rec = buf + (size_t)(poff[e] + r) * a.nattr;
    for (int t = 0; t < a.nattr; t++) rec[t] = a.attrs[t][src * a.astride[t]];
    a.dead[src] = 1;  // the sender gives the slot up at once (core/mpi/sync_particles_3d.c:573-577)
}

// one thread per arriving particle of patch p: boundaries in enum order, then sender slot order; k-th arrival ->
// k-th dead slot (ascending)
__global__ void __launch_bounds__(T) k_remote_unpack(MigArgs a, const int *__restrict__ recv_peer, const i64 *__restrict__ rcnt,
                                                     const i64 *__restrict__ rpoff, const i64 *__restrict__ incoming,
                                                     PeerRecs bufs, int blocks_per_patch) {
    const int p = blockIdx.x / blocks_per_patch;
    const i64 k = (i64)(blockIdx.x - p * blocks_per_patch) * blockDim.x + threadIdx.x;
    if (k >= incoming[p] || k >= a.ndead[p]) return;
    i64 run = 0;
    const double *rec = nullptr;
    for (int b = 0; b < a.nb; b++) {
        const size_t key = (size_t)p * a.nb + b;
        if (recv_peer[key] < 0) continue;
        const i64 c = rcnt[key];
        if (k < run + c) { rec = bufs.p[recv_peer[key]] + (size_t)(rpoff[key] + (k - run)) * a.nattr; break; }
        run += c;
    }
    if (!rec) return;
    const i64 np = a.npart[p];
    const i64 dst = a.off[p] + a.la[a.off[p] + np - 1 - k];
    const double *bx = a.box + 6 * (size_t)p;
    double kpx = 0.0, kpy = 0.0, kpz = 0.0;
    for (int t = 0; t < a.nattr; t++) {
        double v = rec[t];
        const int d = t == a.ia_x ? 0 : (t == a.ia_y ? 1 : (t == a.ia_z ? 2 : -1));
        if (d >= 0) {
            const double lo = a.glob[2 * d], hi = a.glob[2 * d + 1], L = hi - lo, c0 = v;
            if (c0 > hi && fabs(bx[2 * d] - lo) < a.cell[d]) v -= L;
            if (c0 < lo && fabs(bx[2 * d + 1] - hi) < a.cell[d]) v += L;
            if (d == 0) kpx = v; else if (d == 1) kpy = v; else kpz = v;
        }
        a.attrs[t][dst * a.astride[t]] = v;
    }
    a.dead[dst] = 0;
    if (a.kcache)
        a.kcache[dst] = sort_bucket_key(a.kp, a.kp.org[p], a.kp.org[a.kp.npatch + p], a.kp.org[2 * a.kp.npatch + p], kpx, kpy, kpz);
}

}  // namespace

extern "C" int lpic_particle_record_words(lpic_ctx *c, int ispec) {
    DeviceGuard dg(c);
    if (ispec < 0 || ispec >= c->nspec || !c->spec[ispec].allocated) return -1;
    int n = 0;  // resident attributes (a pure query: it must not touch the scratch epoch between prepare and pack)
    for (int t = 0; t < LPIC_NPATTR; t++) n += c->spec[ispec].attr[t] != nullptr;
    return n;
}

extern "C" int lpic_remote_migrate_prepare(lpic_ctx *c, int ispec, int64_t *send_counts, int64_t *ndead) {
    DeviceGuard dg(c);
    HaloPlan *h = c->halo;
    REQUIRE(h, "no exchange plan");
    MigArgs a;
    if (int r = make_args(c, ispec, a)) return r;
    Species &sp = c->spec[ispec];
    const i64 n = c->g.npatch;
    const int nb = c->g.nb;
    k_lists<<<(unsigned)n, T, 0, c->stream>>>(a);  // counts + lists in one classification pass
    c->spec[ispec].sort.keys_written = a.kcache != nullptr;
    LAUNCHED(1);
    KERNEL_CHECK();
    sp.remote_epoch = c->scratch_epoch;  // pack / unpack insist that nobody else used the shared scratch lists in between
    std::vector<i64> out(n * nb);
    CUDA_TRY(cudaMemcpyAsync(out.data(), sp.d_out, sizeof(i64) * n * nb, cudaMemcpyDeviceToHost, c->stream));
    if (ndead) CUDA_TRY(cudaMemcpyAsync(ndead, sp.d_ndead, sizeof(i64) * n, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    std::vector<i64> poff(h->nsend_total + 1, 0);
    for (int s = 0; s < h->npeers; s++) {
        i64 run = 0;
        for (i64 e = h->send_first[s]; e < h->send_first[s + 1]; e++) {
            const i64 cnt = out[(size_t)h->h_send_patch[e] * nb + h->h_send_b[e]];
            h->h_mig_send_cnt[e] = cnt;
            if (send_counts) send_counts[e] = cnt;
            poff[e] = run;
            run += cnt;
        }
    }
    CUDA_TRY(cudaMemcpyAsync(h->d_mig_send_cnt, h->h_mig_send_cnt, sizeof(i64) * h->nsend_total, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(h->d_mig_send_poff, poff.data(), sizeof(i64) * h->nsend_total, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int lpic_remote_migrate_relist(lpic_ctx *c, int ispec) {
    DeviceGuard dg(c);
    MigArgs a;
    if (int r = make_args(c, ispec, a)) return r;
    k_lists<<<(unsigned)c->g.npatch, T, 0, c->stream>>>(a);
    c->spec[ispec].sort.keys_written = a.kcache != nullptr;
    LAUNCHED(1);
    KERNEL_CHECK();
    c->spec[ispec].remote_epoch = c->scratch_epoch;
    return 0;
}

extern "C" int lpic_remote_migrate_pack(lpic_ctx *c, int ispec, int slot, double *dev_send, int64_t *nparticles) {
    DeviceGuard dg(c);
    HaloPlan *h = c->halo;
    REQUIRE(h && slot >= 0 && slot < h->npeers, "no exchange plan / bad peer slot");
    MigArgs a;
    const unsigned long long epoch = c->scratch_epoch;
    if (int r = make_args(c, ispec, a)) return r;
    REQUIRE(c->spec[ispec].remote_epoch == epoch,
            "lpic_remote_migrate_pack: the leaver lists of lpic_remote_migrate_prepare are stale (another operator used the scratch lists)");
    c->spec[ispec].remote_epoch = c->scratch_epoch;
    i64 total = 0, mx = 0;
    for (i64 e = h->send_first[slot]; e < h->send_first[slot + 1]; e++) { total += h->h_mig_send_cnt[e]; mx = std::max(mx, h->h_mig_send_cnt[e]); }
    if (nparticles) *nparticles = total;
    if (total == 0) return 0;
    dim3 grid(div_up(mx, T), (unsigned)(h->send_first[slot + 1] - h->send_first[slot]));
    k_remote_pack<<<grid, T, 0, c->stream>>>(a, h->d_send_patch, h->d_send_b, h->d_mig_send_cnt, h->d_mig_send_poff,
                                             h->send_first[slot], dev_send);
    LAUNCHED(1);
    KERNEL_CHECK();
    c->spec[ispec].sort.valid = false;
    c->spec[ispec].lists_valid = false;
    return 0;
}

extern "C" int lpic_remote_migrate_unpack(lpic_ctx *c, int ispec, const int64_t *recv_counts, const double *const *dev_recv) {
    DeviceGuard dg(c);
    HaloPlan *h = c->halo;
    REQUIRE(h, "no exchange plan");
    MigArgs a;
    const unsigned long long epoch = c->scratch_epoch;
    if (int r = make_args(c, ispec, a)) return r;
    REQUIRE(c->spec[ispec].remote_epoch == epoch,
            "lpic_remote_migrate_unpack: the dead-slot lists of lpic_remote_migrate_prepare are stale (another operator used the scratch lists)");
    const i64 n = c->g.npatch;
    const int nb = c->g.nb;
    std::vector<i64> rcnt(n * nb, 0), rpoff(n * nb, 0), incoming(n, 0);
    i64 mx = 0;
    for (int s = 0; s < h->npeers; s++) {
        i64 run = 0;
        for (i64 e = h->recv_first[s]; e < h->recv_first[s + 1]; e++) {
            const size_t key = (size_t)h->h_recv_patch[e] * nb + h->h_recv_b[e];
            rcnt[key] = recv_counts[e];
            rpoff[key] = run;
            run += recv_counts[e];
            incoming[h->h_recv_patch[e]] += recv_counts[e];
        }
    }
    for (i64 p = 0; p < n; p++) mx = std::max(mx, incoming[p]);
    if (mx == 0) return 0;
    CUDA_TRY(cudaMemcpyAsync(h->d_mig_recv_cnt, rcnt.data(), sizeof(i64) * n * nb, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(h->d_mig_recv_poff, rpoff.data(), sizeof(i64) * n * nb, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(h->d_mig_incoming, incoming.data(), sizeof(i64) * n, cudaMemcpyHostToDevice, c->stream));
    PeerRecs bufs;
    for (int s = 0; s < h->npeers; s++) bufs.p[s] = dev_recv[s];
    const int bpp = (int)div_up(mx, T);
    k_remote_unpack<<<(unsigned)((i64)bpp * n), T, 0, c->stream>>>(a, h->d_recv_peer, h->d_mig_recv_cnt, h->d_mig_recv_poff,
                                                                 h->d_mig_incoming, bufs, bpp);
    LAUNCHED(1);
    KERNEL_CHECK();
    CUDA_TRY(cudaStreamSynchronize(c->stream));  // the host vectors above must outlive the copies
    c->spec[ispec].sort.valid = false;
    c->spec[ispec].lists_valid = false;
    return 0;
}

// ---- pieces of the in-library NCCL path (comm.cu) -------------------------------------------------------------------------
// classification pass + intra-rank plan; the lists stay valid until the arrays grow
int lpic_mig_classify(lpic_ctx *c, int ispec, const i64 *) {
    MigArgs a;
    if (int r = make_args(c, ispec, a)) return r;
    Species &sp = c->spec[ispec];
    k_lists<<<(unsigned)c->g.npatch, T, 0, c->stream>>>(a);
    c->spec[ispec].sort.keys_written = a.kcache != nullptr;
    k_plan<<<div_up(c->g.npatch, 128), 128, 0, c->stream>>>(a);
    LAUNCHED(2);
    KERNEL_CHECK();
    sp.lists_valid = true;
    sp.lists_epoch = c->scratch_epoch;
    return 0;
}

int lpic_mig_pack(lpic_ctx *c, int ispec, int slot, i64 max_entry, double *dev_send) {
    HaloPlan *h = c->halo;
    MigArgs a;
    const unsigned long long epoch = c->scratch_epoch;
    if (int r = make_args(c, ispec, a)) return r;
    REQUIRE(c->spec[ispec].lists_valid && c->spec[ispec].lists_epoch == epoch, "lpic_mig_pack: the leaver lists are stale");
    c->spec[ispec].lists_epoch = c->scratch_epoch;
    dim3 grid(div_up(std::max<i64>(max_entry, 1), T), (unsigned)(h->send_first[slot + 1] - h->send_first[slot]));
    k_remote_pack<<<grid, T, 0, c->stream>>>(a, h->d_send_patch, h->d_send_b, h->d_mig_send_cnt, h->d_mig_send_poff,
                                             h->send_first[slot], dev_send);
    LAUNCHED(1);
    KERNEL_CHECK();
    return 0;
}

int lpic_mig_local_fill(lpic_ctx *c, int ispec, i64 max_local) {
    MigArgs a;
    const unsigned long long epoch = c->scratch_epoch;
    if (int r = make_args(c, ispec, a)) return r;
    REQUIRE(c->spec[ispec].lists_valid && c->spec[ispec].lists_epoch == epoch, "lpic_mig_local_fill: the lists are stale");
    c->spec[ispec].lists_epoch = c->scratch_epoch;
    a.remote_in = c->comm_remote_in;
    const int bpf = (int)div_up(max_local, T);
    if (bpf > 0) {
        k_fill<<<(unsigned)((i64)bpf * c->g.npatch), T, 0, c->stream>>>(a, bpf);
        LAUNCHED(1);
        KERNEL_CHECK();
    }
    return 0;
}

int lpic_mig_unpack_mark(lpic_ctx *c, int ispec, const double *const *dev_recv, i64 max_remote, const i64 *d_remote_in) {
    HaloPlan *h = c->halo;
    MigArgs a;
    const unsigned long long epoch = c->scratch_epoch;
    if (int r = make_args(c, ispec, a)) return r;
    Species &sp = c->spec[ispec];
    REQUIRE(sp.lists_valid && sp.lists_epoch == epoch, "lpic_mig_unpack_mark: the lists are stale");
    a.remote_in = d_remote_in;
    if (max_remote > 0) {
        PeerRecs bufs;
        for (int s = 0; s < h->npeers; s++) bufs.p[s] = dev_recv[s];
        const int bpp = (int)div_up(max_remote, T);
        k_remote_unpack<<<(unsigned)((i64)bpp * c->g.npatch), T, 0, c->stream>>>(a, h->d_recv_peer, h->d_mig_recv_cnt, h->d_mig_recv_poff,
                                                                               d_remote_in, bufs, bpp);
        LAUNCHED(1);
    }
    k_mark<<<(unsigned)c->g.npatch, T, 0, c->stream>>>(a);
    LAUNCHED(1);
    KERNEL_CHECK();
    sp.sort.valid = false;
    sp.lists_valid = false;
    sp.sort.keys_valid = sp.sort.keys_written && a.kcache != nullptr;  // the migration is complete: every alive slot carries its bucket key
    return 0;
}

extern "C" int lpic_migrate_count(lpic_ctx *c, int ispec, int64_t *to_extend, int64_t *incoming, int64_t *outgoing,
                                  int64_t *alive) {
    DeviceGuard dg(c);
    MigArgs a;
    if (int r = make_args(c, ispec, a)) return r;
    Species &sp = c->spec[ispec];
    const i64 n = c->g.npatch;
    k_lists<<<(unsigned)n, T, 0, c->stream>>>(a);  // counts + leaver / dead-slot lists in one classification pass
    c->spec[ispec].sort.keys_written = a.kcache != nullptr;
    LAUNCHED(1);
    sp.lists_valid = true;  // until the arrays grow (lpic_species_extend), anything else touches the slots ...
    sp.lists_epoch = c->scratch_epoch;  // ... or another operator uses the shared scratch lists
    k_plan<<<div_up(n, 128), 128, 0, c->stream>>>(a);
    LAUNCHED(1);
    KERNEL_CHECK();
    if (to_extend) CUDA_TRY(cudaMemcpyAsync(to_extend, sp.d_extend, sizeof(i64) * n, cudaMemcpyDeviceToHost, c->stream));
    std::vector<i64> h_in;
    if (!incoming) { h_in.resize(n); incoming = h_in.data(); }
    CUDA_TRY(cudaMemcpyAsync(incoming, sp.d_incoming, sizeof(i64) * n, cudaMemcpyDeviceToHost, c->stream));
    if (outgoing) CUDA_TRY(cudaMemcpyAsync(outgoing, sp.d_out, sizeof(i64) * n * c->g.nb, cudaMemcpyDeviceToHost, c->stream));
    if (alive) CUDA_TRY(cudaMemcpyAsync(alive, sp.d_alive, sizeof(i64) * n, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    sp.max_incoming = 0;  // sizes the grid of k_fill: one thread per newcomer, not per slot
    for (i64 p = 0; p < n; p++) sp.max_incoming = std::max<i64>(sp.max_incoming, incoming[p]);
    return 0;
}

extern "C" int lpic_migrate_fill(lpic_ctx *c, int ispec) {
    DeviceGuard dg(c);
    MigArgs a;
    const unsigned long long epoch = c->scratch_epoch;
    if (int r = make_args(c, ispec, a)) return r;
    Species &sp = c->spec[ispec];
    const i64 n = c->g.npatch;
    if (sp.max_npart == 0) return 0;
    if (!sp.lists_valid || sp.lists_epoch != epoch) {  // e.g. the host grew some patches after the count: the new (dead) slots must enter the lists
        k_lists<<<(unsigned)n, T, 0, c->stream>>>(a);
        c->spec[ispec].sort.keys_written = a.kcache != nullptr;
        LAUNCHED(1);
    }
    sp.lists_valid = false;
    const int bpp = (int)div_up(sp.max_npart, T);
    const int bpf = sp.max_incoming >= 0 ? (int)div_up(sp.max_incoming, T) : bpp;
    sp.max_incoming = -1;  // valid for one fill only (remote newcomers change d_incoming through other entry points)
    if (bpf > 0) {
        k_fill<<<(unsigned)((i64)bpf * n), T, 0, c->stream>>>(a, bpf);
        LAUNCHED(1);
    }
    k_mark<<<(unsigned)n, T, 0, c->stream>>>(a);
    LAUNCHED(1);
    KERNEL_CHECK();
    sp.sort.valid = false;
    sp.lists_valid = false;
    sp.sort.keys_valid = sp.sort.keys_written && a.kcache != nullptr;  // the migration is complete: every alive slot carries its bucket key
    return 0;
}
