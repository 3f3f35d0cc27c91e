// Inter-rank transport inside the library: NCCL point-to-point on a dedicated stream, ordered against the compute stream
// with CUDA events only -- no host synchronisation per exchange phase.
//
// Reference behaviour restated (not copied): core/mpi/mpi_manager.py:9-298 (one communicator per Simulation, a
// start()/wait() pair per exchange), core/mpi/sync_fields3d.c:713-996 (guard copy and current reduce, MPI_Isend/Irecv
// posted in start, completed in wait while the intra-rank copy runs in between: simulation/simulation.py:948-952) and
// core/mpi/sync_particles_3d.c:413-745 (counts first, then the particle payload).
//
//   lpic_halo_start   pack kernels (compute stream) -> event -> comm stream: ncclGroupStart, one ncclSend + one ncclRecv
//                     per peer GPU, ncclGroupEnd -> event.  Returns at once; the caller runs the intra-rank guard copy /
//                     current reduce on the compute stream meanwhile.
//   lpic_halo_wait    compute stream waits for the receive event, then ONE unpack kernel (copy for E/B, `+=` in the
//                     reference's boundary order for J/rho).
//   lpic_migrate_remote_start / _wait   the same pattern for particles: classification (shared with the intra-rank
//                     pass), per-entry counts exchanged device to device as int64, ONE host synchronisation to learn the
//                     message sizes (NCCL needs them on the host, as MPI does), payload exchange overlapped with the
//                     intra-rank fill.
// NCCL is opened with dlopen at lpic_comm_init: a process that already holds NCCL (torch.distributed) shares that copy,
// and the library keeps loading on hosts without NCCL (single-GPU use, CPU-side ABI tests).
#include <dlfcn.h>
#include <algorithm>
#include <vector>
#include "lpic_common.cuh"
#include "comm.cuh"

namespace {

// the part of nccl.h this file needs (stable since NCCL 2.7)
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclInt64 = 4, ncclFloat64 = 8 };
struct NcclApi {
    void *lib = nullptr;
    int (*GetUniqueId)(ncclUniqueId *) = nullptr;
    int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    int (*GetVersion)(int *) = nullptr;
};
NcclApi g_nccl;

int load_nccl() {
    if (g_nccl.lib) return 0;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    void *h = nullptr;
    for (const char *n : names)
        if ((h = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
    REQUIRE(h, "NCCL not found: %s", dlerror());
#define SYM(field, name)                                           \
    *(void **)(&g_nccl.field) = dlsym(h, name);                    \
    REQUIRE(g_nccl.field, "NCCL symbol %s missing", name)
    SYM(GetUniqueId, "ncclGetUniqueId"); SYM(CommInitRank, "ncclCommInitRank"); SYM(CommDestroy, "ncclCommDestroy");
    SYM(Send, "ncclSend"); SYM(Recv, "ncclRecv"); SYM(GroupStart, "ncclGroupStart"); SYM(GroupEnd, "ncclGroupEnd");
    SYM(GetErrorString, "ncclGetErrorString"); SYM(GetVersion, "ncclGetVersion");
#undef SYM
    g_nccl.lib = h;
    return 0;
}
#define NCCL_TRY(expr)                                                                                   \
    do {                                                                                                 \
        int _r = (expr);                                                                                 \
        if (_r != 0) {                                                                                   \
            lpic_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, g_nccl.GetErrorString(_r));     \
            return -1;                                                                                   \
        }                                                                                                \
    } while (0)

}  // namespace

struct CommState {
    ncclComm_t comm = nullptr;
    int rank = 0, nranks = 1;
    int peer_rank[LPIC_MAX_PEERS] = {0};
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_packed = nullptr, ev_recv = nullptr, ev_unpacked = nullptr;     // field halos
    cudaEvent_t ev_ppacked = nullptr, ev_precv = nullptr, ev_punpacked = nullptr;  // particles (may overlap a pending current reduce)
    // field halos: persistent staging, 4 attributes (the widest phase: jx jy jz rho) per peer
    double *fsend[LPIC_MAX_PEERS] = {nullptr}, *frecv[LPIC_MAX_PEERS] = {nullptr};
    uint32_t pend_mask = 0;
    int pend_reduce = 0;
    bool pending = false;
    // particles
    i64 *d_recv_cnt_ent = nullptr;                 // counts as received, entry order of the receive plan
    i64 *d_peer_tot = nullptr, *h_peer_tot = nullptr;  // [3][MAX_PEERS]: particles sent to / received from each peer, largest send entry (h: pinned)
    i64 *d_remote_in = nullptr;                    // per patch: arrivals from other ranks
    i64 *h_patch = nullptr;                        // pinned [3][npatch]: remote_in, local incoming, ndead
    double *psend[LPIC_MAX_PEERS] = {nullptr}, *precv[LPIC_MAX_PEERS] = {nullptr};
    i64 psend_cap[LPIC_MAX_PEERS] = {0}, precv_cap[LPIC_MAX_PEERS] = {0};
    bool mig_pending = false;
    bool plan_stale = false;  // lpic_halo_plan replaced the plan: lpic_comm_update has to re-size the staging first
    i64 mig_max_remote = 0;
    long long bytes_sent = 0;
};

void lpic_free_comm(lpic_ctx *c) {
    CommState *m = c->comm;
    if (!m) return;
    if (m->stream) cudaStreamSynchronize(m->stream);
    for (int s = 0; s < LPIC_MAX_PEERS; s++) { cudaFree(m->fsend[s]); cudaFree(m->frecv[s]); cudaFree(m->psend[s]); cudaFree(m->precv[s]); }
    cudaFree(m->d_recv_cnt_ent); cudaFree(m->d_peer_tot); cudaFree(m->d_remote_in);
    if (m->h_peer_tot) cudaFreeHost(m->h_peer_tot);
    if (m->h_patch) cudaFreeHost(m->h_patch);
    if (m->ev_packed) cudaEventDestroy(m->ev_packed);
    if (m->ev_recv) cudaEventDestroy(m->ev_recv);
    if (m->ev_unpacked) cudaEventDestroy(m->ev_unpacked);
    if (m->ev_ppacked) cudaEventDestroy(m->ev_ppacked);
    if (m->ev_precv) cudaEventDestroy(m->ev_precv);
    if (m->ev_punpacked) cudaEventDestroy(m->ev_punpacked);
    if (m->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(m->comm);
    if (m->stream) cudaStreamDestroy(m->stream);
    delete m;
    c->comm = nullptr;
}

void lpic_comm_plan_changed(lpic_ctx *c) {
    if (c->comm) c->comm->plan_stale = true;
}

static int alloc_plan_buffers(lpic_ctx *c, CommState *m, const int64_t *peer_rank) {
    HaloPlan *h = c->halo;
    if (m->stream) CUDA_TRY(cudaStreamSynchronize(m->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    for (int s = 0; s < LPIC_MAX_PEERS; s++) {
        cudaFree(m->fsend[s]); cudaFree(m->frecv[s]);
        m->fsend[s] = m->frecv[s] = nullptr;
    }
    cudaFree(m->d_recv_cnt_ent);
    m->d_recv_cnt_ent = nullptr;
    for (int s = 0; s < h->npeers; s++) {
        REQUIRE(peer_rank[s] >= 0 && peer_rank[s] < m->nranks && peer_rank[s] != m->rank, "bad peer rank in slot %d", s);
        m->peer_rank[s] = (int)peer_rank[s];
        CUDA_TRY(cudaMalloc(&m->fsend[s], sizeof(double) * 4 * std::max<i64>(h->send_words[s], 1)));
        CUDA_TRY(cudaMalloc(&m->frecv[s], sizeof(double) * 4 * std::max<i64>(h->recv_words[s], 1)));
    }
    CUDA_TRY(cudaMalloc(&m->d_recv_cnt_ent, sizeof(i64) * (h->nrecv_total + 1)));
    m->plan_stale = false;
    m->pending = m->mig_pending = false;
    return 0;
}

// After lpic_halo_plan replaced the plan of a context that already has a communicator (MovingWindow: the neighbour tables
// change with every shift): new peer table, staging re-sized; the NCCL communicator, stream and events are kept.
extern "C" int lpic_comm_update(lpic_ctx *c, const int64_t *peer_rank) {
    DeviceGuard dg(c);
    REQUIRE(c->comm && c->halo, "lpic_comm_update: no communicator / plan");
    return alloc_plan_buffers(c, c->comm, peer_rank);
}

extern "C" int lpic_comm_unique_id(void *id128) {
    if (int r = load_nccl()) return r;
    ncclUniqueId id;
    NCCL_TRY(g_nccl.GetUniqueId(&id));
    memcpy(id128, &id, sizeof(id));
    return 0;
}

extern "C" int lpic_comm_nccl_version(void) {
    if (load_nccl()) return -1;
    int v = 0;
    g_nccl.GetVersion(&v);
    return v;
}

// After lpic_halo_plan.  peer_rank[s] = rank of plan slot s.  Collective over all ranks of the communicator.
extern "C" int lpic_comm_init(lpic_ctx *c, const void *id128, int rank, int nranks, const int64_t *peer_rank) {
    DeviceGuard dg(c);
    HaloPlan *h = c->halo;
    REQUIRE(h, "lpic_comm_init needs the exchange plan (lpic_halo_plan) first");
    if (int r = load_nccl()) return r;
    lpic_free_comm(c);
    CommState *m = new CommState();
    c->comm = m;
    m->rank = rank; m->nranks = nranks;
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    NCCL_TRY(g_nccl.CommInitRank(&m->comm, nranks, id, rank));
    CUDA_TRY(cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreateWithFlags(&m->ev_packed, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&m->ev_recv, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&m->ev_unpacked, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&m->ev_ppacked, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&m->ev_precv, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&m->ev_punpacked, cudaEventDisableTiming));
    if (int r = alloc_plan_buffers(c, m, peer_rank)) return r;
    const i64 n = c->g.npatch;
    CUDA_TRY(cudaMalloc(&m->d_peer_tot, sizeof(i64) * 3 * LPIC_MAX_PEERS));
    CUDA_TRY(cudaMalloc(&m->d_remote_in, sizeof(i64) * n));
    CUDA_TRY(cudaMemset(m->d_remote_in, 0, sizeof(i64) * n));
    CUDA_TRY(cudaMallocHost(&m->h_peer_tot, sizeof(i64) * 3 * LPIC_MAX_PEERS));
    CUDA_TRY(cudaMallocHost(&m->h_patch, sizeof(i64) * 3 * n));
    // the first event wait of the comm stream needs a recorded event
    CUDA_TRY(cudaEventRecord(m->ev_unpacked, c->stream));
    CUDA_TRY(cudaEventRecord(m->ev_punpacked, c->stream));
    return 0;
}

extern "C" int64_t lpic_comm_bytes_sent(const lpic_ctx *c) { return c->comm ? c->comm->bytes_sent : 0; }

// ---- field halos -------------------------------------------------------------------------------------------------
int lpic_halo_pack_to(lpic_ctx *c, int slot, uint32_t mask, int reduce, double *dev_send);                    // halo.cu
int lpic_halo_unpack_from(lpic_ctx *c, uint32_t mask, int reduce, const double *const *dev_recv, cudaStream_t st);  // halo.cu

extern "C" int lpic_halo_start(lpic_ctx *c, uint32_t mask, int reduce) {
    DeviceGuard dg(c);
    CommState *m = c->comm;
    HaloPlan *h = c->halo;
    REQUIRE(m && h, "lpic_halo_start: no communicator (lpic_comm_init)");
    REQUIRE(!m->plan_stale, "lpic_halo_start: the exchange plan changed, call lpic_comm_update first");
    REQUIRE(!m->pending, "lpic_halo_start: the previous exchange has not been waited for");
    int nattr = 0;
    for (int a = 0; a < LPIC_NFIELD; a++) nattr += (mask >> a) & 1;
    REQUIRE(nattr >= 1 && nattr <= 4, "lpic_halo_start moves 1..4 attributes per phase");
    for (int s = 0; s < h->npeers; s++)
        if (int r = lpic_halo_pack_to(c, s, mask, reduce, m->fsend[s])) return r;
    CUDA_TRY(cudaEventRecord(m->ev_packed, c->stream));
    CUDA_TRY(cudaStreamWaitEvent(m->stream, m->ev_packed, 0));
    // the receive buffers may still be read by the previous phase's unpack kernel on the compute stream
    CUDA_TRY(cudaStreamWaitEvent(m->stream, m->ev_unpacked, 0));
    NCCL_TRY(g_nccl.GroupStart());
    for (int s = 0; s < h->npeers; s++) {
        if (h->send_words[s] > 0) NCCL_TRY(g_nccl.Send(m->fsend[s], (size_t)h->send_words[s] * nattr, ncclFloat64, m->peer_rank[s], m->comm, m->stream));
        if (h->recv_words[s] > 0) NCCL_TRY(g_nccl.Recv(m->frecv[s], (size_t)h->recv_words[s] * nattr, ncclFloat64, m->peer_rank[s], m->comm, m->stream));
        m->bytes_sent += 8ll * h->send_words[s] * nattr;
    }
    NCCL_TRY(g_nccl.GroupEnd());
    CUDA_TRY(cudaEventRecord(m->ev_recv, m->stream));
    m->pend_mask = mask; m->pend_reduce = reduce; m->pending = true;
    return 0;
}

extern "C" int lpic_halo_wait(lpic_ctx *c) {
    DeviceGuard dg(c);
    CommState *m = c->comm;
    REQUIRE(m && m->pending, "lpic_halo_wait without lpic_halo_start");
    CUDA_TRY(cudaStreamWaitEvent(c->stream, m->ev_recv, 0));
    const double *ptrs[LPIC_MAX_PEERS];
    for (int s = 0; s < c->halo->npeers; s++) ptrs[s] = m->frecv[s];
    if (int r = lpic_halo_unpack_from(c, m->pend_mask, m->pend_reduce, ptrs, c->stream)) return r;
    CUDA_TRY(cudaEventRecord(m->ev_unpacked, c->stream));
    m->pending = false;
    return 0;
}

// ---- particles ---------------------------------------------------------------------------------------------------
namespace {

// one thread per peer: counts of this peer's send entries taken from the classification (out[patch][boundary]), their
// exclusive offsets inside the peer's payload, and the peer's total
__global__ void k_send_counts(int npeers, int nb, const i64 *__restrict__ out, const int *__restrict__ ent_patch,
                              const int *__restrict__ ent_b, PeerRanges rg, i64 *__restrict__ cnt, i64 *__restrict__ poff,
                              i64 *__restrict__ peer_tot) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= npeers) return;
    i64 run = 0, mx = 0;
    for (i64 e = rg.send_first[s]; e < rg.send_first[s + 1]; e++) {
        const i64 n = out[(size_t)ent_patch[e] * nb + ent_b[e]];
        cnt[e] = n;
        poff[e] = run;
        run += n;
        mx = n > mx ? n : mx;
    }
    peer_tot[s] = run;
    peer_tot[2 * LPIC_MAX_PEERS + s] = mx;
}

// one thread per peer: received per-entry counts -> (patch, boundary)-keyed counts and payload offsets, peer totals;
// remote_in[patch] accumulated with atomics (a patch may receive from several peers)
__global__ void k_recv_counts(int npeers, int nb, const i64 *__restrict__ cnt_ent, const int *__restrict__ ent_patch,
                              const int *__restrict__ ent_b, PeerRanges rg, i64 *__restrict__ rcnt, i64 *__restrict__ rpoff,
                              i64 *__restrict__ remote_in, i64 *__restrict__ peer_tot) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= npeers) return;
    i64 run = 0;
    for (i64 e = rg.recv_first[s]; e < rg.recv_first[s + 1]; e++) {
        const size_t key = (size_t)ent_patch[e] * nb + ent_b[e];
        const i64 n = cnt_ent[e];
        rcnt[key] = n;
        rpoff[key] = run;
        run += n;
        if (n) atomicAdd((unsigned long long *)&remote_in[ent_patch[e]], (unsigned long long)n);
    }
    peer_tot[LPIC_MAX_PEERS + s] = run;
}

int grow(double **buf, i64 *cap, i64 words) {
    if (words <= *cap) return 0;
    cudaFree(*buf);
    *buf = nullptr; *cap = 0;
    const i64 want = words + words / 2 + 4096;
    CUDA_TRY(cudaMalloc(buf, sizeof(double) * want));
    *cap = want;
    return 0;
}

}  // namespace

// migrate.cu
int lpic_mig_classify(lpic_ctx *c, int ispec, const i64 *d_remote_in);
int lpic_mig_pack(lpic_ctx *c, int ispec, int slot, i64 max_entry, double *dev_send);
int lpic_mig_local_fill(lpic_ctx *c, int ispec, i64 max_local);
int lpic_mig_unpack_mark(lpic_ctx *c, int ispec, const double *const *dev_recv, i64 max_remote, const i64 *d_remote_in);

// Phase 1 (resume = 0): classify, exchange the counts, decide whether any patch has to grow.  to_extend[p] > 0 somewhere:
// the caller grows the arrays (lpic_species_extend) and calls again with resume = 1.  Otherwise (and always with resume = 1)
// the payload exchange is started and the intra-rank fill is queued behind the classification.
// info[0..3] = particles sent to / received from other ranks, largest per-patch remote / local arrival count.
extern "C" int lpic_migrate_remote_start(lpic_ctx *c, int ispec, int resume, int64_t *to_extend, int64_t *info) {
    DeviceGuard dg(c);
    CommState *m = c->comm;
    HaloPlan *h = c->halo;
    REQUIRE(m && h, "lpic_migrate_remote_start: no communicator (lpic_comm_init)");
    REQUIRE(!m->plan_stale, "lpic_migrate_remote_start: the exchange plan changed, call lpic_comm_update first");
    REQUIRE(ispec >= 0 && ispec < c->nspec && c->spec[ispec].allocated, "species %d not allocated", ispec);
    REQUIRE(!m->mig_pending, "lpic_migrate_remote_start: the previous particle exchange has not been waited for");
    Species &sp = c->spec[ispec];
    const Geom &g = c->g;
    const i64 n = g.npatch;
    PeerRanges rg;
    for (int s = 0; s <= h->npeers; s++) { rg.send_first[s] = h->send_first[s]; rg.recv_first[s] = h->recv_first[s]; }
    const unsigned pgrid = div_up(std::max(h->npeers, 1), 32);
    const int nw = lpic_particle_record_words(c, ispec);  // (before the classification: it touches the shared scratch epoch)
    if (!resume) {
        CUDA_TRY(cudaMemsetAsync(m->d_remote_in, 0, sizeof(i64) * n, c->stream));
        if (int r = lpic_mig_classify(c, ispec, nullptr)) return r;  // lists + out + ndead + local incoming (k_plan)
        k_send_counts<<<pgrid, 32, 0, c->stream>>>(h->npeers, g.nb, sp.d_out, h->d_send_patch, h->d_send_b, rg, h->d_mig_send_cnt,
                                                   h->d_mig_send_poff, m->d_peer_tot);
        LAUNCHED(1);
        CUDA_TRY(cudaEventRecord(m->ev_ppacked, c->stream));
        CUDA_TRY(cudaStreamWaitEvent(m->stream, m->ev_ppacked, 0));
        NCCL_TRY(g_nccl.GroupStart());
        for (int s = 0; s < h->npeers; s++) {
            const i64 ns = h->send_first[s + 1] - h->send_first[s], nr = h->recv_first[s + 1] - h->recv_first[s];
            if (ns) NCCL_TRY(g_nccl.Send(h->d_mig_send_cnt + h->send_first[s], (size_t)ns, ncclInt64, m->peer_rank[s], m->comm, m->stream));
            if (nr) NCCL_TRY(g_nccl.Recv(m->d_recv_cnt_ent + h->recv_first[s], (size_t)nr, ncclInt64, m->peer_rank[s], m->comm, m->stream));
        }
        NCCL_TRY(g_nccl.GroupEnd());
        CUDA_TRY(cudaEventRecord(m->ev_precv, m->stream));
        CUDA_TRY(cudaStreamWaitEvent(c->stream, m->ev_precv, 0));
        k_recv_counts<<<pgrid, 32, 0, c->stream>>>(h->npeers, g.nb, m->d_recv_cnt_ent, h->d_recv_patch, h->d_recv_b, rg,
                                                   h->d_mig_recv_cnt, h->d_mig_recv_poff, m->d_remote_in, m->d_peer_tot);
        LAUNCHED(1);
        KERNEL_CHECK();
        CUDA_TRY(cudaMemcpyAsync(m->h_peer_tot, m->d_peer_tot, sizeof(i64) * 3 * LPIC_MAX_PEERS, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaMemcpyAsync(m->h_patch, m->d_remote_in, sizeof(i64) * n, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaMemcpyAsync(m->h_patch + n, sp.d_incoming, sizeof(i64) * n, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaMemcpyAsync(m->h_patch + 2 * n, sp.d_ndead, sizeof(i64) * n, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));  // the ONE host synchronisation: NCCL needs the payload sizes
        bool any = false;
        for (i64 p = 0; p < n; p++) {
            const i64 in = m->h_patch[p] + m->h_patch[n + p], nd = m->h_patch[2 * n + p];
            // one growth rule for both kinds of arrival (core/patch/sync_particles_3d.c:468-473)
            const i64 ext = in - nd > 0 ? in - nd + (i64)((double)sp.h_npart[p] * 0.25) : 0;
            if (to_extend) to_extend[p] = ext;
            any = any || ext > 0;
        }
        if (any) {
            REQUIRE(to_extend, "a patch has to grow but the caller passed no to_extend array");
            return 1;  // grow, then call again with resume = 1
        }
    } else {
        // the arrays grew: the lists have to contain the new (dead) slots; counts and offsets are unchanged
        if (int r = lpic_mig_classify(c, ispec, nullptr)) return r;
    }
    i64 sent = 0, recvd = 0, max_remote = 0, max_local = 0;
    for (i64 p = 0; p < n; p++) { max_remote = std::max(max_remote, m->h_patch[p]); max_local = std::max(max_local, m->h_patch[n + p]); }
    for (int s = 0; s < h->npeers; s++) {
        const i64 ns = m->h_peer_tot[s], nr = m->h_peer_tot[LPIC_MAX_PEERS + s];
        sent += ns; recvd += nr;
        if (grow(&m->psend[s], &m->psend_cap[s], ns * nw) || grow(&m->precv[s], &m->precv_cap[s], nr * nw)) return -1;
        if (ns) {
            if (int r = lpic_mig_pack(c, ispec, s, m->h_peer_tot[2 * LPIC_MAX_PEERS + s], m->psend[s])) return r;
        }
    }
    CUDA_TRY(cudaEventRecord(m->ev_ppacked, c->stream));
    CUDA_TRY(cudaStreamWaitEvent(m->stream, m->ev_ppacked, 0));
    CUDA_TRY(cudaStreamWaitEvent(m->stream, m->ev_punpacked, 0));
    NCCL_TRY(g_nccl.GroupStart());
    for (int s = 0; s < h->npeers; s++) {
        const i64 ns = m->h_peer_tot[s], nr = m->h_peer_tot[LPIC_MAX_PEERS + s];
        if (ns) NCCL_TRY(g_nccl.Send(m->psend[s], (size_t)ns * nw, ncclFloat64, m->peer_rank[s], m->comm, m->stream));
        if (nr) NCCL_TRY(g_nccl.Recv(m->precv[s], (size_t)nr * nw, ncclFloat64, m->peer_rank[s], m->comm, m->stream));
        m->bytes_sent += 8ll * ns * nw;
    }
    NCCL_TRY(g_nccl.GroupEnd());
    CUDA_TRY(cudaEventRecord(m->ev_precv, m->stream));
    // intra-rank arrivals go behind the remote ones in the dead-slot list; they copy while the payload is on the wire
    c->comm_remote_in = m->d_remote_in;
    if (int r = lpic_mig_local_fill(c, ispec, max_local)) return r;
    m->mig_pending = true;
    m->mig_max_remote = max_remote;
    if (info) { info[0] = sent; info[1] = recvd; info[2] = max_remote; info[3] = max_local; }
    return 0;
}

extern "C" int lpic_migrate_remote_wait(lpic_ctx *c, int ispec) {
    DeviceGuard dg(c);
    CommState *m = c->comm;
    REQUIRE(m && m->mig_pending, "lpic_migrate_remote_wait without lpic_migrate_remote_start");
    CUDA_TRY(cudaStreamWaitEvent(c->stream, m->ev_precv, 0));
    const double *ptrs[LPIC_MAX_PEERS];
    for (int s = 0; s < c->halo->npeers; s++) ptrs[s] = m->precv[s];
    if (int r = lpic_mig_unpack_mark(c, ispec, ptrs, m->mig_max_remote, m->d_remote_in)) return r;
    CUDA_TRY(cudaEventRecord(m->ev_punpacked, c->stream));
    m->mig_pending = false;
    return 0;
}
