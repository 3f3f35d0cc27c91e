// Inter-rank guard-cell exchange: pack / unpack kernels around an NCCL send/recv issued by the host.
//
// Reference behaviour restated (not copied): core/mpi/sync_fields3d.c:883-996 (E/B guards: MPI subarray types, one
// message per (patch, boundary, attribute)) and :713-866 (J/rho: packed 4 x region buffers, `+=` on receive);
// 2D twins core/mpi/sync_fields2d.c.  B200 design: ONE contiguous staging buffer per peer GPU and phase, filled
// by a single pack kernel, so a phase costs <= (peers) NVLink messages instead of npatch x 26 x nattr.
// The regions and their element order are exactly those of the intra-rank kernels in fields.cu.
#include <vector>
#include "lpic_common.cuh"

namespace {

struct Region {
    int len[3];
};
__host__ __device__ inline Region region_of(const Geom &g, int b) {
    Region r;
    const int n[3] = {g.nx, g.ny, g.nz}, ngs[3] = {g.ng, g.ng, g.ngz};
    for (int a = 0; a < 3; a++) r.len[a] = dir_component(g.dim, b, a) == 0 ? n[a] : ngs[a];
    return r;
}
__host__ __device__ inline i64 region_words(const Geom &g, int b) {
    const Region r = region_of(g, b);
    return (i64)r.len[0] * r.len[1] * r.len[2];
}
// first logical index of the strip along one axis.  role 0: interior strip next to side s (guard-copy source);
// 1: guard strip on side s (guard-copy destination / current-reduce source); 2: interior strip (reduce destination)
__device__ __forceinline__ int strip_start(int s, int n, int ng, int role) {
    if (s == 0) return 0;
    if (role == 1) return s < 0 ? -ng : n;
    return s < 0 ? 0 : n - ng;
}
__device__ __forceinline__ int sidx(const Geom &g, int i, int j, int k) {
    return wrapneg(k, g.NZ) + g.NZ * (wrapneg(j, g.NY) + g.NY * wrapneg(i, g.NX));
}

struct AttrList {
    int n;
    int a[LPIC_NFIELD];
};
struct PeerBufs {
    const double *p[LPIC_MAX_PEERS];
    i64 words[LPIC_MAX_PEERS];
};

// grid: (chunks of the largest region, entries of this peer)
__global__ void __launch_bounds__(256) k_halo_pack(Geom g, double *__restrict__ F, const int *__restrict__ ent_patch,
                                                   const int *__restrict__ ent_b, const i64 *__restrict__ ent_woff,
                                                   i64 first, i64 words_per_attr, AttrList attrs, int reduce,
                                                   double *__restrict__ buf) {
    const i64 e = first + blockIdx.y;
    const int p = ent_patch[e], b = ent_b[e];
    const Region r = region_of(g, b);
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= r.len[0] * r.len[1] * r.len[2]) return;
    const int k = w % r.len[2], j = (w / r.len[2]) % r.len[1], i = w / (r.len[2] * r.len[1]);
    const int role = reduce ? 1 : 0;
    const int li = strip_start(dir_component(g.dim, b, 0), g.nx, g.ng, role) + i;
    const int lj = strip_start(dir_component(g.dim, b, 1), g.ny, g.ng, role) + j;
    const int lk = strip_start(dir_component(g.dim, b, 2), g.nz, g.ngz, role) + k;
    const size_t cell = (size_t)p * g.ncell + sidx(g, li, lj, lk);
    for (int t = 0; t < attrs.n; t++) {
        double *src = F + (size_t)attrs.a[t] * g.npatch * g.ncell + cell;
        buf[(size_t)t * words_per_attr + ent_woff[e] + w] = *src;
        if (reduce) *src = 0.0;  // the consumer zeroes the guard it reduced (core/patch/sync_fields3d.c:124-125)
    }
}

__device__ __forceinline__ int logical(int s, int n, int ng) { return s < n + ng ? s : s - (n + 2 * ng); }

// guard copy on receive: one thread per (attribute, patch, padded cell), same cell -> boundary map as k_sync_guard
__global__ void __launch_bounds__(256) k_halo_unpack_copy(Geom g, double *__restrict__ F, const int *__restrict__ recv_peer,
                                                          const i64 *__restrict__ recv_woff, AttrList attrs, PeerBufs bufs) {
    const i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (i64)g.npatch * g.ncell) return;
    const int p = (int)(t / g.ncell);
    int r = (int)(t - (i64)p * g.ncell);
    const int sk = r % g.NZ;
    r /= g.NZ;
    const int sj = r % g.NY, si = r / g.NY;
    const int li = logical(si, g.nx, g.ng), lj = logical(sj, g.ny, g.ng), lk = logical(sk, g.nz, g.ngz);
    const int sx = li < 0 ? -1 : (li >= g.nx ? 1 : 0), sy = lj < 0 ? -1 : (lj >= g.ny ? 1 : 0),
              sz = lk < 0 ? -1 : (lk >= g.nz ? 1 : 0);
    if (sx == 0 && sy == 0 && sz == 0) return;
    const int b = dir_lookup(g.dim, sx, sy, sz);
    const int slot = recv_peer[(size_t)p * g.nb + b];
    if (slot < 0) return;
    const Region rg = region_of(g, b);
    const int i = li - strip_start(sx, g.nx, g.ng, 1), j = lj - strip_start(sy, g.ny, g.ng, 1),
              k = lk - strip_start(sz, g.nz, g.ngz, 1);
    const i64 w = recv_woff[(size_t)p * g.nb + b] + ((i64)i * rg.len[1] + j) * rg.len[2] + k;
    const int a = blockIdx.y;
    F[(size_t)attrs.a[a] * g.npatch * g.ncell + (size_t)p * g.ncell + sk + g.NZ * (sj + g.NY * si)] =
        bufs.p[slot][(size_t)a * bufs.words[slot] + w];
}

// current reduce on receive: one thread per (attribute, patch, interior cell); boundaries in enum order
__global__ void __launch_bounds__(256) k_halo_unpack_reduce(Geom g, double *__restrict__ F, const int *__restrict__ recv_peer,
                                                            const i64 *__restrict__ recv_woff, AttrList attrs, PeerBufs bufs) {
    const i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    const int per = g.nx * g.ny * g.nz;
    if (t >= (i64)per * g.npatch) return;
    const int p = (int)(t / per);
    int r = (int)(t - (i64)p * per);
    const int k = r % g.nz;
    r /= g.nz;
    const int j = r % g.ny, i = r / g.ny;
    const bool lox = i < g.ng, hix = i >= g.nx - g.ng, loy = j < g.ng, hiy = j >= g.ny - g.ng;
    const bool loz = g.dim == 3 && k < g.ng, hiz = g.dim == 3 && k >= g.nz - g.ng;
    if (!(lox || hix || loy || hiy || loz || hiz)) return;
    const int a = blockIdx.y;
    double *dst = F + (size_t)attrs.a[a] * g.npatch * g.ncell + (size_t)p * g.ncell + k + g.NZ * (j + g.NY * i);
    double acc = *dst;
    bool touched = false;
    for (int b = 0; b < g.nb; b++) {
        const int sx = dir_component(g.dim, b, 0), sy = dir_component(g.dim, b, 1), sz = dir_component(g.dim, b, 2);
        if ((sx < 0 && !lox) || (sx > 0 && !hix) || (sy < 0 && !loy) || (sy > 0 && !hiy) || (sz < 0 && !loz) || (sz > 0 && !hiz))
            continue;
        const int slot = recv_peer[(size_t)p * g.nb + b];
        if (slot < 0) continue;
        const Region rg = region_of(g, b);
        const int ri = i - strip_start(sx, g.nx, g.ng, 2), rj = j - strip_start(sy, g.ny, g.ng, 2),
                  rk = k - strip_start(sz, g.nz, g.ngz, 2);
        const i64 w = recv_woff[(size_t)p * g.nb + b] + ((i64)ri * rg.len[1] + rj) * rg.len[2] + rk;
        acc = __dadd_rn(acc, bufs.p[slot][(size_t)a * bufs.words[slot] + w]);
        touched = true;
    }
    if (touched) *dst = acc;
}

template <typename T>
int to_device(T **dst, const std::vector<T> &v) {
    CUDA_TRY(cudaMalloc(dst, sizeof(T) * (v.size() ? v.size() : 1)));
    if (!v.empty()) CUDA_TRY(cudaMemcpy(*dst, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice));
    return 0;
}

AttrList attr_list(uint32_t mask) {
    AttrList l;
    l.n = 0;
    for (int a = 0; a < LPIC_NFIELD; a++)
        if (mask & (1u << a)) l.a[l.n++] = a;
    return l;
}

}  // namespace

void lpic_free_comm(lpic_ctx *c);
void lpic_comm_plan_changed(lpic_ctx *c);
// the exchange plan alone (lpic_halo_plan replaces it after a MovingWindow shift; the NCCL communicator survives)
static void free_plan(lpic_ctx *c);
void lpic_free_peers(lpic_ctx *c) {
    lpic_free_comm(c);
    free_plan(c);
}
static void free_plan(lpic_ctx *c) {
    HaloPlan *h = c->halo;
    if (!h) return;
    delete[] h->h_send_patch; delete[] h->h_send_b; delete[] h->h_recv_patch; delete[] h->h_recv_b; delete[] h->h_mig_send_cnt;
    cudaFree(h->d_send_patch); cudaFree(h->d_send_b); cudaFree(h->d_recv_patch); cudaFree(h->d_recv_b); cudaFree(h->d_send_woff); cudaFree(h->d_recv_peer); cudaFree(h->d_recv_woff);
    cudaFree(h->d_mig_send_cnt); cudaFree(h->d_mig_send_poff); cudaFree(h->d_mig_recv_cnt); cudaFree(h->d_mig_recv_poff);
    cudaFree(h->d_mig_incoming);
    delete h;
    c->halo = nullptr;
}

extern "C" int lpic_halo_plan(lpic_ctx *c, int npeers, const int64_t *nsend, const int64_t *send_patch, const int64_t *send_b,
                              const int64_t *nrecv, const int64_t *recv_patch, const int64_t *recv_b) {
    DeviceGuard dg(c);
    REQUIRE(npeers >= 0 && npeers <= LPIC_MAX_PEERS, "at most %d peer ranks are supported", LPIC_MAX_PEERS);
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    free_plan(c);
    lpic_comm_plan_changed(c);  // an existing communicator keeps its NCCL state; its staging is re-sized by lpic_comm_update
    const Geom &g = c->g;
    HaloPlan *h = new HaloPlan();
    c->halo = h;
    h->npeers = npeers;
    for (int s = 0; s < npeers; s++) {
        h->send_first[s + 1] = h->send_first[s] + nsend[s];
        h->recv_first[s + 1] = h->recv_first[s] + nrecv[s];
    }
    h->nsend_total = h->send_first[npeers];
    h->nrecv_total = h->recv_first[npeers];
    h->h_send_patch = new int[h->nsend_total + 1]; h->h_send_b = new int[h->nsend_total + 1];
    h->h_recv_patch = new int[h->nrecv_total + 1]; h->h_recv_b = new int[h->nrecv_total + 1];
    h->h_mig_send_cnt = new i64[h->nsend_total + 1];
    std::vector<int> sp(h->nsend_total), sb(h->nsend_total), rpeer((size_t)g.npatch * g.nb, -1), rp(h->nrecv_total), rb(h->nrecv_total);
    std::vector<i64> swoff(h->nsend_total), rwoff((size_t)g.npatch * g.nb, 0);
    for (int s = 0; s < npeers; s++) {
        i64 w = 0;
        for (i64 e = h->send_first[s]; e < h->send_first[s + 1]; e++) {
            REQUIRE(send_patch[e] >= 0 && send_patch[e] < g.npatch && send_b[e] >= 0 && send_b[e] < g.nb, "bad send entry");
            sp[e] = h->h_send_patch[e] = (int)send_patch[e];
            sb[e] = h->h_send_b[e] = (int)send_b[e];
            swoff[e] = w;
            w += region_words(g, (int)send_b[e]);
        }
        h->send_words[s] = w;
        w = 0;
        for (i64 e = h->recv_first[s]; e < h->recv_first[s + 1]; e++) {
            REQUIRE(recv_patch[e] >= 0 && recv_patch[e] < g.npatch && recv_b[e] >= 0 && recv_b[e] < g.nb, "bad recv entry");
            rp[e] = h->h_recv_patch[e] = (int)recv_patch[e];
            rb[e] = h->h_recv_b[e] = (int)recv_b[e];
            const size_t key = (size_t)recv_patch[e] * g.nb + recv_b[e];
            REQUIRE(rpeer[key] < 0, "boundary listed twice in the receive plan");
            rpeer[key] = s;
            rwoff[key] = w;
            w += region_words(g, (int)recv_b[e]);
        }
        h->recv_words[s] = w;
    }
    if (to_device(&h->d_send_patch, sp) || to_device(&h->d_send_b, sb) || to_device(&h->d_send_woff, swoff) ||
        to_device(&h->d_recv_patch, rp) || to_device(&h->d_recv_b, rb) ||
        to_device(&h->d_recv_peer, rpeer) || to_device(&h->d_recv_woff, rwoff))
        return -1;
    CUDA_TRY(cudaMalloc(&h->d_mig_send_cnt, sizeof(i64) * (h->nsend_total + 1)));
    CUDA_TRY(cudaMalloc(&h->d_mig_send_poff, sizeof(i64) * (h->nsend_total + 1)));
    CUDA_TRY(cudaMalloc(&h->d_mig_recv_cnt, sizeof(i64) * g.npatch * g.nb));
    CUDA_TRY(cudaMalloc(&h->d_mig_recv_poff, sizeof(i64) * g.npatch * g.nb));
    CUDA_TRY(cudaMalloc(&h->d_mig_incoming, sizeof(i64) * g.npatch));
    return 0;
}

extern "C" int64_t lpic_halo_words(lpic_ctx *c, int slot, int recv) {
    DeviceGuard dg(c);
    if (!c->halo || slot < 0 || slot >= c->halo->npeers) return -1;
    return recv ? c->halo->recv_words[slot] : c->halo->send_words[slot];
}

int lpic_halo_pack_to(lpic_ctx *c, int slot, uint32_t mask, int reduce, double *dev_send);
extern "C" int lpic_halo_pack(lpic_ctx *c, int slot, uint32_t mask, int reduce, double *dev_send) {
    DeviceGuard dg(c);
    return lpic_halo_pack_to(c, slot, mask, reduce, dev_send);
}
int lpic_halo_pack_to(lpic_ctx *c, int slot, uint32_t mask, int reduce, double *dev_send) {
    HaloPlan *h = c->halo;
    REQUIRE(h && slot >= 0 && slot < h->npeers, "no exchange plan / bad peer slot");
    const i64 ne = h->send_first[slot + 1] - h->send_first[slot];
    if (ne == 0) return 0;
    const Geom &g = c->g;
    i64 mx = 0;
    for (int b = 0; b < g.nb; b++) mx = std::max(mx, region_words(g, b));
    dim3 grid(div_up(mx, 256), (unsigned)ne);
    k_halo_pack<<<grid, 256, 0, c->stream>>>(g, c->fields, h->d_send_patch, h->d_send_b, h->d_send_woff, h->send_first[slot],
                                             h->send_words[slot], attr_list(mask), reduce, dev_send);
    LAUNCHED(1);
    KERNEL_CHECK();
    return 0;
}

int lpic_halo_unpack_from(lpic_ctx *c, uint32_t mask, int reduce, const double *const *dev_recv, cudaStream_t st);
extern "C" int lpic_halo_unpack(lpic_ctx *c, uint32_t mask, int reduce, const double *const *dev_recv) {
    DeviceGuard dg(c);
    return lpic_halo_unpack_from(c, mask, reduce, dev_recv, c->stream);
}
int lpic_halo_unpack_from(lpic_ctx *c, uint32_t mask, int reduce, const double *const *dev_recv, cudaStream_t st) {
    HaloPlan *h = c->halo;
    REQUIRE(h, "no exchange plan");
    if (h->nrecv_total == 0) return 0;
    const Geom &g = c->g;
    PeerBufs bufs;
    for (int s = 0; s < h->npeers; s++) { bufs.p[s] = dev_recv[s]; bufs.words[s] = h->recv_words[s]; }
    const AttrList attrs = attr_list(mask);
    if (reduce) {
        dim3 grid(div_up((i64)g.npatch * g.nx * g.ny * g.nz, 256), attrs.n);
        k_halo_unpack_reduce<<<grid, 256, 0, st>>>(g, c->fields, h->d_recv_peer, h->d_recv_woff, attrs, bufs);
    } else {
        dim3 grid(div_up((i64)g.npatch * g.ncell, 256), attrs.n);
        k_halo_unpack_copy<<<grid, 256, 0, st>>>(g, c->fields, h->d_recv_peer, h->d_recv_woff, attrs, bufs);
    }
    LAUNCHED(1);
    KERNEL_CHECK();
    return 0;
}
