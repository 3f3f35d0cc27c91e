// Context, device arenas and host<->device mirrors of the C-ABI (include/lpic_b200.h).
#include <stdarg.h>
#include <atomic>
#include <mutex>
#include <stdlib.h>
#include <algorithm>
#include <vector>
#include "lpic_common.cuh"

static thread_local char g_err[1024] = "";
std::atomic<long long> g_lpic_launches{0};

void lpic_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

struct AttrIds {
    int a[LPIC_NFIELD];
};

extern "C" const char *lpic_last_error(void) { return g_err; }
extern "C" const char *lpic_version(void) { return "lpic_b200 0.1 sm_100a"; }

extern "C" int lpic_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

// pinned allocations are remembered so lpic_host_free knows how to release them
static std::vector<void *> g_pinned;
static std::mutex g_pinned_mutex;  // contexts may be driven from several threads

extern "C" void *lpic_host_alloc(int64_t bytes) {
    if (bytes <= 0) bytes = 8;
    void *p = nullptr;
    if (lpic_device_count() > 0 && cudaHostAlloc(&p, (size_t)bytes, cudaHostAllocDefault) == cudaSuccess) {
        std::lock_guard<std::mutex> lock(g_pinned_mutex);
        g_pinned.push_back(p);
        return p;
    }
    cudaGetLastError();
    return malloc((size_t)bytes);
}

extern "C" void lpic_host_free(void *p) {
    if (!p) return;
    std::unique_lock<std::mutex> lock(g_pinned_mutex);
    auto it = std::find(g_pinned.begin(), g_pinned.end(), p);
    if (it != g_pinned.end()) {
        g_pinned.erase(it);
        cudaFreeHost(p);
    } else {
        free(p);
    }
}

extern "C" lpic_ctx *lpic_create(int dim, int64_t npatch, int64_t nx, int64_t ny, int64_t nz, int64_t ng, double dx,
                                 double dy, double dz, int nspec, int device) {
    if (!(dim == 2 || dim == 3) || npatch <= 0 || nx <= 0 || ny <= 0 || ng < 0 || nspec < 0 || (dim == 3 && nz <= 0)) {
        lpic_set_error("lpic_create: bad arguments");
        return nullptr;
    }
    if (lpic_device_count() <= device) {
        lpic_set_error("lpic_create: CUDA device %d not available (this library has no CPU fallback)", device);
        return nullptr;
    }
    if (cudaSetDevice(device) != cudaSuccess) {
        lpic_set_error("cudaSetDevice(%d) failed", device);
        return nullptr;
    }
    lpic_ctx *c = new lpic_ctx();
    Geom &g = c->g;
    g.dim = dim; g.nb = dim == 3 ? 26 : 8;
    g.nx = (int)nx; g.ny = (int)ny; g.nz = dim == 3 ? (int)nz : 1; g.ng = (int)ng; g.ngz = dim == 3 ? (int)ng : 0;
    g.NX = g.nx + 2 * g.ng; g.NY = g.ny + 2 * g.ng; g.NZ = g.nz + 2 * g.ngz;
    g.ncell = g.NX * g.NY * g.NZ;
    g.npatch = (int)npatch;
    g.dx = dx; g.dy = dy; g.dz = dim == 3 ? dz : 1.0;
    c->device = device;
    c->nspec = nspec;
    c->spec = new Species[nspec > 0 ? nspec : 1];
    bool ok = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) == cudaSuccess;
    const size_t fbytes = sizeof(double) * LPIC_NFIELD * (size_t)npatch * g.ncell;
    ok = ok && cudaMalloc(&c->fields, fbytes) == cudaSuccess;
    ok = ok && cudaMemset(c->fields, 0, fbytes) == cudaSuccess;
    ok = ok && cudaMalloc(&c->d_x0, sizeof(double) * npatch) == cudaSuccess;
    ok = ok && cudaMalloc(&c->d_y0, sizeof(double) * npatch) == cudaSuccess;
    ok = ok && cudaMalloc(&c->d_z0, sizeof(double) * npatch) == cudaSuccess;
    ok = ok && cudaMalloc(&c->d_nbr, sizeof(i64) * npatch * g.nb) == cudaSuccess;
    ok = ok && cudaMalloc(&c->d_box, sizeof(double) * npatch * 6) == cudaSuccess;
    ok = ok && cudaMalloc(&c->d_sort_org, sizeof(double) * npatch * 3) == cudaSuccess;
    ok = ok && cudaMalloc(&c->d_tmp64, sizeof(i64) * (64 + 16 * npatch)) == cudaSuccess;
    ok = ok && cudaMalloc(&c->d_tmpf, sizeof(double) * 64) == cudaSuccess;
    if (!ok) {
        lpic_set_error("lpic_create: device allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
        lpic_destroy(c);
        return nullptr;
    }
    c->h_x0 = new double[npatch]; c->h_y0 = new double[npatch]; c->h_z0 = new double[npatch];
    c->h_nbr = new i64[npatch * g.nb];
    c->h_patch_index = new i64[npatch];
    for (i64 p = 0; p < npatch; p++) c->h_patch_index[p] = p;
    return c;
}

static void free_species(Species &sp) {
    for (int a = 0; a < LPIC_NPATTR; a++) {
        if (!(sp.rec && a < LPIC_NREC)) cudaFree(sp.attr[a]);  // attr[0..7] point into the record arena when there is one
        sp.attr[a] = nullptr;
    }
    cudaFree(sp.rec); sp.rec = nullptr;
    sp.pstride = 1;
    cudaFree(sp.dead); sp.dead = nullptr;
    cudaFree(sp.d_off); cudaFree(sp.d_npart); sp.d_off = sp.d_npart = nullptr;
    cudaFree(sp.sort.bucket_count); cudaFree(sp.sort.bound_min); cudaFree(sp.sort.bound_max); cudaFree(sp.sort.pidx);
    cudaFree(sp.sort.kcache); cudaFree(sp.sort.d_korg); delete[] sp.sort.h_korg;
    sp.sort = SortState();
    cudaFree(sp.d_out); cudaFree(sp.d_ndead); cudaFree(sp.d_incoming); cudaFree(sp.d_extend); cudaFree(sp.d_alive);
    sp.d_out = sp.d_ndead = sp.d_incoming = sp.d_extend = sp.d_alive = nullptr;
    delete[] sp.h_off; delete[] sp.h_pcap; delete[] sp.h_npart;
    sp.h_off = sp.h_pcap = sp.h_npart = nullptr;
    sp.allocated = false;
}

void lpic_free_peers(lpic_ctx *c);
void lpic_free_pml(lpic_ctx *c);
static void free_xfer(lpic_ctx *c);
int lpic_pml_zero_psi_of(lpic_ctx *c, i64 n, const int64_t *patches);  // fields.cu

extern "C" void lpic_destroy(lpic_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    for (int s = 0; s < c->nspec; s++) free_species(c->spec[s]);
    delete[] c->spec;
    lpic_free_peers(c);
    lpic_free_pml(c);
    free_xfer(c);
    cudaFree(c->fields); cudaFree(c->d_x0); cudaFree(c->d_y0); cudaFree(c->d_z0); cudaFree(c->d_nbr); cudaFree(c->d_box);
    cudaFree(c->scr_a); cudaFree(c->scr_b); cudaFree(c->scr_buf); cudaFree(c->d_sort_org); cudaFree(c->d_tmp64); cudaFree(c->d_tmpf); cudaFree(c->d_tile_start); cudaFree(c->d_slice); cudaFree(c->d_slice_k); cudaFree(c->d_laser_i); cudaFree(c->d_laser_s); cudaFree(c->d_sort_hist); cudaFree(c->d_ext_tab); cudaFree(c->d_ext_attrs);
    delete[] c->h_x0; delete[] c->h_y0; delete[] c->h_z0; delete[] c->h_nbr; delete[] c->h_patch_index;
    if (c->events) {
        for (int i = 0; i < 4096; i++)
            if (c->events[i]) cudaEventDestroy(c->events[i]);
        delete[] c->events;
    }
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

extern "C" int lpic_set_patch_geometry(lpic_ctx *c, const double *x0, const double *y0, const double *z0,
                                       const int64_t *nbr, const double *box, const double *glob, int64_t rank,
                                       const int64_t *patch_index) {
    DeviceGuard dg(c);
    const Geom &g = c->g;
    const i64 n = g.npatch;
    std::vector<double> zeros(n, 0.0);
    if (!z0) z0 = zeros.data();
    memcpy(c->h_x0, x0, sizeof(double) * n); memcpy(c->h_y0, y0, sizeof(double) * n); memcpy(c->h_z0, z0, sizeof(double) * n);
    memcpy(c->h_nbr, nbr, sizeof(i64) * n * g.nb);
    for (i64 i = 0; i < n * g.nb; i++) REQUIRE(nbr[i] < n, "neighbor_ipatch[%lld] = %lld out of range", (long long)i, (long long)nbr[i]);
    memcpy(c->glob, glob, sizeof(double) * 6);
    c->rank = rank;
    if (patch_index) memcpy(c->h_patch_index, patch_index, sizeof(i64) * n);
    CUDA_TRY(cudaMemcpyAsync(c->d_x0, c->h_x0, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(c->d_y0, c->h_y0, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(c->d_z0, c->h_z0, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(c->d_nbr, c->h_nbr, sizeof(i64) * n * g.nb, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(c->d_box, box, sizeof(double) * n * 6, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int lpic_sync(lpic_ctx *c) {
    DeviceGuard dg(c);
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return 0;
}
extern "C" void *lpic_stream(lpic_ctx *c) {
    DeviceGuard dg(c); return (void *)c->stream; }
extern "C" int64_t lpic_field_cells(const lpic_ctx *c) {
    DeviceGuard dg(c); return c->g.ncell; }

// ---- fields --------------------------------------------------------------------------------------------------
static int copy_fields(lpic_ctx *c, uint32_t mask, double *host, bool up) {
    const size_t n = (size_t)c->g.npatch * c->g.ncell;
    int a = 0;
    while (a < LPIC_NFIELD) {  // coalesce runs of consecutive attributes into one copy
        if (!(mask & (1u << a))) { a++; continue; }
        int b = a;
        while (b + 1 < LPIC_NFIELD && (mask & (1u << (b + 1)))) b++;
        const size_t bytes = sizeof(double) * n * (b - a + 1);
        if (up) CUDA_TRY(cudaMemcpyAsync(field_ptr(c, a), host + n * a, bytes, cudaMemcpyHostToDevice, c->stream));
        else CUDA_TRY(cudaMemcpyAsync(host + n * a, field_ptr(c, a), bytes, cudaMemcpyDeviceToHost, c->stream));
        a = b + 1;
    }
    if (!up) CUDA_TRY(cudaStreamSynchronize(c->stream));
    return 0;
}
extern "C" int lpic_upload_fields(lpic_ctx *c, uint32_t mask, const double *host) {
    DeviceGuard dg(c); return copy_fields(c, mask, (double *)host, true); }
extern "C" int lpic_download_fields(lpic_ctx *c, uint32_t mask, double *host) {
    DeviceGuard dg(c); return copy_fields(c, mask, host, false); }

extern "C" int lpic_upload_field_ptrs(lpic_ctx *c, int attr, const double *const *ptrs) {
    DeviceGuard dg(c);
    REQUIRE(attr >= 0 && attr < LPIC_NFIELD, "bad field attribute %d", attr);
    for (int p = 0; p < c->g.npatch; p++)
        CUDA_TRY(cudaMemcpyAsync(field_ptr(c, attr) + (size_t)p * c->g.ncell, ptrs[p], sizeof(double) * c->g.ncell,
                                 cudaMemcpyHostToDevice, c->stream));
    return 0;
}
extern "C" int lpic_download_field_ptrs(lpic_ctx *c, int attr, double *const *ptrs) {
    DeviceGuard dg(c);
    REQUIRE(attr >= 0 && attr < LPIC_NFIELD, "bad field attribute %d", attr);
    for (int p = 0; p < c->g.npatch; p++)
        CUDA_TRY(cudaMemcpyAsync(ptrs[p], field_ptr(c, attr) + (size_t)p * c->g.ncell, sizeof(double) * c->g.ncell,
                                 cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return 0;
}

// ---- sliced download: one z-plane of the interior per patch (callback/utils.py:125-230 get_fields_3d slices on the host
// after every rank has copied whole patches; here only the plane crosses PCIe) -------------------------------------------
__global__ void __launch_bounds__(256) k_gather_zplane(Geom g, const double *__restrict__ F, const int *__restrict__ kz, int nattr,
                                                       AttrIds attrs, double *__restrict__ out) {
    const int plane = g.nx * g.ny;
    const i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (i64)g.npatch * plane) return;
    const int p = (int)(t / plane), r = (int)(t - (i64)p * plane);
    const int i = r / g.ny, j = r - i * g.ny, k = kz[p];
    if (k < 0) return;  // this patch does not contain the plane
    const size_t cell = (size_t)p * g.ncell + (size_t)(i * g.NY + j) * g.NZ + k;
    for (int a = 0; a < nattr; a++)
        out[((size_t)a * g.npatch + p) * plane + r] = F[(size_t)attrs.a[a] * g.npatch * g.ncell + cell];
}
extern "C" int lpic_download_field_slice(lpic_ctx *c, uint32_t mask, const int64_t *kz, double *host) {
    DeviceGuard dg(c);
    const Geom &g = c->g;
    REQUIRE(g.dim == 3, "lpic_download_field_slice is for 3D grids (2D fields are their own slice)");
    AttrIds attrs;
    int na = 0;
    for (int a = 0; a < LPIC_NFIELD; a++)
        if (mask & (1u << a)) attrs.a[na++] = a;
    if (!na) return 0;
    const size_t plane = (size_t)g.nx * g.ny, words = (size_t)na * g.npatch * plane;
    if (words > c->slice_cap) {
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        cudaFree(c->d_slice); cudaFree(c->d_slice_k); cudaFree(c->d_laser_i); cudaFree(c->d_laser_s); cudaFree(c->d_sort_hist); cudaFree(c->d_ext_tab); cudaFree(c->d_ext_attrs);
        c->d_slice = nullptr; c->d_slice_k = nullptr; c->slice_cap = 0;
        CUDA_TRY(cudaMalloc(&c->d_slice, sizeof(double) * (size_t)LPIC_NFIELD * g.npatch * plane));
        CUDA_TRY(cudaMalloc(&c->d_slice_k, sizeof(int) * g.npatch));
        c->slice_cap = (size_t)LPIC_NFIELD * g.npatch * plane;
    }
    std::vector<int> hk(g.npatch);
    for (int p = 0; p < g.npatch; p++) {
        REQUIRE(kz[p] < g.nz, "slice index %lld outside patch %d", (long long)kz[p], p);
        hk[p] = kz[p] < 0 ? -1 : (int)kz[p];
    }
    CUDA_TRY(cudaMemcpyAsync(c->d_slice_k, hk.data(), sizeof(int) * g.npatch, cudaMemcpyHostToDevice, c->stream));
    k_gather_zplane<<<div_up((i64)g.npatch * plane, 256), 256, 0, c->stream>>>(g, c->fields, c->d_slice_k, na, attrs, c->d_slice);
    LAUNCHED(1);
    KERNEL_CHECK();
    CUDA_TRY(cudaMemcpyAsync(host, c->d_slice, sizeof(double) * words, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return 0;
}

// ---- particles -----------------------------------------------------------------------------------------------
static bool attr_resident(const Species &sp, int a) {
    return a < LPIC_P_EX_PART || a > LPIC_P_BZ_PART || sp.with_part;
}

static i64 phys_cap(i64 npart, double slack, i64 min_extra) {
    i64 c = std::max((i64)(npart * slack), npart + min_extra);
    return ((c + 31) / 32) * 32;
}

static int upload_layout(lpic_ctx *c, Species &sp) {
    const i64 n = c->g.npatch;
    sp.max_npart = 0;
    for (i64 p = 0; p < n; p++) sp.max_npart = std::max(sp.max_npart, sp.h_npart[p]);
    CUDA_TRY(cudaMemcpyAsync(sp.d_off, sp.h_off, sizeof(i64) * n, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(sp.d_npart, sp.h_npart, sizeof(i64) * n, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));  // h_* may be rewritten by the caller right after
    return 0;
}

extern "C" int lpic_species_alloc(lpic_ctx *c, int ispec, const int64_t *npart, double slack, int64_t min_extra,
                                  int with_part) {
    DeviceGuard dg(c);
    REQUIRE(ispec >= 0 && ispec < c->nspec, "bad species %d", ispec);
    Species &sp = c->spec[ispec];
    const i64 n = c->g.npatch;
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    free_species(sp);
    sp.with_part = with_part != 0;
    sp.h_off = new i64[n]; sp.h_pcap = new i64[n]; sp.h_npart = new i64[n];
    i64 total = 0;
    for (i64 p = 0; p < n; p++) {
        REQUIRE(npart[p] >= 0, "negative npart");
        sp.h_npart[p] = npart[p];
        sp.h_pcap[p] = phys_cap(npart[p], slack < 1.0 ? 1.0 : slack, min_extra < 0 ? 0 : min_extra);
        sp.h_off[p] = total;
        total += sp.h_pcap[p];
    }
    sp.total = total;
    const size_t slots = (size_t)std::max<i64>(total, 1);
    const char *layout = getenv("LPIC_PARTICLE_LAYOUT");  // "soa": eight separate arrays (A/B runs); default: 64-byte records
    if (!(layout && strcmp(layout, "soa") == 0)) {
        CUDA_TRY(cudaMalloc(&sp.rec, sizeof(double) * LPIC_NREC * slots));
        CUDA_TRY(cudaMemsetAsync(sp.rec, 0, sizeof(double) * LPIC_NREC * slots, c->stream));
        sp.pstride = LPIC_NREC;
        for (int a = 0; a < LPIC_NREC; a++) sp.attr[a] = sp.rec + a;
    }
    for (int a = sp.rec ? LPIC_NREC : 0; a < LPIC_NPATTR; a++)
        if (attr_resident(sp, a)) {
            CUDA_TRY(cudaMalloc(&sp.attr[a], sizeof(double) * slots));
            CUDA_TRY(cudaMemsetAsync(sp.attr[a], 0, sizeof(double) * slots, c->stream));  // no stale bit patterns in unused slots
        }
    CUDA_TRY(cudaMalloc(&sp.dead, slots));
    CUDA_TRY(cudaMemsetAsync(sp.dead, 1, slots, c->stream));
    CUDA_TRY(cudaMalloc(&sp.sort.pidx, sizeof(int) * slots));  // particle_index starts at -1 (particle_sort.py:120-131)
    CUDA_TRY(cudaMemsetAsync(sp.sort.pidx, 0xff, sizeof(int) * slots, c->stream));
    sp.sort.pidx_cap = (i64)slots;
    CUDA_TRY(cudaMalloc(&sp.d_off, sizeof(i64) * n));
    CUDA_TRY(cudaMalloc(&sp.d_npart, sizeof(i64) * n));
    CUDA_TRY(cudaMalloc(&sp.d_out, sizeof(i64) * n * c->g.nb));
    CUDA_TRY(cudaMalloc(&sp.d_ndead, sizeof(i64) * n));
    CUDA_TRY(cudaMalloc(&sp.d_incoming, sizeof(i64) * n));
    CUDA_TRY(cudaMalloc(&sp.d_extend, sizeof(i64) * n));
    CUDA_TRY(cudaMalloc(&sp.d_alive, sizeof(i64) * n));
    sp.allocated = true;
    return upload_layout(c, sp);
}

extern "C" int lpic_species_layout(const lpic_ctx *c, int ispec, int64_t *off, int64_t *pcap, int64_t *npart, int64_t *total) {
    DeviceGuard dg(c);
    REQUIRE(ispec >= 0 && ispec < c->nspec && c->spec[ispec].allocated, "species %d not allocated", ispec);
    const Species &sp = c->spec[ispec];
    const i64 n = c->g.npatch;
    if (off) memcpy(off, sp.h_off, sizeof(i64) * n);
    if (pcap) memcpy(pcap, sp.h_pcap, sizeof(i64) * n);
    if (npart) memcpy(npart, sp.h_npart, sizeof(i64) * n);
    if (total) *total = sp.total;
    return 0;
}

static int particle_array(lpic_ctx *c, int ispec, int attr, void **dev, size_t *esz) {
    REQUIRE(ispec >= 0 && ispec < c->nspec && c->spec[ispec].allocated, "species %d not allocated", ispec);
    Species &sp = c->spec[ispec];
    if (attr == LPIC_P_IS_DEAD) { *dev = sp.dead; *esz = 1; return 0; }
    REQUIRE(attr >= 0 && attr < LPIC_NPATTR, "bad particle attribute %d", attr);
    REQUIRE(attr_resident(sp, attr), "attribute %d is not resident (species allocated without *_part)", attr);
    *dev = sp.attr[attr]; *esz = sizeof(double);
    return 0;
}

// ---- host <-> device moves of the record attributes ------------------------------------------------------------------
// The host mirrors keep the reference's one-array-per-attribute layout; the device keeps 64-byte records.  Chunks of slots
// go through two staging buffers [8][chunk] on two auxiliary streams: the (un)packing kernel of one chunk runs under the
// PCIe copies of the other, and every copy is one contiguous cudaMemcpyAsync per attribute and chunk.
struct XferState {
    double *buf[2] = {nullptr, nullptr};
    i64 chunk = 0;
    cudaStream_t st[2] = {nullptr, nullptr};
    cudaEvent_t ev_begin = nullptr, ev_end[2] = {nullptr, nullptr};
};
static const i64 kXferChunkMax = 1ll << 21;  // slots per chunk: 2 x 128 MB of staging at most

__global__ void __launch_bounds__(256) k_rec_unpack(const double *__restrict__ rec, i64 first, i64 n, double *__restrict__ stage,
                                                    i64 chunk, unsigned mask) {
    const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double2 *r = reinterpret_cast<const double2 *>(rec + (first + i) * LPIC_NREC);
    const double2 a = r[0], b = r[1], cc = r[2], d = r[3];
    const double v[LPIC_NREC] = {a.x, a.y, b.x, b.y, cc.x, cc.y, d.x, d.y};
#pragma unroll
    for (int t = 0; t < LPIC_NREC; t++)
        if (mask >> t & 1u) stage[t * chunk + i] = v[t];
}
__global__ void __launch_bounds__(256) k_rec_pack(double *__restrict__ rec, i64 first, i64 n, const double *__restrict__ stage,
                                                  i64 chunk, unsigned mask) {
    const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double *r = rec + (first + i) * LPIC_NREC;
    if (mask == 0xffu) {
        double2 *r2 = reinterpret_cast<double2 *>(r);
        r2[0] = make_double2(stage[i], stage[chunk + i]);
        r2[1] = make_double2(stage[2 * chunk + i], stage[3 * chunk + i]);
        r2[2] = make_double2(stage[4 * chunk + i], stage[5 * chunk + i]);
        r2[3] = make_double2(stage[6 * chunk + i], stage[7 * chunk + i]);
        return;
    }
#pragma unroll
    for (int t = 0; t < LPIC_NREC; t++)
        if (mask >> t & 1u) r[t] = stage[t * chunk + i];
}

static void free_xfer(lpic_ctx *c) {
    XferState *x = c->xfer;
    if (!x) return;
    for (int i = 0; i < 2; i++) {
        cudaFree(x->buf[i]);
        if (x->st[i]) cudaStreamDestroy(x->st[i]);
        if (x->ev_end[i]) cudaEventDestroy(x->ev_end[i]);
    }
    if (x->ev_begin) cudaEventDestroy(x->ev_begin);
    delete x;
    c->xfer = nullptr;
}

// host[t] (t in mask) = base of an arena-layout array of attribute t; slots [first, first + nslots) move
static int xfer_records(lpic_ctx *c, Species &sp, unsigned mask, double *const *host, i64 first, i64 nslots, bool to_device) {
    if (nslots <= 0 || !(mask & 0xffu)) return 0;
    if (!c->xfer) c->xfer = new XferState();
    XferState *x = c->xfer;
    const i64 want = std::min<i64>(kXferChunkMax, ((nslots + 255) / 256) * 256);
    if (want > x->chunk) {
        CUDA_TRY(cudaDeviceSynchronize());
        for (int i = 0; i < 2; i++) {
            cudaFree(x->buf[i]);
            x->buf[i] = nullptr;
        }
        x->chunk = 0;
        for (int i = 0; i < 2; i++) CUDA_TRY(cudaMalloc(&x->buf[i], sizeof(double) * LPIC_NREC * want));
        x->chunk = want;
    }
    if (!x->st[0]) {
        for (int i = 0; i < 2; i++) {
            CUDA_TRY(cudaStreamCreateWithFlags(&x->st[i], cudaStreamNonBlocking));
            CUDA_TRY(cudaEventCreateWithFlags(&x->ev_end[i], cudaEventDisableTiming));
        }
        CUDA_TRY(cudaEventCreateWithFlags(&x->ev_begin, cudaEventDisableTiming));
    }
    CUDA_TRY(cudaEventRecord(x->ev_begin, c->stream));
    for (int i = 0; i < 2; i++) CUDA_TRY(cudaStreamWaitEvent(x->st[i], x->ev_begin, 0));
    const i64 ch = x->chunk;
    int k = 0;
    for (i64 done = 0; done < nslots; done += ch, k ^= 1) {
        const i64 n = std::min<i64>(ch, nslots - done), f = first + done;
        cudaStream_t st = x->st[k];
        double *B = x->buf[k];
        if (to_device) {
            for (int t = 0; t < LPIC_NREC; t++)
                if (mask >> t & 1u) CUDA_TRY(cudaMemcpyAsync(B + t * ch, host[t] + f, sizeof(double) * n, cudaMemcpyHostToDevice, st));
            k_rec_pack<<<div_up(n, 256), 256, 0, st>>>(sp.rec, f, n, B, ch, mask & 0xffu);
        } else {
            k_rec_unpack<<<div_up(n, 256), 256, 0, st>>>(sp.rec, f, n, B, ch, mask & 0xffu);
            for (int t = 0; t < LPIC_NREC; t++)
                if (mask >> t & 1u) CUDA_TRY(cudaMemcpyAsync(host[t] + f, B + t * ch, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
        }
        LAUNCHED(1);
    }
    KERNEL_CHECK();
    for (int i = 0; i < 2; i++) {
        CUDA_TRY(cudaEventRecord(x->ev_end[i], x->st[i]));
        CUDA_TRY(cudaStreamWaitEvent(c->stream, x->ev_end[i], 0));
    }
    return 0;
}
static int xfer_one(lpic_ctx *c, Species &sp, int attr, void *host, i64 first, i64 nslots, bool to_device) {
    double *tab[LPIC_NREC] = {nullptr};
    tab[attr] = (double *)host;
    return xfer_records(c, sp, 1u << attr, tab, first, nslots, to_device);
}
static bool in_record(const Species &sp, int attr) { return sp.rec && attr >= 0 && attr < LPIC_NREC; }

extern "C" int lpic_upload_particles(lpic_ctx *c, int ispec, int attr, const void *host) {
    DeviceGuard dg(c);
    void *dev; size_t esz;
    if (int r = particle_array(c, ispec, attr, &dev, &esz)) return r;
    Species &sp = c->spec[ispec];
    if (in_record(sp, attr)) { if (int r = xfer_one(c, sp, attr, (void *)host, 0, sp.total, true)) return r; }
    else CUDA_TRY(cudaMemcpyAsync(dev, host, esz * sp.total, cudaMemcpyHostToDevice, c->stream));
    sp.sort.valid = false;
    sp.lists_valid = false; sp.sort.keys_valid = sp.sort.keys_written = false;
    return 0;
}
// The record attributes named by mask (bit a = attribute a < 8) in one pass: host[a] = base of the arena-layout array of
// attribute a.  This is what Simulation.run's entry / exit copies use; the per-attribute calls above remain for the rest.
extern "C" int lpic_upload_particle_records(lpic_ctx *c, int ispec, uint32_t mask, const double *const *host) {
    DeviceGuard dg(c);
    REQUIRE(ispec >= 0 && ispec < c->nspec && c->spec[ispec].allocated, "species %d not allocated", ispec);
    Species &sp = c->spec[ispec];
    if (sp.rec) { if (int r = xfer_records(c, sp, mask, (double *const *)host, 0, sp.total, true)) return r; }
    else
        for (int t = 0; t < LPIC_NREC; t++)
            if (mask >> t & 1u) CUDA_TRY(cudaMemcpyAsync(sp.attr[t], host[t], sizeof(double) * sp.total, cudaMemcpyHostToDevice, c->stream));
    sp.sort.valid = false;
    sp.lists_valid = false; sp.sort.keys_valid = sp.sort.keys_written = false;
    return 0;
}
extern "C" int lpic_download_particle_records(lpic_ctx *c, int ispec, uint32_t mask, double *const *host) {
    DeviceGuard dg(c);
    REQUIRE(ispec >= 0 && ispec < c->nspec && c->spec[ispec].allocated, "species %d not allocated", ispec);
    Species &sp = c->spec[ispec];
    if (sp.rec) { if (int r = xfer_records(c, sp, mask, host, 0, sp.total, false)) return r; }
    else
        for (int t = 0; t < LPIC_NREC; t++)
            if (mask >> t & 1u) CUDA_TRY(cudaMemcpyAsync(host[t], sp.attr[t], sizeof(double) * sp.total, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return 0;
}
// One patch's slots [off[p], off[p] + npart[p]) of one attribute, from an arena with the device's layout (MovingWindow:
// only the recycled patches' particles cross PCIe)
extern "C" int lpic_upload_particles_patch(lpic_ctx *c, int ispec, int attr, int64_t patch, const void *host_arena) {
    DeviceGuard dg(c);
    void *dev; size_t esz;
    if (int r = particle_array(c, ispec, attr, &dev, &esz)) return r;
    Species &sp = c->spec[ispec];
    REQUIRE(patch >= 0 && patch < c->g.npatch, "bad patch %lld", (long long)patch);
    const size_t first = (size_t)sp.h_off[patch] * esz, bytes = (size_t)sp.h_npart[patch] * esz;
    if (in_record(sp, attr)) { if (int r = xfer_one(c, sp, attr, (void *)host_arena, sp.h_off[patch], sp.h_npart[patch], true)) return r; }
    else if (bytes) CUDA_TRY(cudaMemcpyAsync((char *)dev + first, (const char *)host_arena + first, bytes, cudaMemcpyHostToDevice, c->stream));
    sp.sort.valid = false;
    sp.lists_valid = false; sp.sort.keys_valid = sp.sort.keys_written = false;
    return 0;
}
// Vacuum fields (all ten attributes) and psi arrays of the listed patches, on the device (callback/utils.py:790-800)
extern "C" int lpic_zero_patches(lpic_ctx *c, int64_t n, const int64_t *patches) {
    DeviceGuard dg(c);
    const Geom &g = c->g;
    for (i64 i = 0; i < n; i++) {
        REQUIRE(patches[i] >= 0 && patches[i] < g.npatch, "bad patch in the list");
        for (int a = 0; a < LPIC_NFIELD; a++)
            CUDA_TRY(cudaMemsetAsync(field_ptr(c, a) + (size_t)patches[i] * g.ncell, 0, sizeof(double) * g.ncell, c->stream));
    }
    return lpic_pml_zero_psi_of(c, n, patches);
}
extern "C" int lpic_download_particles(lpic_ctx *c, int ispec, int attr, void *host) {
    DeviceGuard dg(c);
    void *dev; size_t esz;
    if (int r = particle_array(c, ispec, attr, &dev, &esz)) return r;
    Species &sp = c->spec[ispec];
    if (in_record(sp, attr)) { if (int r = xfer_one(c, sp, attr, host, 0, sp.total, false)) return r; }
    else CUDA_TRY(cudaMemcpyAsync(host, dev, esz * sp.total, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return 0;
}
// per-patch host pointers (ptrs[p] = the patch's own array): a record attribute is staged patch by patch
extern "C" int lpic_upload_particle_ptrs(lpic_ctx *c, int ispec, int attr, const void *const *ptrs) {
    DeviceGuard dg(c);
    void *dev; size_t esz;
    if (int r = particle_array(c, ispec, attr, &dev, &esz)) return r;
    Species &sp = c->spec[ispec];
    for (int p = 0; p < c->g.npatch; p++) {
        if (sp.h_npart[p] <= 0) continue;
        if (in_record(sp, attr)) {  // xfer_one indexes host + first: hand it the base such an arena would have
            if (int r = xfer_one(c, sp, attr, (double *)ptrs[p] - sp.h_off[p], sp.h_off[p], sp.h_npart[p], true)) return r;
        } else {
            CUDA_TRY(cudaMemcpyAsync((char *)dev + esz * sp.h_off[p], ptrs[p], esz * sp.h_npart[p], cudaMemcpyHostToDevice, c->stream));
        }
    }
    sp.sort.valid = false;
    sp.lists_valid = false; sp.sort.keys_valid = sp.sort.keys_written = false;
    return 0;
}
extern "C" int lpic_download_particle_ptrs(lpic_ctx *c, int ispec, int attr, void *const *ptrs) {
    DeviceGuard dg(c);
    void *dev; size_t esz;
    if (int r = particle_array(c, ispec, attr, &dev, &esz)) return r;
    Species &sp = c->spec[ispec];
    for (int p = 0; p < c->g.npatch; p++) {
        if (sp.h_npart[p] <= 0) continue;
        if (in_record(sp, attr)) {
            if (int r = xfer_one(c, sp, attr, (double *)ptrs[p] - sp.h_off[p], sp.h_off[p], sp.h_npart[p], false)) return r;
        } else {
            CUDA_TRY(cudaMemcpyAsync(ptrs[p], (char *)dev + esz * sp.h_off[p], esz * sp.h_npart[p], cudaMemcpyDeviceToHost, c->stream));
        }
    }
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return 0;
}

// ParticlesBase.extend (core/particles.py:141-168): new slots are NaN, w = 0, dead, with fresh ids.
__global__ void __launch_bounds__(256) k_extend_init(int npatch, const i64 *__restrict__ off, const i64 *__restrict__ old_npart,
                                                     const i64 *__restrict__ ext, const u64 *__restrict__ id_first,
                                                     double *const *attrs, int nattr, int nrec, int ia_w, int ia_id, u8 *dead,
                                                     int blocks_per_patch) {  // attrs[a], a < nrec, are indexed with the record stride
    const int p = blockIdx.x / blocks_per_patch;
    const i64 t = (i64)(blockIdx.x - p * blocks_per_patch) * blockDim.x + threadIdx.x;
    if (t >= ext[p]) return;
    const i64 ip = off[p] + old_npart[p] + t;
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    for (int a = 0; a < nattr; a++) {
        double v = nan;
        if (a == ia_w) v = 0.0;
        if (a == ia_id) v = __longlong_as_double((long long)(id_first[p] + (u64)t));
        attrs[a][a < nrec ? ip * LPIC_NREC : ip] = v;
    }
    dead[ip] = 1;
}

__global__ void __launch_bounds__(256) k_pidx_reset(const i64 *__restrict__ off, const i64 *__restrict__ old_npart,
                                                    const i64 *__restrict__ ext, int *__restrict__ pidx, int blocks_per_patch) {
    const int p = blockIdx.x / blocks_per_patch;
    const i64 t = (i64)(blockIdx.x - p * blocks_per_patch) * blockDim.x + threadIdx.x;
    if (ext[p] <= 0 || t >= old_npart[p] + ext[p]) return;
    pidx[off[p] + t] = -1;
}

struct __align__(16) Rec64 { double2 v[4]; };
// move every patch segment from the old arena offsets to the new ones (one attribute at a time)
template <typename T>
__global__ void __launch_bounds__(256) k_relayout(const T *__restrict__ src, T *__restrict__ dst, const i64 *__restrict__ old_off,
                                                  const i64 *__restrict__ new_off, const i64 *__restrict__ npart,
                                                  int blocks_per_patch) {
    const int p = blockIdx.x / blocks_per_patch;
    const i64 t = (i64)(blockIdx.x - p * blocks_per_patch) * blockDim.x + threadIdx.x;
    if (t >= npart[p]) return;
    dst[new_off[p] + t] = src[old_off[p] + t];
}

extern "C" int lpic_species_extend(lpic_ctx *c, int ispec, const int64_t *ext, const uint64_t *id_first, int *relayout) {
    DeviceGuard dg(c);
    REQUIRE(ispec >= 0 && ispec < c->nspec && c->spec[ispec].allocated, "species %d not allocated", ispec);
    Species &sp = c->spec[ispec];
    const i64 n = c->g.npatch;
    if (relayout) *relayout = 0;
    i64 max_ext = 0;
    bool fits = true;
    for (i64 p = 0; p < n; p++) {
        REQUIRE(ext[p] >= 0, "negative extension");
        max_ext = std::max(max_ext, (i64)ext[p]);
        if (sp.h_npart[p] + ext[p] > sp.h_pcap[p]) fits = false;
    }
    if (max_ext == 0) return 0;
    // small device tables: old npart is still in d_npart; ext / id_first go to scratch
    // (persistent per context: cudaMalloc / cudaFree in a call that sparse moving-window runs make every step would
    // synchronise the whole device each time)
    if (!c->d_ext_tab) {
        CUDA_TRY(cudaMalloc(&c->d_ext_tab, sizeof(i64) * n * 3));
        CUDA_TRY(cudaMalloc(&c->d_ext_attrs, sizeof(double *) * LPIC_NPATTR));
    }
    i64 *d_ext = c->d_ext_tab;
    u64 *d_idf = (u64 *)(d_ext + n);
    i64 *d_newoff = d_ext + 2 * n;
    CUDA_TRY(cudaMemcpyAsync(d_ext, ext, sizeof(i64) * n, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(d_idf, id_first, sizeof(u64) * n, cudaMemcpyHostToDevice, c->stream));
    if (!fits) {
        // re-layout: every segment gets room for its new size plus the same relative slack as before
        std::vector<i64> new_off(n), new_pcap(n);
        i64 total = 0;
        for (i64 p = 0; p < n; p++) {
            const i64 want = sp.h_npart[p] + ext[p];
            new_pcap[p] = std::max(sp.h_pcap[p], phys_cap(want, 1.25, 64));
            new_off[p] = total;
            total += new_pcap[p];
        }
        CUDA_TRY(cudaMemcpyAsync(d_newoff, new_off.data(), sizeof(i64) * n, cudaMemcpyHostToDevice, c->stream));
        const int bpp = (int)div_up(std::max<i64>(sp.max_npart, 1), 256);
        if (sp.rec) {  // the record arena moves as 64-byte elements
            double *fresh = nullptr;
            CUDA_TRY(cudaMalloc(&fresh, sizeof(Rec64) * total));
            CUDA_TRY(cudaMemsetAsync(fresh, 0, sizeof(Rec64) * total, c->stream));
            k_relayout<Rec64><<<(unsigned)((i64)bpp * n), 256, 0, c->stream>>>((const Rec64 *)sp.rec, (Rec64 *)fresh, sp.d_off, d_newoff, sp.d_npart, bpp);
            LAUNCHED(1);
            KERNEL_CHECK();
            CUDA_TRY(cudaStreamSynchronize(c->stream));
            cudaFree(sp.rec);
            sp.rec = fresh;
            for (int a = 0; a < LPIC_NREC; a++) sp.attr[a] = sp.rec + a;
        }
        for (int a = sp.rec ? LPIC_NREC : 0; a < LPIC_NPATTR; a++) {
            if (!attr_resident(sp, a)) continue;
            double *fresh = nullptr;
            CUDA_TRY(cudaMalloc(&fresh, sizeof(double) * total));
            CUDA_TRY(cudaMemsetAsync(fresh, 0, sizeof(double) * total, c->stream));
            k_relayout<double><<<(unsigned)((i64)bpp * n), 256, 0, c->stream>>>(sp.attr[a], fresh, sp.d_off, d_newoff, sp.d_npart, bpp);
            LAUNCHED(1);
            KERNEL_CHECK();
            CUDA_TRY(cudaStreamSynchronize(c->stream));
            cudaFree(sp.attr[a]);
            sp.attr[a] = fresh;
        }
        u8 *fresh_dead = nullptr;
        CUDA_TRY(cudaMalloc(&fresh_dead, (size_t)total));
        CUDA_TRY(cudaMemsetAsync(fresh_dead, 1, (size_t)total, c->stream));
        k_relayout<u8><<<(unsigned)((i64)bpp * n), 256, 0, c->stream>>>(sp.dead, fresh_dead, sp.d_off, d_newoff, sp.d_npart, bpp);
        LAUNCHED(1);
        KERNEL_CHECK();
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        cudaFree(sp.dead);
        sp.dead = fresh_dead;
        int *fresh_pidx = nullptr;
        CUDA_TRY(cudaMalloc(&fresh_pidx, sizeof(int) * (size_t)total));
        CUDA_TRY(cudaMemsetAsync(fresh_pidx, 0xff, sizeof(int) * (size_t)total, c->stream));
        k_relayout<int><<<(unsigned)((i64)bpp * n), 256, 0, c->stream>>>(sp.sort.pidx, fresh_pidx, sp.d_off, d_newoff, sp.d_npart, bpp);
        LAUNCHED(1);
        KERNEL_CHECK();
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        cudaFree(sp.sort.pidx);
        sp.sort.pidx = fresh_pidx;
        sp.sort.pidx_cap = total;
        for (i64 p = 0; p < n; p++) { sp.h_off[p] = new_off[p]; sp.h_pcap[p] = new_pcap[p]; }
        sp.total = total;
        CUDA_TRY(cudaMemcpyAsync(sp.d_off, sp.h_off, sizeof(i64) * n, cudaMemcpyHostToDevice, c->stream));
        if (relayout) *relayout = 1;
    }
    // initialise the appended slots
    double *h_attrs[LPIC_NPATTR];
    int na = 0, ia_w = -1, ia_id = -1;
    for (int a = 0; a < LPIC_NPATTR; a++) {
        if (!attr_resident(sp, a)) continue;
        if (a == LPIC_P_W) ia_w = na;
        if (a == LPIC_P_ID) ia_id = na;
        h_attrs[na++] = sp.attr[a];
    }
    double **d_attrs = c->d_ext_attrs;
    CUDA_TRY(cudaMemcpyAsync(d_attrs, h_attrs, sizeof(double *) * na, cudaMemcpyHostToDevice, c->stream));
    const int bpp = (int)div_up(max_ext, 256);
    k_extend_init<<<(unsigned)((i64)bpp * n), 256, 0, c->stream>>>((int)n, sp.d_off, sp.d_npart, d_ext, d_idf, d_attrs, na,
                                                                 sp.rec ? LPIC_NREC : 0, ia_w, ia_id, sp.dead, bpp);
    LAUNCHED(1);
    KERNEL_CHECK();
    for (i64 p = 0; p < n; p++) sp.h_npart[p] += ext[p];
    {   // the reference re-creates the sorter's index arrays (-1) for every extended patch (simulation.py:781-824)
        i64 m = 0;
        for (i64 p = 0; p < n; p++) m = std::max(m, sp.h_npart[p]);
        const int bpp2 = (int)div_up(std::max<i64>(m, 1), 256);
        k_pidx_reset<<<(unsigned)((i64)bpp2 * n), 256, 0, c->stream>>>(sp.d_off, sp.d_npart, d_ext, sp.sort.pidx, bpp2);
        LAUNCHED(1);
        KERNEL_CHECK();
    }
    int r = upload_layout(c, sp);  // synchronises the stream: the host tables above may go out of scope
    sp.sort.valid = false;
    sp.lists_valid = false; sp.sort.keys_valid = sp.sort.keys_written = false;
    return r;
}

// New slot counts inside the existing segments (every npart[p] <= capacity of patch p): the host re-initialised some
// patches' particles (ParticlesBase.initialize after a MovingWindow shift) and uploads their values next.
extern "C" int lpic_species_set_npart(lpic_ctx *c, int ispec, const int64_t *npart) {
    DeviceGuard dg(c);
    REQUIRE(ispec >= 0 && ispec < c->nspec && c->spec[ispec].allocated, "species %d not allocated", ispec);
    Species &sp = c->spec[ispec];
    const i64 n = c->g.npatch;
    for (i64 p = 0; p < n; p++)
        REQUIRE(npart[p] >= 0 && npart[p] <= sp.h_pcap[p], "patch %lld: %lld slots do not fit the segment of %lld", (long long)p,
                (long long)npart[p], (long long)sp.h_pcap[p]);
    for (i64 p = 0; p < n; p++) sp.h_npart[p] = npart[p];
    sp.sort.valid = false;
    sp.lists_valid = false; sp.sort.keys_valid = sp.sort.keys_written = false;
    return upload_layout(c, sp);
}

int lpic_ensure_scratch(lpic_ctx *c, i64 slots) {
    c->scratch_epoch++;  // every caller is about to overwrite (part of) the shared lists
    if (slots <= c->scr_cap) return 0;
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    cudaFree(c->scr_a); cudaFree(c->scr_b); cudaFree(c->scr_buf);
    c->scr_a = c->scr_b = nullptr; c->scr_buf = nullptr; c->scr_cap = 0;
    const i64 cap = slots + slots / 8 + 1024;
    CUDA_TRY(cudaMalloc(&c->scr_a, sizeof(int) * cap));
    CUDA_TRY(cudaMalloc(&c->scr_b, sizeof(int) * cap));
    CUDA_TRY(cudaMalloc(&c->scr_buf, sizeof(double) * cap));
    c->scr_cap = cap;
    return 0;
}


// fp64 FMA peak of this device, measured: the co-bound of the fused particle kernel next to the HBM copy bandwidth
// (SURVEY.md 8(d)).  8 independent DFMA chains per thread, 148 x 8 CTAs of 256 threads, best of 5 launches.
__global__ void __launch_bounds__(256) k_fp64_peak(double *out, int iters, double a, double b) {
    double v[8];
#pragma unroll
    for (int j = 0; j < 8; j++) v[j] = threadIdx.x + j;
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int j = 0; j < 8; j++) v[j] = __fma_rn(v[j], a, b);
    }
    double sum = 0.0;
#pragma unroll
    for (int j = 0; j < 8; j++) sum += v[j];
    if (sum == 12345.678) out[0] = sum;  // never true: keeps the chains alive
}
extern "C" int lpic_fp64_peak(lpic_ctx *c, double *tflops) {
    DeviceGuard dg(c);
    const int iters = 4096, grid = 148 * 8;
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    double best = 0.0;
    for (int rep = 0; rep < 6; rep++) {
        CUDA_TRY(cudaEventRecord(e0, c->stream));
        k_fp64_peak<<<grid, 256, 0, c->stream>>>(c->d_tmpf, iters, 0.999999, 1e-9);
        CUDA_TRY(cudaEventRecord(e1, c->stream));
        CUDA_TRY(cudaEventSynchronize(e1));
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        const double flops = 2.0 * 32.0 * iters * 256.0 * grid;
        if (rep > 0) best = std::max(best, flops / (ms * 1e-3) / 1e12);
    }
    LAUNCHED(6);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *tflops = best;
    return 0;
}

// PCI bus id of the context's device ("0000:1b:00.0"): lets the host side place its pinned mirrors on the GPU's NUMA node
extern "C" int lpic_device_pci_bus_id(lpic_ctx *c, char *out, int len) {
    CUDA_TRY(cudaDeviceGetPCIBusId(out, len, c->device));
    return 0;
}

static const int kEventSlots = 4096;
extern "C" int lpic_event_record(lpic_ctx *c, int slot) {
    DeviceGuard dg(c);
    REQUIRE(slot >= 0 && slot < kEventSlots, "event slot %d out of range", slot);
    if (!c->events) c->events = new cudaEvent_t[kEventSlots]();
    if (!c->events[slot]) CUDA_TRY(cudaEventCreate(&c->events[slot]));
    CUDA_TRY(cudaEventRecord(c->events[slot], c->stream));
    return 0;
}
extern "C" int lpic_event_elapsed_ms(lpic_ctx *c, int a, int b, double *ms) {
    DeviceGuard dg(c);
    REQUIRE(c->events && a >= 0 && b >= 0 && a < kEventSlots && b < kEventSlots && c->events[a] && c->events[b], "events not recorded");
    CUDA_TRY(cudaEventSynchronize(c->events[b]));
    float f = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&f, c->events[a], c->events[b]));
    *ms = f;
    return 0;
}
extern "C" int64_t lpic_launch_count(void) { return g_lpic_launches; }
