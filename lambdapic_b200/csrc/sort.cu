// Per-patch in-place bucket sort of the particle arenas (records + plain arrays), reproducing the reference's slot permutation bit-for-bit.
//
// Reference behaviour restated (not copied): core/sort/cpu3d.c:8-156 (calculate_cell_index, calculate_bucket_bound,
// bucket_sort_3d), driver :214-299, 2D twin core/sort/cpu2d.c; facade core/sort/particle_sort.py.
//   1. key[ip]  = bucket of slot ip; dead slots inherit the key of the previous slot (a running value that starts
//                 at 0 for every patch); alive out-of-range particles go to the last bucket.
//   2. count / exclusive prefix -> bucket_bound_min/max; owner[ip] = bucket whose range contains slot ip.
//   3. the misplaced slots (key != owner), listed ascending, are permuted among themselves: the values are laid
//      out stably by key, then written back in list order.  Equivalent closed form used here:
//         new[T_i] = old[T_j]  with  dest(j) = start[key_j] + #{j' < j : key_j' = key_j} = i.
// One CTA owns one patch (all scans are block scans with a running carry), then two grid-wide kernels per
// attribute move the values through a staging buffer.  Integer results are identical to the reference by construction.
#include <limits.h>
#include <algorithm>
#include <vector>
#include "lpic_common.cuh"

namespace {

constexpr int T = 256;         // threads per sort CTA
constexpr int IT = 4;          // consecutive slots per thread and scan step
constexpr int SMEM_BINS = 2048;  // histograms up to this many buckets live in shared memory

__device__ __forceinline__ int warp_incl_sum(int v) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int n = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += n;
    }
    return v;
}
__device__ __forceinline__ int warp_incl_max(int v) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int n = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v = max(v, n);
    }
    return v;
}
// inclusive block scans over T threads; `total` receives the block aggregate. sw: int[T/32] scratch.
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
__device__ __forceinline__ int block_incl_max(int v, int *sw) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    v = warp_incl_max(v);
    __syncthreads();
    if (lane == 31) sw[w] = v;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < T / 32; i++)
        if (i < w) v = max(v, sw[i]);
    return v;
}

// warp-aggregated counter increment: lanes with equal key elect a leader that adds the group size.
// Returns the value of the counter before this warp's group was added plus the lane's rank inside the group.
__device__ __forceinline__ int grouped_fetch_add(int *ctr, int key, bool active) {
    const unsigned lane = threadIdx.x & 31;
    const int k = active ? key : -2 - (int)lane;
    const unsigned peers = __match_any_sync(0xffffffffu, k);
    const int leader = __ffs(peers) - 1;
    int base = 0;
    if (active && (int)lane == leader) base = atomicAdd(ctr + key, __popc(peers));
    base = __shfl_sync(0xffffffffu, base, leader);
    return base + __popc(peers & ((1u << lane) - 1u));
}

struct SortArgs {
    SortKeyParams kp;     // the same parameters as a view for sort_bucket_key
    const int *kcache;    // keys left behind by the migration (null: compute them from the positions)
    int ps;  // stride of x, y, z (8: records, 1: separate arrays)
    const double *x, *y, *z;
    const u8 *dead;
    const i64 *off, *npart;
    const double *org;  // (3, npatch) bucket origins
    int npatch, nxb, nyb, nzb, nbin, reverse_x, dim;
    double dxb, dyb, dzb;
    int *pidx;                              // arena
    i64 *bucket_count, *bound_min, *bound_max;  // (npatch, nbin)
    int *g_hist, *g_cur;                    // (npatch, nbin) global fallbacks when nbin > SMEM_BINS
    int *tgt, *src_of;                      // arena-sized lists (local slot numbers)
    i64 *nbuf;                              // (npatch)
};

// Single-bucket axes (ny_buckets = nz_buckets = 1, the default): floor(v / d) is 0 exactly when 0 <= v < d -- for
// doubles v < d the rounded quotient is at most 1 - 2^-53 < 1 -- so the reference's division is replaced by two
// comparisons there without changing any result.
__global__ void __launch_bounds__(T) k_sort_index(SortArgs a) {
    __shared__ int skey[T];
    __shared__ int sinc[T];
    __shared__ int sw[T / 32];
    __shared__ int s_hist[SMEM_BINS];
    __shared__ int s_keys[T * IT];
    __shared__ int s_carry;
    const int p = blockIdx.x, tid = threadIdx.x;
    const i64 off = a.off[p];
    const int np = (int)a.npart[p];
    const int nbin = a.nbin;
    const bool use_smem = nbin <= SMEM_BINS;
    int *hist = use_smem ? s_hist : a.g_hist + (size_t)p * nbin;
    int *cur = use_smem ? s_hist : a.g_cur + (size_t)p * nbin;
    if (use_smem)
        for (int b = tid; b < nbin; b += T) s_hist[b] = 0;
    if (tid == 0) s_carry = 0;  // icell = 0 at the start of every patch (cpu3d.c:19)
    __syncthreads();
    const double x0 = a.org[p], y0 = a.org[a.npatch + p], z0 = a.org[2 * a.npatch + p];

    // ---- 1. keys with inheritance, histogram (IT consecutive slots per thread: one block scan per T*IT slots) ----
    for (int base = 0; base < np; base += T * IT) {
        const int ip0 = base + tid * IT;
        int keys[IT];
        bool valid[IT];
        int lastv = -1;
        // Keys are computed with adjacent lanes on adjacent slots (a warp covers its 32 * IT consecutive slots in IT strips:
        // coalesced whether the positions sit in separate arrays or in 64-byte records) and handed to the thread that owns
        // IT consecutive slots through shared memory.  -1: dead or past the end.
        {
            const int wbase = base + (tid >> 5) * (32 * IT), lane = tid & 31;
#pragma unroll
            for (int m = 0; m < IT; m++) {
                const int ip = wbase + 32 * m + lane;
                int key = -1;
                if (ip < np && !a.dead[off + ip]) {
                    if (a.kcache) {
                        key = a.kcache[off + ip];
                    } else {
                        double px, py, pz;
                        if (a.ps == LPIC_NREC) {
                            const double2 xy = *reinterpret_cast<const double2 *>(a.x + (off + ip) * LPIC_NREC);
                            px = xy.x; py = xy.y; pz = a.dim == 3 ? a.z[(off + ip) * LPIC_NREC] : 0.0;
                        } else {
                            px = a.x[off + ip]; py = a.y[off + ip]; pz = a.dim == 3 ? a.z[off + ip] : 0.0;
                        }
                        key = sort_bucket_key(a.kp, x0, y0, z0, px, py, pz);
                    }
                }
                s_keys[(tid >> 5) * (32 * IT) + 32 * m + lane] = key;
            }
            __syncwarp();
#pragma unroll
            for (int j = 0; j < IT; j++) {
                keys[j] = s_keys[tid * IT + j];
                valid[j] = keys[j] >= 0;
                if (valid[j]) lastv = j; else keys[j] = 0;
            }
            __syncwarp();
        }
        int mylast = 0;
#pragma unroll
        for (int j = 0; j < IT; j++)
            if (j == lastv) mylast = keys[j];
        skey[tid] = mylast;
        const int carry = s_carry;
        sinc[tid] = block_incl_max(lastv >= 0 ? tid : -1, sw);  // syncs inside make skey visible
        __syncthreads();
        const int prev = tid > 0 ? sinc[tid - 1] : -1;  // last thread before me that holds an alive slot
        int run = prev >= 0 ? skey[prev] : carry;
        int rk = -1, rc = 0;  // run-length compressed histogram update (sorted data: usually one run per thread)
#pragma unroll
        for (int j = 0; j < IT; j++) {
            const int ip = ip0 + j;
            if (valid[j]) run = keys[j];
            if (ip < np) {
                a.pidx[off + ip] = run;
                if (run == rk) rc++;
                else {
                    if (rc) atomicAdd(&hist[rk], rc);
                    rk = run;
                    rc = 1;
                }
            }
        }
        if (rc) atomicAdd(&hist[rk], rc);
        __syncthreads();
        if (tid == T - 1) s_carry = run;
        __syncthreads();
    }

    // ---- 2. bucket bounds (exclusive prefix of the counts) ----------------------------------------------
    {
        int run = 0;
        for (int base = 0; base < nbin; base += T) {
            const int b = base + tid;
            const int cnt = b < nbin ? hist[b] : 0;
            int tot;
            const int incl = block_incl_sum(cnt, sw, tot);
            if (b < nbin) {
                a.bucket_count[(size_t)p * nbin + b] = cnt;
                a.bound_min[(size_t)p * nbin + b] = run + incl - cnt;
                a.bound_max[(size_t)p * nbin + b] = run + incl;
            }
            run += tot;
            __syncthreads();
        }
    }
    __syncthreads();
    if (use_smem) {
        for (int b = tid; b < nbin; b += T) s_hist[b] = 0;
    } else {
        for (int b = tid; b < nbin; b += T) cur[b] = 0;
    }
    __threadfence_block();
    __syncthreads();

    // ---- 3. misplaced slots, ascending; per-bucket counts of them ------------------------------------------
    const i64 *bmax = a.bound_max + (size_t)p * nbin;
    int nbuf = 0;
    for (int base = 0; base < np; base += T * IT) {
        const int ip0 = base + tid * IT;
        int key[IT];
        bool miss[IT];
        int nmiss = 0, owner = 0;
        if (ip0 < np) {
            int lo = 0, hi = nbin - 1;  // owner = first bucket with bound_max > ip0
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (bmax[mid] > ip0) hi = mid; else lo = mid + 1;
            }
            owner = lo;
        }
#pragma unroll
        for (int j = 0; j < IT; j++) {
            const int ip = ip0 + j;
            miss[j] = false;
            key[j] = 0;
            if (ip < np) {
                while (bmax[owner] <= ip) owner++;  // owners are non-decreasing along the slots
                key[j] = a.pidx[off + ip];
                miss[j] = key[j] != owner;
                nmiss += miss[j];
            }
        }
        int tot;
        int pos = nbuf + block_incl_sum(nmiss, sw, tot) - nmiss;
#pragma unroll
        for (int j = 0; j < IT; j++)
            if (miss[j]) {
                a.tgt[off + pos++] = ip0 + j;
                atomicAdd(&cur[key[j]], 1);  // counts by key == counts by owner (see header comment)
            }
        nbuf += tot;
        __syncthreads();
    }
    if (tid == 0) a.nbuf[p] = nbuf;
    if (nbuf == 0) return;
    __syncthreads();

    // ---- 4. start[b] = exclusive prefix of the misplaced counts (in place) -----------------------------------
    {
        int run = 0;
        for (int base = 0; base < nbin; base += T) {
            const int b = base + tid;
            const int cnt = b < nbin ? cur[b] : 0;
            int tot;
            const int incl = block_incl_sum(cnt, sw, tot);
            __syncthreads();
            if (b < nbin) cur[b] = run + incl - cnt;
            run += tot;
            __syncthreads();
        }
    }
    __threadfence_block();
    __syncthreads();

    // ---- 5. stable placement: warps take turns so that counters advance in list order ------------------------
    for (int base = 0; base < nbuf; base += T) {
        const int i = base + tid;
        const bool act = i < nbuf;
        int slot = 0, key = 0;
        if (act) {
            slot = a.tgt[off + i];
            key = a.pidx[off + slot];
        }
        for (int w = 0; w < T / 32; w++) {
            if ((tid >> 5) == w) {
                const int dest = grouped_fetch_add(cur, key, act);
                if (act) a.src_of[off + dest] = slot;
            }
            __syncthreads();
        }
    }
}

struct MoveArgs {
    double *a[LPIC_NPATTR];
    int stride[LPIC_NPATTR];  // 1, or 8 for a word of the record arena moved on its own
    int n;                    // number of 8-byte attributes in this group
    double *rec;              // record arena: 16-byte pieces [piece0, piece0 + npiece) of every moved record ride with the group
    int piece0, npiece;
    u8 *dead;                 // moved with the group if not null
};

// Value move of the misplaced slots, all attributes of a group in one thread: thread i of the compact list
// [0, sum nbuf) finds its patch by bisection of the prefix poff, then issues one independent load per attribute.
// gather: staging[attr][i] = attr[src_of[i]]; scatter: attr[tgt[i]] = staging[attr][i].  Staging: [npiece][total] 16-byte
// record pieces, then [n][total] words, then the is_dead bytes.
__device__ __forceinline__ int patch_of(const i64 *__restrict__ poff, int npatch, i64 i) {
    int lo = 0, hi = npatch - 1;  // last patch with poff[p] <= i
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (poff[mid] <= i) lo = mid; else hi = mid - 1;
    }
    return lo;
}
// (the launch covers the patches [p0, p0 + npatch) whose moved slots are [base, base + total) of the compact list)
__global__ void __launch_bounds__(256) k_sort_gather_all(MoveArgs A, double *__restrict__ buf, i64 total, const int *__restrict__ src_of,
                                                         const i64 *__restrict__ off, const i64 *__restrict__ poff, int p0, int npatch,
                                                         i64 base) {
    const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int p = p0 + patch_of(poff + p0, npatch, i + base);
    const i64 src = off[p] + src_of[off[p] + (i + base - poff[p])];
    double v[LPIC_NPATTR];
    double2 r[4];
#pragma unroll
    for (int k = 0; k < 4; k++)
        if (k < A.npiece) r[k] = reinterpret_cast<const double2 *>(A.rec + src * LPIC_NREC)[A.piece0 + k];
#pragma unroll
    for (int y = 0; y < LPIC_NPATTR; y++)
        if (y < A.n) v[y] = A.a[y][src * A.stride[y]];
    double2 *pieces = reinterpret_cast<double2 *>(buf);
#pragma unroll
    for (int k = 0; k < 4; k++)
        if (k < A.npiece) pieces[(size_t)k * total + i] = r[k];
    double *words = buf + (size_t)2 * A.npiece * total;
#pragma unroll
    for (int y = 0; y < LPIC_NPATTR; y++)
        if (y < A.n) words[(size_t)y * total + i] = v[y];
    if (A.dead) ((u8 *)(words + (size_t)A.n * total))[i] = A.dead[src];
}
__global__ void __launch_bounds__(256) k_sort_scatter_all(MoveArgs A, const double *__restrict__ buf, i64 total, const int *__restrict__ tgt,
                                                          const i64 *__restrict__ off, const i64 *__restrict__ poff, int p0, int npatch,
                                                          i64 base) {
    const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int p = p0 + patch_of(poff + p0, npatch, i + base);
    const i64 dst = off[p] + tgt[off[p] + (i + base - poff[p])];
    const double2 *pieces = reinterpret_cast<const double2 *>(buf);
#pragma unroll
    for (int k = 0; k < 4; k++)
        if (k < A.npiece) reinterpret_cast<double2 *>(A.rec + dst * LPIC_NREC)[A.piece0 + k] = pieces[(size_t)k * total + i];
    const double *words = buf + (size_t)2 * A.npiece * total;
#pragma unroll
    for (int y = 0; y < LPIC_NPATTR; y++)
        if (y < A.n) A.a[y][dst * A.stride[y]] = words[(size_t)y * total + i];
    if (A.dead) A.dead[dst] = ((const u8 *)(words + (size_t)A.n * total))[i];
}

__global__ void k_widen(const int *__restrict__ src, i64 *__restrict__ dst, i64 n) {
    const i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) dst[t] = src[t];
}

}  // namespace

extern "C" int lpic_sort(lpic_ctx *c, int ispec, int reverse_x, int64_t nxb, int64_t nyb, int64_t nzb, double dxb,
                         double dyb, double dzb, const double *x0s, const double *y0s, const double *z0s,
                         int64_t *nbuf_total) {
    DeviceGuard dg(c);
    REQUIRE(ispec >= 0 && ispec < c->nspec && c->spec[ispec].allocated, "species %d not allocated", ispec);
    Species &sp = c->spec[ispec];
    const Geom &g = c->g;
    const i64 n = g.npatch;
    if (g.dim == 2) { nzb = 1; dzb = 1.0; }
    const i64 nbin = nxb * nyb * nzb;
    REQUIRE(nbin > 0 && nbin < (1ll << 30), "bad bucket grid");
    REQUIRE(sp.max_npart < (1ll << 31), "patch too large for 32-bit slot numbers");
    SortState &st = sp.sort;
    if (st.nbin != nbin) {
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        cudaFree(st.bucket_count); cudaFree(st.bound_min); cudaFree(st.bound_max);
        CUDA_TRY(cudaMalloc(&st.bucket_count, sizeof(i64) * n * nbin));
        CUDA_TRY(cudaMalloc(&st.bound_min, sizeof(i64) * n * nbin));
        CUDA_TRY(cudaMalloc(&st.bound_max, sizeof(i64) * n * nbin));
        st.nbin = nbin;
    }
    st.nxb = nxb; st.nyb = nyb; st.nzb = nzb;
    if (int r = lpic_ensure_scratch(c, sp.total)) return r;
    // bucket origins
    std::vector<double> org(3 * n, 0.0);
    for (i64 p = 0; p < n; p++) { org[p] = x0s[p]; org[n + p] = y0s[p]; org[2 * n + p] = (g.dim == 3 && z0s) ? z0s[p] : 0.0; }
    CUDA_TRY(cudaMemcpyAsync(c->d_sort_org, org.data(), sizeof(double) * 3 * n, cudaMemcpyHostToDevice, c->stream));
    int *g_hist = nullptr, *g_cur = nullptr;
    if (nbin > SMEM_BINS) {  // persistent, grown on demand
        const size_t need = (size_t)n * nbin * 2;
        if (need > c->sort_hist_cap) {
            CUDA_TRY(cudaStreamSynchronize(c->stream));
            cudaFree(c->d_sort_hist);
            c->d_sort_hist = nullptr; c->sort_hist_cap = 0;
            CUDA_TRY(cudaMalloc(&c->d_sort_hist, sizeof(int) * need));
            c->sort_hist_cap = need;
        }
        g_hist = c->d_sort_hist;
        g_cur = g_hist + n * nbin;
        CUDA_TRY(cudaMemsetAsync(g_hist, 0, sizeof(int) * n * nbin * 2, c->stream));
    }
    i64 *d_nbuf = c->d_tmp64 + 64;
    // ---- key cache bookkeeping: are the keys the migration left behind computed with exactly these parameters? --------
    bool use_cache = false;
    {
        // Off by default: measured at 128^3 x 32+32 ppc, the sorter gains 0.63 ms per step and the migration's passes lose
        // 0.36 ms computing the keys (net 0.9 % of the step); LPIC_SORT_KEY_CACHE=1 turns it on (all GPU tests pass with it).
        static const bool enabled = getenv("LPIC_SORT_KEY_CACHE") != nullptr;
        SortKeyParams kp;
        kp.org = nullptr; kp.npatch = (int)n; kp.nxb = (int)nxb; kp.nyb = (int)nyb; kp.nzb = (int)nzb; kp.nbin = (int)nbin;
        kp.reverse_x = reverse_x; kp.dim = g.dim; kp.dxb = dxb; kp.dyb = dyb; kp.dzb = dzb;
        if (enabled) {
            if (!st.h_korg) {
                st.h_korg = new double[3 * n];
                CUDA_TRY(cudaMalloc(&st.d_korg, sizeof(double) * 3 * n));
                st.have_kp = false;
            }
            kp.org = st.d_korg;
            const bool same = st.have_kp && memcmp(st.h_korg, org.data(), sizeof(double) * 3 * n) == 0 && st.kp.nxb == kp.nxb &&
                              st.kp.nyb == kp.nyb && st.kp.nzb == kp.nzb && st.kp.reverse_x == kp.reverse_x && st.kp.dim == kp.dim &&
                              st.kp.dxb == kp.dxb && st.kp.dyb == kp.dyb && st.kp.dzb == kp.dzb;
            use_cache = same && st.keys_valid && st.kcache && st.kcache_cap >= sp.total;
            if (!same) {
                memcpy(st.h_korg, org.data(), sizeof(double) * 3 * n);
                CUDA_TRY(cudaMemcpyAsync(st.d_korg, st.h_korg, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, c->stream));
            }
            if (st.kcache_cap < sp.total) {  // (re)allocated here, filled by the next migration
                CUDA_TRY(cudaStreamSynchronize(c->stream));
                cudaFree(st.kcache);
                st.kcache = nullptr; st.kcache_cap = 0;
                CUDA_TRY(cudaMalloc(&st.kcache, sizeof(int) * (size_t)sp.total));
                st.kcache_cap = sp.total;
                use_cache = false;
            }
        }
        st.kp = kp;          // (the sorter's own key computation reads these too)
        st.have_kp = enabled;
    }
    SortArgs a;
    a.kp = st.kp; a.kp.org = c->d_sort_org;
    a.kcache = use_cache ? st.kcache : nullptr;
    a.x = sp.attr[LPIC_P_X]; a.y = sp.attr[LPIC_P_Y]; a.z = sp.attr[LPIC_P_Z]; a.dead = sp.dead; a.ps = sp.pstride;
    a.off = sp.d_off; a.npart = sp.d_npart; a.org = c->d_sort_org;
    a.npatch = (int)n; a.nxb = (int)nxb; a.nyb = (int)nyb; a.nzb = (int)nzb; a.nbin = (int)nbin;
    a.reverse_x = reverse_x; a.dim = g.dim; a.dxb = dxb; a.dyb = dyb; a.dzb = dzb;
    a.pidx = st.pidx; a.bucket_count = st.bucket_count; a.bound_min = st.bound_min; a.bound_max = st.bound_max;
    a.g_hist = g_hist; a.g_cur = g_cur; a.tgt = c->scr_a; a.src_of = c->scr_b; a.nbuf = d_nbuf;
    k_sort_index<<<(unsigned)n, T, 0, c->stream>>>(a);
    LAUNCHED(1);
    KERNEL_CHECK();
    std::vector<i64> h_nbuf(n);
    CUDA_TRY(cudaMemcpyAsync(h_nbuf.data(), d_nbuf, sizeof(i64) * n, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    i64 total = 0;
    for (i64 p = 0; p < n; p++) total += h_nbuf[p];
    if (nbuf_total) *nbuf_total = total;
    if (total > 0) {
        std::vector<i64> poff(n);
        i64 run = 0;
        for (i64 p = 0; p < n; p++) { poff[p] = run; run += h_nbuf[p]; }
        i64 *d_poff = d_nbuf + n;
        CUDA_TRY(cudaMemcpyAsync(d_poff, poff.data(), sizeof(i64) * n, cudaMemcpyHostToDevice, c->stream));
        // The permutation never leaves a patch, so the move is done for one group of consecutive patches at a time, each
        // group small enough that ALL its attributes fit the staging buffer at once: whole 64-byte records (two full
        // sectors per moved slot) instead of four passes over half-used sectors, and the slot lists are read once.  That is
        // what the first steps of a run need, when the reference's growth rule makes up to half of all slots change
        // bucket.  A single patch too large for that (few-patch runs) falls back to moving its attributes in turns.
        struct Item { double *ptr; int stride; };
        std::vector<Item> plain;
        for (int at = sp.rec ? LPIC_NREC : 0; at < LPIC_NPATTR; at++)
            if (sp.attr[at]) plain.push_back({sp.attr[at], 1});
        const i64 rows_all = (sp.rec ? LPIC_NREC : 0) + (i64)plain.size() + 1;  // 8-byte words per moved slot, is_dead included
        i64 p0 = 0;
        while (p0 < n) {
            i64 p1 = p0, cnt = 0;
            while (p1 < n && (cnt == 0 || (cnt + h_nbuf[p1]) * rows_all <= c->scr_cap)) cnt += h_nbuf[p1++];
            const i64 base = poff[p0];
            if (cnt > 0) {
                // rows of the staging buffer (one row = one word per moved slot of the group): the record arena as four
                // 16-byte pieces of two rows each -- or, if not even one piece fits, as eight strided words -- then the
                // plain arrays, then the is_dead bytes (less than a row; alone if no row is left).  Every pass is gathered
                // completely before it is scattered: sources and targets are the same set of slots.
                const i64 room = std::max<i64>(1, c->scr_cap / cnt);
                std::vector<Item> words;
                int pieces_left = 0;
                if (sp.rec) {
                    if (room >= 2) pieces_left = 4;
                    else for (int at = 0; at < LPIC_NREC; at++) words.push_back({sp.rec + at, LPIC_NREC});
                }
                words.insert(words.end(), plain.begin(), plain.end());
                const unsigned grid = (unsigned)div_up(cnt, 256);
                size_t a0 = 0;
                bool dead_done = false;
                while (a0 < words.size() || pieces_left > 0 || !dead_done) {
                    MoveArgs A;
                    A.n = 0; A.dead = nullptr; A.rec = sp.rec; A.piece0 = 4 - pieces_left; A.npiece = 0;
                    i64 rows = 0;
                    while (pieces_left > 0 && rows + 2 <= room) { A.npiece++; pieces_left--; rows += 2; }
                    while (pieces_left == 0 && a0 < words.size() && rows + 1 <= room && A.n < LPIC_NPATTR) {
                        A.a[A.n] = words[a0].ptr; A.stride[A.n] = words[a0].stride; A.n++; a0++; rows++;
                    }
                    if (pieces_left == 0 && a0 == words.size() && (rows + 1 <= room || rows == 0)) { A.dead = sp.dead; dead_done = true; }
                    k_sort_gather_all<<<grid, 256, 0, c->stream>>>(A, c->scr_buf, cnt, c->scr_b, sp.d_off, d_poff, (int)p0, (int)(p1 - p0), base);
                    k_sort_scatter_all<<<grid, 256, 0, c->stream>>>(A, c->scr_buf, cnt, c->scr_a, sp.d_off, d_poff, (int)p0, (int)(p1 - p0), base);
                    LAUNCHED(2);
                }
            }
            p0 = p1;
        }
        KERNEL_CHECK();
    }
    st.valid = true;
    st.keys_valid = st.keys_written = false;  // the slots were permuted: the cached keys are consumed
    sp.lists_valid = false;  // slots were permuted
    return 0;
}

extern "C" int lpic_sort_download(lpic_ctx *c, int ispec, int which, int64_t *out) {
    DeviceGuard dg(c);
    REQUIRE(ispec >= 0 && ispec < c->nspec && c->spec[ispec].allocated, "species %d not allocated", ispec);
    Species &sp = c->spec[ispec];
    SortState &st = sp.sort;
    const i64 n = c->g.npatch;
    if (which == LPIC_SORT_PARTICLE_INDEX) {
        if (int r = lpic_ensure_scratch(c, sp.total)) return r;
        i64 *tmp = (i64 *)c->scr_buf;  // the sort's fp64 staging buffer: one 8-byte word per slot, idle outside lpic_sort
        k_widen<<<div_up(std::max<i64>(sp.total, 1), 256), 256, 0, c->stream>>>(st.pidx, tmp, sp.total);
        LAUNCHED(1);
        CUDA_TRY(cudaMemcpyAsync(out, tmp, sizeof(i64) * sp.total, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        return 0;
    }
    REQUIRE(st.nbin > 0, "sorter of species %d has not run yet", ispec);
    const i64 *src = which == LPIC_SORT_BUCKET_COUNT ? st.bucket_count : which == LPIC_SORT_BOUND_MIN ? st.bound_min
                     : which == LPIC_SORT_BOUND_MAX ? st.bound_max : nullptr;
    REQUIRE(src != nullptr, "bad sorter array id %d", which);
    CUDA_TRY(cudaMemcpyAsync(out, src, sizeof(i64) * n * st.nbin, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return 0;
}
