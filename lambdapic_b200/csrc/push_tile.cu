// Tile kernel: the fused particle step (half push, 6 x 27-point gather, Boris, half push, Esirkepov deposit) of the 3D
// path, organised around a CTA-owned block of cells.
//
// Reference behaviour restated (not copied): core/pusher/unified/unified_pusher_3d.c:281-431 (strip loop: half push,
// interpolation, Boris, half push), core/current/current_deposit.h:275-440 (charge-conserving deposit).
//
//   1. k_tile_perm   one CTA per patch.  Every alive slot gets the key (tile, cell in tile) of the node nearest to its
//                    position AFTER the first half push -- the cell it gathers from, and where its deposit starts.  Shared-
//                    memory histogram -> block scan -> scatter gives a permutation of the slots in (tile, cell) order plus
//                    the first position of every tile.  The slot order in memory stays the reference's (it is the
//                    parity contract of sort and migration); only the ORDER OF PROCESSING changes.  Particles whose
//                    nearest node lies outside the patch's own cells (a boundary layer of v dt/2) go to the patch's
//                    list instead.
//   2. k_push_tile   one CTA per (patch, tile of TX x TY x TZ cells).  The CTA stages the E/B values of the tile plus its
//                    halo (6 x (TX+3)(TY+3)(TZ+3) doubles) in shared memory in LOGICAL order -- the wrapped-guard
//                    arithmetic of the global layout is paid once per staged value instead of once per gather load --
//                    and every gather becomes an LDS with an immediate offset.  Each warp then walks a contiguous run
//                    of the tile's cell-ordered particles.  Deposit: for one x-plane of the 3x3x3 stencil at a time every
//                    lane stores its 30 values as a column of a [30][33] tile, lane l then owns row l (one stencil
//                    point of one of rho/jx/jy/jz) and adds the source lanes of a cell; the sum stays in the owner's
//                    registers ACROSS warp iterations until the cell changes, so a cell costs 81 fp64 REDs per species
//                    however many particles it holds.  Particles that change cell during the step are listed.
//   3. k_list_particles  the listed particles: boundary-layer ones get the whole step with gathers from global memory,
//                    cell-crossing ones only the general 125-point deposit.
//
// The Esirkepov sums are factored (jx(i,j,k) = cumx[i] * Wx[j][k] etc.), which is the reference's running sums
// re-associated: results agree to a few ulp of each particle's largest term (parity bar 1e-12 of the array's max-abs).
#include <stdlib.h>
#include <algorithm>
#include "lpic_common.cuh"
#include "particle_math.cuh"
#include "push_tile.cuh"

namespace {

constexpr int PT = 256;  // threads of the permutation CTA

__device__ __forceinline__ int warp_incl_sum(int v) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int n = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += n;
    }
    return v;
}

// Position after the first half push and its grid coordinate: explicit intrinsics, so that the permutation kernel and
// the particle kernel agree bit for bit on the cell a particle belongs to.
__device__ __forceinline__ double half_push(double x, double cdt, double ig, double u) { return __fma_rn(__dmul_rn(cdt, ig), u, x); }
__device__ __forceinline__ double grid_coord(double x, double x0, double inv_d) { return __dmul_rn(__dsub_rn(x, x0), inv_d); }
__device__ __forceinline__ double nearest(double X) { return floor(__dadd_rn(X, 0.5)); }
// n / d from the rounded reciprocal r = 1/d plus one correction step: the correctly rounded quotient (Markstein) at 3 FMA-pipe
// operations instead of the ~25-instruction division sequence
__device__ __forceinline__ double div_rn(double n, double d, double r) {
    const double q = __dmul_rn(n, r);
    return __fma_rn(__fma_rn(-q, d, n), r, q);
}

struct TilePermArgs {
    int ps;  // stride of the attribute pointers (8: 64-byte records, 1: separate arrays)
    const double *x, *y, *z, *ux, *uy, *uz, *ig;
    const u8 *dead;
    const i64 *off, *npart;
    const double *x0, *y0, *z0;
    double cdt, inv_dx, inv_dy, inv_dz;
    int nx, ny, nz, nty, ntz, ntile;
    int *keys;        // arena scratch: key of every slot (-1: not in a tile)
    int *perm;        // arena: local slot numbers in (tile, cell) order
    int *tile_start;  // (npatch, ntile + 1)
    int *list;        // arena: the patch's list of particles for k_list_particles
    int *nlist;       // per patch
};

template <int TX, int TY, int TZ, int DIM>
__global__ void __launch_bounds__(PT) k_tile_perm(TilePermArgs a) {
    extern __shared__ int hist[];
    __shared__ int sw[PT / 32];
    constexpr int TC = TX * TY * TZ;
    const int p = blockIdx.x, tid = threadIdx.x;
    const i64 off = a.off[p];
    const int np = (int)a.npart[p];
    const int nkey = a.ntile * TC;
    for (int b = tid; b < nkey; b += PT) hist[b] = 0;
    __syncthreads();
    const double x0 = a.x0[p], y0 = a.y0[p], z0 = a.z0[p];
    auto key_of = [&](int ip) -> int {
        if (a.dead[off + ip]) return -1;
        double x, y, z, ux, uy, uz, ig;
        if (a.ps == LPIC_NREC) {  // one 64-byte record: four 16-byte loads from one or two lines
            const double2 *r = reinterpret_cast<const double2 *>(a.x + (off + ip) * LPIC_NREC);
            const double2 r0 = __ldg(r), r1 = __ldg(r + 1), r2 = __ldg(r + 2), r3 = __ldg(r + 3);
            x = r0.x; y = r0.y; z = DIM == 3 ? r1.x : 0.0; ux = r2.x; uy = r2.y; uz = r3.x; ig = r3.y;
        } else {
            const i64 ir = (off + ip) * a.ps;
            x = a.x[ir]; y = a.y[ir]; z = DIM == 3 ? a.z[ir] : 0.0; ux = a.ux[ir]; uy = a.uy[ir]; uz = DIM == 3 ? a.uz[ir] : 0.0; ig = a.ig[ir];
        }
        if (isnan(x) || isnan(y) || isnan(z)) return -1;
        const int ix = (int)nearest(grid_coord(half_push(x, a.cdt, ig, ux), x0, a.inv_dx));
        const int iy = (int)nearest(grid_coord(half_push(y, a.cdt, ig, uy), y0, a.inv_dy));
        const int iz = DIM == 3 ? (int)nearest(grid_coord(half_push(z, a.cdt, ig, uz), z0, a.inv_dz)) : 0;
        if ((unsigned)ix >= (unsigned)a.nx || (unsigned)iy >= (unsigned)a.ny || (unsigned)iz >= (unsigned)a.nz) return -2;
        const int tx = ix / TX, ty = iy / TY, tz = iz / TZ;
        return ((tx * a.nty + ty) * a.ntz + tz) * TC + ((ix - tx * TX) * TY + (iy - ty * TY)) * TZ + (iz - tz * TZ);
    };
    // four slots per thread and iteration: the loads of the four are independent and in flight together
    for (int base = 0; base < np; base += 4 * PT) {
        int k[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int ip = base + u * PT + tid;
            k[u] = ip < np ? key_of(ip) : -1;
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int ip = base + u * PT + tid;
            if (ip < np) a.keys[off + ip] = k[u];
            if (k[u] >= 0) atomicAdd(&hist[k[u]], 1);
            if (k[u] == -2) a.list[off + atomicAdd(&a.nlist[p], 1)] = ip | LIST_WHOLE_STEP;
        }
    }
    __syncthreads();
    // exclusive scan of the histogram in place
    int run = 0;
    for (int base = 0; base < nkey; base += PT) {
        const int b = base + tid;
        const int cnt = b < nkey ? hist[b] : 0;
        int v = warp_incl_sum(cnt);
        if ((tid & 31) == 31) sw[tid >> 5] = v;
        __syncthreads();
        int add = 0, tot = 0;
#pragma unroll
        for (int i = 0; i < PT / 32; i++) {
            if (i < (tid >> 5)) add += sw[i];
            tot += sw[i];
        }
        if (b < nkey) hist[b] = run + v + add - cnt;
        run += tot;
        __syncthreads();
    }
    int *ts = a.tile_start + (size_t)p * (a.ntile + 1);
    for (int t = tid; t < a.ntile; t += PT) ts[t] = hist[t * TC];
    if (tid == 0) ts[a.ntile] = run;
    __syncthreads();
    for (int base = 0; base < np; base += 4 * PT) {
        int k[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int ip = base + u * PT + tid;
            k[u] = ip < np ? a.keys[off + ip] : -1;
        }
#pragma unroll
        for (int u = 0; u < 4; u++)
            if (k[u] >= 0) a.perm[off + atomicAdd(&hist[k[u]], 1)] = base + u * PT + tid;
    }
}

// ---- record access of one warp iteration (particles t0 .. t0+31 of the processing order) ------------------------------
// A quad of lanes fetches one 64-byte record per load instruction (lane q of quad Q: piece q of particle 4Q + j), so a
// warp-level load touches 8 records = 16 full sectors however the slots are ordered.  The pieces go through the warp's
// (idle) reduction tile, 80 bytes per particle: conflict-free for the quad stores and for the owners' loads.  The quad's
// four slot numbers are parked behind them for the store phase, and w stays in the tile until the deposit: registers are
// what these kernels are short of (their spills miss the 28 KB of L1 left beside the shared memory: an L2 round trip each).
__device__ __forceinline__ void rec_load_warp(const double *__restrict__ rec, const int *__restrict__ perm, i64 off, int t0,
                                              int wlast, int lane, double *red) {
    const int q = lane & 3, qb = lane & ~3;
    double2 *stg = reinterpret_cast<double2 *>(red);
    int4 lq = make_int4(0, 0, 0, 0);
    if (t0 + qb < wlast) lq.x = perm[off + t0 + qb];
    if (t0 + qb + 1 < wlast) lq.y = perm[off + t0 + qb + 1];
    if (t0 + qb + 2 < wlast) lq.z = perm[off + t0 + qb + 2];
    if (t0 + qb + 3 < wlast) lq.w = perm[off + t0 + qb + 3];
    if (t0 + qb < wlast) stg[qb * 5 + q] = __ldg(reinterpret_cast<const double2 *>(rec + (off + lq.x) * LPIC_NREC) + q);
    if (t0 + qb + 1 < wlast) stg[(qb + 1) * 5 + q] = __ldg(reinterpret_cast<const double2 *>(rec + (off + lq.y) * LPIC_NREC) + q);
    if (t0 + qb + 2 < wlast) stg[(qb + 2) * 5 + q] = __ldg(reinterpret_cast<const double2 *>(rec + (off + lq.z) * LPIC_NREC) + q);
    if (t0 + qb + 3 < wlast) stg[(qb + 3) * 5 + q] = __ldg(reinterpret_cast<const double2 *>(rec + (off + lq.w) * LPIC_NREC) + q);
    reinterpret_cast<int4 *>(red + 32 * 10)[lane] = lq;
    __syncwarp();
}
// whole records back (z and w as they are in the tile): every store instruction writes full 32-byte sectors.  The caller
// has put this lane's new values into red[lane * 10 + ...] before.
__device__ __forceinline__ void rec_store_warp(double *__restrict__ rec, i64 off, int t0, int wlast, int lane, double *red) {
    const int q = lane & 3, qb = lane & ~3;
    const double2 *stg = reinterpret_cast<const double2 *>(red);
    __syncwarp();
    const int *lq = reinterpret_cast<const int *>(red + 32 * 10) + lane * 4;
#pragma unroll 1  // one record at a time: stores need no memory-level parallelism, and four in flight cost 28 registers
    for (int j = 0; j < 4; j++)
        if (t0 + qb + j < wlast) reinterpret_cast<double2 *>(rec + (off + lq[j]) * LPIC_NREC)[q] = stg[(qb + j) * 5 + q];
    __syncwarp();  // the deposit reuses the tile
}

struct TileArgs {
    Geom g;
    double *F;
    double *rec;  // records (x y z w | ux uy uz inv_gamma) of 64 bytes per slot, REC kernels only
    bool prefetch;
    const double *px0, *py0, *pz0;
    Slots s;
    const int *perm, *tile_start;
    int *list, *nlist;
    int nty, ntz, ntile;
    double dt, cdt, efactor, bfactor, inv_dx, inv_dy, inv_dz, q_dV, q_dydzdt, q_dxdzdt, q_dxdydt;
};

__device__ __forceinline__ void spline3(double d, double &a, double &b, double &c) {  // quadratic spline at offsets -1, 0, +1
    const double d2 = d * d;
    a = 0.5 * (0.25 + d2 + d);
    b = 0.75 - d2;
    c = 0.5 * (0.25 + d2 - d);
}

// 27-point weighted sum from the staged tile.  The three z-neighbours of every (x, y) stencil row are fetched as one aligned
// 16-byte pair plus a single: a broadcast-style LDS.128 costs 1.07 crossbar cycles per double against 1.52 for an LDS.64
// (profiles/microbench/lds_bench.cu).  tp = address of the pair, ts = address of the single, (wp0, wp1, ws) the z-weights
// in that order (see ZSplit).  Same 27 products as interp_field_safe_3d (unified_pusher_3d.c:79-106), summed z first.
template <int SX, int SY>
__device__ __forceinline__ double gather_tile(const double *__restrict__ tp, const double *__restrict__ ts, double fx0, double fx1,
                                              double fx2, double fy0, double fy1, double fy2, double wp0, double wp1, double ws) {
    double ax[3];
#pragma unroll
    for (int a = 0; a < 3; a++) {
        double ay[3];
#pragma unroll
        for (int b = 0; b < 3; b++) {
            const double2 v = *reinterpret_cast<const double2 *>(tp + a * SX + b * SY);
            ay[b] = wp0 * v.x + wp1 * v.y + ws * ts[a * SX + b * SY];
        }
        ax[a] = fy0 * ay[0] + fy1 * ay[1] + fy2 * ay[2];
    }
    return fx0 * ax[0] + fx1 * ax[1] + fx2 * ax[2];
}

// z-stencil starting at staged index bz (weights w0, w1, w2 for bz, bz+1, bz+2): even start -> pair (bz, bz+1) + single bz+2,
// odd start -> single bz + pair (bz+1, bz+2).  The staged rows have an even length, so the parity is that of bz alone.
struct ZSplit {
    int pair, single;    // offsets relative to the row's first staged z index
    double wp0, wp1, ws;
};
__device__ __forceinline__ ZSplit zsplit(int bz, double w0, double w1, double w2) {
    ZSplit z;
    const bool odd = bz & 1;
    z.pair = bz + (odd ? 1 : 0);
    z.single = bz + (odd ? 0 : 2);
    z.wp0 = odd ? w1 : w0;
    z.wp1 = odd ? w2 : w1;
    z.ws = odd ? w0 : w2;
    return z;
}

// What an owner lane needs to turn (cell inside the tile, x-plane) into the address of its stencil point
struct RowOwner {
    double *dst;     // this patch's jx / jy / jz / rho block
    int ax, ay, az;  // tile origin + the owner's stencil offset - 1 (ax is advanced by one per x-plane)
};
struct PadDims {  // padded patch sizes: passed along from the kernel parameters (constant bank) instead of living in registers
    int NX, NY, NZ;
};

// fp64 reduction into global memory without a return value.  Spelled as PTX: inside a non-inlined function the compiler
// no longer knows that the pointer is a global one and emits a generic ATOM with a shared-memory CAS fallback.
__device__ __forceinline__ void red_add(double *p, double v) {
    asm volatile("red.global.add.f64 [%0], %1;" ::"l"(__cvta_generic_to_global(p)), "d"(v) : "memory");
}

// one RED for a finished cell (code = cell inside the tile; -1: nothing carried yet)
template <int TY, int TZ>
__device__ __forceinline__ void flush_cell(const RowOwner &o, const PadDims &d, int code, double sum) {
    if (code < 0) return;
    const int lz = code % TZ, ly = (code / TZ) % TY, lx = code / (TZ * TY);
    const int id = (wrapneg(o.ax + lx, d.NX) * d.NY + wrapneg(o.ay + ly, d.NY)) * d.NZ + wrapneg(o.az + lz, d.NZ);
    red_add(o.dst + id, sum);
}

// Owner lane: add this iteration's 32 source lanes of one row to the running sum of the current cell; a set bit in
// `heads` marks a source lane that starts a new cell: the finished cell is flushed and the sum restarts.  One copy of
// this in the instruction stream serves the three x-planes.  Straight-line per group of four source lanes; a group
// that contains a head splits its four values around it.
template <int TY, int TZ>
__device__ __noinline__ double row_sum(double acc, unsigned heads, const double *__restrict__ tile, int cur, RowOwner o, PadDims d) {
    // tile = the warp's [30][33] reduction tile followed by its 32 cell codes; this lane's row and the codes are derived here
    // rather than handed in: four registers fewer to keep alive around the call
    const double *__restrict__ row = tile + (threadIdx.x & 31) * 33;
    const int *__restrict__ codes = reinterpret_cast<const int *>(tile + 30 * 33);
#pragma unroll
    for (int g4 = 0; g4 < 32; g4 += 4) {
        const unsigned mm = (heads >> g4) & 0xFu;
        const double v0 = row[g4], v1 = row[g4 + 1], v2 = row[g4 + 2], v3 = row[g4 + 3];
        if (mm == 0u) {
            acc += (v0 + v1) + (v2 + v3);
        } else if ((mm & (mm - 1u)) == 0u) {  // one head, at position t: sources before it close the current cell
            const int t = __ffs(mm) - 1;
            const double lo = (t > 0 ? v0 : 0.0) + (t > 1 ? v1 : 0.0) + (t > 2 ? v2 : 0.0);
            flush_cell<TY, TZ>(o, d, cur, acc + lo);
            cur = codes[g4 + t];
            acc = ((t > 0 ? 0.0 : v0) + (t > 1 ? 0.0 : v1)) + ((t > 2 ? 0.0 : v2) + v3);
        } else {  // several cells start inside the group (fewer than 4 particles per cell): rolled, re-reads the row
#pragma unroll 1
            for (int t = 0; t < 4; t++) {
                if ((mm >> t) & 1u) {
                    flush_cell<TY, TZ>(o, d, cur, acc);
                    acc = 0.0;
                    cur = codes[g4 + t];
                }
                acc += row[g4 + t];
            }
        }
    }
    return acc;
}

template <int TX, int TY, int TZ, int NW, bool WRITE_PART, bool REC>
__global__ void __launch_bounds__(NW * 32, TILE_MIN_CTAS) k_push_tile(const TileArgs a) {
    const bool PREFETCH = a.prefetch;
    constexpr int EX = TX + 3, EY = TY + 3, EZ = TZ + 4, EN = EX * EY * EZ;  // EZ: TZ + 3 nodes, padded to an even row length
    constexpr int SX = EY * EZ, SY = EZ;  // strides (in doubles) of the staged tile
    extern __shared__ __align__(16) double smem[];
    double *eb = smem;                                            // [6][EX][EY][EZ]
    double *red = smem + 6 * EN + (threadIdx.x >> 5) * (30 * 33 + 16);  // this warp's [30][33] reduction tile + its 32 cell codes
    int *codes = (int *)(red + 30 * 33);
    const Geom &g = a.g;
    const int p = blockIdx.x / a.ntile, tile = blockIdx.x - p * a.ntile;
    const int *ts = a.tile_start + (size_t)p * (a.ntile + 1) + tile;
    const int first = ts[0], last = ts[1];
    if (first == last) return;
    const int tz = tile % a.ntz, ty = (tile / a.ntz) % a.nty, tx = tile / (a.ntz * a.nty);
    const int ox = tx * TX, oy = ty * TY, oz = tz * TZ;
    const size_t stride = (size_t)g.npatch * g.ncell;
    double *Fp = a.F + (size_t)p * g.ncell;
    __shared__ double s_org[3];
    if (threadIdx.x == 0) { s_org[0] = a.px0[p]; s_org[1] = a.py0[p]; s_org[2] = a.pz0[p]; }
    // ---- stage E/B of the tile and its halo: nodes [o-2, o+T] per axis, logical order --------------------------------
    for (int idx = threadIdx.x; idx < EN; idx += NW * 32) {
        const int lz = idx % EZ, ly = (idx / EZ) % EY, lx = idx / (EZ * EY);  // (the pad column lz = TZ + 3 is loaded too: one more guard node)
        const int gx = ox - 2 + lx, gy = oy - 2 + ly, gz = oz - 2 + lz;
        if (gx < g.nx + g.ng && gy < g.ny + g.ng && gz < g.nz + g.ng) {  // (the low side is always inside: ng >= 2)
            const double *src = Fp + (wrapneg(gx, g.NX) * g.NY + wrapneg(gy, g.NY)) * g.NZ + wrapneg(gz, g.NZ);
#pragma unroll
            for (int c = 0; c < 6; c++) eb[c * EN + idx] = __ldg(src + c * stride);
        }
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int per = ((last - first + NW * 32 - 1) / (NW * 32)) * 32;  // particles per warp, a multiple of 32
    const int wfirst = first + warp * per, wlast = min(last, wfirst + per);
    if (wfirst >= wlast) return;
    const i64 off = a.s.off[p];
    // (the patch origin is read from shared memory where it is used: three doubles fewer to keep in registers across the
    // iteration -- whatever this kernel spills misses the small L1 and costs an L2 round trip)
    const volatile double *org = s_org;
    // which row of the reduction tile does this lane own?  [0,9) rho(j,k)  [9,15) jy(j<2,k)  [15,21) jz(j,k<2)  [21,30) jx(j,k)
    int comp, sj, sk;
    if (lane < 9) { comp = LPIC_RHO; sj = lane / 3; sk = lane - 3 * sj; }
    else if (lane < 15) { comp = LPIC_JY; sj = (lane - 9) / 3; sk = (lane - 9) - 3 * sj; }
    else if (lane < 21) { comp = LPIC_JZ; sj = (lane - 15) >> 1; sk = (lane - 15) & 1; }
    else { comp = LPIC_JX; sj = (lane - 21) / 3; sk = (lane - 21) - 3 * sj; }
    RowOwner own;
    own.dst = Fp + comp * stride;
    own.ax = ox - 1; own.ay = oy + sj - 1; own.az = oz + sk - 1;
    PadDims pd;
    pd.NX = g.NX; pd.NY = g.NY; pd.NZ = g.NZ;
    double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0;  // the owner's running sums of the current cell, one per x-plane
    int ccode = -1;                             // that cell (-1: none yet)
    for (int t0 = wfirst; t0 < wlast; t0 += 32) {
        const bool active = t0 + lane < wlast;
        double x = 0, y = 0, z = 0, ux = 0, uy = 0, uz = 0, ig = 1, w = 0;
        int local = 0, cx = 0, cy = 0, cz = 0;
        if (active && (!REC || WRITE_PART)) local = a.perm[off + t0 + lane];
        if (REC) {
            rec_load_warp(a.rec, a.perm, off, t0, wlast, lane, red);
            if (PREFETCH && t0 + 32 + lane < wlast)  // the next iteration's record on its way into L2 while this one computes
                asm volatile("prefetch.global.L2 [%0];" ::"l"(a.rec + (off + a.perm[off + t0 + 32 + lane]) * LPIC_NREC));
            if (active) {
                const double2 *stg = reinterpret_cast<const double2 *>(red);
                const double2 r0 = stg[lane * 5], r2 = stg[lane * 5 + 2], r3 = stg[lane * 5 + 3];
                x = r0.x; y = r0.y; z = red[lane * 10 + 2]; ux = r2.x; uy = r2.y; uz = r3.x; ig = r3.y;
            }
        }
        if (active) {
            const i64 ip = off + local;
            if (!REC) {
                x = a.s.x[ip * a.s.ps]; y = a.s.y[ip * a.s.ps]; z = a.s.z[ip * a.s.ps];
                ux = a.s.ux[ip * a.s.ps]; uy = a.s.uy[ip * a.s.ps]; uz = a.s.uz[ip * a.s.ps]; ig = a.s.ig[ip * a.s.ps];
                w = a.s.w[ip * a.s.ps];
            }
            x = half_push(x, a.cdt, ig, ux); y = half_push(y, a.cdt, ig, uy); z = half_push(z, a.cdt, ig, uz);
            const double X = grid_coord(x, org[0], a.inv_dx), Y = grid_coord(y, org[1], a.inv_dy), Z = grid_coord(z, org[2], a.inv_dz);
            const double rX = nearest(X), rY = nearest(Y), rZ = nearest(Z), fX = floor(X), fY = floor(Y), fZ = floor(Z);
            cx = (int)rX; cy = (int)rY; cz = (int)rZ;
            double gx0, gx1, gx2, gy0, gy1, gy2, gz0, gz1, gz2, hx0, hx1, hx2, hy0, hy1, hy2, hz0, hz1, hz2;
            spline3(rX - X, gx0, gx1, gx2); spline3(fX - X + 0.5, hx0, hx1, hx2);
            spline3(rY - Y, gy0, gy1, gy2); spline3(fY - Y + 0.5, hy0, hy1, hy2);
            spline3(rZ - Z, gz0, gz1, gz2); spline3(fZ - Z + 0.5, hz0, hz1, hz2);
            // staged index of the first stencil node: node-centred (g) r-1 -> r-1-(o-2), cell-centred (h) f-1 -> f-1-(o-2)
            const int bgx = (cx - ox + 1) * SX, bhx = ((int)fX - ox + 1) * SX;
            const int bgy = (cy - oy + 1) * SY, bhy = ((int)fY - oy + 1) * SY;
            const ZSplit zg = zsplit(cz - oz + 1, gz0, gz1, gz2), zh = zsplit((int)fZ - oz + 1, hz0, hz1, hz2);
            // ex(h,g,g) ey(g,h,g) ez(g,g,h) bx(g,h,h) by(h,g,h) bz(h,h,g)  (unified_pusher_3d.c:190-195)
            double f[6];
            const double *t;
            t = eb + 0 * EN + bhx + bgy; f[0] = gather_tile<SX, SY>(t + zg.pair, t + zg.single, hx0, hx1, hx2, gy0, gy1, gy2, zg.wp0, zg.wp1, zg.ws);
            t = eb + 1 * EN + bgx + bhy; f[1] = gather_tile<SX, SY>(t + zg.pair, t + zg.single, gx0, gx1, gx2, hy0, hy1, hy2, zg.wp0, zg.wp1, zg.ws);
            t = eb + 2 * EN + bgx + bgy; f[2] = gather_tile<SX, SY>(t + zh.pair, t + zh.single, gx0, gx1, gx2, gy0, gy1, gy2, zh.wp0, zh.wp1, zh.ws);
            t = eb + 3 * EN + bgx + bhy; f[3] = gather_tile<SX, SY>(t + zh.pair, t + zh.single, gx0, gx1, gx2, hy0, hy1, hy2, zh.wp0, zh.wp1, zh.ws);
            t = eb + 4 * EN + bhx + bgy; f[4] = gather_tile<SX, SY>(t + zh.pair, t + zh.single, hx0, hx1, hx2, gy0, gy1, gy2, zh.wp0, zh.wp1, zh.ws);
            t = eb + 5 * EN + bhx + bhy; f[5] = gather_tile<SX, SY>(t + zg.pair, t + zg.single, hx0, hx1, hx2, hy0, hy1, hy2, zg.wp0, zg.wp1, zg.ws);
            if (WRITE_PART) {
#pragma unroll
                for (int c = 0; c < 6; c++) a.s.part[c][ip] = f[c];
            }
            boris_kick(ux, uy, uz, ig, f, a.efactor, a.bfactor);
            x += a.cdt * ig * ux; y += a.cdt * ig * uy; z += a.cdt * ig * uz;
            if (!REC) {
                a.s.ux[ip * a.s.ps] = ux; a.s.uy[ip * a.s.ps] = uy; a.s.uz[ip * a.s.ps] = uz; a.s.ig[ip * a.s.ps] = ig;
                a.s.x[ip * a.s.ps] = x; a.s.y[ip * a.s.ps] = y; a.s.z[ip * a.s.ps] = z;
            }
        }
        if (REC) {
            if (active) {
                double2 *stg = reinterpret_cast<double2 *>(red);
                stg[lane * 5] = make_double2(x, y); red[lane * 10 + 2] = z;
                stg[lane * 5 + 2] = make_double2(ux, uy); stg[lane * 5 + 3] = make_double2(uz, ig);
                w = red[lane * 10 + 3];
            }
            rec_store_warp(a.rec, off, t0, wlast, lane, red);
        }
        // ---- deposit set-up (current_deposit.h:341-373): the path from x - v dt/2 to x + v dt/2 ------------------------
        // Same expressions as the reference (and k_particles): a slow particle's current is proportional to X1 - X0, the
        // difference of two separately rounded coordinates, so the quotients have to round the way a division does.
        const double vx = ux * LPIC_C_LIGHT * ig, vy = uy * LPIC_C_LIGHT * ig, vz = uz * LPIC_C_LIGHT * ig;
        const double x0 = org[0], y0 = org[1], z0 = org[2];
        const double X0 = div_rn(x - vx * 0.5 * a.dt - x0, g.dx, a.inv_dx), X1 = div_rn(x + vx * 0.5 * a.dt - x0, g.dx, a.inv_dx);
        const double Y0 = div_rn(y - vy * 0.5 * a.dt - y0, g.dy, a.inv_dy), Y1 = div_rn(y + vy * 0.5 * a.dt - y0, g.dy, a.inv_dy);
        const double Z0 = div_rn(z - vz * 0.5 * a.dt - z0, g.dz, a.inv_dz), Z1 = div_rn(z + vz * 0.5 * a.dt - z0, g.dz, a.inv_dz);
        const double rX0 = floor(X0 + 0.5), rY0 = floor(Y0 + 0.5), rZ0 = floor(Z0 + 0.5);
        // fast: the nearest node does not change along the path AND it is the node the particle was keyed by (recomputing
        // the start from the end position can round to the other side of a cell boundary)
        const bool fast = active && floor(X1 + 0.5) == rX0 && floor(Y1 + 0.5) == rY0 && floor(Z1 + 0.5) == rZ0 &&
                          (int)rX0 == cx && (int)rY0 == cy && (int)rZ0 == cz;
        {   // everything else takes the general deposit (k_list_particles); one warp-aggregated append
            const unsigned cm = __ballot_sync(0xffffffffu, active && !fast);
            if (cm) {
                int basepos = 0;
                if (lane == __ffs(cm) - 1) basepos = atomicAdd(&a.nlist[p], __popc(cm));
                basepos = __shfl_sync(0xffffffffu, basepos, __ffs(cm) - 1);
                if (active && !fast) a.list[off + basepos + __popc(cm & ((1u << lane) - 1u))] = REC ? a.perm[off + t0 + lane] : local;
            }
        }
        // cells = runs of consecutive fast lanes with the same code; lanes outside the fast path add zeros and never
        // start a run; the first fast lane continues the run carried over from the previous iteration if the cell matches
        const int code = fast ? ((cx - ox) * TY + (cy - oy)) * TZ + (cz - oz) : -1;
        const unsigned fm = __ballot_sync(0xffffffffu, fast);
        if (fm == 0u) continue;
        const unsigned before = fm & ((1u << lane) - 1u);
        const int pf = before ? 31 - __clz(before) : 0;
        int pcode = __shfl_sync(0xffffffffu, code, pf);
        if (!before) pcode = ccode;
        const unsigned heads = __ballot_sync(0xffffffffu, fast && code != pcode);
        codes[lane] = code;
        const int lastcode = __shfl_sync(0xffffffffu, code, 31 - __clz(fm));
        // shape factors at the start (S0) and their change (DS); no cell crossing: both live on the same 3 nodes
        double S0x[3], S0y[3], S0z[3], DSx[3], DSy[3], DSz[3];
        {
            double s1[3];
            spline3(rX0 - X0, S0x[0], S0x[1], S0x[2]); spline3(rX0 - X1, s1[0], s1[1], s1[2]);
#pragma unroll
            for (int i = 0; i < 3; i++) DSx[i] = s1[i] - S0x[i];
            spline3(rY0 - Y0, S0y[0], S0y[1], S0y[2]); spline3(rY0 - Y1, s1[0], s1[1], s1[2]);
#pragma unroll
            for (int i = 0; i < 3; i++) DSy[i] = s1[i] - S0y[i];
            spline3(rZ0 - Z0, S0z[0], S0z[1], S0z[2]); spline3(rZ0 - Z1, s1[0], s1[1], s1[2]);
#pragma unroll
            for (int i = 0; i < 3; i++) DSz[i] = s1[i] - S0z[i];
        }
        const double wq = fast ? w : 0.0;
        const double cd = a.q_dV * wq, fdx = a.q_dydzdt * wq, fdy = a.q_dxdzdt * wq, fdz = a.q_dxdydt * wq;
        // running sums of the reference's jyb / jzb loops, factored: jy(i,j,k) = cumy[j] * Wy_i[k], jz(i,j,k) = cumz[k] * Tz_i[j]
        const double cumy0 = -fdy * DSy[0], cumy1 = cumy0 - fdy * DSy[1];
        const double cumz0 = -fdz * DSz[0], cumz1 = cumz0 - fdz * DSz[1];
        double cumx = 0.0;
#pragma unroll
        for (int i = 0; i < 3; i++) {
            const double ax = S0x[i] + 0.5 * DSx[i], cxx = 0.5 * S0x[i] + LPIC_ONE_THIRD * DSx[i];
            const double rx = cd * (S0x[i] + DSx[i]);
            cumx -= fdx * DSx[i];
            double wy[3];
#pragma unroll
            for (int k = 0; k < 3; k++) wy[k] = ax * S0z[k] + cxx * DSz[k];
#pragma unroll
            for (int j = 0; j < 3; j++) {
                const double rxy = rx * (S0y[j] + DSy[j]);
                const double tzj = ax * S0y[j] + cxx * DSy[j];
                const double ay = S0y[j] + 0.5 * DSy[j], cyy = 0.5 * S0y[j] + LPIC_ONE_THIRD * DSy[j];
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    red[(j * 3 + k) * 33 + lane] = rxy * (S0z[k] + DSz[k]);
                    if (i < 2) red[(21 + j * 3 + k) * 33 + lane] = cumx * (ay * S0z[k] + cyy * DSz[k]);
                    if (j < 2) red[(9 + j * 3 + k) * 33 + lane] = (j == 0 ? cumy0 : cumy1) * wy[k];
                    if (k < 2) red[(15 + j * 2 + k) * 33 + lane] = (k == 0 ? cumz0 : cumz1) * tzj;
                }
            }
            __syncwarp();
            if (lane < (i < 2 ? 30 : 21)) {
                RowOwner o = own;
                o.ax += i;
                const double acc = row_sum<TY, TZ>(i == 0 ? acc0 : (i == 1 ? acc1 : acc2), heads, red, ccode, o, pd);
                if (i == 0) acc0 = acc; else if (i == 1) acc1 = acc; else acc2 = acc;
            }
            __syncwarp();
        }
        ccode = lastcode;
    }
    flush_cell<TY, TZ>(own, pd, ccode, acc0);
    own.ax++;
    flush_cell<TY, TZ>(own, pd, ccode, acc1);
    own.ax++;
    if (lane < 21) flush_cell<TY, TZ>(own, pd, ccode, acc2);
}

// The listed particles of every patch: entries tagged LIST_WHOLE_STEP get the whole step with gathers from global memory
// (they sit in the boundary layer outside the tiles), the others were pushed by k_push_tile and only need the general
// deposit.  Persistent grid: CTAs stride over the patches, threads over each patch's (short) list.
template <bool WRITE_PART, int DIM>
__global__ void __launch_bounds__(128) k_list_particles(Geom g, double *__restrict__ F, const double *__restrict__ px0,
                                                        const double *__restrict__ py0, const double *__restrict__ pz0, Slots s,
                                                        const int *__restrict__ list, const int *__restrict__ nlist, double dt,
                                                        double q, double m) {
    DepositCoef3 k;
    k.q_dV = q / (g.dx * g.dy * g.dz); k.q_dydzdt = q / (g.dy * g.dz * dt);
    k.q_dxdzdt = q / (g.dx * g.dz * dt); k.q_dxdydt = q / (g.dx * g.dy * dt); k.dt = dt;
    DepositCoef2 k2;
    k2.q_dxdy = q / (g.dx * g.dy); k2.q_dydt = q / (g.dy * dt); k2.q_dxdt = q / (g.dx * dt); k2.dt = dt;
    const double cdt = LPIC_C_LIGHT * 0.5 * dt, efactor = q * dt / (2 * m * LPIC_C_LIGHT), bfactor = q * dt / (2 * m);
    for (int p = blockIdx.x; p < g.npatch; p += gridDim.x) {
        const int n = nlist[p];
        if (n == 0) continue;
        const PatchView v = patch_view(g, F, px0, py0, pz0, p);
        for (int t = threadIdx.x; t < n; t += blockDim.x) {
            const int e = list[s.off[p] + t];
            const i64 ip = s.off[p] + (e & ~LIST_WHOLE_STEP);
            double x, y, z, ux, uy, uz, ig, w;
            double2 *r = reinterpret_cast<double2 *>(s.x + ip * LPIC_NREC);  // (records only)
            if (s.ps == LPIC_NREC) {
                const double2 r0 = r[0], r1 = r[1], r2 = r[2], r3 = r[3];
                x = r0.x; y = r0.y; z = DIM == 3 ? r1.x : 0.0; w = r1.y; ux = r2.x; uy = r2.y; uz = r3.x; ig = r3.y;
            } else {
                x = s.x[ip]; y = s.y[ip]; z = DIM == 3 ? s.z[ip] : 0.0; ux = s.ux[ip]; uy = s.uy[ip]; uz = s.uz[ip]; ig = s.ig[ip]; w = s.w[ip];
            }
            if (e & LIST_WHOLE_STEP) {
                x = half_push(x, cdt, ig, ux); y = half_push(y, cdt, ig, uy);
                if (DIM == 3) z = half_push(z, cdt, ig, uz);
                double eb[6];
                gather_eb<DIM>(g, v, x, y, z, eb);
                if (WRITE_PART) {
#pragma unroll
                    for (int c = 0; c < 6; c++) s.part[c][ip] = eb[c];
                }
                boris_kick(ux, uy, uz, ig, eb, efactor, bfactor);
                x += cdt * ig * ux; y += cdt * ig * uy;
                if (DIM == 3) z += cdt * ig * uz;
                if (s.ps == LPIC_NREC) {
                    r[0] = make_double2(x, y);
                    if (DIM == 3) r[1] = make_double2(z, w);
                    r[2] = make_double2(ux, uy); r[3] = make_double2(uz, ig);
                } else {
                    s.ux[ip] = ux; s.uy[ip] = uy; s.uz[ip] = uz; s.ig[ip] = ig;
                    s.x[ip] = x; s.y[ip] = y;
                    if (DIM == 3) s.z[ip] = z;
                }
            }
            if (DIM == 3) deposit3(g, v, k, x, y, z, ux, uy, uz, ig, w);
            else deposit2(g, v, k2, x, y, ux, uy, uz, ig, w);
        }
    }
}

// ---- 2D twin -----------------------------------------------------------------------------------------------------------
// Same organisation in two dimensions (reference: core/pusher/unified/unified_pusher_2d.c:157-330,
// core/current/current_deposit.h:150-268): tile = TX x TY cells (y contiguous), 6 x 9 gather points from the staged tile
// (the three y-neighbours of a stencil row as an aligned LDS.128 + LDS.64), and the whole 3x3 stencil of a particle that
// stays in its cell is ONE round of the [30][33] reduction tile -- rho 9 rows, jx 6 (the last x row is sum(DSx) = 0 up to
// rounding), jy 6 (last y column likewise), jz 9 -- with one carried sum per owner lane.
template <int TX, int TY, int NW, bool WRITE_PART, bool REC>
__global__ void __launch_bounds__(NW * 32, 4) k_push_tile2d(const TileArgs a) {
    constexpr int EX = TX + 3, EY = TY + 4, EN = EX * EY;  // EY: TY + 3 nodes, padded to an even row length
    constexpr int SX = EY;
    extern __shared__ __align__(16) double smem[];
    double *eb = smem;                                            // [6][EX][EY]
    double *red = smem + 6 * EN + (threadIdx.x >> 5) * (30 * 33 + 16);  // this warp's [30][33] reduction tile + its 32 cell codes
    int *codes = (int *)(red + 30 * 33);
    const Geom &g = a.g;
    const int p = blockIdx.x / a.ntile, tile = blockIdx.x - p * a.ntile;
    const int *ts = a.tile_start + (size_t)p * (a.ntile + 1) + tile;
    const int first = ts[0], last = ts[1];
    if (first == last) return;
    const int ty = tile % a.nty, tx = tile / a.nty;
    const int ox = tx * TX, oy = ty * TY;
    const size_t stride = (size_t)g.npatch * g.ncell;
    double *Fp = a.F + (size_t)p * g.ncell;
    for (int idx = threadIdx.x; idx < EN; idx += NW * 32) {  // stage E/B: nodes [o-2, o+T] per axis, logical order
        const int ly = idx % EY, lx = idx / EY;
        const int gx = ox - 2 + lx, gy = oy - 2 + ly;
        if (gx < g.nx + g.ng && gy < g.ny + g.ng) {
            const double *src = Fp + wrapneg(gx, g.NX) * g.NY + wrapneg(gy, g.NY);
#pragma unroll
            for (int c = 0; c < 6; c++) eb[c * EN + idx] = __ldg(src + c * stride);
        }
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int per = ((last - first + NW * 32 - 1) / (NW * 32)) * 32;
    const int wfirst = first + warp * per, wlast = min(last, wfirst + per);
    if (wfirst >= wlast) return;
    const i64 off = a.s.off[p];
    const double x0 = a.px0[p], y0 = a.py0[p];
    // rows: [0,9) rho(i,j)  [9,15) jx(i<2,j)  [15,21) jy(i,j<2)  [21,30) jz(i,j)
    int comp, si, sj;
    if (lane < 9) { comp = LPIC_RHO; si = lane / 3; sj = lane - 3 * si; }
    else if (lane < 15) { comp = LPIC_JX; si = (lane - 9) / 3; sj = (lane - 9) - 3 * si; }
    else if (lane < 21) { comp = LPIC_JY; si = (lane - 15) >> 1; sj = (lane - 15) & 1; }
    else { comp = LPIC_JZ; si = (lane - 21) / 3; sj = (lane - 21) - 3 * si; }
    RowOwner own;
    own.dst = Fp + comp * stride;
    own.ax = ox + si - 1; own.ay = oy + sj - 1; own.az = 0;
    PadDims pd;
    pd.NX = g.NX; pd.NY = g.NY; pd.NZ = 1;
    double acc = 0.0;
    int ccode = -1;
    for (int t0 = wfirst; t0 < wlast; t0 += 32) {
        const bool active = t0 + lane < wlast;
        double x = 0, y = 0, ux = 0, uy = 0, uz = 0, ig = 1, w = 0;
        int local = 0, cx = 0, cy = 0;
        if (REC) {
            rec_load_warp(a.rec, a.perm, off, t0, wlast, lane, red);
            if (active) {
                const double2 *stg = reinterpret_cast<const double2 *>(red);
                const double2 r0 = stg[lane * 5], r2 = stg[lane * 5 + 2], r3 = stg[lane * 5 + 3];
                x = r0.x; y = r0.y; ux = r2.x; uy = r2.y; uz = r3.x; ig = r3.y;
            }
        }
        if (active) {
            local = a.perm[off + t0 + lane];
            const i64 ip = off + local;
            if (!REC) {
                x = a.s.x[ip * a.s.ps]; y = a.s.y[ip * a.s.ps];
                ux = a.s.ux[ip * a.s.ps]; uy = a.s.uy[ip * a.s.ps]; uz = a.s.uz[ip * a.s.ps]; ig = a.s.ig[ip * a.s.ps];
                w = a.s.w[ip * a.s.ps];
            }
            x = half_push(x, a.cdt, ig, ux); y = half_push(y, a.cdt, ig, uy);
            const double X = grid_coord(x, x0, a.inv_dx), Y = grid_coord(y, y0, a.inv_dy);
            const double rX = nearest(X), rY = nearest(Y), fX = floor(X), fY = floor(Y);
            cx = (int)rX; cy = (int)rY;
            double gx[3], gy0, gy1, gy2, hx[3], hy0, hy1, hy2;
            spline3(rX - X, gx[0], gx[1], gx[2]); spline3(fX - X + 0.5, hx[0], hx[1], hx[2]);
            spline3(rY - Y, gy0, gy1, gy2); spline3(fY - Y + 0.5, hy0, hy1, hy2);
            const int bgx = (cx - ox + 1) * SX, bhx = ((int)fX - ox + 1) * SX;
            const ZSplit yg = zsplit(cy - oy + 1, gy0, gy1, gy2), yh = zsplit((int)fY - oy + 1, hy0, hy1, hy2);
            auto gather9 = [&](const double *t, const double *fx, const ZSplit &ys) -> double {
                double s = 0.0;
#pragma unroll
                for (int i = 0; i < 3; i++) {
                    const double2 v = *reinterpret_cast<const double2 *>(t + i * SX + ys.pair);
                    s += fx[i] * (ys.wp0 * v.x + ys.wp1 * v.y + ys.ws * t[i * SX + ys.single]);
                }
                return s;
            };
            // ex(h,g) ey(g,h) ez(g,g) bx(g,h) by(h,g) bz(h,h)  (unified_pusher_2d.c:120-150)
            double f[6];
            f[0] = gather9(eb + 0 * EN + bhx, hx, yg);
            f[1] = gather9(eb + 1 * EN + bgx, gx, yh);
            f[2] = gather9(eb + 2 * EN + bgx, gx, yg);
            f[3] = gather9(eb + 3 * EN + bgx, gx, yh);
            f[4] = gather9(eb + 4 * EN + bhx, hx, yg);
            f[5] = gather9(eb + 5 * EN + bhx, hx, yh);
            if (WRITE_PART) {
#pragma unroll
                for (int c = 0; c < 6; c++) a.s.part[c][ip] = f[c];
            }
            boris_kick(ux, uy, uz, ig, f, a.efactor, a.bfactor);
            x += a.cdt * ig * ux; y += a.cdt * ig * uy;
            if (!REC) {
                a.s.ux[ip * a.s.ps] = ux; a.s.uy[ip * a.s.ps] = uy; a.s.uz[ip * a.s.ps] = uz; a.s.ig[ip * a.s.ps] = ig;
                a.s.x[ip * a.s.ps] = x; a.s.y[ip * a.s.ps] = y;
            }
        }
        if (REC) {
            if (active) {
                double2 *stg = reinterpret_cast<double2 *>(red);
                stg[lane * 5] = make_double2(x, y);
                stg[lane * 5 + 2] = make_double2(ux, uy); stg[lane * 5 + 3] = make_double2(uz, ig);
                w = red[lane * 10 + 3];
            }
            rec_store_warp(a.rec, off, t0, wlast, lane, red);
        }
        // ---- deposit set-up (current_deposit.h:196-222) ----------------------------------------------------------------
        const double vx = ux * LPIC_C_LIGHT * ig, vy = uy * LPIC_C_LIGHT * ig, vz = uz * LPIC_C_LIGHT * ig;
        const double X0 = div_rn(x - vx * 0.5 * a.dt - x0, g.dx, a.inv_dx), X1 = div_rn(x + vx * 0.5 * a.dt - x0, g.dx, a.inv_dx);
        const double Y0 = div_rn(y - vy * 0.5 * a.dt - y0, g.dy, a.inv_dy), Y1 = div_rn(y + vy * 0.5 * a.dt - y0, g.dy, a.inv_dy);
        const double rX0 = floor(X0 + 0.5), rY0 = floor(Y0 + 0.5);
        const bool fast = active && floor(X1 + 0.5) == rX0 && floor(Y1 + 0.5) == rY0 && (int)rX0 == cx && (int)rY0 == cy;
        {
            const unsigned cm = __ballot_sync(0xffffffffu, active && !fast);
            if (cm) {
                int basepos = 0;
                if (lane == __ffs(cm) - 1) basepos = atomicAdd(&a.nlist[p], __popc(cm));
                basepos = __shfl_sync(0xffffffffu, basepos, __ffs(cm) - 1);
                if (active && !fast) a.list[off + basepos + __popc(cm & ((1u << lane) - 1u))] = local;
            }
        }
        const int code = fast ? (cx - ox) * TY + (cy - oy) : -1;
        const unsigned fm = __ballot_sync(0xffffffffu, fast);
        if (fm == 0u) continue;
        const unsigned before = fm & ((1u << lane) - 1u);
        int pcode = __shfl_sync(0xffffffffu, code, before ? 31 - __clz(before) : 0);
        if (!before) pcode = ccode;
        const unsigned heads = __ballot_sync(0xffffffffu, fast && code != pcode);
        codes[lane] = code;
        const int lastcode = __shfl_sync(0xffffffffu, code, 31 - __clz(fm));
        double S0x[3], S0y[3], DSx[3], DSy[3];
        {
            double s1[3];
            spline3(rX0 - X0, S0x[0], S0x[1], S0x[2]); spline3(rX0 - X1, s1[0], s1[1], s1[2]);
#pragma unroll
            for (int i = 0; i < 3; i++) DSx[i] = s1[i] - S0x[i];
            spline3(rY0 - Y0, S0y[0], S0y[1], S0y[2]); spline3(rY0 - Y1, s1[0], s1[1], s1[2]);
#pragma unroll
            for (int i = 0; i < 3; i++) DSy[i] = s1[i] - S0y[i];
        }
        const double wq = fast ? w : 0.0;
        // 2D prefactors: q/(dx dy), q/(dy dt), q/(dx dt) (DepositCoef2), passed in q_dV, q_dydzdt, q_dxdzdt
        const double cd = a.q_dV * wq, fdx = a.q_dydzdt * wq, fdy = a.q_dxdzdt * wq, fvz = cd * vz;
        const double one_twelfth = 1.0 / 12.0;
        const double cumy0 = -fdy * DSy[0], cumy1 = cumy0 - fdy * DSy[1];
        double cumx = 0.0;
#pragma unroll
        for (int i = 0; i < 3; i++) {
            const double ai = S0x[i] + 0.5 * DSx[i], t12 = one_twelfth * DSx[i], rx = cd * (S0x[i] + DSx[i]);
            cumx -= fdx * DSx[i];
#pragma unroll
            for (int j = 0; j < 3; j++) {
                const double bj = S0y[j] + 0.5 * DSy[j];
                red[(i * 3 + j) * 33 + lane] = rx * (S0y[j] + DSy[j]);
                if (i < 2) red[(9 + i * 3 + j) * 33 + lane] = cumx * bj;
                if (j < 2) red[(15 + i * 2 + j) * 33 + lane] = (j == 0 ? cumy0 : cumy1) * ai;
                red[(21 + i * 3 + j) * 33 + lane] = fvz * (ai * bj + t12 * DSy[j]);
            }
        }
        __syncwarp();
        if (lane < 30) acc = row_sum<TY, 1>(acc, heads, red, ccode, own, pd);
        __syncwarp();
        ccode = lastcode;
    }
    if (lane < 30) flush_cell<TY, 1>(own, pd, ccode, acc);
}

template <int TX, int TY, int NW>
int launch_tiles2d(lpic_ctx *c, Species &sp, double dt, double q, double m, bool write_part) {
    const Geom &g = c->g;
    constexpr int TC = TX * TY;
    const int ntx = (g.nx + TX - 1) / TX, nty = (g.ny + TY - 1) / TY;
    const int ntile = ntx * nty;
    const size_t perm_smem = sizeof(int) * (size_t)ntile * TC;
    if (perm_smem > TILE_PERM_SMEM_LIMIT) return 1;
    if (g.ng < 2) return 1;
    if (int r = lpic_ensure_scratch(c, sp.total)) return r;
    const size_t need = (size_t)g.npatch * (ntile + 1);
    if (need > c->tile_start_cap) {
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        cudaFree(c->d_tile_start);
        c->d_tile_start = nullptr; c->tile_start_cap = 0;
        CUDA_TRY(cudaMalloc(&c->d_tile_start, sizeof(int) * need));
        c->tile_start_cap = need;
    }
    int *d_nlist = (int *)(c->d_tmp64 + 64 + g.npatch);
    TilePermArgs pa;
    pa.x = sp.attr[LPIC_P_X]; pa.y = sp.attr[LPIC_P_Y]; pa.z = sp.attr[LPIC_P_Z];
    pa.ux = sp.attr[LPIC_P_UX]; pa.uy = sp.attr[LPIC_P_UY]; pa.uz = sp.attr[LPIC_P_UZ]; pa.ig = sp.attr[LPIC_P_INV_GAMMA];
    pa.ps = sp.pstride; pa.dead = sp.dead; pa.off = sp.d_off; pa.npart = sp.d_npart; pa.x0 = c->d_x0; pa.y0 = c->d_y0; pa.z0 = c->d_z0;
    pa.cdt = LPIC_C_LIGHT * 0.5 * dt; pa.inv_dx = 1.0 / g.dx; pa.inv_dy = 1.0 / g.dy; pa.inv_dz = 1.0;
    pa.nx = g.nx; pa.ny = g.ny; pa.nz = 1; pa.nty = nty; pa.ntz = 1; pa.ntile = ntile;  // key = (tx * nty + ty) * TC + lx * TY + ly
    pa.keys = (int *)c->scr_buf;
    pa.perm = c->scr_b; pa.tile_start = c->d_tile_start; pa.list = c->scr_a; pa.nlist = d_nlist;
    constexpr size_t push_smem = sizeof(double) * (6 * (TX + 3) * (TY + 4) + NW * 30 * 33) + sizeof(int) * NW * 32;
    if (!c->tile2d_attr_set) {
        CUDA_TRY(cudaFuncSetAttribute(k_tile_perm<TX, TY, 1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE_PERM_SMEM_LIMIT));
        CUDA_TRY(cudaFuncSetAttribute(k_push_tile2d<TX, TY, NW, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)push_smem));
        CUDA_TRY(cudaFuncSetAttribute(k_push_tile2d<TX, TY, NW, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)push_smem));
        CUDA_TRY(cudaFuncSetAttribute(k_push_tile2d<TX, TY, NW, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)push_smem));
        CUDA_TRY(cudaFuncSetAttribute(k_push_tile2d<TX, TY, NW, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)push_smem));
        c->tile2d_attr_set = true;
    }
    CUDA_TRY(cudaMemsetAsync(d_nlist, 0, sizeof(int) * g.npatch, c->stream));
    k_tile_perm<TX, TY, 1, 2><<<g.npatch, PT, perm_smem, c->stream>>>(pa);
    LAUNCHED(1);
    TileArgs ta;
    ta.g = g; ta.F = c->fields; ta.rec = nullptr; ta.prefetch = getenv("LPIC_REC_NO_PREFETCH") == nullptr; ta.px0 = c->d_x0; ta.py0 = c->d_y0; ta.pz0 = c->d_z0; ta.s = make_slots(sp);
    ta.perm = c->scr_b; ta.tile_start = c->d_tile_start; ta.list = c->scr_a; ta.nlist = d_nlist;
    ta.nty = nty; ta.ntz = 1; ta.ntile = ntile;
    ta.dt = dt; ta.cdt = pa.cdt; ta.efactor = q * dt / (2 * m * LPIC_C_LIGHT); ta.bfactor = q * dt / (2 * m);
    ta.inv_dx = pa.inv_dx; ta.inv_dy = pa.inv_dy; ta.inv_dz = 1.0;
    ta.q_dV = q / (g.dx * g.dy); ta.q_dydzdt = q / (g.dy * dt); ta.q_dxdzdt = q / (g.dx * dt); ta.q_dxdydt = 0.0;
    const unsigned grid = (unsigned)((i64)g.npatch * ntile);
    ta.rec = sp.rec;
    if (sp.rec) {
        if (write_part) k_push_tile2d<TX, TY, NW, true, true><<<grid, NW * 32, push_smem, c->stream>>>(ta);
        else k_push_tile2d<TX, TY, NW, false, true><<<grid, NW * 32, push_smem, c->stream>>>(ta);
    } else {
        if (write_part) k_push_tile2d<TX, TY, NW, true, false><<<grid, NW * 32, push_smem, c->stream>>>(ta);
        else k_push_tile2d<TX, TY, NW, false, false><<<grid, NW * 32, push_smem, c->stream>>>(ta);
    }
    LAUNCHED(1);
    const unsigned lgrid = (unsigned)std::min<i64>(g.npatch, 148 * 8);
    if (write_part) k_list_particles<true, 2><<<lgrid, 128, 0, c->stream>>>(g, c->fields, c->d_x0, c->d_y0, c->d_z0, ta.s, c->scr_a, d_nlist, dt, q, m);
    else k_list_particles<false, 2><<<lgrid, 128, 0, c->stream>>>(g, c->fields, c->d_x0, c->d_y0, c->d_z0, ta.s, c->scr_a, d_nlist, dt, q, m);
    LAUNCHED(1);
    KERNEL_CHECK();
    return 0;
}

template <int TX, int TY, int TZ, int NW>
int launch_tiles(lpic_ctx *c, Species &sp, double dt, double q, double m, bool write_part) {
    const Geom &g = c->g;
    constexpr int TC = TX * TY * TZ;
    const int ntx = (g.nx + TX - 1) / TX, nty = (g.ny + TY - 1) / TY, ntz = (g.nz + TZ - 1) / TZ;
    const int ntile = ntx * nty * ntz;
    const size_t perm_smem = sizeof(int) * (size_t)ntile * TC;
    if (perm_smem > TILE_PERM_SMEM_LIMIT) return 1;
    if (g.ng < 2) return 1;  // the staged halo reaches node -2
    if (int r = lpic_ensure_scratch(c, sp.total)) return r;
    const size_t need = (size_t)g.npatch * (ntile + 1);
    if (need > c->tile_start_cap) {
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        cudaFree(c->d_tile_start);
        c->d_tile_start = nullptr; c->tile_start_cap = 0;
        CUDA_TRY(cudaMalloc(&c->d_tile_start, sizeof(int) * need));
        c->tile_start_cap = need;
    }
    int *d_nlist = (int *)(c->d_tmp64 + 64 + g.npatch);
    TilePermArgs pa;
    pa.x = sp.attr[LPIC_P_X]; pa.y = sp.attr[LPIC_P_Y]; pa.z = sp.attr[LPIC_P_Z];
    pa.ux = sp.attr[LPIC_P_UX]; pa.uy = sp.attr[LPIC_P_UY]; pa.uz = sp.attr[LPIC_P_UZ]; pa.ig = sp.attr[LPIC_P_INV_GAMMA];
    pa.ps = sp.pstride; pa.dead = sp.dead; pa.off = sp.d_off; pa.npart = sp.d_npart; pa.x0 = c->d_x0; pa.y0 = c->d_y0; pa.z0 = c->d_z0;
    pa.cdt = LPIC_C_LIGHT * 0.5 * dt; pa.inv_dx = 1.0 / g.dx; pa.inv_dy = 1.0 / g.dy; pa.inv_dz = 1.0 / g.dz;
    pa.nx = g.nx; pa.ny = g.ny; pa.nz = g.nz; pa.nty = nty; pa.ntz = ntz; pa.ntile = ntile;
    pa.keys = (int *)c->scr_buf;  // the sort's staging buffer is idle during the push
    pa.perm = c->scr_b; pa.tile_start = c->d_tile_start; pa.list = c->scr_a; pa.nlist = d_nlist;
    constexpr size_t push_smem = sizeof(double) * (6 * (TX + 3) * (TY + 3) * (TZ + 4) + NW * 30 * 33) + sizeof(int) * NW * 32;
    if (!c->tile_attr_set) {  // per context: function attributes are per device, and a process may drive several
        CUDA_TRY(cudaFuncSetAttribute(k_tile_perm<TX, TY, TZ, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE_PERM_SMEM_LIMIT));
        CUDA_TRY(cudaFuncSetAttribute(k_push_tile<TX, TY, TZ, NW, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)push_smem));
        CUDA_TRY(cudaFuncSetAttribute(k_push_tile<TX, TY, TZ, NW, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)push_smem));
        CUDA_TRY(cudaFuncSetAttribute(k_push_tile<TX, TY, TZ, NW, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)push_smem));
        CUDA_TRY(cudaFuncSetAttribute(k_push_tile<TX, TY, TZ, NW, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)push_smem));
        c->tile_attr_set = true;
    }
    CUDA_TRY(cudaMemsetAsync(d_nlist, 0, sizeof(int) * g.npatch, c->stream));
    k_tile_perm<TX, TY, TZ, 3><<<g.npatch, PT, perm_smem, c->stream>>>(pa);
    LAUNCHED(1);
    TileArgs ta;
    ta.g = g; ta.F = c->fields; ta.rec = nullptr; ta.prefetch = getenv("LPIC_REC_NO_PREFETCH") == nullptr; ta.px0 = c->d_x0; ta.py0 = c->d_y0; ta.pz0 = c->d_z0; ta.s = make_slots(sp);
    ta.perm = c->scr_b; ta.tile_start = c->d_tile_start; ta.list = c->scr_a; ta.nlist = d_nlist;
    ta.nty = nty; ta.ntz = ntz; ta.ntile = ntile;
    ta.dt = dt; ta.cdt = pa.cdt; ta.efactor = q * dt / (2 * m * LPIC_C_LIGHT); ta.bfactor = q * dt / (2 * m);
    ta.inv_dx = pa.inv_dx; ta.inv_dy = pa.inv_dy; ta.inv_dz = pa.inv_dz;
    ta.q_dV = q / (g.dx * g.dy * g.dz); ta.q_dydzdt = q / (g.dy * g.dz * dt);
    ta.q_dxdzdt = q / (g.dx * g.dz * dt); ta.q_dxdydt = q / (g.dx * g.dy * dt);
    const unsigned grid = (unsigned)((i64)g.npatch * ntile);
    ta.rec = sp.rec;
    if (sp.rec) {
        if (write_part) k_push_tile<TX, TY, TZ, NW, true, true><<<grid, NW * 32, push_smem, c->stream>>>(ta);
        else k_push_tile<TX, TY, TZ, NW, false, true><<<grid, NW * 32, push_smem, c->stream>>>(ta);
    } else {
        if (write_part) k_push_tile<TX, TY, TZ, NW, true, false><<<grid, NW * 32, push_smem, c->stream>>>(ta);
        else k_push_tile<TX, TY, TZ, NW, false, false><<<grid, NW * 32, push_smem, c->stream>>>(ta);
    }
    LAUNCHED(1);
    const unsigned lgrid = (unsigned)std::min<i64>(g.npatch, 148 * 8);
    if (write_part) k_list_particles<true, 3><<<lgrid, 128, 0, c->stream>>>(g, c->fields, c->d_x0, c->d_y0, c->d_z0, ta.s, c->scr_a, d_nlist, dt, q, m);
    else k_list_particles<false, 3><<<lgrid, 128, 0, c->stream>>>(g, c->fields, c->d_x0, c->d_y0, c->d_z0, ta.s, c->scr_a, d_nlist, dt, q, m);
    LAUNCHED(1);
    KERNEL_CHECK();
    return 0;
}

}  // namespace

// fused push + deposit, tile kernels (3D and 2D); returns 1 if this path does not apply (caller falls back)
int lpic_push_deposit_tiles(lpic_ctx *c, int ispec, double dt, double q, double m, bool write_part) {
    const Geom &g = c->g;
    Species &sp = c->spec[ispec];
    if (sp.max_npart == 0) return 0;
    if (g.dim == 2) return launch_tiles2d<8, 16, 4>(c, sp, dt, q, m, write_part);
    return launch_tiles<4, 4, 16, 8>(c, sp, dt, q, m, write_part);
}
