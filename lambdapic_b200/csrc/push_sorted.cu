// Cell-ordered, warp-cooperative version of the fused particle kernel (3D; the 2D twin k_push_sorted2d is further down).
//
// Same arithmetic as k_particles<3,FUSED> (particles.cu; reference: core/pusher/unified/unified_pusher_3d.c:281-431,
// core/current/current_deposit.h:275-440) but organised for the memory system of the B200:
//   1. k_cell_perm   one CTA per patch builds, once per species and step, a permutation of the ALIVE slots ordered
//                    by cell (z fastest): shared-memory histogram -> block scan -> scatter.  The slot order in memory
//                    stays the reference's (x-column buckets, bit-exact sort/migration indices); only the ORDER OF
//                    PROCESSING changes.  Dead slots never reach the particle kernel.
//   2. k_push_sorted one thread per alive particle in cell order.  A warp now works on 1-3 neighbouring cells, so its
//                    6x27 gather loads coalesce into a handful of L1-resident sectors, and the 27x4 stencil values
//                    of the lanes that share a start cell are summed across the warp (transposed through shared
//                    memory) before a single lane issues the fp64 RED: ~10x fewer L2 atomics per particle.
//                    Particles that cross a cell boundary during the step (a few %) are appended to a list ...
//   3. k_deposit_list ... and deposited by the general 125-point routine, one thread each.
// The instruction stream of the fully unrolled kernel (5.4 k SASS instructions, 86 KB) does not fit the 32 KB L1.5
// instruction cache; template parameter COMPACT keeps the component / plane / source-lane loops rolled (2.1 k
// instructions) with bit-identical arithmetic.  Round-1 measurements (profiles/r1_push_variants.txt): no_instruction
// stalls vanish, the extra loop overhead eats the gain -- unrolled stays the default, LPIC_PUSH_COMPACT=1 selects the other.
// Shared-memory fp64 atomics are a CAS loop on sm_100a (ATOMS.CAST.SPIN.64), which is why the reduction happens in
// registers and the accumulation uses native REDG.E.ADD.F64 in L2.
#include <stdlib.h>
#include <algorithm>
#include "lpic_common.cuh"
#include "particle_math.cuh"

namespace {

constexpr int PUSH_WARPS = 4;      // warps per CTA of the particle kernel (128 threads)
constexpr int PT = 256;          // threads of the permutation CTA
constexpr int KEY_LIMIT = 57344;  // cells per patch that fit the shared-memory histogram (224 KB of the 227 KB a CTA may use)

__device__ __forceinline__ int warp_incl_sum(int v) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int n = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += n;
    }
    return v;
}

// per-launch constants evaluated once on the host (same IEEE operations as the per-thread expressions they replace)
struct PushConst {
    double cdt, efactor, bfactor, inv_dx, inv_dy, inv_dz, q_dV, q_dydzdt, q_dxdzdt, q_dxdydt;
};

struct PermArgs {
    int ps;  // stride of the attribute pointers below (8: records, 1: separate arrays)
    const double *x, *y, *z;
    const double *ux, *uy, *uz, *ig;  // predict != 0: order by the cell of x + v dt/2 (where the particle gathers AND where
    double cdt;                       // its deposit starts: x_end - v_new dt/2 == x + v_old dt/2), padded by one cell per side
    int predict;
    const u8 *dead;
    const i64 *off, *npart;
    const double *x0, *y0, *z0;
    int dim, nx, ny, nz, kx, ky, kz;  // kx,ky,kz: key grid (nz -> 1 etc. when a patch has more cells than KEY_LIMIT; 2D: kz = 1)
    double dx, dy, dz;
    int *keys;     // arena scratch: cell key of every slot (-1: dead), written by the counting pass, read by the scatter pass
    int *perm;     // arena: local slot numbers of the alive particles in cell order
    i64 *nalive;   // per patch
};

__device__ __forceinline__ int node_of(double x, double x0, double d, int n) {
    const int i = (int)floor((x - x0) / d + 0.5);
    return i < 0 ? 0 : (i >= n ? n - 1 : i);
}

// nearest node clamped to [-1, n], shifted by one: a particle may sit up to half a cell outside its patch box
__device__ __forceinline__ int node_pad(double x, double x0, double d, int n) {
    const int i = (int)floor((x - x0) / d + 0.5);
    return (i < -1 ? -1 : (i > n ? n : i)) + 1;
}

__global__ void __launch_bounds__(PT) k_cell_perm(PermArgs a) {
    extern __shared__ int hist[];
    __shared__ int sw[PT / 32];
    const int p = blockIdx.x, tid = threadIdx.x;
    const i64 off = a.off[p];
    const int np = (int)a.npart[p];
    const int nkey = a.kx * a.ky * a.kz;
    for (int b = tid; b < nkey; b += PT) hist[b] = 0;
    __syncthreads();
    const double x0 = a.x0[p], y0 = a.y0[p], z0 = a.z0[p];
    auto key_of = [&](int ip) -> int {
        if (a.dead[off + ip]) return -1;
        const i64 ir = (off + ip) * a.ps;
        const double x = a.x[ir], y = a.y[ir], z = a.dim == 3 ? a.z[ir] : 0.0;
        if (isnan(x) || isnan(y) || isnan(z)) return -1;
        if (a.predict) {
            const double h = a.cdt * a.ig[ir];
            const int ix = node_pad(x + h * a.ux[ir], x0, a.dx, a.nx), iy = node_pad(y + h * a.uy[ir], y0, a.dy, a.ny),
                      iz = a.dim == 3 ? node_pad(z + h * a.uz[ir], z0, a.dz, a.nz) : 0;
            return iz + a.kz * (iy + a.ky * ix);
        }
        const int ix = node_of(x, x0, a.dx, a.nx), iy = a.ky > 1 ? node_of(y, y0, a.dy, a.ny) : 0,
                  iz = a.kz > 1 ? node_of(z, z0, a.dz, a.nz) : 0;
        return iz + a.kz * (iy + a.ky * ix);
    };
    // four slots per thread and iteration: the 4 x 8 loads are independent and in flight together
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
    if (tid == 0) a.nalive[p] = run;
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

__device__ __forceinline__ void shape3(double delta, double *S) {  // calculate_S0 restricted to its 3 non-zeros
    const double d2 = delta * delta;
    S[0] = 0.5 * (d2 + delta + 0.25);
    S[1] = 0.75 - d2;
    S[2] = 0.5 * (d2 - delta + 0.25);
}

// one RED for a finished segment: the segment head's start cell (sb = bx0, by0, bz0; bz0 < 0: nothing to deposit)
__device__ __noinline__ void flush_red(double *dst, const int *sb, int i, int sj, int sk, int NX, int NY, int NZ, double sum) {
    if (sb[2] < 0) return;
    const int id = wrap_once(sb[0] + i - 1, NX) * NY * NZ + wrap_once(sb[1] + sj - 1, NY) * NZ + wrap_once(sb[2] + sk - 1, NZ);
    atomicAdd(dst + id, sum);
}

template <typename T>
__device__ __forceinline__ T sel3(const T *a, int i) { return i == 0 ? a[0] : (i == 1 ? a[1] : a[2]); }

// gather_eb<3> with the six components as a loop that the COMPACT kernel keeps rolled (one 27-point body in the
// instruction stream instead of six): which of the node-centred (g) / cell-centred (h) weight sets a component uses
// along x, y, z is a bit of the masks below (ex(h,g,g) ey(g,h,g) ez(g,g,h) bx(g,h,h) by(h,g,h) bz(h,h,g)).
template <bool COMPACT>
__device__ __forceinline__ void gather_eb_loop(const Geom &g, const PatchView &v, double x, double y, double z, double *eb,
                                               const PushConst &k) {
    const double X = (x - v.x0) * k.inv_dx, Y = (y - v.y0) * k.inv_dy, Z = (z - v.z0) * k.inv_dz;
    const double fX = floor(X), fY = floor(Y), fZ = floor(Z), rX = floor(X + 0.5), rY = floor(Y + 0.5), rZ = floor(Z + 0.5);
    double gx[3], gy[3], gz[3], hx[3], hy[3], hz[3];
    tsc3(rX - X, gx); tsc3(fX - X + 0.5, hx);
    tsc3(rY - Y, gy); tsc3(fY - Y + 0.5, hy);
    tsc3(rZ - Z, gz); tsc3(fZ - Z + 0.5, hz);
    int ogx[3], ohx[3], ogy[3], ohy[3], ogz[3], ohz[3];
    offsets3((int)rX, g.NX, g.NY * g.NZ, ogx); offsets3((int)fX, g.NX, g.NY * g.NZ, ohx);
    offsets3((int)rY, g.NY, g.NZ, ogy); offsets3((int)fY, g.NY, g.NZ, ohy);
    offsets3((int)rZ, g.NZ, 1, ogz); offsets3((int)fZ, g.NZ, 1, ohz);
    const size_t stride = (size_t)g.npatch * g.ncell;
#pragma unroll(COMPACT ? 1 : 6)
    for (int c = 0; c < 6; c++) {
        const bool sx = (0x31 >> c) & 1, sy = (0x2A >> c) & 1, sz = (0x1C >> c) & 1;
        double fx[3], fy[3], fz[3];
        int ox[3], oy[3], oz[3];
#pragma unroll
        for (int a = 0; a < 3; a++) {
            fx[a] = sx ? hx[a] : gx[a]; ox[a] = sx ? ohx[a] : ogx[a];
            fy[a] = sy ? hy[a] : gy[a]; oy[a] = sy ? ohy[a] : ogy[a];
            fz[a] = sz ? hz[a] : gz[a]; oz[a] = sz ? ohz[a] : ogz[a];
        }
        const double val = gather27(v.ex + c * stride, fx, fy, fz, ox, oy, oz);
#pragma unroll
        for (int a = 0; a < 6; a++) eb[a] = c == a ? val : eb[a];
    }
}

// One warp-iteration of the fused step for the particles perm[off + t], t < n.
template <bool WRITE_PART, bool COMPACT>
__device__ __forceinline__ void push_body(const Geom &g, double *__restrict__ F, const double *__restrict__ px0,
                                          const double *__restrict__ py0, const double *__restrict__ pz0, const Slots &s,
                                          const int *__restrict__ perm, int *__restrict__ cross, int *__restrict__ ncross,
                                          double dt, double q, double m, int p, i64 t, i64 n, const PushConst &k) {
    const int lane = threadIdx.x & 31;
    const bool active = t < n;
    const i64 off = s.off[p];
    const PatchView v = patch_view(g, F, px0, py0, pz0, p);
    const double cdt = k.cdt;
    double x = 0, y = 0, z = 0, ux = 0, uy = 0, uz = 0, ig = 1, w = 0;
    i64 ip = 0;
    int local = 0;
    if (active) {
        local = perm[off + t];
        ip = off + local;
        x = s.x[ip * s.ps]; y = s.y[ip * s.ps]; z = s.z[ip * s.ps];
        ux = s.ux[ip * s.ps]; uy = s.uy[ip * s.ps]; uz = s.uz[ip * s.ps]; ig = s.ig[ip * s.ps];
        w = s.w[ip * s.ps];
        x += cdt * ig * ux; y += cdt * ig * uy; z += cdt * ig * uz;
        double eb[6] = {0, 0, 0, 0, 0, 0};
        gather_eb_loop<COMPACT>(g, v, x, y, z, eb, k);
        if (WRITE_PART) {
#pragma unroll
            for (int a = 0; a < 6; a++) s.part[a][ip] = eb[a];
        }
        boris_kick(ux, uy, uz, ig, eb, k.efactor, k.bfactor);
        s.ux[ip * s.ps] = ux; s.uy[ip * s.ps] = uy; s.uz[ip * s.ps] = uz; s.ig[ip * s.ps] = ig;
        x += cdt * ig * ux; y += cdt * ig * uy; z += cdt * ig * uz;
        s.x[ip * s.ps] = x; s.y[ip * s.ps] = y; s.z[ip * s.ps] = z;
    }
    // ---- deposit set-up (current_deposit.h:341-373) ------------------------------------------------------------
    const double vx = ux * LPIC_C_LIGHT * ig, vy = uy * LPIC_C_LIGHT * ig, vz = uz * LPIC_C_LIGHT * ig;
    const double X0 = (x - vx * 0.5 * dt - v.x0) / g.dx, X1 = (x + vx * 0.5 * dt - v.x0) / g.dx;
    const double Y0 = (y - vy * 0.5 * dt - v.y0) / g.dy, Y1 = (y + vy * 0.5 * dt - v.y0) / g.dy;
    const double Z0 = (z - vz * 0.5 * dt - v.z0) / g.dz, Z1 = (z + vz * 0.5 * dt - v.z0) / g.dz;
    const int ix0 = (int)floor(X0 + 0.5), iy0 = (int)floor(Y0 + 0.5), iz0 = (int)floor(Z0 + 0.5);
    const int ix1 = (int)floor(X1 + 0.5), iy1 = (int)floor(Y1 + 0.5), iz1 = (int)floor(Z1 + 0.5);
    const bool fast = active && ix1 == ix0 && iy1 == iy0 && iz1 == iz0;
    {   // particles that change cell go to the general routine (second kernel); one warp-aggregated append keeps the
        // list in (roughly) cell order, so that kernel's neighbouring lanes hit neighbouring cells
        const unsigned cm = __ballot_sync(0xffffffffu, active && !fast);
        if (cm) {
            int basepos = 0;
            if (lane == __ffs(cm) - 1) basepos = atomicAdd(&ncross[p], __popc(cm));
            basepos = __shfl_sync(0xffffffffu, basepos, __ffs(cm) - 1);
            if (active && !fast) cross[off + basepos + __popc(cm & ((1u << lane) - 1u))] = local;
        }
    }
    // segments = runs of consecutive lanes that start in the same cell
    const int bx0 = wrap_base(ix0, g.NX), by0 = wrap_base(iy0, g.NY), bz0 = wrap_base(iz0, g.NZ);
    // Lanes outside the fast path add zeros, so they must not split a run: a segment starts at lane 0 and at every
    // fast lane whose start cell differs from the previous FAST lane's.
    const int key = bz0 + g.NZ * (by0 + g.NY * bx0);
    const unsigned fm = __ballot_sync(0xffffffffu, fast);
    const unsigned before = fm & ((1u << lane) - 1u);
    const int pf = before ? 31 - __clz(before) : -1;
    const int pkey = __shfl_sync(0xffffffffu, key, pf < 0 ? 0 : pf);
    const unsigned heads = __ballot_sync(0xffffffffu, lane == 0 || (fast && (pf < 0 || key != pkey)));
    if (!__any_sync(0xffffffffu, fast)) return;
    double S0x[3], S0y[3], S0z[3], DSx[3], DSy[3], DSz[3];
    shape3(ix0 - X0, S0x); shape3(iy0 - Y0, S0y); shape3(iz0 - Z0, S0z);
    shape3(ix1 - X1, DSx); shape3(iy1 - Y1, DSy); shape3(iz1 - Z1, DSz);  // S1 (no cell crossing: same support)
    double S1x[3], S1y[3], S1z[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        S1x[i] = DSx[i]; S1y[i] = DSy[i]; S1z[i] = DSz[i];
        DSx[i] -= S0x[i]; DSy[i] -= S0y[i]; DSz[i] -= S0z[i];
    }
    const double wq = fast ? w : 0.0;  // lanes outside the fast path add zeros
    const double cd = k.q_dV * wq, fdx = k.q_dydzdt * wq, fdy = k.q_dxdzdt * wq, fdz = k.q_dxdydt * wq;
    // Warp-level reduction through shared memory, one x-plane of the 3x3x3 stencil at a time.  Every lane stores
    // its values of the plane as rows of a [30][33] tile (row stride 33 doubles: conflict-free both ways); then lane l
    // owns row l and adds up the 32 source lanes, issuing ONE fp64 RED per segment (= run of lanes that start in the
    // same cell).  Rows per plane: rho 9, jy 6, jz 6, jx 9 -- the last jx plane, jy row and jz column of a particle
    // that stays in its cell are sum(DS) = 0 up to rounding (|.| <= 4 eps of the particle's largest term) and are not
    // deposited, which makes every plane fit one round of 32 lanes.
    __shared__ double red[PUSH_WARPS][30 * 33];
    __shared__ int segbase[PUSH_WARPS][32][3];
    double *rtile = red[threadIdx.x >> 5];
    int(*sb)[3] = segbase[threadIdx.x >> 5];
    sb[lane][0] = bx0; sb[lane][1] = by0; sb[lane][2] = fast ? bz0 : -1;
    // which row does this lane own?  [0,9) rho(j,k)  [9,15) jy(j<2,k)  [15,21) jz(j,k<2)  [21,30) jx(j,k)
    int comp, sj, sk;
    if (lane < 9) { comp = 3; sj = lane / 3; sk = lane - 3 * sj; }
    else if (lane < 15) { comp = 1; sj = (lane - 9) / 3; sk = (lane - 9) - 3 * sj; }
    else if (lane < 21) { comp = 2; sj = (lane - 15) >> 1; sk = (lane - 15) & 1; }
    else { comp = 0; sj = (lane - 21) / 3; sk = (lane - 21) - 3 * sj; }
    double *dst = comp == 0 ? v.jx : (comp == 1 ? v.jy : (comp == 2 ? v.jz : v.rho));
    const double *row = rtile + lane * 33;
    double jxb[3][3];
#pragma unroll
    for (int a = 0; a < 3; a++)
#pragma unroll
        for (int b = 0; b < 3; b++) jxb[a][b] = 0.0;
#pragma unroll(COMPACT ? 1 : 3)
    for (int i = 0; i < 3; i++) {
        const double s0x = sel3(S0x, i), dsx = sel3(DSx, i), s1x = sel3(S1x, i);
        const double ax = s0x + 0.5 * dsx, cx = 0.5 * s0x + LPIC_ONE_THIRD * dsx, fx = fdx * dsx;
        double jyb[3] = {0.0, 0.0, 0.0};
#pragma unroll
        for (int j = 0; j < 3; j++) {
            const double ay = S0y[j] + 0.5 * DSy[j], cy = 0.5 * S0y[j] + LPIC_ONE_THIRD * DSy[j], fy = fdy * DSy[j];
            const double tz = ax * S0y[j] + cx * DSy[j];
            const double rxy = cd * s1x * S1y[j];
            double jzb = 0.0;
#pragma unroll
            for (int k = 0; k < 3; k++) {
                rtile[(j * 3 + k) * 33 + lane] = rxy * S1z[k];
                if (COMPACT || i < 2) {  // rolled: the last plane's rows 21..29 are written but never read
                    jxb[k][j] -= fx * (ay * S0z[k] + cy * DSz[k]);
                    rtile[(21 + j * 3 + k) * 33 + lane] = jxb[k][j];
                }
                if (j < 2) {
                    jyb[k] -= fy * (ax * S0z[k] + cx * DSz[k]);
                    rtile[(9 + j * 3 + k) * 33 + lane] = jyb[k];
                }
                if (k < 2) {
                    jzb -= fdz * DSz[k] * tz;
                    rtile[(15 + j * 2 + k) * 33 + lane] = jzb;
                }
            }
        }
        __syncwarp();
        if (lane < (i < 2 ? 30 : 21)) {
            double acc = 0.0;
            int seg = 0;  // lane index of the current segment's head
#pragma unroll(COMPACT ? 1 : 8)
            for (int g4 = 0; g4 < 32; g4 += 4) {
                const unsigned mm = (heads >> g4) & 0xFu;
                if ((mm & 0xEu) == 0u) {  // no segment starts strictly inside this group of 4 source lanes
                    if (g4 > 0 && (mm & 1u)) {
                        flush_red(dst, sb[seg], i, sj, sk, g.NX, g.NY, g.NZ, acc);
                        acc = 0.0;
                        seg = g4;
                    }
                    acc += row[g4];
                    acc += row[g4 + 1];
                    acc += row[g4 + 2];
                    acc += row[g4 + 3];
                } else {
                    for (int t = 0; t < 4; t++) {
                        const int src = g4 + t;
                        if (src > 0 && ((mm >> t) & 1u)) {
                            flush_red(dst, sb[seg], i, sj, sk, g.NX, g.NY, g.NZ, acc);
                            acc = 0.0;
                            seg = src;
                        }
                        acc += row[src];
                    }
                }
            }
            flush_red(dst, sb[seg], i, sj, sk, g.NX, g.NY, g.NZ, acc);
        }
        __syncwarp();
    }
}

template <bool WRITE_PART, bool COMPACT>
__global__ void __launch_bounds__(128, 4) k_push_sorted(Geom g, double *__restrict__ F, const double *__restrict__ px0,
                                                     const double *__restrict__ py0, const double *__restrict__ pz0, Slots s,
                                                     const int *__restrict__ perm, const i64 *__restrict__ nalive,
                                                     int *__restrict__ cross, int *__restrict__ ncross, int blocks_per_patch,
                                                     double dt, double q, double m, PushConst k) {
    const int p = blockIdx.x / blocks_per_patch;
    const i64 t = (i64)(blockIdx.x - p * blocks_per_patch) * blockDim.x + threadIdx.x;
    const i64 n = nalive[p];
    if (t - (threadIdx.x & 31) >= n) return;  // whole warp beyond the alive particles of this patch
    push_body<WRITE_PART, COMPACT>(g, F, px0, py0, pz0, s, perm, cross, ncross, dt, q, m, p, t, n, k);
}

// general deposit for the particles that changed cell during the step
__global__ void __launch_bounds__(128) k_deposit_list(Geom g, double *__restrict__ F, const double *__restrict__ px0,
                                                      const double *__restrict__ py0, const double *__restrict__ pz0, Slots s,
                                                      const int *__restrict__ cross, const int *__restrict__ ncross,
                                                      int blocks_per_patch, double dt, double q) {
    // persistent grid: CTAs stride over the patches, threads over each patch's (short) list
    DepositCoef3 k;
    k.q_dV = q / (g.dx * g.dy * g.dz); k.q_dydzdt = q / (g.dy * g.dz * dt);
    k.q_dxdzdt = q / (g.dx * g.dz * dt); k.q_dxdydt = q / (g.dx * g.dy * dt); k.dt = dt;
    for (int p = blockIdx.x; p < g.npatch; p += gridDim.x) {
        const int n = ncross[p];
        if (n == 0) continue;
        const PatchView v = patch_view(g, F, px0, py0, pz0, p);
        for (int t = threadIdx.x; t < n; t += blockDim.x) {
            const i64 ip = s.off[p] + cross[s.off[p] + t];
            deposit3(g, v, k, s.x[ip * s.ps], s.y[ip * s.ps], s.z[ip * s.ps], s.ux[ip * s.ps], s.uy[ip * s.ps], s.uz[ip * s.ps], s.ig[ip * s.ps], s.w[ip * s.ps]);
        }
    }
}


// ---- 2D twin --------------------------------------------------------------------------------------------------------
// Same organisation in two dimensions (reference: core/pusher/unified/unified_pusher_2d.c:157-330,
// core/current/current_deposit.h:150-268): 6 x 9 gather points per particle, and the whole 3x3 stencil of a particle that
// stays in its cell is ONE round of the [30][33] reduction tile -- rho 9 rows, jx 6 (the last x row is sum(DSx) = 0 up to
// rounding), jy 6 (last y column likewise), jz 9.  RED address of row (i, j): wrap(bx0 + i - 1) * NY + wrap(by0 + j - 1).
template <bool WRITE_PART>
__global__ void __launch_bounds__(128, 6) k_push_sorted2d(Geom g, double *__restrict__ F, const double *__restrict__ px0,
                                                       const double *__restrict__ py0, const double *__restrict__ pz0, Slots s,
                                                       const int *__restrict__ perm, const i64 *__restrict__ nalive,
                                                       int *__restrict__ cross, int *__restrict__ ncross, int blocks_per_patch,
                                                       double dt, PushConst k) {
    const int p = blockIdx.x / blocks_per_patch;
    const i64 t = (i64)(blockIdx.x - p * blocks_per_patch) * blockDim.x + threadIdx.x;
    const i64 n = nalive[p];
    const int lane = threadIdx.x & 31;
    if (t - lane >= n) return;  // whole warp beyond the alive particles of this patch
    const bool active = t < n;
    const i64 off = s.off[p];
    const PatchView v = patch_view(g, F, px0, py0, pz0, p);
    const double cdt = k.cdt;
    double x = 0, y = 0, ux = 0, uy = 0, uz = 0, ig = 1, w = 0;
    i64 ip = 0;
    int local = 0;
    if (active) {
        local = perm[off + t];
        ip = off + local;
        x = s.x[ip * s.ps]; y = s.y[ip * s.ps];
        ux = s.ux[ip * s.ps]; uy = s.uy[ip * s.ps]; uz = s.uz[ip * s.ps]; ig = s.ig[ip * s.ps];
        w = s.w[ip * s.ps];
        x += cdt * ig * ux; y += cdt * ig * uy;
        double eb[6];
        gather_eb<2>(g, v, x, y, 0.0, eb);
        if (WRITE_PART) {
#pragma unroll
            for (int a = 0; a < 6; a++) s.part[a][ip] = eb[a];
        }
        boris_kick(ux, uy, uz, ig, eb, k.efactor, k.bfactor);
        s.ux[ip * s.ps] = ux; s.uy[ip * s.ps] = uy; s.uz[ip * s.ps] = uz; s.ig[ip * s.ps] = ig;
        x += cdt * ig * ux; y += cdt * ig * uy;
        s.x[ip * s.ps] = x; s.y[ip * s.ps] = y;
    }
    // ---- deposit set-up (current_deposit.h:196-222) ----------------------------------------------------------------
    const double vx = ux * LPIC_C_LIGHT * ig, vy = uy * LPIC_C_LIGHT * ig, vz = uz * LPIC_C_LIGHT * ig;
    const double X0 = (x - vx * 0.5 * dt - v.x0) / g.dx, X1 = (x + vx * 0.5 * dt - v.x0) / g.dx;
    const double Y0 = (y - vy * 0.5 * dt - v.y0) / g.dy, Y1 = (y + vy * 0.5 * dt - v.y0) / g.dy;
    const int ix0 = (int)floor(X0 + 0.5), iy0 = (int)floor(Y0 + 0.5);
    const int ix1 = (int)floor(X1 + 0.5), iy1 = (int)floor(Y1 + 0.5);
    const bool fast = active && ix1 == ix0 && iy1 == iy0;
    {   // particles that change cell go to the general routine (k_deposit_list2d)
        const unsigned cm = __ballot_sync(0xffffffffu, active && !fast);
        if (cm) {
            int basepos = 0;
            if (lane == __ffs(cm) - 1) basepos = atomicAdd(&ncross[p], __popc(cm));
            basepos = __shfl_sync(0xffffffffu, basepos, __ffs(cm) - 1);
            if (active && !fast) cross[off + basepos + __popc(cm & ((1u << lane) - 1u))] = local;
        }
    }
    const int bx0 = wrap_base(ix0, g.NX), by0 = wrap_base(iy0, g.NY);
    const int key = by0 + g.NY * bx0;
    const unsigned fm = __ballot_sync(0xffffffffu, fast);
    const unsigned before = fm & ((1u << lane) - 1u);
    const int pf = before ? 31 - __clz(before) : -1;
    const int pkey = __shfl_sync(0xffffffffu, key, pf < 0 ? 0 : pf);
    const unsigned heads = __ballot_sync(0xffffffffu, lane == 0 || (fast && (pf < 0 || key != pkey)));
    if (!fm) return;
    double S0x[3], S0y[3], S1x[3], S1y[3], DSx[3], DSy[3];
    shape3(ix0 - X0, S0x); shape3(iy0 - Y0, S0y);
    shape3(ix1 - X1, S1x); shape3(iy1 - Y1, S1y);  // no cell crossing: same support as S0
#pragma unroll
    for (int i = 0; i < 3; i++) { DSx[i] = S1x[i] - S0x[i]; DSy[i] = S1y[i] - S0y[i]; }
    const double wq = fast ? w : 0.0;  // lanes outside the fast path add zeros
    const double cd = k.q_dV * wq, fdx = k.q_dydzdt * wq, fdy = k.q_dxdzdt * wq, fvz = cd * vz;  // 2D: q/(dx dy), q/(dy dt), q/(dx dt)
    const double one_twelfth = 1.0 / 12.0;
    __shared__ double red[PUSH_WARPS][30 * 33];
    __shared__ int segbase[PUSH_WARPS][32][2];
    double *rtile = red[threadIdx.x >> 5];
    int(*sb)[2] = segbase[threadIdx.x >> 5];
    sb[lane][0] = bx0; sb[lane][1] = fast ? by0 : -1;
    // rows: [0,9) rho(i,j)  [9,15) jx(i<2,j)  [15,21) jy(i,j<2)  [21,30) jz(i,j)
    double jxb[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const double a = S0x[i] + 0.5 * DSx[i], fxi = fdx * DSx[i], t12 = one_twelfth * DSx[i];
        double jyb = 0.0;
#pragma unroll
        for (int j = 0; j < 3; j++) {
            const double b = S0y[j] + 0.5 * DSy[j];
            rtile[(i * 3 + j) * 33 + lane] = cd * S1x[i] * S1y[j];
            if (i < 2) {
                jxb[j] -= fxi * b;
                rtile[(9 + i * 3 + j) * 33 + lane] = jxb[j];
            }
            if (j < 2) {
                jyb -= fdy * (DSy[j] * a);
                rtile[(15 + i * 2 + j) * 33 + lane] = jyb;
            }
            rtile[(21 + i * 3 + j) * 33 + lane] = fvz * (a * b + t12 * DSy[j]);
        }
    }
    __syncwarp();
    if (lane < 30) {
        int comp, si, sj;
        if (lane < 9) { comp = 3; si = lane / 3; sj = lane - 3 * si; }
        else if (lane < 15) { comp = 0; si = (lane - 9) / 3; sj = (lane - 9) - 3 * si; }
        else if (lane < 21) { comp = 1; si = (lane - 15) >> 1; sj = (lane - 15) & 1; }
        else { comp = 2; si = (lane - 21) / 3; sj = (lane - 21) - 3 * si; }
        double *dst = comp == 0 ? v.jx : (comp == 1 ? v.jy : (comp == 2 ? v.jz : v.rho));
        const double *row = rtile + lane * 33;
        double acc = 0.0;
        int seg = 0;  // lane index of the current segment's head
        auto flush = [&]() {
            if (sb[seg][1] >= 0) atomicAdd(dst + wrap_once(sb[seg][0] + si - 1, g.NX) * g.NY + wrap_once(sb[seg][1] + sj - 1, g.NY), acc);
        };
#pragma unroll 4
        for (int src = 0; src < 32; src++) {
            if (src > 0 && ((heads >> src) & 1u)) {
                flush();
                acc = 0.0;
                seg = src;
            }
            acc += row[src];
        }
        flush();
    }
}

__global__ void __launch_bounds__(128) k_deposit_list2d(Geom g, double *__restrict__ F, const double *__restrict__ px0,
                                                        const double *__restrict__ py0, const double *__restrict__ pz0, Slots s,
                                                        const int *__restrict__ cross, const int *__restrict__ ncross, double dt,
                                                        double q) {
    DepositCoef2 k;
    k.q_dxdy = q / (g.dx * g.dy); k.q_dydt = q / (g.dy * dt); k.q_dxdt = q / (g.dx * dt); k.dt = dt;
    for (int p = blockIdx.x; p < g.npatch; p += gridDim.x) {
        const int n = ncross[p];
        if (n == 0) continue;
        const PatchView v = patch_view(g, F, px0, py0, pz0, p);
        for (int t = threadIdx.x; t < n; t += blockDim.x) {
            const i64 ip = s.off[p] + cross[s.off[p] + t];
            deposit2(g, v, k, s.x[ip * s.ps], s.y[ip * s.ps], s.ux[ip * s.ps], s.uy[ip * s.ps], s.uz[ip * s.ps], s.ig[ip * s.ps], s.w[ip * s.ps]);
        }
    }
}

}  // namespace

// fused push + deposit in cell order; returns 1 if this path does not apply (caller falls back to k_particles)
int lpic_push_deposit_sorted(lpic_ctx *c, int ispec, double dt, double q, double m, bool write_part) {
    const Geom &g = c->g;
    Species &sp = c->spec[ispec];
    if (sp.max_npart == 0) return 0;
    const bool three = g.dim == 3;
    if (int r = lpic_ensure_scratch(c, sp.total)) return r;
    PermArgs a;
    a.x = sp.attr[LPIC_P_X]; a.y = sp.attr[LPIC_P_Y]; a.z = sp.attr[LPIC_P_Z]; a.dead = sp.dead; a.ps = sp.pstride;
    a.off = sp.d_off; a.npart = sp.d_npart; a.x0 = c->d_x0; a.y0 = c->d_y0; a.z0 = c->d_z0;
    a.ux = sp.attr[LPIC_P_UX]; a.uy = sp.attr[LPIC_P_UY]; a.uz = sp.attr[LPIC_P_UZ]; a.ig = sp.attr[LPIC_P_INV_GAMMA];
    a.cdt = LPIC_C_LIGHT * 0.5 * dt;
    a.dim = g.dim; a.nx = g.nx; a.ny = g.ny; a.nz = three ? g.nz : 1;
    static const bool no_predict = getenv("LPIC_PERM_CURRENT_CELL") != nullptr;  // tuning knob: order by the current cell
    a.predict = !no_predict && (i64)(g.nx + 2) * (g.ny + 2) * (three ? g.nz + 2 : 1) <= KEY_LIMIT;
    if (a.predict) {
        a.kx = g.nx + 2; a.ky = g.ny + 2; a.kz = three ? g.nz + 2 : 1;
    } else {
        a.kx = g.nx; a.ky = g.ny; a.kz = a.nz;
        if ((i64)a.kx * a.ky * a.kz > KEY_LIMIT) a.kz = 1;
        if ((i64)a.kx * a.ky * a.kz > KEY_LIMIT) a.ky = 1;
        if ((i64)a.kx * a.ky * a.kz > KEY_LIMIT) return 1;
    }
    a.dx = g.dx; a.dy = g.dy; a.dz = g.dz;
    a.perm = c->scr_b;
    a.keys = (int *)c->scr_buf;  // the sort's staging buffer is idle during the push
    i64 *d_nalive = c->d_tmp64 + 64;
    int *d_ncross = (int *)(c->d_tmp64 + 64 + g.npatch);
    a.nalive = d_nalive;
    const size_t smem = sizeof(int) * (size_t)a.kx * a.ky * a.kz;
    if (!c->perm_attr_set) {  // per context: function attributes are per device, and a process may drive several
        CUDA_TRY(cudaFuncSetAttribute(k_cell_perm, cudaFuncAttributeMaxDynamicSharedMemorySize, KEY_LIMIT * (int)sizeof(int)));
        c->perm_attr_set = true;
    }
    CUDA_TRY(cudaMemsetAsync(d_ncross, 0, sizeof(int) * g.npatch, c->stream));
    k_cell_perm<<<g.npatch, PT, smem, c->stream>>>(a);
    LAUNCHED(1);
    const int B = 128;
    Slots s = make_slots(sp);
    PushConst pk;
    pk.cdt = LPIC_C_LIGHT * 0.5 * dt;
    pk.efactor = q * dt / (2 * m * LPIC_C_LIGHT); pk.bfactor = q * dt / (2 * m);
    pk.inv_dx = 1.0 / g.dx; pk.inv_dy = 1.0 / g.dy; pk.inv_dz = 1.0 / g.dz;
    pk.q_dV = q / (g.dx * g.dy * g.dz); pk.q_dydzdt = q / (g.dy * g.dz * dt);
    pk.q_dxdzdt = q / (g.dx * g.dz * dt); pk.q_dxdydt = q / (g.dx * g.dy * dt);
    if (!three) {
        pk.q_dV = q / (g.dx * g.dy); pk.q_dydzdt = q / (g.dy * dt); pk.q_dxdzdt = q / (g.dx * dt);  // DepositCoef2's q_dxdy, q_dydt, q_dxdt
        const int bpp = (int)div_up(sp.max_npart, B);
        const unsigned grid = (unsigned)((i64)bpp * g.npatch);
        if (write_part)
            k_push_sorted2d<true><<<grid, B, 0, c->stream>>>(g, c->fields, c->d_x0, c->d_y0, c->d_z0, s, c->scr_b, d_nalive, c->scr_a,
                                                            d_ncross, bpp, dt, pk);
        else
            k_push_sorted2d<false><<<grid, B, 0, c->stream>>>(g, c->fields, c->d_x0, c->d_y0, c->d_z0, s, c->scr_b, d_nalive, c->scr_a,
                                                             d_ncross, bpp, dt, pk);
        LAUNCHED(1);
        k_deposit_list2d<<<(unsigned)std::min<i64>(g.npatch, 148 * 8), B, 0, c->stream>>>(g, c->fields, c->d_x0, c->d_y0, c->d_z0, s,
                                                                                        c->scr_a, d_ncross, dt, q);
        LAUNCHED(1);
        KERNEL_CHECK();
        return 0;
    }
    {
        const int bpp = (int)div_up(sp.max_npart, B);
        const unsigned grid = (unsigned)((i64)bpp * g.npatch);
        // LPIC_PUSH_COMPACT: keep the loops over components / stencil planes / source lanes rolled (2.1 k SASS instructions,
        // fits the 32 KB L1.5 instruction cache: no_instruction stall 3.9 -> 0 cycles/issue, but 35 % more instructions
        // executed; measured within 2 % of the unrolled kernel, which stays the default)
        const bool compact = getenv("LPIC_PUSH_COMPACT") != nullptr;  // read per call: tests toggle it
#define SORTED_LAUNCH(W, C)                                                                                                \
    k_push_sorted<W, C><<<grid, B, 0, c->stream>>>(g, c->fields, c->d_x0, c->d_y0, c->d_z0, s, c->scr_b, d_nalive, c->scr_a, \
                                                  d_ncross, bpp, dt, q, m, pk)
        if (write_part) { if (compact) SORTED_LAUNCH(true, true); else SORTED_LAUNCH(true, false); }
        else { if (compact) SORTED_LAUNCH(false, true); else SORTED_LAUNCH(false, false); }
#undef SORTED_LAUNCH
    }
    LAUNCHED(1);
    k_deposit_list<<<(unsigned)std::min<i64>(g.npatch, 148 * 8), B, 0, c->stream>>>(g, c->fields, c->d_x0, c->d_y0, c->d_z0, s, c->scr_a, d_ncross, 0, dt, q);
    LAUNCHED(1);
    KERNEL_CHECK();
    return 0;
}
